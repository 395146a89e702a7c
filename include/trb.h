/*
 * trb.h - C ABI of the B200 rasterization backend ("tinyrenderder-b200").
 *
 * This is the drop-in boundary for the hot path of AnnaUshnova/tinyrenderder:
 * the reference has no FFI layer, its boundary is the C++ header our_gl.h plus
 * the Model accessors.  Every entry point below names the reference interface
 * it replaces (file:line relative to the reference checkout).
 *
 * Conventions
 *   - every call returns int: 0 = ok, <0 = error class (TRB_E_*);
 *     trb_last_error(ctx) returns the CUDA / argument error string.
 *   - the caller owns every host pointer; the library owns device memory behind
 *     opaque handles.  Host buffers may be pageable or pinned; a PINNED buffer handed to
 *     trb_upload_* must stay unchanged until the next synchronising call (trb_read_*,
 *     trb_get_stats, trb_readback_wait, trb_synchronize), pageable ones are consumed
 *     before the call returns.
 *   - matrices are row-major double[16], exactly mat<4,4>::rows of the
 *     reference (geometry.h:155-166), column-vector convention (M*v).
 *   - colours are BGR bytes (TGAColor layout, tgaimage.h:29-63); row y=0 of the
 *     framebuffer is the bottom of the picture (tgaimage.cpp:176).
 *   - one context per GPU; a context is NOT thread-safe (the reference keeps its
 *     state in unsynchronised globals, our_gl.cpp:12-22); different contexts may
 *     be driven from different host threads.  Rendering of a context is queued on
 *     one CUDA stream (uploads and pipelined read-backs on two more, ordered by
 *     events); draw calls never wait for the device.  trb_read_* / trb_get_stats /
 *     trb_encode_tga / trb_readback_wait / trb_synchronize are the synchronising calls.
 *   - there is no CPU fallback: every entry point fails with TRB_E_CUDA when no
 *     sm_100 device is usable.
 *
 * The CPU oracle (oracle/, test infrastructure only) implements the SAME
 * signatures under the prefix orc_ by defining TRB_FN before including this
 * header; the product library exports trb_*.
 */
#ifndef TRB_H_
#define TRB_H_

#include <stddef.h>
#include <stdint.h>

#ifndef TRB_FN
#define TRB_FN(name) trb_##name
#endif

#if defined(__GNUC__)
#define TRB_EXPORT __attribute__((visibility("default")))
#else
#define TRB_EXPORT
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct TrbCtx TrbCtx;
typedef uint64_t TrbMesh; /* 0 = invalid */
typedef uint64_t TrbTex;  /* 0 = "texture absent" (model.cpp:416,429,447 fallbacks) */

enum TrbError {
    TRB_OK = 0,
    TRB_E_ARG = -1,      /* bad argument / call out of sequence */
    TRB_E_CUDA = -2,     /* CUDA runtime error, see trb_last_error */
    TRB_E_NOMEM = -3,    /* device or host allocation failed */
    TRB_E_SHADER = -4,   /* shader kind unknown to the device (no CPU fallback) */
    TRB_E_COMM = -5      /* multi-GPU composite misuse */
};

/* Device-resident fragment shaders.  A GPU cannot call the host virtual
 * IShader::fragment (our_gl.h:51), so the shaders of the reference are
 * identified by kind + a POD uniform block. */
enum TrbShaderKind {
    TRB_SHADER_FLAT_BARY = 0, /* BGR = (u8)min(255,max(0,255*pc[i])), test shader of SURVEY K1-K7 / C5 */
    TRB_SHADER_PHONG = 1,     /* PhongShader, main.cpp:39-171 */
    TRB_SHADER_EYE = 2,       /* EyeShader,   main.cpp:176-262 */
    TRB_SHADER_DEPTH = 3,     /* depth-only pass (no colour write), shadow pass 1 of config 2 */
    TRB_SHADER_SHADOW_PHONG = 4, /* Phong + shadow-map lookup, config 2 (authored here) */
    TRB_SHADER_GOURAUD = 5    /* per-vertex diffuse intensity interpolated, config 2 family */
};

/* Uniform block of PHONG / EYE / GOURAUD.  Light directions are ALREADY in eye
 * space and normalised, i.e. the result of PhongShader::initLightDirections
 * (main.cpp:55-69) / EyeShader::initLightDirections (main.cpp:187-197); use
 * trb_light_dir_eye() to compute them with the reference's operation order. */
typedef struct TrbPhongUniforms {
    double key_dir_eye[3];
    double fill_dir_eye[3];        /* unused by EYE */
    double rim_dir_eye[3];
    double normal_map_strength;    /* PhongShader::normal_map_strength, main.cpp:51 */
    TrbTex diffuse;                /* Model::diffuse  source, model.cpp:415-425 */
    TrbTex normal;                 /* Model::normal   source, model.cpp:428-444 */
    TrbTex specular;               /* Model::specular source, model.cpp:446-459 */
} TrbPhongUniforms;

/* Uniform block of SHADOW_PHONG: Phong block + the light-space transform used
 * by the depth pass and the shadow map it produced (a depth snapshot handle). */
typedef struct TrbShadowUniforms {
    TrbPhongUniforms phong;
    double light_modelview[16];    /* ModelView of the depth pass */
    double light_perspective[16];  /* Perspective of the depth pass */
    double light_viewport[16];     /* Viewport of the depth pass */
    double shadow_bias;            /* NDC-z bias; lit if z_light <= shadow_z + bias */
    double shadow_darkening;       /* multiplier applied to diffuse+specular when occluded */
    int32_t shadow_map;            /* index returned by trb_keep_depth_as_shadow_map */
    int32_t shadow_w, shadow_h;
    int32_t _pad;
} TrbShadowUniforms;

/* Counters.  triangles_submitted / bbox / z_min mirror the statics of
 * our_gl.cpp:18-22 printed by print_render_stats (our_gl.cpp:204-210).
 * The reference's fragments_drawn and max_z depend on submission order
 * (SURVEY K5), so the order-independent equivalents are reported instead. */
typedef struct TrbStats {
    uint64_t triangles_submitted;  /* == triangles_rasterized, our_gl.cpp:90 */
    uint64_t triangles_binned;     /* survived all rejects of our_gl.cpp:94-135 */
    uint64_t tile_entries;         /* R: (triangle,16x16 tile) pairs written by binning */
    uint64_t fragments_covered;    /* samples passing our_gl.cpp:152 and :160 */
    uint64_t pixels_shaded;        /* pixels whose colour was written by flush */
    uint64_t visible_triangles;    /* distinct winning triangles at flush (0 if not counted) */
    int32_t bbox_min_x, bbox_min_y, bbox_max_x, bbox_max_y; /* our_gl.cpp:138-141 */
    double z_min;                  /* == min_z of our_gl.cpp:197 (order independent) */
    double z_max_covered;          /* largest finite depth left in the z-buffer (visible surfaces) */
    uint64_t fragments_drawn_ref;  /* oracle only: the reference's order-dependent counter */
    double z_max_ref;              /* oracle only: the reference's order-dependent max_z */
} TrbStats;

/* ---- contexts -------------------------------------------------------------------- */
TRB_EXPORT int TRB_FN(create)(int device, TrbCtx** out);
TRB_EXPORT int TRB_FN(destroy)(TrbCtx* ctx);
TRB_EXPORT const char* TRB_FN(last_error)(TrbCtx* ctx);
TRB_EXPORT const char* TRB_FN(backend_name)(void); /* "cuda-sm100a", "oracle-ref", "oracle-port" */

/* ---- resources: upload once (replaces Model's host arrays, model.h:118-121, and the
 *      accessors Model::vert/normal/uv(iface,nthvert), model.cpp:396-412) --------------
 * pos3/nrm3: nverts*3 floats, uv2: nverts*2 floats (Assimp hands back floats,
 * model.cpp:160-185).  nrm3 / uv2 may be NULL -> (0,0,1) / (0,0) like the accessors'
 * fallbacks.  idx: nidx (multiple of 3) vertex indices; NULL = implicit 0..nidx-1.
 * Uploads run on their own stream, so they overlap rendering that is already queued.
 * Pageable arrays are consumed (staged through pinned chunks) when the call returns.
 * Page-locked arrays (cudaHostAlloc / cudaHostRegister) are read by the copy engine directly,
 * without a CPU copy: the caller keeps them alive and unchanged until the next synchronising
 * call (synchronize, read_*, readback_wait, get_stats).  Same for upload_texture.
 * Indexed meshes of 2 M triangles or more (TRB_MESH_ORDER_MIN_TRIS overrides; 0 = never) also get a PROCESSING ORDER
 * at upload: the triangles sorted by the Morton code of their centroids (device-side, on the upload stream).  Draws
 * visit the triangles in that order so that the vertex records a triangle gathers are the ones its predecessors just
 * used; triangle ids - and with them every output bit - stay those of the index buffer.  The vertices of such a mesh are
 * renumbered in the same spirit (internal: no entry point exposes vertex numbers), and a SOUP of that size (idx == NULL)
 * has its vertex arrays put into the processing order themselves. */
TRB_EXPORT int TRB_FN(upload_mesh)(TrbCtx* ctx, const float* pos3, const float* nrm3,
                                   const float* uv2, uint32_t nverts, const uint32_t* idx,
                                   uint64_t nidx, TrbMesh* out);
/* Lifetime: a mesh / texture may be freed at any time, also between a draw that uses it and the flush
 * that shades it - free_* then resolves the pending draws first (an implicit trb_flush), so the
 * deferred shade pass never reads a recycled block. */
TRB_EXPORT int TRB_FN(free_mesh)(TrbCtx* ctx, TrbMesh mesh);
/* texels: h rows of w texels of bpp (1,3,4) bytes in TGAImage memory order
 * ((x+y*w)*bpp, BGR(A), tgaimage.cpp:24-30); sampled nearest with trunc+clamp like
 * model.cpp:420-423. */
TRB_EXPORT int TRB_FN(upload_texture)(TrbCtx* ctx, const uint8_t* texels, int w, int h, int bpp,
                                      TrbTex* out);
TRB_EXPORT int TRB_FN(free_texture)(TrbCtx* ctx, TrbTex tex);

/* ---- frame ------------------------------------------------------------------------- */
/* init_zbuffer(w,h) (our_gl.cpp:72-74) + TGAImage framebuffer(w,h,RGB) (main.cpp:606):
 * depth = +inf, colour = clear colour (default 0,0,0), counters reset.
 * nviews > 1 renders a batch of independent frames (camera orbit) in one launch set. */
TRB_EXPORT int TRB_FN(begin_frame)(TrbCtx* ctx, int width, int height);
TRB_EXPORT int TRB_FN(begin_batch)(TrbCtx* ctx, int width, int height, int nviews);
TRB_EXPORT int TRB_FN(set_clear_color)(TrbCtx* ctx, uint8_t b, uint8_t g, uint8_t r);
/* the global Viewport (our_gl.h:19) as set by init_viewport (our_gl.cpp:59-69) */
TRB_EXPORT int TRB_FN(set_viewport)(TrbCtx* ctx, const double viewport[16]);

/* ---- draw: one call replaces the per-face loop main.cpp:660-666 / 692-698 / 715-721 -- */
/* for face in [first_tri, first_tri+ntris): clip[v] = shader.vertex(face,v);
 * rasterize(clip, shader, framebuffer).  modelview/perspective are the globals
 * ModelView / Perspective (our_gl.h:17-18) at the time of the loop. */
TRB_EXPORT int TRB_FN(draw)(TrbCtx* ctx, TrbMesh mesh, const double modelview[16],
                            const double perspective[16], int shader_kind, const void* uniforms,
                            size_t uniform_bytes, uint64_t first_tri, uint64_t ntris);
/* batch form: modelview / perspective are nviews*16 doubles, uniforms nviews blocks */
TRB_EXPORT int TRB_FN(draw_batch)(TrbCtx* ctx, TrbMesh mesh, const double* modelview,
                                  const double* perspective, int shader_kind,
                                  const void* uniforms, size_t uniform_bytes, uint64_t first_tri,
                                  uint64_t ntris);
/* One rank's share of a mesh that shard_count ranks draw together into one picture (sort-last, config 4): the whole
 * mesh is submitted, this context rasterises the triangles that fall to shard_rank.  Triangle ids are those of the whole
 * mesh on every rank (the draw consumes the mesh's id range; trb_set_triangle_id_base is not needed), so after
 * trb_composite the picture is bit for bit what one context drawing the mesh with trb_draw produces, depth ties included
 * (our_gl.cpp:160-166).  Which triangles a rank gets is the backend's choice: meshes that carry a processing order
 * (large indexed meshes, see trb_upload_mesh) are dealt out in blocks of 4096 consecutive positions of that order -
 * every rank's share is spatially coherent and evenly spread over the surface; other meshes are split into contiguous
 * ranges of the index buffer.  Single-view frames only; replaces the loop main.cpp:660-666 on N devices. */
TRB_EXPORT int TRB_FN(draw_shard)(TrbCtx* ctx, TrbMesh mesh, const double modelview[16], const double perspective[16],
                                  int shader_kind, const void* uniforms, size_t uniform_bytes, int shard_rank,
                                  int shard_count);
/* immediate mode behind rasterize(const Triangle&, const IShader&, TGAImage&)
 * (our_gl.cpp:89): n clip-space triangles, clip12 = n*12 doubles (3 x vec4).
 * varyings (may be NULL for FLAT_BARY/DEPTH): n*24 doubles per triangle =
 * 3 x {uv.x, uv.y, position_eye.xyz, normal_eye.xyz}, the state PhongShader::vertex
 * leaves in the shader object (main.cpp:75-87). modelview is read by
 * PhongShader::fragment (main.cpp:116). */
TRB_EXPORT int TRB_FN(submit_clip_triangles)(TrbCtx* ctx, const double* clip12,
                                             const double* varyings, uint64_t n,
                                             const double modelview[16], int shader_kind,
                                             const void* uniforms, size_t uniform_bytes);

/* z-buffer copy at a draw boundary (std::vector<double> zbuffer_before_eyes = zbuffer,
 * main.cpp:700) and roll-back (zbuffer = zbuffer_before_eyes, main.cpp:730).  Restore
 * shades everything drawn so far first, so colours persist like in the reference.
 * Cost: the snapshot is tile granular - the draws between the two calls save the 16x16 tiles they are about to change
 * and the restore copies those back, so a few small draws (the eyes) cost a few tiles, not two passes over the plane.
 * Results are those of the full copy (TRB_LAZY_SNAPSHOT=0 selects it). */
TRB_EXPORT int TRB_FN(depth_snapshot)(TrbCtx* ctx);
TRB_EXPORT int TRB_FN(depth_restore)(TrbCtx* ctx);
/* keep the current depth buffer of view 0 as a shadow map for later frames */
TRB_EXPORT int TRB_FN(keep_depth_as_shadow_map)(TrbCtx* ctx, int32_t* out_index);
/* drop every kept shadow map (indices restart at 0) */
TRB_EXPORT int TRB_FN(release_shadow_maps)(TrbCtx* ctx);

/* resolve visibility -> colour (the fragment() calls of our_gl.cpp:187-192, run once per
 * visible pixel).  Implied by end_frame / read_color / depth_restore. */
TRB_EXPORT int TRB_FN(flush)(TrbCtx* ctx);
TRB_EXPORT int TRB_FN(end_frame)(TrbCtx* ctx);

/* ---- frame recordings: a launch-bound frame loop as one CUDA graph launch per frame -------------------------
 * A small frame (config 1: 2 520 triangles at 800x800; config 2: two passes at 2048x2048) costs ~40 kernel launches
 * of a few microseconds each: the GPU waits for the host, not the other way round.  A recording captures everything
 * the calls between record_begin and record_end queue (begin_frame / begin_batch, set_viewport, draw*, depth_snapshot
 * / restore, keep_depth_as_shadow_map, flush) into a CUDA graph; trb_replay then re-runs the whole frame with ONE
 * launch - optionally with new matrices and uniform blocks, i.e. the next camera of main.cpp's frame loop
 * (main.cpp:647-730 with a moving camera).  Results are bit for bit those of issuing the calls again.
 *   - The frame must have been rendered once by plain calls before it is recorded (same sizes, same meshes): while
 *     recording, no buffer may grow (TRB_E_ARG "render the frame once before recording it").
 *   - Calls that synchronise, upload, read back or use other streams are refused inside a recording; record_end
 *     resolves the frame (an implicit flush), runs it once, and leaves the context as the calls would have.
 *   - A recording holds device addresses: it goes stale (trb_replay -> TRB_E_ARG) when a mesh or texture is freed
 *     or any working buffer of the context is reallocated (a larger frame, a larger draw) after it was made.
 *   - replay resolves the frame in flight, then IS a frame: read_* / readback_async / encode_tga* work on it as
 *     usual, and a shadow map the recorded frame kept is held again (release it as after the recorded frame).
 *   - `draws` (may be NULL = replay unchanged): one entry per trb_draw* call of the recording, in order; NULL
 *     members keep what was recorded.  Matrices / uniform blocks are [nviews] arrays as in trb_draw_batch; texture
 *     handles and shadow-map indices in a new uniform block are resolved again.  Rewriting parameters waits for the
 *     previous replay of the same recording to have started its last copy (host-side, microseconds). */
typedef uint64_t TrbRecording; /* 0 = invalid */
typedef struct TrbReplayDraw {
    const double* modelview;    /* [nviews][16] or NULL */
    const double* perspective;  /* [nviews][16] or NULL */
    const void* uniforms;       /* [nviews] uniform blocks of the draw's shader kind, or NULL */
    size_t uniform_bytes;       /* size of ONE block (as passed to trb_draw) */
} TrbReplayDraw;
TRB_EXPORT int TRB_FN(record_begin)(TrbCtx* ctx);
TRB_EXPORT int TRB_FN(record_end)(TrbCtx* ctx, TrbRecording* out);
TRB_EXPORT int TRB_FN(replay)(TrbCtx* ctx, TrbRecording recording, const TrbReplayDraw* draws, int ndraws);
TRB_EXPORT int TRB_FN(recording_free)(TrbCtx* ctx, TrbRecording recording);

/* ---- post passes on the resident z-buffer ------------------------- */
/* compute_ssao_at over the frame (main.cpp:324-362, 756-763): ao[x+y*w] = (u8)(255*ao) */
TRB_EXPORT int TRB_FN(ssao)(TrbCtx* ctx, int view, uint8_t* ao_out);
/* save_zbuffer_image grey map (main.cpp:269-314) without the file write */
TRB_EXPORT int TRB_FN(depth_image)(TrbCtx* ctx, int view, uint8_t* grey_out);
/* final = phong * ao (main.cpp:768-783), BGR out */
TRB_EXPORT int TRB_FN(composite_ao)(TrbCtx* ctx, int view, uint8_t* bgr_out);

/* ---- TGA files of the frame ---------------------------------------- */
/* The images main.cpp writes out: framebuffer.tga (main.cpp:743), zbuffer.tga (:312), ssao.tga
 * (:765), final.tga (:785). */
enum TrbImage {
    TRB_IMAGE_COLOR = 0,  /* BGR framebuffer, 24 bit */
    TRB_IMAGE_DEPTH = 1,  /* save_zbuffer_image grey map, stored as TGAColor(v,v,v): 24 bit like the reference's file */
    TRB_IMAGE_SSAO = 2,   /* ambient-occlusion map, likewise 24 bit */
    TRB_IMAGE_FINAL = 3   /* framebuffer * ao, 24 bit */
};
/* Replaces TGAImage::write_tga_file(name, vflip = true, rle = true) (tgaimage.cpp:160-191) and its
 * packetiser unload_rle_data (tgaimage.cpp:193-242) for every view of the batch: out[v] receives the
 * complete file image - 18-byte header + run-length packets, byte for byte what the reference
 * writes - and sizes[v] its length.  The packets are built on the device, so only the compressed
 * bytes cross PCIe.  out[v] may be NULL (size query only); `capacity` is the size of each out[v]
 * (width*height*bpp + width*height/2 + 19 always suffices:
 * the shortest packet the encoder emits away from the image end is a raw packet of two pixels).  Synchronises. */
TRB_EXPORT int TRB_FN(encode_tga)(TrbCtx* ctx, int which, uint8_t* const* out, uint64_t capacity, uint64_t* sizes);
/* The asynchronous frame writer (what a 1024-frame orbit that ends in framebuffer.write_tga_file, main.cpp:743, needs):
 * same output as trb_encode_tga, but the call only queues work.  The packets are built behind the frame on the
 * context's stream; the per-view sizes travel to pinned host memory and, as soon as the host has them - at the latest
 * inside the next trb_encode_tga_async or trb_readback_wait - the packets follow on the copy stream, overlapping the
 * rendering of the next frames.  out[v] (page-locked for real overlap) and sizes[v] must stay valid and untouched until
 * trb_readback_wait returns; two encodes may be in flight, a third call first waits for the oldest. */
TRB_EXPORT int TRB_FN(encode_tga_async)(TrbCtx* ctx, int which, uint8_t* const* out, uint64_t capacity, uint64_t* sizes);

/* ---- readback ------------------------------------------------------------------------ */
/* framebuffer bytes, BGR, (x+y*w)*3 like TGAImage(w,h,RGB) (tgaimage.cpp:32-39) */
TRB_EXPORT int TRB_FN(read_color)(TrbCtx* ctx, int view, uint8_t* bgr_out);
/* The inverse: replace the view's framebuffer by a host image in TGAImage order (the caller owns a
 * TGAImage in the reference and may set() pixels itself, tgaimage.cpp:32-39; this is the bulk form
 * of that).  Pending draws are resolved first.  Depth is untouched. */
TRB_EXPORT int TRB_FN(write_color)(TrbCtx* ctx, int view, const uint8_t* bgr);
/* the global zbuffer (our_gl.h:20): w*h doubles, +inf where nothing was drawn */
TRB_EXPORT int TRB_FN(read_depth)(TrbCtx* ctx, int view, double* z_out);
/* winning triangle id per pixel BEFORE flush: 0xFFFFFFFF = none, 0 = already shaded,
 * else 1 + global submission index over all draws since begin_frame (diagnostic) */
TRB_EXPORT int TRB_FN(read_visibility)(TrbCtx* ctx, int view, uint32_t* id_out);
/* Pipelined readback of EVERY view of the frame: BGR planes into color_out[v] and f64 depths into
 * depth_out[v] (either array may be NULL).  The frame is resolved and snapshotted into a device
 * staging area on the context's stream; the device->host copies run on a second stream, so they
 * overlap the next frame's rendering.  Returns after queueing: the host buffers (pinned, for real
 * overlap) are valid after trb_readback_wait.  Three readbacks may be in flight; a fourth call first
 * waits for the oldest. */
TRB_EXPORT int TRB_FN(readback_async)(TrbCtx* ctx, uint8_t* const* color_out, double* const* depth_out);
TRB_EXPORT int TRB_FN(readback_wait)(TrbCtx* ctx);
TRB_EXPORT int TRB_FN(get_stats)(TrbCtx* ctx, int view, TrbStats* out);
TRB_EXPORT int TRB_FN(synchronize)(TrbCtx* ctx);

/* device-side timing of everything queued between the two marks (CUDA events on the
 * context's stream); used by bench.py. */
TRB_EXPORT int TRB_FN(timer_start)(TrbCtx* ctx);
TRB_EXPORT int TRB_FN(timer_stop_ms)(TrbCtx* ctx, float* ms_out);
/* per-kernel accumulated device time and launch counts since the last reset
 * (CUDA events around each launch; enabled by trb_profile_enable(ctx,1)). */
typedef struct TrbKernelTime {
    char name[32];
    uint64_t launches;
    double ms;
} TrbKernelTime;
TRB_EXPORT int TRB_FN(profile_enable)(TrbCtx* ctx, int on);
TRB_EXPORT int TRB_FN(profile_read)(TrbCtx* ctx, TrbKernelTime* out, int capacity, int* n_out,
                                    int reset);
TRB_EXPORT uint64_t TRB_FN(launch_count)(TrbCtx* ctx); /* kernels launched since create */

/* ---- multi-GPU sort-last composite (config 4) ----------------------------------------
 * Each rank draws a triangle range with global ids (trb_set_triangle_id_base) and does NOT
 * flush.  Then, with the raw device pointers of view 0's depth-key (uint64) and
 * visibility-id (uint32) planes from trb_device_planes:
 *   1. trb_composite_save_local_depth  keeps a copy of the local keys and rewrites the plane as
 *                                      int64-sortable so that a signed MIN all-reduce is exact;
 *   2. host: all-reduce(MIN, int64) over the key plane (NCCL over NVLink);
 *   3. trb_composite_mask              restores the key encoding; ids whose local key lost (or
 *                                      that are empty) become 0x7FFFFFFF (largest int32);
 *   4. host: all-reduce(MIN, int32) over the id plane - lowest id = first submitted, the
 *                                      reference's tie rule (our_gl.cpp:165);
 *   5. trb_composite_finish            0x7FFFFFFF -> "none"; trb_set_shade_rows + trb_flush shade
 *                                      the slice this rank owns (every rank holds all meshes). */
TRB_EXPORT int TRB_FN(device_planes)(TrbCtx* ctx, uint64_t* depth_key_ptr, uint64_t* vis_id_ptr,
                                     uint64_t* npixels);
TRB_EXPORT int TRB_FN(set_triangle_id_base)(TrbCtx* ctx, uint64_t base);
TRB_EXPORT int TRB_FN(composite_save_local_depth)(TrbCtx* ctx);
TRB_EXPORT int TRB_FN(composite_mask)(TrbCtx* ctx);
TRB_EXPORT int TRB_FN(composite_finish)(TrbCtx* ctx);
/* Fused NVLink composite (the sm_100a-native form of steps 1-5): every rank exports CUDA IPC
 * handles of its two planes (trb_ipc_export_planes, 64 bytes each), the host exchanges them
 * (e.g. torch.distributed.all_gather_object) and every rank opens its peers' planes once
 * (trb_ipc_open_peers; handles[r] of the calling rank is ignored).  Then, per frame, after ALL ranks
 * finished their draws (host barrier), trb_composite_shade_p2p runs ONE kernel that, for the rows
 * [y0,y1) this rank owns, loads the n candidate (depth key, id) pairs straight from the peers' HBM
 * over NVLink, keeps the exact lexicographic minimum (depth, then lowest id = first submitted) and
 * shades the winner in the same pass.  Peers must not start their next frame before everybody is
 * done reading (second host barrier). */
TRB_EXPORT int TRB_FN(ipc_export_planes)(TrbCtx* ctx, void* key_handle64, void* vis_handle64);
TRB_EXPORT int TRB_FN(ipc_open_peers)(TrbCtx* ctx, const void* key_handles, const void* vis_handles, int n,
                                      int my_rank);
/* same for ranks that live in ONE process (contexts on one or several GPUs with peer access):
 * plain device pointers as returned by trb_device_planes, no IPC */
TRB_EXPORT int TRB_FN(open_peers_raw)(TrbCtx* ctx, const uint64_t* key_ptrs, const uint64_t* vis_ptrs, int n,
                                      int my_rank);
TRB_EXPORT int TRB_FN(ipc_close_peers)(TrbCtx* ctx);
TRB_EXPORT int TRB_FN(composite_shade_p2p)(TrbCtx* ctx, int y0, int y1);
/* ---- composite groups: trb_comm_init(ctx[], n) / trb_composite(ctx) of SURVEY 8(b) -----------------------
 * The fused NVLink composite above without any host-side barrier.  A group is n ranks (one context each) that
 * render triangle ranges of ONE picture; every member must have begun a single-view frame of the final size before the
 * group is formed, and keeps that size while the group exists.
 *   one process, n contexts (one per GPU, or several on one GPU):   trb_comm_init(ctxs, n)
 *   one process per rank:   trb_comm_export(ctx, blob)  ->  exchange the TRB_COMM_BLOB_BYTES-byte blobs by any means
 *                           (MPI, files, torch.distributed) into an array ordered by rank  ->  trb_comm_open(ctx, blobs, n, rank)
 * Per frame every rank: trb_begin_frame, trb_comm_shard -> trb_set_triangle_id_base(first) + trb_draw(first, count)
 * (ids stay global, so ties resolve like one sequential submission, our_gl.cpp:165), then trb_composite(ctx): the
 * rank's stream publishes "drawn" in a counter in its own HBM, waits (a one-thread kernel) until every peer has
 * published the same frame, runs ONE kernel that reads the n candidate (depth key, id) pairs of each pixel of the rows
 * trb_comm_rows gives it straight from the peers' HBM, keeps the exact lexicographic minimum and shades it, and
 * publishes "done reading"; the next trb_begin_frame waits for the peers' "done" before it clears the planes.  No call
 * blocks the host; trb_read_color etc. of the owned rows are stream ordered as usual.  A peer that stays silent for
 * 20 s makes the next call fail with TRB_E_COMM instead of hanging the GPU.
 * Contexts that share a GPU (tests, small machines) cannot wait for each other with kernels: composite them with
 * trb_composite_group(ctxs, n), which orders the streams with events - one host thread, all members in one call. */
#define TRB_COMM_BLOB_BYTES 256
TRB_EXPORT int TRB_FN(comm_init)(TrbCtx* const* ctxs, int n);
TRB_EXPORT int TRB_FN(comm_export)(TrbCtx* ctx, void* blob, size_t blob_bytes);
TRB_EXPORT int TRB_FN(comm_open)(TrbCtx* ctx, const void* blobs, int n, int rank);
TRB_EXPORT int TRB_FN(comm_close)(TrbCtx* ctx);
TRB_EXPORT int TRB_FN(comm_shard)(TrbCtx* ctx, uint64_t total_triangles, uint64_t* first, uint64_t* count);
TRB_EXPORT int TRB_FN(comm_rows)(TrbCtx* ctx, int* y0, int* y1);
TRB_EXPORT int TRB_FN(composite)(TrbCtx* ctx);
TRB_EXPORT int TRB_FN(composite_group)(TrbCtx* const* ctxs, int n);
/* restrict flush to rows [y0,y1) (the screen slice this rank owns after the composite) */
TRB_EXPORT int TRB_FN(set_shade_rows)(TrbCtx* ctx, int y0, int y1);

/* ---- host helpers with the reference's operation order (no device work) --------------- */
/* normalized(ModelView[0..2][0..2] * dir_world), main.cpp:59-68 */
TRB_EXPORT void TRB_FN(light_dir_eye)(const double modelview[16], const double dir_world[3],
                                      double out[3]);
/* lookat (our_gl.cpp:25-41), init_perspective (our_gl.cpp:44-56), init_viewport
 * (our_gl.cpp:59-69) writing a row-major 4x4 */
TRB_EXPORT void TRB_FN(lookat)(const double eye[3], const double center[3], const double up[3],
                               double out[16]);
TRB_EXPORT void TRB_FN(perspective)(double fov_deg, double aspect, double znear, double zfar,
                                    double out[16]);
TRB_EXPORT void TRB_FN(viewport)(int x, int y, int w, int h, double out[16]);
/* batch forms for many views at once (camera orbits): out[v] = a[v] * b, and the eye-space direction
 * of one world direction under every modelview */
TRB_EXPORT void TRB_FN(mat4_mul_batch)(const double* a, int n, const double b[16], double* out);
TRB_EXPORT void TRB_FN(light_dir_eye_batch)(const double* modelviews, int n, const double dir_world[3],
                                            double* out);
/* Model-level frustum culling, bug-for-bug with the reference (host scalars, no device work).
 * frustum_planes = Frustum::createFromMatrix (our_gl.cpp:212-261, which reads the planes of the TRANSPOSED matrix):
 * six planes as {nx, ny, nz, d}, order left, right, bottom, top, near, far.  frustum_intersects = Frustum::intersects
 * (our_gl.cpp:263-280) -> 1 / 0.  aabb_transform = AABB::transform (geometry.h:297-327: eight corners through the
 * matrix with the divide by w).  cull_batch answers, for n cameras at once, what main() asks per model:
 * Frustum::createFromMatrix(Perspective * view[i]).intersects(box) (main.cpp:623-624, 647, 680, 706). */
TRB_EXPORT void TRB_FN(frustum_planes)(const double view_projection[16], double planes24[24]);
TRB_EXPORT int TRB_FN(frustum_intersects)(const double planes24[24], const double box_min[3], const double box_max[3]);
TRB_EXPORT void TRB_FN(aabb_transform)(const double box_min[3], const double box_max[3], const double m[16],
                                       double out_min[3], double out_max[3]);
TRB_EXPORT void TRB_FN(cull_batch)(const double perspective[16], const double* views, int n, const double box_min[3],
                                   const double box_max[3], uint8_t* visible_out);
/* mat<4,4> * mat<4,4> (geometry.h:195-205) */
TRB_EXPORT void TRB_FN(mat4_mul)(const double a[16], const double b[16], double out[16]);

#ifdef __cplusplus
}
#endif
#endif /* TRB_H_ */
