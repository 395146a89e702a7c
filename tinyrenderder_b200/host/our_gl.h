// our_gl.h - the reference's pipeline header (our_gl.h:17-86) re-implemented over the B200 backend.
// Same globals, same function names and argument meaning; what changes for a caller is listed in
// INTEGRATION.md:
//   * a GPU cannot call the host virtual IShader::fragment, so shaders the device knows describe
//     themselves through IShader::device_shader(); rasterize() with any other shader throws
//     (there is no CPU fallback);
//   * rasterize() calls are batched; gl_flush(framebuffer) resolves them and refreshes the
//     caller-visible framebuffer and the global zbuffer (the reference writes both immediately);
//   * gl_draw_model() replaces a whole per-face loop (main.cpp:660-666) by one device draw call.
#pragma once
#include <string>
#define TRB_DEVICE_BACKEND 1
#include <geometry.h>
#include <tgaimage.h>

#include <limits>
#include <utility>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

extern mat<4, 4> ModelView;          // our_gl.h:17
extern mat<4, 4> Perspective;        // our_gl.h:18
extern mat<4, 4> Viewport;           // our_gl.h:19
extern std::vector<double> zbuffer;  // our_gl.h:20 - refreshed by gl_flush / gl_zbuffer_restore

void lookat(const vec3 eye, const vec3 center, const vec3 up);                         // our_gl.cpp:25-41
void init_perspective(double fov_deg, double aspect, double znear, double zfar);       // our_gl.cpp:44-56
void init_viewport(int x, int y, int w, int h);                                        // our_gl.cpp:59-69
void init_zbuffer(int width, int height);  // our_gl.cpp:72-74; also begins a device frame of that size

class Model;

// What a device-resident shader tells the backend about itself.
struct TrbDeviceShader {
    int kind = -1;                     // TrbShaderKind
    double key_dir_eye[3] = {0, 0, 0}, fill_dir_eye[3] = {0, 0, 0}, rim_dir_eye[3] = {0, 0, 0};
    double normal_map_strength = 1.0;
    const Model* model = nullptr;      // mesh + textures (materials[0], model.cpp:416-459)
    const double* varyings = nullptr;  // immediate mode: 3 x {uv.x, uv.y, position_eye.xyz, normal_eye.xyz}
};

struct IShader {
    static TGAColor sample2D(const TGAImage& img, const vec2& uv) {  // our_gl.h:38-44
        int x = std::min<int>(img.width() - 1, std::max<int>(0, int(uv.x * img.width())));
        int y = std::min<int>(img.height() - 1, std::max<int>(0, int(uv.y * img.height())));
        return img.get(x, y);
    }
    virtual vec4 vertex(int face_index, int vertex_index) { (void)face_index; (void)vertex_index; return vec4(); }
    virtual std::pair<bool, TGAColor> fragment(const vec3 barycentric) const = 0;
    // B200 backend: fill `out` and return true when this shader has a device implementation
    virtual bool device_shader(TrbDeviceShader& out) const { (void)out; return false; }
    virtual ~IShader() {}
};

typedef vec<4> Triangle[3];  // our_gl.h:55

// our_gl.cpp:89-201.  Batched; throws std::runtime_error for a shader without device_shader().
void rasterize(const Triangle& clip, const IShader& shader, TGAImage& framebuffer);
void print_render_stats();   // our_gl.cpp:204-210 (order-independent counters, see trb.h TrbStats)

// Upstream-tinyrenderer spellings of the same three entry points (BASELINE north_star names them; the fork
// renamed them).  Thin aliases: same arguments as the fork's functions; the z-buffer the
// backend tests against is the device-resident one behind the global `zbuffer`, so that is the only
// vector triangle() accepts.
inline void projection(double fov_deg, double aspect, double znear, double zfar) { init_perspective(fov_deg, aspect, znear, zfar); }
inline void viewport(int x, int y, int w, int h) { init_viewport(x, y, w, h); }
void triangle(const Triangle& clip_verts, const IShader& shader, TGAImage& image, std::vector<double>& zbuffer_arg);

// ---- additions of the backend ---------------------------------------------------------------
// for face in model: clip[v] = shader.vertex(face, v); rasterize(clip, shader, framebuffer)  -> one draw
void gl_draw_model(const Model& model, const IShader& shader, TGAImage& framebuffer);
// resolve everything submitted so far: colours into `framebuffer`, depths into the global zbuffer
void gl_flush(TGAImage& framebuffer);
// `std::vector<double> saved = zbuffer;` ... `zbuffer = saved;` around a draw (main.cpp:700,730)
void gl_zbuffer_snapshot();
void gl_zbuffer_restore(TGAImage& framebuffer);
// post passes of main.cpp:269-362, 756-786 on the device-resident z-buffer
void gl_ssao(TGAImage& ao_map);
void gl_zbuffer_image(TGAImage& grey);
void gl_composite_ao(TGAImage& final_result);
// framebuffer.write_tga_file(name) (main.cpp:743) and the three post-pass files (main.cpp:312, 765, 785)
// without bringing the pixels to the host: the RLE packets of tgaimage.cpp:193-242 are built on the
// device, only the file image crosses PCIe.  image: 0 framebuffer, 1 z-buffer grey map, 2 ssao, 3 final.
bool gl_write_tga_file(int image, const std::string& filename);
// ---- several cameras in one launch set (a camera orbit, config 3) ----------------------------------
// gl_begin_views is init_zbuffer + the framebuffer clear for views.size() independent frames of one size.
// Until the next init_zbuffer / gl_begin_views, gl_draw_model_views draws a model into every frame (the
// per-face loop of main.cpp:652-666 with ModelView_i = views[i] * model_matrix and the WORLD-space light
// directions taken to each view's eye space like PhongShader::initLightDirections, main.cpp:55-69), and
// gl_zbuffer_snapshot / gl_zbuffer_restore act on all frames.  One launch set renders them all, which is
// what lifts small frames out of the launch-bound regime.  kind: 1 = PhongShader, 2 = EyeShader.
void gl_begin_views(const std::vector<mat<4, 4>>& views, int width, int height);
void gl_draw_model_views(const Model& model, int kind, const mat<4, 4>& model_matrix, const vec3& key_world,
                         const vec3& fill_world, const vec3& rim_world, double normal_map_strength);
// colours of frame `view` into `framebuffer` (resized to the frame), its depths into the global zbuffer
void gl_read_view(int view, TGAImage& framebuffer);
// gl_write_tga_file for every frame of the batch: filenames[i] receives frame i
bool gl_write_tga_files(int image, const std::vector<std::string>& filenames);
// Frame recordings of a batch (trb_record_begin / trb_record_end / trb_replay, trb.h): after a batch of this shape has been
// rendered once, everything between gl_record_views_begin and gl_record_views_end (gl_begin_views, gl_draw_model_views,
// gl_zbuffer_snapshot / restore) is captured into one CUDA graph per context; gl_replay_views re-runs it for other cameras
// (as many as were recorded) with ONE launch per context - the frame loop of main.cpp:606-730 for frames so small that the
// kernel launches cost more than the kernels.  gl_read_view / gl_write_tga_files then work on the replayed batch.
void gl_record_views_begin();
int gl_record_views_end();
void gl_replay_views(int recording, const std::vector<mat<4, 4>>& views);
struct TrbCtx;
TrbCtx* gl_context();        // the process-wide device context (device = $TRB_DEVICE, default 0)
// ---- several GPUs from one C++ program ---------------------------------------------------
// gl_set_devices({0, 1, 2, 3}) before the first frame: one context per entry (an entry may repeat - two contexts on one
// GPU, used by the tests).  From then on
//   * gl_begin_views splits the cameras over the contexts in contiguous blocks (config 3: frames are independent,
//     nothing is exchanged); gl_draw_model_views / gl_zbuffer_snapshot / restore act on every block, gl_read_view and
//     gl_write_tga_files address frames by their global index;
//   * init_zbuffer begins ONE picture on all contexts (config 4): gl_draw_model makes context k draw triangle range k of
//     the model with global submission indices, and gl_composite (implied by gl_flush) runs the sort-last composite of
//     the C ABI (trb_comm_init once, trb_composite_group per picture) and gathers the rows into the framebuffer and the
//     global zbuffer.  The result is bit-identical to the single-GPU picture.
void gl_set_devices(const std::vector<int>& devices);
int gl_device_count();
void gl_composite(TGAImage& framebuffer);

// Frustum culling of whole models (our_gl.h:68-86, our_gl.cpp:212-280): host side, bug-for-bug -
// the planes are extracted from the matrix as if it were transposed.
struct Frustum {
    Plane planes[6];
    enum PlaneIndex { LEFT = 0, RIGHT = 1, BOTTOM = 2, TOP = 3, NEAR = 4, FAR = 5 };
    static Frustum createFromMatrix(const mat<4, 4>& matrix);
    bool intersects(const AABB& aabb) const;
};
