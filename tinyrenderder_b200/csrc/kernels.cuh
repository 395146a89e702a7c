// kernels.cuh - CUDA kernels of the rasterization path for sm_100a (B200).
//
// Pipeline per draw call (one call == the per-face loop of main.cpp:660-666):
//   k_vertex_*     per unique vertex: ModelView/Perspective/perspective divide/Viewport -> VRec (32 B)
//   k_setup_count  per triangle: rejects of our_gl.cpp:94-135, pixel bbox -> 16x16 tile range,
//                  warp-aggregated per-tile counts, statistics bbox
//   k_scan_*       exclusive prefix sum of the tile counts -> bin offsets
//   k_fill         per triangle: warp-aggregated slot claim, triangle id written into its tiles' bins
//   k_raster       one CTA per 16x16 tile: the tile's depth keys + ids staged in shared memory,
//                  exact per-sample evaluation, order-independent (depth, id) resolve
//   k_shade        (flush) one thread per pixel: winner's barycentrics recomputed, fragment shader run
//                  once per visible pixel, BGR written
// All arithmetic is in exact.cuh; nothing here reorders a floating-point operation.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>      // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, nothing links libcuda)
#include <cassert>
#include "exact.cuh"
// -DTRB_DEBUG_CHECKS builds a variant whose kernels assert their own indexing invariants (bin slots,
// shared-memory tile indices, packet offsets); tests/test_gpu_parity.py runs against it when
// TRB_CUDA_LIB points at that build.  compute-sanitizer is not available on the GPU pool.
#if defined(TRB_DEBUG_CHECKS)
#define TRB_CHECK(cond) assert(cond)
#else
#define TRB_CHECK(cond) ((void)0)
#endif
#include "fastshade.cuh"

namespace trbk {
using namespace trbx;

constexpr int MAX_PEERS = 16;
constexpr int MAX_PEERS_COMM = MAX_PEERS;
constexpr int TILE = 16;
constexpr int TILE_SHIFT = 4;
constexpr int TPB = 256;               // threads per block everywhere (== pixels per tile)
constexpr uint32_t VIS_NONE = 0xFFFFFFFFu;
constexpr uint32_t VIS_SHADED = 0u;    // pixel already shaded by an earlier flush; ties keep it
constexpr int QCAP = 1024;             // candidate queue entries per CTA
constexpr int BIG_NS_DEFAULT = 16;     // triangles covering >= this many samples of a tile take the
                                       // pixel-owner path (no atomics); TRB_BIG_NS overrides

struct DevStats {
    unsigned long long tri_binned, tile_entries, frag_covered, pixels_shaded;
    unsigned long long touched;   // upper bound of the pixels with an unshaded winner (picks the shade kernel)
    unsigned long long list_len;  // length of the compacted pixel list of the current flush
    int shade_mode;               // 1: sparse frame, shade through the list; 0: dense, one thread per pixel
    int bx0, by0, bx1, by1;
    unsigned long long zmin_key, zmax_key;
};

struct FrameDev {
    int W, H, tw, th, ntiles, nviews;
    unsigned long long npix;
    unsigned long long* zkey;  // [nviews][npix] order-preserving depth keys
    uint32_t* vis;             // [nviews][npix] winning triangle id
    uint8_t* color;            // [nviews][npix][3] BGR
    DevStats* stats;           // [nviews]
    double viewport[16];
};

struct ShadowUniformsDev {
    LitUniforms lit;
    ShadowParams shadow;
};

struct DrawDev {               // one draw call, kept until flush (the shade kernel needs it)
    uint32_t id_base;          // vis id of local triangle t is id_base + t + 1
    uint32_t ntris;
    uint32_t first_tri;        // offset (in triangles) into idx
    uint32_t nverts;
    const uint32_t* idx;       // nullptr: implicit soup, vertex 3t+k
    const float* attr8;        // [nverts][8] pos,nrm,uv (mesh draws)
    const VRec* vrec;          // [nviews][nverts]
    const double* mats;        // [nviews][32] ModelView, Perspective
    const void* uniforms;      // [nviews] LitUniforms (PHONG, EYE, GOURAUD) or ShadowUniformsDev (SHADOW_PHONG)
    const double* varyings;    // immediate mode: [ntris][24]
    const void* litf;          // [nviews] trbf::LitF: fp32 copies for the fp32 lighting path (mesh draws of lit kinds)
    int kind;
    uint32_t mesh_ntris;       // triangles of the whole mesh (ids of ranges other ranks drew map here too)
    long long mesh_id_base;    // id of mesh triangle g is mesh_id_base + g + 1  (= id_base - first_tri)
    const uint32_t* inv_perm;  // ordered soup (idx == nullptr): triangle g sits in slot inv_perm[g], vertices 3 * slot + k
    // trb_draw_shard inside a composite group: the vertex stage only ran for the vertices of this rank's share
    // (vmark[v] != 0); the composite's shade pass computes the record of any other vertex it meets from pos4
    const uint8_t* vmark;
    const float4* pos4;
};

__device__ __forceinline__ uint32_t vertex_index(const uint32_t* idx, uint32_t first_tri, uint32_t t, int k) {
    size_t e = ((size_t)first_tri + t) * 3 + k;
    return idx ? __ldg(idx + e) : (uint32_t)e;
}
__device__ __forceinline__ VRec load_vrec(const VRec* p) {
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = __ldg(q), b = __ldg(q + 1);
    VRec r;
    r.sx = a.x; r.sy = a.y; r.z = b.x; r.iw = b.y;
    return r;
}
__device__ __forceinline__ void store_vrec(VRec* p, const VRec& r) {
    double2* q = reinterpret_cast<double2*>(p);
    q[0] = make_double2(r.sx, r.sy);
    q[1] = make_double2(r.z, r.iw);
}

// ---------------------------------------------------------------------------------------------
// frame clear: init_zbuffer (our_gl.cpp:72-74) + TGAImage(w,h,RGB,clear) (tgaimage.cpp:8-17)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_clear(FrameDev f, uint8_t cb, uint8_t cg, uint8_t cr) {
    unsigned long long total = f.npix * f.nviews;
    unsigned long long stride = (unsigned long long)gridDim.x * TPB;
    for (unsigned long long i = (unsigned long long)blockIdx.x * TPB + threadIdx.x; i < total; i += stride) {
        f.zkey[i] = KEY_PLUS_INF;
        f.vis[i] = VIS_NONE;
        f.color[3 * i] = cb;
        f.color[3 * i + 1] = cg;
        f.color[3 * i + 2] = cr;
    }
    for (int v = threadIdx.x; blockIdx.x == 0 && v < f.nviews; v += TPB) {
        DevStats s;
        s.tri_binned = s.tile_entries = s.frag_covered = s.pixels_shaded = 0;
        s.touched = s.list_len = 0;
        s.shade_mode = 0;
        s.bx0 = s.by0 = INT_MAX;
        s.bx1 = s.by1 = INT_MIN;
        s.zmin_key = ~0ull;
        s.zmax_key = 0ull;
        f.stats[v] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// mesh upload from pinned host arrays: the raw pos3 / nrm3 / uv2 arrays are DMA'd as they are and
// interleaved here into the two device layouts (float4 positions, 32-byte attribute records)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_interleave_mesh(const float* __restrict__ pos3, const float* __restrict__ nrm3,
                                                         const float* __restrict__ uv2, uint32_t nverts,
                                                         float4* __restrict__ pos4, float* __restrict__ attr8) {
    const uint32_t v = blockIdx.x * TPB + threadIdx.x;
    if (v >= nverts) return;
    const float x = pos3[3 * (size_t)v], y = pos3[3 * (size_t)v + 1], z = pos3[3 * (size_t)v + 2];
    pos4[v] = make_float4(x, y, z, 1.0f);
    float4 a0, a1;
    a0.x = x; a0.y = y; a0.z = z;
    a0.w = nrm3 ? nrm3[3 * (size_t)v] : 0.f;              // Model::normal fallback (0,0,1), model.cpp:404
    a1.x = nrm3 ? nrm3[3 * (size_t)v + 1] : 0.f;
    a1.y = nrm3 ? nrm3[3 * (size_t)v + 2] : 1.f;
    a1.z = uv2 ? uv2[2 * (size_t)v] : 0.f;
    a1.w = uv2 ? uv2[2 * (size_t)v + 1] : 0.f;
    float4* a = reinterpret_cast<float4*>(attr8 + (size_t)v * 8);
    a[0] = a0;
    a[1] = a1;
}

// ---------------------------------------------------------------------------------------------
// vertex stage
// ---------------------------------------------------------------------------------------------
// mark != nullptr: only the vertices with mark[v] != 0 (one rank's share of a mesh, k_mark_share_vertices)
__global__ void __launch_bounds__(TPB) k_vertex_mesh(FrameDev f, const float4* __restrict__ pos4, uint32_t nverts,
                                                     const double* __restrict__ mats, VRec* __restrict__ out,
                                                     const uint8_t* __restrict__ mark) {
    __shared__ double m[32];
    const int view = blockIdx.y;
    if (threadIdx.x < 32) m[threadIdx.x] = mats[view * 32 + threadIdx.x];
    __syncthreads();
    uint32_t v = blockIdx.x * TPB + threadIdx.x;
    if (v >= nverts) return;
    if (mark && !__ldg(mark + v)) return;
    float4 p = __ldg(pos4 + v);
    VRec r = vrec_from_position(m, m + 16, f.viewport, (double)p.x, (double)p.y, (double)p.z);
    store_vrec(out + (size_t)view * nverts + v, r);
}

__global__ void __launch_bounds__(TPB) k_vertex_clip(FrameDev f, const double* __restrict__ clip4, uint32_t nverts,
                                                     VRec* __restrict__ out) {
    uint32_t v = blockIdx.x * TPB + threadIdx.x;
    if (v >= nverts) return;
    const double2* q = reinterpret_cast<const double2*>(clip4) + (size_t)v * 2;
    double2 a = __ldg(q), b = __ldg(q + 1);
    store_vrec(out + v, vrec_from_clip(f.viewport, a.x, a.y, b.x, b.y));
}

// ---------------------------------------------------------------------------------------------
// setup + per-tile count
// ---------------------------------------------------------------------------------------------
struct GeomArgs {
    const uint32_t* idx;
    uint32_t first_tri, ntris, nverts, id_base;
    const VRec* vrec;          // [nviews][nverts]
    // Coherent processing order of a large indexed mesh (built once at upload, trb.cu mesh_order): slot j of the draw
    // holds mesh triangle perm[j], idx_perm[3j..3j+2] are its vertex indices.  The draw then runs over the nslots
    // slots of the whole mesh; slots whose triangle lies outside [first_tri, first_tri + ntris) are rejected.  Everything
    // per-draw (tribox, trirec, bins, direct_list) is indexed by SLOT; the id of slot j is id_base + (perm[j] - first_tri) + 1,
    // so depth ties still go to the triangle submitted first.  perm == nullptr: slot == triangle of the range.
    const uint32_t* perm;
    const uint32_t* idx_perm;
    uint32_t nslots;           // perm ? slots this draw visits (the whole mesh, or this rank's share of it) : ntris
    // trb_draw_shard: rank shard_r of shard_n draws the blocks b = shard_r (mod shard_n) of 2^shard_shift consecutive
    // positions of the processing order - a spatially coherent AND evenly spread share of the mesh (a contiguous run
    // of the order would be one patch of the surface, all of it back-facing for some ranks).  shard_n <= 1: no sharding
    uint32_t shard_n, shard_r, shard_shift;
    uint32_t nperm;            // entries of perm
};
// position in the processing order of the draw's slot `slot` (identity unless the draw is one rank's share)
__device__ __forceinline__ uint32_t slot_position(uint32_t slot, uint32_t shard_n, uint32_t shard_r, uint32_t shard_shift) {
    return shard_n > 1u ? ((((slot >> shard_shift) * shard_n + shard_r) << shard_shift) | (slot & ((1u << shard_shift) - 1u))) : slot;
}
struct SlotIds {               // slot -> triangle id for the raster kernels
    const uint32_t* perm;      // nullptr: id = id_off + slot + 1
    uint32_t id_off;           // id_base - first_tri (mod 2^32) with perm, else id_base
    uint32_t shard_n, shard_r, shard_shift;
};
__device__ __forceinline__ uint32_t slot_gid(const SlotIds& m, uint32_t slot) {
    return m.id_off + (m.perm ? __ldg(m.perm + slot_position(slot, m.shard_n, m.shard_r, m.shard_shift)) : slot) + 1u;
}

// the vertices one rank's share of an ordered indexed mesh refers to (trb_draw_shard): mark[v] = 1
__global__ void __launch_bounds__(TPB) k_mark_share_vertices(const uint32_t* __restrict__ idx_perm, uint32_t nperm, uint32_t nslots,
                                                             uint32_t shard_n, uint32_t shard_r, uint32_t shard_shift,
                                                             uint8_t* __restrict__ mark) {
    const uint32_t slot = blockIdx.x * TPB + threadIdx.x;
    if (slot >= nslots) return;
    const uint32_t j = slot_position(slot, shard_n, shard_r, shard_shift);
    if (j >= nperm) return;
    const uint32_t* q = idx_perm + (size_t)j * 3;
    mark[__ldg(q)] = 1; mark[__ldg(q + 1)] = 1; mark[__ldg(q + 2)] = 1;
}

constexpr uint32_t BOX_NONE = 0xFFFFFFFFu;

// Per-triangle raster record written once by k_setup_count and gathered by k_raster through the
// bins: everything eval_sample needs plus the clamped pixel bbox.  96 bytes = 3 DRAM sectors.
struct __align__(32) TriRec {
    double ax, ay, s00, s01, s10, s11, uz, z0, z1, z2, ruz;
    unsigned short x0, y0, x1, y1;
};
static_assert(sizeof(TriRec) == 96, "TriRec must be 96 bytes");

__device__ __forceinline__ void store_trirec(TriRec* p, const TriSetup& t) {
    double2* q = reinterpret_cast<double2*>(p);
    q[0] = make_double2(t.ax, t.ay);
    q[1] = make_double2(t.s00, t.s01);
    q[2] = make_double2(t.s10, t.s11);
    q[3] = make_double2(t.uz, t.z0);
    q[4] = make_double2(t.z1, t.z2);
    uint2 b;
    b.x = (uint32_t)t.x0 | ((uint32_t)t.y0 << 16);
    b.y = (uint32_t)t.x1 | ((uint32_t)t.y1 << 16);
    double2 last;
    last.x = t.ruz;
    last.y = __longlong_as_double((long long)(((unsigned long long)b.y << 32) | b.x));
    q[5] = last;
}
__device__ __forceinline__ void load_trirec(const TriRec* p, TriSetup& t) {
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3), e = __ldg(q + 4);
    double2 l = __ldg(q + 5);
    const unsigned long long bbw = (unsigned long long)__double_as_longlong(l.y);
    const uint32_t bx = (uint32_t)bbw, by = (uint32_t)(bbw >> 32);
    t.ax = a.x; t.ay = a.y; t.s00 = b.x; t.s01 = b.y; t.s10 = c.x; t.s11 = c.y;
    t.uz = d.x; t.z0 = d.y; t.z1 = e.x; t.z2 = e.y; t.ruz = l.x;
    t.x0 = bx & 0xffff; t.y0 = bx >> 16; t.x1 = by & 0xffff; t.y1 = by >> 16;
}

__device__ __forceinline__ int block_reduce_min(int v, int* sh) {
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < TPB / 32 ? sh[threadIdx.x] : INT_MAX;
        for (int o = 4; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    }
    __syncthreads();
    return v;  // valid in thread 0
}
__device__ __forceinline__ unsigned long long block_reduce_sum(unsigned long long v, unsigned long long* sh) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < TPB / 32 ? sh[threadIdx.x] : 0ull;
        for (int o = 4; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    __syncthreads();
    return v;  // valid in thread 0
}
__device__ __forceinline__ unsigned long long block_reduce_min64(unsigned long long v, unsigned long long* sh) {
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < TPB / 32 ? sh[threadIdx.x] : ~0ull;
        for (int o = 4; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    }
    __syncthreads();
    return v;
}

// Triangles whose WHOLE pixel bbox holds at most `direct_area` samples (the sub-pixel triangles of
// configs 4 and 5) skip the bins: their fragments go straight to the global depth-key plane with a
// 64-bit atomicMin, here in the setup kernel; a fragment that strictly lowers a pixel's depth
// invalidates its id, and k_direct_resolve then gives every pixel the lowest id among the fragments
// that reached its final depth - the same exact (depth, id) minimum as the tile path, no bins, no
// barriers.  The tile kernels of the same draw run afterwards and see these pixels as older state.
constexpr uint32_t BOX_DIRECT = 0xFFFFFFFEu;   // tribox.x of a direct triangle
#ifndef TRB_SETUP_MIN_BLOCKS
#define TRB_SETUP_MIN_BLOCKS 4
#endif
#ifndef TRB_SETUP_CHUNKS
#define TRB_SETUP_CHUNKS 4
#endif
constexpr int SETUP_CHUNKS = TRB_SETUP_CHUNKS;   // chunks of TPB triangles per CTA of k_setup_count
constexpr int DIRECT_AREA_DEFAULT = 16;

__global__ void __launch_bounds__(TPB, TRB_SETUP_MIN_BLOCKS) k_setup_count(FrameDev f, GeomArgs g, uint2* __restrict__ tribox,
                                                     TriRec* __restrict__ trirec, uint32_t* __restrict__ tile_count,
                                                     int direct_area, uint32_t* __restrict__ direct_list,
                                                     uint32_t* __restrict__ direct_n, int direct_by_pixel) {
    __shared__ int sh_i[4 * 8];
    __shared__ unsigned sh_w[5 * 8];
    static_assert(TPB / 32 == 8, "block totals are laid out for 8 warps");
    const int view = blockIdx.y;
    const VRec* vr = g.vrec + (size_t)view * g.nverts;
    const unsigned lane_id = threadIdx.x & 31;
    uint32_t* cnt = tile_count + (size_t)view * f.ntiles;
    // block totals, accumulated over the CTA's chunks and reduced once at the end
    int acc_x0 = INT_MAX, acc_y0 = INT_MAX, acc_x1 = INT_MIN, acc_y1 = INT_MIN;
    unsigned acc_nb = 0, acc_ne = 0;
    unsigned long long direct_cov = 0, direct_zmin = ~0ull;
    // A CTA takes SETUP_CHUNKS consecutive chunks of TPB triangles.  The vertex indices of the NEXT chunk are
    // requested before the current chunk's vertex records are gathered, so the dependent chain
    // index -> record -> setup only pays one memory latency per chunk instead of two.
    // t = the SLOT this thread handles (== the triangle of the range unless the mesh carries a processing order)
    const uint32_t t_first = blockIdx.x * (TPB * SETUP_CHUNKS) + threadIdx.x;
    const uint32_t nslots = g.nslots;
    uint32_t n0 = 0, n1 = 0, n2 = 0, nm = 0;
    auto fetch_slot = [&](uint32_t slot) {
        if (g.perm) {
            const uint32_t j = slot_position(slot, g.shard_n, g.shard_r, g.shard_shift);
            if (j >= g.nperm) { nm = 0xffffffffu; n0 = n1 = n2 = 0u; return; }   // past the end of the last block
            nm = __ldg(g.perm + j) - g.first_tri;                  // triangle inside the range, or >= ntris (wraps)
            if (g.idx_perm) {
                const uint32_t* q = g.idx_perm + (size_t)j * 3;
                n0 = __ldg(q); n1 = __ldg(q + 1); n2 = __ldg(q + 2);
            } else {                                               // ordered soup: the vertex arrays are in slot order
                n0 = 3u * j; n1 = 3u * j + 1u; n2 = 3u * j + 2u;
            }
        } else {
            nm = slot;
            n0 = vertex_index(g.idx, g.first_tri, slot, 0);
            n1 = vertex_index(g.idx, g.first_tri, slot, 1);
            n2 = vertex_index(g.idx, g.first_tri, slot, 2);
        }
    };
    if (t_first < nslots) fetch_slot(t_first);
    #pragma unroll 1
    for (int chunk = 0; chunk < SETUP_CHUNKS; ++chunk) {
        const uint32_t t = t_first + chunk * TPB;
        if (blockIdx.x * (TPB * SETUP_CHUNKS) + chunk * TPB >= nslots) break;   // uniform over the CTA
        const uint32_t i0 = n0, i1 = n1, i2 = n2;
        const bool mine = t < nslots && nm < g.ntris;
        if (chunk + 1 < SETUP_CHUNKS && t + TPB < nslots) fetch_slot(t + TPB);
        int res = SETUP_REJECT;
        TriSetup ts;
        if (mine) {
            VRec a = load_vrec(vr + i0);
            VRec b = load_vrec(vr + i1);
            VRec c = load_vrec(vr + i2);
            res = setup_triangle(a, b, c, f.W, f.H, ts);
        }
        // statistics bbox + "survived the rejects" count, our_gl.cpp:138-141
        if (res != SETUP_REJECT) {
            acc_x0 = min(acc_x0, ts.x0); acc_y0 = min(acc_y0, ts.y0);
            acc_x1 = max(acc_x1, ts.x1); acc_y1 = max(acc_y1, ts.y1);
            ++acc_nb;
        }
        uint2 box = make_uint2(BOX_NONE, 0u);
        uint32_t ntile = 0;
        int tx0 = 0, ty0 = 0, tx1 = -1, ty1 = -1;
        // Direct only when (nearly) the whole warp holds small triangles in DISTINCT tiles (a random
        // soup).  Measured on B200: with half of the lanes idle (back faces of a closed mesh) or lanes
        // hitting the same pixels the two direct passes cost more than the compacted bins of the tile path
        // (config 4: 1.0 ms binned vs 1.55 ms direct at level 9; config 5: 6.3 ms binned vs 2.7 ms direct).
        const bool small_tri = res == SETUP_DRAW && (ts.x1 - ts.x0 + 1) * (ts.y1 - ts.y0 + 1) <= direct_area;
        // "distinct" = tile, or - for a soup that is processed in a coherent order (mesh_order.cu: direct_by_pixel) - first
        // pixel of the bbox: its warps sit in a handful of tiles but still in different pixels, and its atomics hit lines
        // that are in L2.  Measured: ordered 100 M soup 18.0 ms with the tile vote (everything binned), 13.5 ms with the
        // pixel vote, 14.3 ms unordered; the pixel vote on indexed meshes costs config 4 6 % and config 3 1 % (not used).
        // (31-bit keys: two pixels of a frame beyond 2^31 pixels may alias, which only makes lanes look less lonely;
        // a 64-bit MATCH costs the set-up kernel 14 % on configs 3 and 4)
        const unsigned tkey = !small_tri ? (0x80000000u | lane_id)
                              : direct_by_pixel ? (((unsigned)ts.y0 * (unsigned)f.W + (unsigned)ts.x0) & 0x7fffffffu)
                                                : (unsigned)((ts.y0 >> TILE_SHIFT) * f.tw + (ts.x0 >> TILE_SHIFT));
        const unsigned same_tile = __match_any_sync(0xffffffffu, tkey);   // every lane takes part: no short-circuit
        const bool lonely = small_tri && __popc(same_tile) <= 2;
        const unsigned m_lonely = __ballot_sync(0xffffffffu, lonely);
        if (small_tri && __popc(m_lonely) >= 24) {
            unsigned long long* zk = f.zkey + (size_t)view * f.npix;
            uint32_t* vis = f.vis + (size_t)view * f.npix;
            uint32_t candidate = 0;
            for (int y = ts.y0; y <= ts.y1; ++y)
                for (int x = ts.x0; x <= ts.x1; ++x) {
                    double b[3], z;
                    if (!eval_sample(ts, x, y, b, z)) continue;
                    const unsigned long long key = fragment_key(z);
                    const size_t p = (size_t)y * f.W + x;
                    ++direct_cov;
                    direct_zmin = min(direct_zmin, key);
                    if (key <= zk[p]) {
                        const unsigned long long old = atomicMin(zk + p, key);
                        if (old > key) vis[p] = VIS_NONE;      // strictly nearer: the old winner is gone
                        if (old >= key) candidate = 1;
                    }
                }
            box = make_uint2(BOX_DIRECT, candidate);
            res = SETUP_NO_COVERAGE;                            // handled: keep it out of the bins
            // triangles that may own a pixel go on the view's list: k_direct_resolve then runs full warps
            // over ~the visible fraction instead of idling through every triangle of the draw
            const unsigned act = __activemask(), cm = __ballot_sync(act, candidate != 0);
            if (cm) {
                const int leader = __ffs(cm) - 1;
                uint32_t base = 0;
                if ((int)lane_id == leader) base = atomicAdd(direct_n + view, (uint32_t)__popc(cm));
                base = __shfl_sync(act, base, leader);
                if (candidate) direct_list[(size_t)view * nslots + base + __popc(cm & ((1u << lane_id) - 1u))] = t;
            }
        }
        if (res == SETUP_DRAW) {
            tx0 = ts.x0 >> TILE_SHIFT; tx1 = ts.x1 >> TILE_SHIFT;
            ty0 = ts.y0 >> TILE_SHIFT; ty1 = ts.y1 >> TILE_SHIFT;
            box = make_uint2((uint32_t)tx0 | ((uint32_t)ty0 << 16), (uint32_t)tx1 | ((uint32_t)ty1 << 16));
            ntile = (uint32_t)(tx1 - tx0 + 1) * (uint32_t)(ty1 - ty0 + 1);
            acc_ne += ntile;
            store_trirec(trirec + (size_t)view * nslots + t, ts);
        }
        if (t < nslots) tribox[(size_t)view * nslots + t] = box;
        // per-tile counts; single-tile triangles (the common case for small triangles) are aggregated
        // across the warp with match_any so that coherent meshes do not serialise on one counter
        unsigned key = 0x80000000u | lane_id;  // unique: no aggregation
        if (ntile == 1) key = (unsigned)(ty0 * f.tw + tx0);
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (ntile == 1) {
            if ((unsigned)(__ffs(peers) - 1) == lane_id) atomicAdd(cnt + key, (uint32_t)__popc(peers));
        } else if (ntile > 1) {
            for (int ty = ty0; ty <= ty1; ++ty)
                for (int tx = tx0; tx <= tx1; ++tx) atomicAdd(cnt + ty * f.tw + tx, 1u);
        }
    }
    {   // block totals: one REDUX per value and warp, one barrier, one more REDUX in warp 0
        const unsigned FULL = 0xffffffffu;
        const unsigned lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5;
        int bx0 = __reduce_min_sync(FULL, acc_x0), by0 = __reduce_min_sync(FULL, acc_y0);
        int bx1 = __reduce_max_sync(FULL, acc_x1), by1 = __reduce_max_sync(FULL, acc_y1);
        unsigned nb = __reduce_add_sync(FULL, acc_nb), ne = __reduce_add_sync(FULL, acc_ne);
        unsigned dcov = __reduce_add_sync(FULL, (unsigned)direct_cov);
        unsigned zhi = __reduce_min_sync(FULL, (unsigned)(direct_zmin >> 32));
        unsigned zlo = __reduce_min_sync(FULL, (unsigned)(direct_zmin >> 32) == zhi ? (unsigned)direct_zmin : 0xffffffffu);
        if (lane_ == 0) {
            sh_i[warp_] = bx0; sh_i[8 + warp_] = by0; sh_i[16 + warp_] = bx1; sh_i[24 + warp_] = by1;
            sh_w[warp_] = nb; sh_w[8 + warp_] = ne; sh_w[16 + warp_] = dcov; sh_w[24 + warp_] = zhi; sh_w[32 + warp_] = zlo;
        }
        __syncthreads();
        if (warp_ == 0) {
            const bool in = lane_ < TPB / 32;
            bx0 = __reduce_min_sync(FULL, in ? sh_i[lane_] : INT_MAX);
            by0 = __reduce_min_sync(FULL, in ? sh_i[8 + lane_] : INT_MAX);
            bx1 = __reduce_max_sync(FULL, in ? sh_i[16 + lane_] : INT_MIN);
            by1 = __reduce_max_sync(FULL, in ? sh_i[24 + lane_] : INT_MIN);
            nb = __reduce_add_sync(FULL, in ? sh_w[lane_] : 0u);
            ne = __reduce_add_sync(FULL, in ? sh_w[8 + lane_] : 0u);
            dcov = __reduce_add_sync(FULL, in ? sh_w[16 + lane_] : 0u);
            const unsigned whi = in ? sh_w[24 + lane_] : 0xffffffffu, wlo = in ? sh_w[32 + lane_] : 0xffffffffu;
            zhi = __reduce_min_sync(FULL, whi);
            zlo = __reduce_min_sync(FULL, whi == zhi ? wlo : 0xffffffffu);
            if (lane_ == 0 && nb) {
                DevStats* s = f.stats + view;
                atomicMin(&s->bx0, bx0); atomicMin(&s->by0, by0);
                atomicMax(&s->bx1, bx1); atomicMax(&s->by1, by1);
                atomicAdd(&s->tri_binned, (unsigned long long)nb);
                if (ne) atomicAdd(&s->tile_entries, (unsigned long long)ne);
                if (dcov) {
                    atomicAdd(&s->frag_covered, (unsigned long long)dcov);
                    atomicAdd(&s->touched, (unsigned long long)dcov);
                    atomicMin(&s->zmin_key, ((unsigned long long)zhi << 32) | zlo);
                }
            }
        }
    }
}

// second half of the direct path: ids of the fragments that sit at a pixel's final depth, over the
// compacted list of the triangles that were (for a moment at least) nearest somewhere
__device__ __forceinline__ void direct_resolve_entry(const FrameDev& f, const GeomArgs& g, int view, uint32_t t) {   // t: a slot
    const VRec* vr = g.vrec + (size_t)view * g.nverts;
    uint32_t v0, v1, v2;
    if (g.perm) {
        const uint32_t j = slot_position(t, g.shard_n, g.shard_r, g.shard_shift);
        if (g.idx_perm) {
            const uint32_t* q = g.idx_perm + (size_t)j * 3;
            v0 = __ldg(q); v1 = __ldg(q + 1); v2 = __ldg(q + 2);
        } else {
            v0 = 3u * j; v1 = 3u * j + 1u; v2 = 3u * j + 2u;      // ordered soup
        }
    } else {
        v0 = vertex_index(g.idx, g.first_tri, t, 0); v1 = vertex_index(g.idx, g.first_tri, t, 1); v2 = vertex_index(g.idx, g.first_tri, t, 2);
    }
    VRec a = load_vrec(vr + v0), b_ = load_vrec(vr + v1), c = load_vrec(vr + v2);
    TriSetup ts;
    setup_triangle(a, b_, c, f.W, f.H, ts);
    const unsigned long long* zk = f.zkey + (size_t)view * f.npix;
    uint32_t* vis = f.vis + (size_t)view * f.npix;
    const uint32_t gid = slot_gid(SlotIds{g.perm, g.id_base - (g.perm ? g.first_tri : 0u), g.shard_n, g.shard_r, g.shard_shift}, t);
    for (int y = ts.y0; y <= ts.y1; ++y)
        for (int x = ts.x0; x <= ts.x1; ++x) {
            double b[3], z;
            if (!eval_sample(ts, x, y, b, z)) continue;
            const size_t p = (size_t)y * f.W + x;
            if (fragment_key(z) == zk[p]) atomicMin(vis + p, gid);   // ties: lowest id = first submitted
        }
}
// One thread per list entry.  The list length is only known on the device.  A context whose draws have had direct
// candidates (host_seen: mapped memory, read by the host at the next draw) launches STRIDE = false on a grid sized for
// the whole draw; any other context launches STRIDE = true on a few CTAs, which is correct for any length and costs
// 3 us instead of 12 us per draw when the list is empty (config 3).  Two instantiations because the loop form costs the
// 100 M soup 16 % (2.73 vs 2.35 ms) even on the full grid, and a capped grid doubles it.
template <bool STRIDE>
__global__ void __launch_bounds__(TPB) k_direct_resolve(FrameDev f, GeomArgs g, const uint32_t* __restrict__ direct_list,
                                                        const uint32_t* __restrict__ direct_n, uint32_t* __restrict__ host_seen) {
    const int view = blockIdx.y;
    const uint32_t n = direct_n[view];
    if (!STRIDE) {
        const uint32_t i = blockIdx.x * TPB + threadIdx.x;
        if (i >= n) return;
        direct_resolve_entry(f, g, view, direct_list[(size_t)view * g.nslots + i]);
        return;
    }
    if (n && blockIdx.x == 0 && threadIdx.x == 0) *host_seen = 1u;
    for (unsigned long long i = (unsigned long long)blockIdx.x * TPB + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * TPB)
        direct_resolve_entry(f, g, view, direct_list[(size_t)view * g.nslots + i]);
}

// ---------------------------------------------------------------------------------------------
// exclusive scan of the tile counts (two small kernels; 2048 elements per block)
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_BLOCK = TPB * SCAN_ITEMS;

// Per-context control block of the draw in flight, zeroed before every draw.  It is what lets a
// draw run without a host round trip: the kernels that follow the scan read their work from it.
struct DrawCtl {
    uint32_t heavy_n;    // tiles whose bin is longer than warp_max: the work list of k_raster
    uint32_t longest;    // longest bin of the draw
    uint32_t total;      // R, the bin entries of the draw
    uint32_t overflow;   // R does not fit the bin buffer: the unbinned kernels take the draw instead
    uint32_t split_n;    // slices of long bins handed to the split flavour of k_raster_warp (may overshoot its capacity)
};
// Long bins are cut into slices of split_s triangles, one warp each (k_raster_warp<.., true>): a bin of 8000 sub-pixel
// triangles on the silhouette of the config-4 sphere would otherwise keep ONE warp busy for 250 batches while the rest
// of the GPU has long finished.  The slices of a tile are consecutive entries of the item list.
struct SplitArgs {
    uint2* items;                 // [cap] {tile slot, slice | slices << 16}; tile slot 0xffffffff: not used
    uint32_t* done;               // [tile slots] slices of the tile that have delivered (zeroed per draw)
    unsigned long long* keys;     // [cap][256] the slices' private tiles
    uint32_t* ids;                // [cap][256]
    uint32_t cap, split_s;
};

// ---------------------------------------------------------------------------------------------
// Tile-granular depth snapshot: `zbuffer_before_eyes = zbuffer` ... `zbuffer = zbuffer_before_eyes` (main.cpp:700, 730)
// without copying the plane.  trb_depth_snapshot only zeroes one byte per tile slot; every draw between the snapshot and
// the restore runs k_snap_save between its scan and its raster kernels, which copies the tiles the draw is about to
// change (per-tile bin count > 0; every tile when the draw overflowed into the unbinned kernels, or when the caller
// asks for `all`) into the snapshot plane ONCE; k_snap_restore copies the saved tiles back.  Invariant: a tile whose
// byte is 0 still holds the snapshot's keys.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void snap_copy_tile(const FrameDev& f, uint32_t slot, const unsigned long long* __restrict__ from,
                                               unsigned long long* __restrict__ to, unsigned lane) {
    const uint32_t view = slot / (uint32_t)f.ntiles, tile = slot - view * (uint32_t)f.ntiles;
    const int tx0 = (int)(tile % (uint32_t)f.tw) << TILE_SHIFT, ty0 = (int)(tile / (uint32_t)f.tw) << TILE_SHIFT;
    const unsigned long long base = (unsigned long long)view * f.npix;
    constexpr int N = TILE * TILE / 32;
    unsigned long long v[N];
    const unsigned long long p0 = base + (unsigned long long)(ty0 + (int)(lane >> TILE_SHIFT)) * f.W + tx0 + (int)(lane & (TILE - 1));
    const unsigned long long step = 2ull * f.W;       // a warp covers two rows of the tile per pass
    if (tx0 + TILE <= f.W && ty0 + TILE <= f.H) {     // whole tile inside the frame: all loads in flight, then the stores
#pragma unroll
        for (int it = 0; it < N; ++it) v[it] = __ldg(from + p0 + it * step);
#pragma unroll
        for (int it = 0; it < N; ++it) to[p0 + it * step] = v[it];
        return;
    }
    const bool in_x = tx0 + (int)(lane & (TILE - 1)) < f.W;
#pragma unroll
    for (int it = 0; it < N; ++it)
        if (in_x && ty0 + (int)(lane >> TILE_SHIFT) + 2 * it < f.H) to[p0 + it * step] = from[p0 + it * step];
}
// the flagged slots of a CTA's 256 are collected in shared memory and dealt out to its 8 warps: the tiles an object touches
// are runs of consecutive slots, and one warp copying a run tile after tile is a chain of memory latencies
__device__ __forceinline__ void snap_copy_flagged(const FrameDev& f, bool need, uint32_t slot, const unsigned long long* __restrict__ from,
                                                  unsigned long long* __restrict__ to) {
    __shared__ uint32_t list[TPB];
    __shared__ uint32_t n_sh;
    if (threadIdx.x == 0) n_sh = 0;
    __syncthreads();
    if (need) list[atomicAdd(&n_sh, 1u)] = slot;
    __syncthreads();
    const uint32_t n = n_sh;
    for (uint32_t i = threadIdx.x >> 5; i < n; i += TPB / 32) snap_copy_tile(f, list[i], from, to, threadIdx.x & 31);
}
__global__ void __launch_bounds__(TPB) k_snap_save(FrameDev f, uint32_t nslots, const uint32_t* __restrict__ counts,
                                                   const DrawCtl* __restrict__ ctl, int all, uint8_t* __restrict__ saved,
                                                   unsigned long long* __restrict__ snap) {
    const uint32_t slot = blockIdx.x * TPB + threadIdx.x;
    const bool every = all || (ctl && ctl->overflow);
    const bool need = slot < nslots && !saved[slot] && (every || counts[slot] != 0);
    if (need) saved[slot] = 1;            // one thread per slot: nobody else looks at this byte inside the launch
    snap_copy_flagged(f, need, slot, f.zkey, snap);
}
__global__ void __launch_bounds__(TPB) k_snap_restore(FrameDev f, uint32_t nslots, const uint8_t* __restrict__ saved,
                                                      const unsigned long long* __restrict__ snap) {
    const uint32_t slot = blockIdx.x * TPB + threadIdx.x;
    snap_copy_flagged(f, slot < nslots && saved[slot], slot, snap, f.zkey);
}

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* sh, uint32_t& total) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (unsigned)o) x += y;
    }
    if (lane == 31) sh[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < TPB / 32 ? sh[lane] : 0u;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= (unsigned)o) w += y;
        }
        if (lane < TPB / 32) sh[lane] = w;
    }
    __syncthreads();
    uint32_t warp_off = warp ? sh[warp - 1] : 0u;
    total = sh[TPB / 32 - 1];
    __syncthreads();
    return warp_off + x - v;
}

__global__ void __launch_bounds__(TPB) k_scan_partial(const uint32_t* __restrict__ in, uint32_t n,
                                                      uint32_t* __restrict__ block_sum, DrawCtl* __restrict__ ctl,
                                                      uint32_t warp_max, uint32_t* __restrict__ heavy_list, SplitArgs sp) {
    __shared__ unsigned long long sh[TPB / 32];
    size_t base = (size_t)blockIdx.x * SCAN_BLOCK;
    unsigned long long s = 0;
    uint32_t m = 0;
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        size_t e = base + (size_t)i * TPB + threadIdx.x;
        if (e < n) {
            const uint32_t v = in[e];
            s += v;
            m = max(m, v);
            if (v > warp_max) {
                // long bin: slices for the split warp kernel while its item list has room, else the CTA-per-tile kernel
                // (split_n counts what the draw asks for even while the list has no room: the host sizes the next one from it)
                const uint32_t K = (v + sp.split_s - 1) / sp.split_s;
                bool split = K <= 0xffffu;
                if (split) {
                    const uint32_t b = atomicAdd(&ctl->split_n, K);
                    split = b + K <= sp.cap && b + K >= b;
                    for (uint32_t j = 0; j < K && b + j < sp.cap && b + j >= b; ++j)
                        sp.items[b + j] = split ? make_uint2((uint32_t)e, j | (K << 16)) : make_uint2(0xffffffffu, 0u);
                }
                if (!split) heavy_list[atomicAdd(&ctl->heavy_n, 1u)] = (uint32_t)e;
            }
        }
    }
    s = block_reduce_sum(s, sh);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&ctl->longest, m);
    if (threadIdx.x == 0) block_sum[blockIdx.x] = (uint32_t)s;
}
// second (and last) pass: every block adds up the partial sums of the blocks before it itself (a few hundred 4-byte
// reads from L2 - cheaper than a single-block kernel in between, whose launch and dependent scan cost ~12 us per draw),
// then scans its own elements.  The last block also knows R: the total, the longest bin and the overflow verdict go to
// ctl and to host_out[0..2] (mapped host memory: the host reads them without a copy).
__global__ void __launch_bounds__(TPB) k_scan_final(const uint32_t* __restrict__ in, uint32_t n,
                                                    const uint32_t* __restrict__ block_sum,
                                                    uint32_t* __restrict__ out, uint32_t* __restrict__ host_out,
                                                    DrawCtl* __restrict__ ctl, uint32_t bin_capacity) {
    __shared__ uint32_t sh[TPB / 32];
    __shared__ unsigned long long shw[TPB / 32];
    __shared__ unsigned long long before_sh;   // R in 64 bits: offsets are 32 bit, a draw beyond that must not wrap silently
    unsigned long long before = 0;
    for (uint32_t e = threadIdx.x; e < blockIdx.x; e += TPB) before += block_sum[e];
    before = block_reduce_sum(before, shw);
    if (threadIdx.x == 0) before_sh = before;
    __syncthreads();
    before = before_sh;
    size_t base = (size_t)blockIdx.x * SCAN_BLOCK + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = base + i < n ? in[base + i] : 0u;
        s += v[i];
    }
    uint32_t tot;
    uint32_t ex = block_exclusive_scan(s, sh, tot) + (uint32_t)before;
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        const unsigned long long wide = before + block_sum[blockIdx.x];
        ctl->total = (uint32_t)wide;
        ctl->overflow = wide > (unsigned long long)bin_capacity ? 1u : 0u;   // also true when R does not fit 32 bits
        host_out[0] = wide > 0xffffffffull ? 0xffffffffu : (uint32_t)wide;
        host_out[1] = ctl->longest;
        host_out[2] = ctl->overflow;
        host_out[3] = ctl->split_n;      // slices the draw asked for: the host sizes the next draw's item list from it
    }
}

// ---------------------------------------------------------------------------------------------
// bin fill: order inside a bin is irrelevant, the resolve is order independent
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_fill(FrameDev f, uint32_t ntris, const uint2* __restrict__ tribox,
                                              const uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursor,
                                              uint32_t* __restrict__ bins, const DrawCtl* __restrict__ ctl) {
    if (ctl->overflow) return;
    const int view = blockIdx.y;
    const uint32_t t = blockIdx.x * TPB + threadIdx.x;
    uint2 box = make_uint2(BOX_NONE, 0u);
    if (t < ntris) box = tribox[(size_t)view * ntris + t];
    const bool has = box.x < BOX_DIRECT;   // neither rejected (BOX_NONE) nor handled by the direct path
    const int tx0 = box.x & 0xffff, ty0 = box.x >> 16, tx1 = box.y & 0xffff, ty1 = box.y >> 16;
    const bool single = has && tx0 == tx1 && ty0 == ty1;
    const size_t vbase = (size_t)view * f.ntiles;
    const unsigned lane = threadIdx.x & 31;
    unsigned key = 0x80000000u | lane;
    if (single) key = (unsigned)(ty0 * f.tw + tx0);
    unsigned peers = __match_any_sync(0xffffffffu, key);
    if (single) {
        int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if ((unsigned)leader == lane) base = atomicAdd(cursor + vbase + key, (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        TRB_CHECK(offsets[vbase + key] + base + rank < ctl->total);
        bins[offsets[vbase + key] + base + rank] = t;
    } else if (has) {
        for (int ty = ty0; ty <= ty1; ++ty)
            for (int tx = tx0; tx <= tx1; ++tx) {
                size_t tile = vbase + (size_t)ty * f.tw + tx;
                uint32_t slot = atomicAdd(cursor + tile, 1u);
                TRB_CHECK(offsets[tile] + slot < ctl->total);
                bins[offsets[tile] + slot] = t;
            }
    }
}

// ---------------------------------------------------------------------------------------------
// fine raster: one CTA owns one 16x16 tile
//
// The tile's depth keys and winner ids live in shared memory / registers for the whole bin.
// Triangles of the bin are taken in chunks of CHUNK records; per chunk every thread gathers one
// triangle record and the chunk's triangles are split by how many samples of the tile their
// clipped bbox holds (ns):
//   tiny  (ns < big_ns, and the chunk has enough of them)  one THREAD per triangle walks its samples;
//   mid   (everything else below large_ns)                 SAMPLE-parallel: the samples of all mid
//         triangles are laid end to end (prefix sum) and dealt out to the 256 threads, so every
//         lane evaluates an in-bbox sample whatever the triangle sizes are;
//   large (ns >= large_ns, about half a tile)              PIXEL-owner: every thread owns one pixel
//         and walks the large triangles from shared memory - no atomics, no barriers.
// Tiny and mid fragments resolve through shared-memory atomicMin on the depth key plus a candidate
// queue and an id atomicMin, which makes the (depth, id) minimum exact whatever the order.
// ---------------------------------------------------------------------------------------------
constexpr int CHUNK = TPB;
constexpr int SMALL_MIN_DEFAULT = 24;   // fewer tiny triangles than this in a chunk: treat them as mid
constexpr int LARGE_NS_DEFAULT = 128;   // clipped-bbox samples from which a triangle is pixel-owner
constexpr int SP_GROUP = 3;             // sample-parallel rounds (of TPB samples) between two resolves

struct RasterArgs {
    uint32_t ntris;           // slots of the draw
    SlotIds ids;              // slot -> triangle id
    const TriRec* trirec;     // [nviews][ntris]
    const uint2* tribox;      // [nviews][ntris] tile ranges (k_setup_count); BOX_NONE / BOX_DIRECT markers
    const uint32_t* counts;   // [nviews][ntiles]
    const uint32_t* offsets;  // [nviews][ntiles]
    const uint32_t* bins;
    int big_ns, small_min, large_ns;
    uint32_t warp_max;        // bins of 1..warp_max triangles go to k_raster_warp, longer ones to k_raster
    const DrawCtl* ctl;
    const uint32_t* heavy_list;   // [ctl->heavy_n] tile slots (view * ntiles + tile) of the long bins
};

struct SpEntry {              // one mid triangle of the chunk in the sample-parallel list
    uint32_t first;           // index of its first sample in the chunk's sample sequence
    uint16_t inv;             // ceil(32768 / bw): l / bw == (l * inv) >> 15 for l < 256, bw <= 16
    uint8_t rec, bw;          // slot in recs[], clipped bbox width
};

#ifndef TRB_RASTER_MIN_BLOCKS
#define TRB_RASTER_MIN_BLOCKS 4
#endif
__device__ __forceinline__ void raster_tile_cta(const FrameDev& f, const RasterArgs& a, uint32_t tslot_) {
    const int tile = (int)(tslot_ % (uint32_t)f.ntiles), view = (int)(tslot_ / (uint32_t)f.ntiles);
    const size_t tslot = tslot_;
    const uint32_t n = a.counts[tslot];
    const uint32_t off = a.offsets[tslot];

    __shared__ unsigned long long zk[TPB];
    __shared__ uint32_t vid[TPB];
    __shared__ unsigned long long qk[QCAP];
    __shared__ uint32_t qid[QCAP];
    __shared__ uint8_t qp[QCAP];
    __shared__ TriRec recs[CHUNK];       // mid + large triangles of the current chunk, bbox clipped to the tile
    __shared__ uint32_t rec_id[CHUNK];
    __shared__ SpEntry sp[CHUNK];
    __shared__ uint8_t big_slot[CHUNK];
    __shared__ uint32_t scan_sh[TPB / 32];
    __shared__ unsigned int qn, nrec[2], nmid[2], nbig[2];
    __shared__ unsigned long long red[TPB / 32];

    const int tid = threadIdx.x;
    const int tx0 = (tile % f.tw) << TILE_SHIFT, ty0 = (tile / f.tw) << TILE_SHIFT;
    const int px = tx0 + (tid & 15), py = ty0 + (tid >> 4);
    const bool pvalid = px < f.W && py < f.H;
    const size_t gp = (size_t)view * f.npix + (size_t)py * f.W + px;
    unsigned long long myk = pvalid ? f.zkey[gp] : 0ull;   // this thread's pixel
    uint32_t myid = pvalid ? f.vis[gp] : VIS_NONE;
    const unsigned long long k_in = myk;
    const uint32_t id_in = myid;
    if (tid == 0) { qn = 0; nrec[0] = nrec[1] = 0; nmid[0] = nmid[1] = 0; nbig[0] = nbig[1] = 0; }
    __syncthreads();

    const TriRec* tr = a.trirec + (size_t)view * a.ntris;
    uint32_t covered = 0;
    int parity = 0;

    // one fragment of the atomic paths: key min, and a queue entry when it is (for now) the winner
    auto fragment = [&](unsigned long long k, int p, uint32_t gid) -> bool {
        if (k <= *(volatile unsigned long long*)&zk[p]) {
            unsigned long long old = atomicMin(&zk[p], k);
            if (old >= k) {  // current minimum or a tie: remember who asked
                unsigned s = atomicAdd(&qn, 1u);
                if (s >= (unsigned)QCAP) return false;
                qk[s] = k; qid[s] = gid; qp[s] = (uint8_t)p;
            }
        }
        return true;
    };
    // barrier-separated resolve of everything queued so far; returns the block-wide OR of `again`
    auto resolve = [&](bool again) -> int {
        __syncthreads();
        if (zk[tid] != myk) { vid[tid] = VIS_NONE; myk = zk[tid]; }  // strictly nearer: forget the old winner
        __syncthreads();
        const unsigned qc = min(qn, (unsigned)QCAP);
        for (unsigned e = tid; e < qc; e += TPB)
            if (zk[qp[e]] == qk[e]) atomicMin(&vid[qp[e]], qid[e]);   // ties: lowest id = first submitted
        const int more = __syncthreads_or(again);
        if (tid == 0) qn = 0;
        __syncthreads();
        return more;
    };

    for (uint32_t base = 0; base < n; base += CHUNK, parity ^= 1) {
        TriSetup ts;
        uint32_t gid = 0;
        int cx0 = 0, cy0 = 0, cx1 = -1, cy1 = -1, ns = 0;
        const bool has = base + tid < n;
        if (has) {
            const uint32_t t = __ldg(a.bins + off + base + tid);
            load_trirec(tr + t, ts);
            gid = slot_gid(a.ids, t);
            cx0 = max(ts.x0, tx0); cx1 = min(ts.x1, tx0 + TILE - 1);
            cy0 = max(ts.y0, ty0); cy1 = min(ts.y1, ty0 + TILE - 1);
            ns = (cx1 - cx0 + 1) * (cy1 - cy0 + 1);
        }
        bool tiny = has && ns < a.big_ns;
        const int ntiny = __syncthreads_count(tiny);        // also: the previous chunk is done with the lists
        if (ntiny < a.small_min) tiny = false;
        const bool large = has && !tiny && ns >= a.large_ns;
        const bool mid = has && !tiny && !large;
        if (mid || large) {
            const unsigned s = atomicAdd(&nrec[parity], 1u);
            TriRec& B = recs[s];
            B.ax = ts.ax; B.ay = ts.ay; B.s00 = ts.s00; B.s01 = ts.s01; B.s10 = ts.s10; B.s11 = ts.s11;
            B.uz = ts.uz; B.z0 = ts.z0; B.z1 = ts.z1; B.z2 = ts.z2; B.ruz = ts.ruz;
            B.x0 = (unsigned short)cx0; B.y0 = (unsigned short)cy0; B.x1 = (unsigned short)cx1; B.y1 = (unsigned short)cy1;
            rec_id[s] = gid;
            if (large) {
                big_slot[atomicAdd(&nbig[parity], 1u)] = (uint8_t)s;
            } else {
                const unsigned m = atomicAdd(&nmid[parity], 1u);
                const int bw = cx1 - cx0 + 1;
                SpEntry e;
                e.first = (uint32_t)ns;                     // sample count for now; prefix-summed below
                e.inv = (uint16_t)((32768 + bw - 1) / bw);
                e.rec = (uint8_t)s;
                e.bw = (uint8_t)bw;
                sp[m] = e;
            }
        }
        __syncthreads();
        const unsigned nm = nmid[parity], nb = nbig[parity];
        if (tid == 0) { nrec[parity ^ 1] = 0; nmid[parity ^ 1] = 0; nbig[parity ^ 1] = 0; }
        const bool atomics = ntiny >= a.small_min || nm > 0;
        uint32_t S = 0;
        if (nm > 0) {   // lay the mid triangles' samples end to end
            const uint32_t cnt = (unsigned)tid < nm ? sp[tid].first : 0u;
            const uint32_t ex = block_exclusive_scan(cnt, scan_sh, S);
            if ((unsigned)tid < nm) sp[tid].first = ex;
        }
        if (atomics) {
            zk[tid] = myk; vid[tid] = myid;
            __syncthreads();
            // ---- tiny triangles: one thread per triangle ------------------------------------------------
            if (ntiny >= a.small_min) {
                bool pending = tiny;
                int sx = cx0, sy = cy0;  // resume position when the candidate queue fills up
                for (;;) {
                    bool full = false;
                    if (pending) {
                        while (sy <= cy1) {
                            double b[3], z;
                            if (eval_sample(ts, sx, sy, b, z)) {
                                if (!fragment(fragment_key(z), ((sy - ty0) << TILE_SHIFT) | (sx - tx0), gid)) { full = true; break; }
                                ++covered;
                            }
                            if (++sx > cx1) { sx = cx0; ++sy; }
                        }
                        if (!full) pending = false;
                    }
                    if (!resolve(full)) break;
                }
            }
            // ---- mid triangles: sample parallel ---------------------------------------------------------
            for (uint32_t s0 = 0; s0 < S; s0 += SP_GROUP * TPB) {
                #pragma unroll 1
                for (int g = 0; g < SP_GROUP; ++g) {
                    const uint32_t s = s0 + g * TPB + tid;
                    if (s >= S) break;
                    unsigned lo = 0, hi = nm;               // last entry with first <= s
                    while (hi - lo > 1) {
                        const unsigned m = (lo + hi) >> 1;
                        if (sp[m].first <= s) lo = m; else hi = m;
                    }
                    const SpEntry e = sp[lo];
                    const TriRec& B = recs[e.rec];
                    const uint32_t l = s - e.first;
                    const uint32_t row = (l * e.inv) >> 15;
                    const int x = (int)B.x0 + (int)(l - row * e.bw), y = (int)B.y0 + (int)row;
                    TriSetup t2;
                    t2.ax = B.ax; t2.ay = B.ay; t2.s00 = B.s00; t2.s01 = B.s01; t2.s10 = B.s10; t2.s11 = B.s11;
                    t2.uz = B.uz; t2.z0 = B.z0; t2.z1 = B.z1; t2.z2 = B.z2; t2.ruz = B.ruz;
                    double b[3], z;
                    if (eval_sample(t2, x, y, b, z)) {
                        fragment(fragment_key(z), ((y - ty0) << TILE_SHIFT) | (x - tx0), rec_id[e.rec]);  // <= 768 per group: fits
                        ++covered;
                    }
                }
                resolve(false);
            }
            myid = vid[tid];
        }
        // ---- large triangles: every thread owns its pixel, no atomics -----------------------------------
        for (unsigned j = 0; j < nb; ++j) {
            const unsigned slot = big_slot[j];
            const TriRec& B = recs[slot];
            if (px < (int)B.x0 || px > (int)B.x1 || py < (int)B.y0 || py > (int)B.y1) continue;
            TriSetup t2;
            t2.ax = B.ax; t2.ay = B.ay; t2.s00 = B.s00; t2.s01 = B.s01; t2.s10 = B.s10; t2.s11 = B.s11;
            t2.uz = B.uz; t2.z0 = B.z0; t2.z1 = B.z1; t2.z2 = B.z2; t2.ruz = B.ruz;
            double b[3], z;
            if (!eval_sample(t2, px, py, b, z)) continue;
            const unsigned long long k = fragment_key(z);
            const uint32_t id = rec_id[slot];
            ++covered;
            if (k < myk) { myk = k; myid = id; }
            else if (k == myk && id < myid) myid = id;
        }
    }
    if (pvalid) { f.zkey[gp] = myk; f.vis[gp] = myid; }
    const unsigned long long total = block_reduce_sum((unsigned long long)covered, red);
    // min_z of our_gl.cpp:197: the smallest drawn depth.  A pixel's new key is the minimum of this
    // draw's fragments there, and fragments that lost to an older, smaller depth cannot be the minimum.
    const unsigned long long zmin = block_reduce_min64(myk != k_in ? myk : ~0ull, red);
    const unsigned long long touched = block_reduce_sum((myid != id_in && pvalid) ? 1ull : 0ull, red);
    if (tid == 0 && total) {
        atomicAdd(&f.stats[view].frag_covered, total);
        atomicMin(&f.stats[view].zmin_key, zmin);
        if (touched) atomicAdd(&f.stats[view].touched, touched);
    }
}
// ---------------------------------------------------------------------------------------------
// unbinned fallback.  A draw is enqueued without a host round trip, so its bin buffer is sized
// from an estimate; when the scan finds that R does not fit (ctl->overflow) the fill and tile
// kernels stand down and the draw's binned triangles are taken straight from their TriRecs:
// one warp per triangle, lanes over the samples of its clamped bbox, 64-bit atomicMin on the global
// key plane (the depth pass: it rides in k_raster's persistent grid, whose tile work is off in that case),
// then the lowest id among the fragments that sit at a pixel's final depth (k_unbinned<ids>) - the same
// exact (depth, id) minimum as the other paths (it is the direct path of k_setup_count with a warp
// per triangle).  Slower than the tile kernels, but only ever a performance cliff, never an error.
// ---------------------------------------------------------------------------------------------
template <bool IDS>
__device__ __noinline__ void unbinned_pass(unsigned long long* zkey, uint32_t* visp, DevStats* stats, size_t npix, int W, int H,
                                           int view0, int view1, uint32_t ntris, const SlotIds& ids,
                                           const uint2* __restrict__ tribox, const TriRec* __restrict__ trirec,
                                           uint32_t w0, uint32_t nwarps) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    for (int view = view0; view < view1; ++view) {
        unsigned long long* zk = zkey + (size_t)view * npix;
        uint32_t* vis = visp + (size_t)view * npix;
        unsigned long long covered = 0, zmin = ~0ull;
        for (uint32_t t = w0; t < ntris; t += nwarps) {
            if (tribox[(size_t)view * ntris + t].x >= BOX_DIRECT) continue;   // rejected, or already drawn by the direct path
            TriSetup ts;
            load_trirec(trirec + (size_t)view * ntris + t, ts);
            const uint32_t bw = (uint32_t)(ts.x1 - ts.x0 + 1), ns = bw * (uint32_t)(ts.y1 - ts.y0 + 1);
            const uint32_t gid = slot_gid(ids, t);
            for (uint32_t s = lane; s < ns; s += 32) {
                const uint32_t row = s / bw;
                const int x = ts.x0 + (int)(s - row * bw), y = ts.y0 + (int)row;
                TRB_CHECK(x >= 0 && x < W && y >= 0 && y < H);
                double b[3], z;
                if (!eval_sample(ts, x, y, b, z)) continue;
                const unsigned long long key = fragment_key(z);
                const size_t p = (size_t)y * W + x;
                if (!IDS) {
                    ++covered;
                    zmin = min(zmin, key);
                    if (key <= zk[p]) {
                        const unsigned long long old = atomicMin(zk + p, key);
                        if (old > key) vis[p] = VIS_NONE;          // strictly nearer: the old winner is gone
                    }
                } else if (key == zk[p]) {
                    atomicMin(vis + p, gid);                       // ties: lowest id = first submitted
                }
            }
        }
        if constexpr (!IDS) {
            for (int o = 16; o; o >>= 1) {
                covered += __shfl_xor_sync(FULL, covered, o);
                zmin = min(zmin, __shfl_xor_sync(FULL, zmin, o));
            }
            if (lane == 0 && covered) {
                atomicAdd(&stats[view].frag_covered, covered);
                atomicAdd(&stats[view].touched, covered);
                atomicMin(&stats[view].zmin_key, zmin);
            }
        }
    }
}

// persistent grid over the (usually short or empty) list of long bins
__global__ void __launch_bounds__(TPB, TRB_RASTER_MIN_BLOCKS) k_raster(FrameDev f, RasterArgs a) {
    if (a.ctl->overflow) {   // the bins were not filled: depth pass of the unbinned fallback, warps over all triangles of all views
        unbinned_pass<false>(f.zkey, f.vis, f.stats, f.npix, f.W, f.H, 0, f.nviews, a.ntris, a.ids, a.tribox, a.trirec,
                             blockIdx.x * (TPB / 32) + (threadIdx.x >> 5), gridDim.x * (TPB / 32));
        return;
    }
    const uint32_t nh = a.ctl->heavy_n;
    for (uint32_t i = blockIdx.x; i < nh; i += gridDim.x) {
        raster_tile_cta(f, a, a.heavy_list[i]);
        __syncthreads();                     // the tile's shared state is reused by the next one
    }
}

// ---------------------------------------------------------------------------------------------
// fine raster, warp flavour: one WARP owns one 16x16 tile (bins of at most warp_max triangles)
//
// No block barrier, no atomics, no candidate queue.  The warp keeps the tile's depth keys and ids
// in its private slice of shared memory and takes the bin in batches of 32 triangles: every lane
// gathers one TriRec into shared memory, the clipped-bbox samples of the batch are laid end to end
// (warp prefix sum) and dealt out 32 at a time, so every lane evaluates an in-bbox sample whatever
// the triangle sizes are.  A lane finds the triangle of its sample with one reduce_or + popc (the
// triangles' start positions inside the 32-sample window are a bit mask).  Fragments are applied
// with a plain read-compare-write of (key, id); lanes of a round that hit the same pixel (only
// possible when the window spans several triangles) take turns (match_any), so the result is the
// lexicographic (depth, id) minimum whatever the order - the reference's first-wins on ties.
// ---------------------------------------------------------------------------------------------
#ifndef TRB_RW_WARPS
#define TRB_RW_WARPS 4
#endif
constexpr int RW_WARPS = TRB_RW_WARPS;      // tiles per CTA.  One: the warp's shared-memory block sits at a constant address
constexpr uint32_t WARP_MAX_DEFAULT = 1024; // longest bin a single warp takes; TRB_WARP_MAX overrides (0: k_raster only)
// A triangle of the current batch as the warp keeps it in shared memory: the eleven doubles eval_sample reads
// plus its id.  96 bytes of payload on a 112-byte pitch: consecutive records start 28 banks apart, so the eight
// lanes of a quarter warp that read different records with one LDS.128 do not collide.
struct __align__(16) SmTri {
    double ax, ay, s00, s01, s10, s11, uz, ruz, z0, z1, z2;
    uint32_t gid;
    uint32_t flags;   // bit 0: u.z inside div_rn's exponent window
    double pad_;
};
static_assert(sizeof(SmTri) == 112, "SmTri pitch");
constexpr int RW_SPAN_CAP = 32 * TILE;      // every row of every triangle of a batch
struct __align__(16) WarpTile {
    unsigned long long zk[TILE * TILE];
    uint32_t vid[TILE * TILE];
    SmTri recs[32];
    uint32_t spans[RW_SPAN_CAP];            // triangle (5 bits) | first column (4) | row (4) | length - 1 (4)
    uint8_t claim[TILE * TILE];             // which lane last announced a fragment for the pixel (collision test)
};
static_assert(sizeof(WarpTile) == 8960 && sizeof(WarpTile) % 128 == 0, "WarpTile size (TMA destinations need 128-byte alignment)");

// Conservative row spans.  A sample of row y can only be covered when the three edge values the reference computes
// (u.x <= 0, u.y <= 0, u.x + u.y >= u.z; our_gl.cpp:77-86, 152) allow it.  In real arithmetic each is linear in the
// column, so each edge bounds the columns of the row from one side: column bound_k(y) = alpha_k + beta_k * s12 with
// s12 = A.y - (y + 0.5).  The reference evaluates the edge values in floating point; alpha_k carries a margin that is
// orders of magnitude larger than those rounding errors (and than the errors of this evaluation itself), always in the
// direction that keeps MORE columns.  Columns outside the span are therefore certain to fail the reference's test and
// are never enumerated; every column inside still goes through the exact evaluation, which alone decides coverage.
// (-DTRB_DEBUG_CHECKS evaluates the skipped columns too and asserts that none of them is covered.)
struct SpanEdges {
    double ay;                       // A.y (0 when the triangle's coordinates are not finite: every edge is then disabled)
    double a0, b0, a1, b1, a2, b2;   // bound_k = a_k + b_k * s12, stored so that floor() applies to all three
    int m0, m1, m2;                  // INT_MIN: edge k bounds the columns from above (x <= floor(bound));
                                     // INT_MAX: from below (x >= -floor(bound), the bound is stored negated)
};
__device__ __forceinline__ void span_edge(double num_const, double num_s12, double B, double centre, double s12max,
                                          bool ge, double& a, double& b, int& m) {
    // constraint  t * B >= num  (ge)  or  t * B <= num  (!ge)  with  t = A.x - px,  num = num_const + num_s12 * s12
    //   ->  px <= A.x - num / B  when the inequality bounds t from below (ge == (B > 0)), else px >= A.x - num / B
    double r;                                                   // 1 / B to ~2^-40: MUFU.RCP64H + one Newton step
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(B));
    r = __fma_rn(__fma_rn(-B, r, 1.0), r, r);
    const bool upper = ge == (B > 0.0);                         // columns bounded from above
    double alpha = centre - num_const * r, beta = -num_s12 * r; // column + 0.5 = A.x - num / B
    const double w = 2.9802322387695312e-08 * (fabs(alpha) + fabs(beta) * s12max + fabs(centre)) + 9.5367431640625e-07;  // 2^-25 rel + 2^-20
    alpha = upper ? alpha + w : alpha - w;
    const bool ok = fabs(r) <= 1.0e300 && fabs(alpha) <= 1.0e300 && fabs(beta) <= 1.0e300 && B != 0.0;   // false for NaN too
    if (!ok) { a = 1.0e300; b = 0.0; m = INT_MIN; return; }      // no bound from this edge (floor saturates to INT_MAX)
    if (upper) { a = alpha; b = beta; m = INT_MIN; }
    else { a = -alpha; b = -beta; m = INT_MAX; }                // x >= ceil(v)  <=>  -x <= floor(-v)
}
__device__ __forceinline__ void span_setup(const double2& r0, const double2& r1, const double2& r2, double uz, int X0, int X1,
                                           int Y0, int Y1, SpanEdges& E) {
    const double ax = r0.x, ay = r0.y, s00 = r1.x, s01 = r1.y, s10 = r2.x, s11 = r2.y;
    // magnitudes over the clipped bbox: |t| = |A.x - px| and |s12| = |A.y - py|
    const double tmax = fmax(fabs(ax - pixel_centre(X0)), fabs(ax - pixel_centre(X1)));
    const double s12max = fmax(fabs(ay - pixel_centre(Y0)), fabs(ay - pixel_centre(Y1)));
    const double K = 1.4210854715202004e-14;                    // 2^-46: 64 x the rounding of a product / sum
    const double e1 = K * (fabs(s01) * s12max + fabs(s11) * tmax), e2 = K * (fabs(s00) * s12max + fabs(s10) * tmax);
    const double thr = fabs(uz) * 1e-290;
    const double e3 = 4.0 * (e1 + e2) + K * fabs(uz);
    const double centre = ax - 0.5;
    // infinite / NaN coordinates make s12max (and with it every margin) non-finite, which disables all three edges;
    // the row evaluation must then not multiply an infinite s12 by a zero slope
    E.ay = fabs(ay) <= 1.0e300 ? ay : 0.0;
    // u.x = s01*s12 - t*s11 <= thr            ->  t * s11 >= s01*s12 - (thr + e1)
    span_edge(-(thr + e1), s01, s11, centre, s12max, true, E.a0, E.b0, E.m0);
    // u.y = t*s10 - s00*s12 <= thr            ->  t * s10 <= s00*s12 + (thr + e2)
    span_edge(thr + e2, s00, s10, centre, s12max, false, E.a1, E.b1, E.m1);
    // u.x + u.y = (s01-s00)*s12 + t*(s10-s11) >= u.z*1.000001 - e3   ->  t * (s10-s11) >= (u.z*1.000001 - e3) - (s01-s00)*s12
    span_edge(uz * 1.000001 - e3, -(s01 - s00), s10 - s11, centre, s12max, true, E.a2, E.b2, E.m2);
}
// columns [xa, xb] (absolute, clamped to [X0, X1]) of row y that can hold a covered sample; empty when xa > xb
__device__ __forceinline__ void span_of_row(const SpanEdges& E, double s12, int X0, int X1, int& xa, int& xb) {
    const int f0 = __double2int_rd(__fma_rn(E.b0, s12, E.a0)), f1 = __double2int_rd(__fma_rn(E.b1, s12, E.a1)),
              f2 = __double2int_rd(__fma_rn(E.b2, s12, E.a2));
    // m_k == INT_MIN: max(f_k, m_k) = f_k takes part in the upper bound, max(f_k, ~m_k) = INT_MAX drops out of the lower
    xb = min(X1, min(min(max(f0, E.m0), max(f1, E.m1)), max(f2, E.m2)));
    const int lo = min(min(max(f0, ~E.m0), max(f1, ~E.m1)), max(f2, ~E.m2));   // min over the lower edges of floor(-v)
    xa = max(X0, lo == INT_MAX ? INT_MIN : -lo);                                // -INT_MIN wraps to INT_MIN: no bound, safe
}

// ---- TMA (cp.async.bulk.tensor) staging of a tile: the depth-key and id planes are 3-D tensors [view][y][x]; a 16x16
// box lands in shared memory exactly in the row-major order the warp uses, rows and columns beyond the frame come in
// as zeros (key 0 never loses) and are clipped on the way back.  One lane issues the copies; their completion is
// counted in bytes on the warp's mbarrier, so the tile travels while the warp already gathers its first batch of
// triangle records.  Frames whose row pitch is not a multiple of 16 bytes (width not a multiple of 4) cannot be
// described to the copy engine and take the LDG / STG path.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 24)) __trap();    // a copy that never lands must fail the launch, not hang the GPU
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int v, uint32_t mbar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(v), "r"(mbar) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int x, int y, int v, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(x), "r"(y), "r"(v) : "memory");
}
struct TileMaps {           // tensor maps of the frame's two planes, passed by value in the kernel parameters
    CUtensorMap key, vis;
};

// MINB = resident CTAs per SM the register allocation aims for.  TRB_RW_BLOCKS picks the instantiation at run time.
constexpr int RW_BLOCKS_DEFAULT = 6;
// SPLIT = false: warp w of CTA (x, view) takes tile x * RW_WARPS + w when its bin holds 1..warp_max triangles, starts from
//   the tile's current contents and writes the tile back.
// SPLIT = true: a warp takes ONE SLICE of a long bin (SplitArgs), starts from an empty tile and delivers its private
//   (key, id) tile to scratch memory; the warp that delivers the last slice of a tile folds all of them into the frame -
//   the lexicographic (depth, id) minimum over the old contents and every slice, i.e. exactly what one warp walking the
//   whole bin would have left.
constexpr unsigned long long KEY_EMPTY = ~0ull;   // above every key a fragment or a cleared pixel can have
template <int MINB, bool SPLIT>
__global__ void __launch_bounds__(RW_WARPS * 32, MINB * (4 / RW_WARPS > 0 ? 4 / RW_WARPS : 1))
k_raster_warp(FrameDev f, RasterArgs a, const __grid_constant__ TileMaps maps, const int use_tma_, SplitArgs sp) {
    __shared__ __align__(128) WarpTile tiles[RW_WARPS];
    __shared__ __align__(8) unsigned long long tile_bar[RW_WARPS];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = RW_WARPS > 1 ? (int)(threadIdx.x >> 5) : 0;
    const int use_tma = SPLIT ? 0 : use_tma_;
    int tile, view;
    size_t tslot;
    uint32_t n, off, item = 0, slice = 0, nslices = 1;
    if constexpr (!SPLIT) {
        tile = blockIdx.x * RW_WARPS + warp; view = blockIdx.y;
        if (tile >= f.ntiles) return;
        tslot = (size_t)view * f.ntiles + tile;
        n = __ldg(a.counts + tslot);
        if (n == 0 || n > a.warp_max || a.ctl->overflow) return;
        off = __ldg(a.offsets + tslot);
    } else {
        item = blockIdx.x * RW_WARPS + warp;
        if (a.ctl->overflow || item >= min(a.ctl->split_n, sp.cap)) return;
        const uint2 it = sp.items[item];
        if (it.x == 0xffffffffu) return;                     // the tail of a bin that found no room in the list
        tslot = it.x;
        slice = it.y & 0xffffu; nslices = it.y >> 16;
        tile = (int)(it.x % (uint32_t)f.ntiles); view = (int)(it.x / (uint32_t)f.ntiles);
        const uint32_t nt = __ldg(a.counts + tslot);
        TRB_CHECK(slice < nslices && (unsigned long long)slice * sp.split_s < nt);
        off = __ldg(a.offsets + tslot) + slice * sp.split_s;
        n = min(sp.split_s, nt - slice * sp.split_s);
    }
    WarpTile& sm = tiles[warp];
    const int tx0 = (tile % f.tw) << TILE_SHIFT, ty0 = (tile / f.tw) << TILE_SHIFT;
    unsigned long long* gz = f.zkey + (size_t)view * f.npix;
    uint32_t* gv = f.vis + (size_t)view * f.npix;
    const TriRec* tr = a.trirec + (size_t)view * a.ntris;
    const unsigned lane_le = FULL >> (31 - lane), lane_lt = lane_le >> 1;

    // first batch's bin entry: in flight while the tile is staged
    TRB_CHECK((unsigned long long)off + n <= a.ctl->total);
    uint32_t t_next = lane < (int)n ? __ldg(a.bins + off + lane) : 0u;
    const uint32_t bar = smem_addr(&tile_bar[warp]);
    if (use_tma) {
        if (lane == 0) {
            mbar_init(bar, 1);
            mbar_expect_tx(bar, (uint32_t)(sizeof(sm.zk) + sizeof(sm.vid)));
            tma_load_3d(smem_addr(sm.zk), &maps.key, tx0, ty0, view, bar);
            tma_load_3d(smem_addr(sm.vid), &maps.vis, tx0, ty0, view, bar);
        }
        __syncwarp();
    } else {
        #pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int p = j * 32 + lane, x = tx0 + (p & 15), y = ty0 + (p >> 4);
            const bool valid = x < f.W && y < f.H;
            const size_t gp = (size_t)y * f.W + x;
            if constexpr (SPLIT) {
                sm.zk[p] = valid ? KEY_EMPTY : 0ull;   // a slice starts from an empty tile
                sm.vid[p] = VIS_NONE;
            } else {
                sm.zk[p] = valid ? gz[gp] : 0ull;      // key 0 never loses: pixels outside the frame stay untouched
                sm.vid[p] = valid ? gv[gp] : VIS_NONE;
            }
        }
        if constexpr (SPLIT) __syncwarp();
    }
    bool tile_landed = !use_tma;
    uint32_t covered = 0, touched = 0;
    unsigned zmin_hi = 0xffffffffu, zmin_lo = 0xffffffffu;   // smallest key written (min_z of our_gl.cpp:197)
    auto apply = [&](int p, unsigned long long key, uint32_t gid) {
        const unsigned long long cur = sm.zk[p];
        if (key < cur || (key == cur && gid < sm.vid[p])) {
            sm.zk[p] = key;
            sm.vid[p] = gid;
            ++touched;                                        // upper bound of the pixels that changed hands
            const unsigned kh = (unsigned)(key >> 32), kl = (unsigned)key;
            if (kh < zmin_hi || (kh == zmin_hi && kl < zmin_lo)) { zmin_hi = kh; zmin_lo = kl; }
        }
    };
    for (uint32_t base = 0; base < n; base += 32) {
        // ---- the batch: every lane takes one triangle of the bin, keeps its record in shared memory and lists the
        //      conservative spans of its rows inside the tile
        const bool has = base + lane < n;
        int X0 = 0, X1 = -1, Y0 = 0, nrows = 0;
        SpanEdges E;
        E.ay = E.a0 = E.a1 = E.a2 = E.b0 = E.b1 = E.b2 = 0.0;
        E.m0 = E.m1 = E.m2 = INT_MIN;
        if (has) {
            const uint32_t t = t_next;
            TRB_CHECK(t < a.ntris);
            const double2* q = reinterpret_cast<const double2*>(tr + t);
            // TriRec: ax ay | s00 s01 | s10 s11 | uz z0 | z1 z2 | ruz bbox
            const double2 r0 = __ldg(q), r1 = __ldg(q + 1), r2 = __ldg(q + 2), r3 = __ldg(q + 3), r4 = __ldg(q + 4), r5 = __ldg(q + 5);
            const uint32_t gid = slot_gid(a.ids, t);
            if (base + 32 + lane < n) t_next = __ldg(a.bins + off + base + 32 + lane);
            const unsigned long long bbw = (unsigned long long)__double_as_longlong(r5.y);
            X0 = max((int)(bbw & 0xffff), tx0); X1 = min((int)((bbw >> 32) & 0xffff), tx0 + TILE - 1);
            Y0 = max((int)((bbw >> 16) & 0xffff), ty0);
            const int Y1 = min((int)(bbw >> 48), ty0 + TILE - 1);
            nrows = Y1 - Y0 + 1;
            TRB_CHECK(X1 >= X0 && nrows >= 1 && nrows <= TILE && X1 - X0 < TILE);
            span_setup(r0, r1, r2, r3.x, X0, X1, Y0, Y1, E);
            double2* d = reinterpret_cast<double2*>(&sm.recs[lane]);
            d[0] = r0; d[1] = r1; d[2] = r2;
            d[3] = make_double2(r3.x, r5.x);             // uz, ruz
            d[4] = make_double2(r3.y, r4.x);             // z0, z1
            const uint32_t flags = exponent_in_window(r3.x) ? 1u : 0u;
            d[5] = make_double2(r4.y, __longlong_as_double((long long)(((unsigned long long)flags << 32) | gid)));
        }
        // ---- the rows of the batch's triangles laid end to end and dealt out 32 at a time (a lane per triangle walking
        //      its own rows would idle two lanes out of three: triangles have ~5 rows, the tallest of a batch 16).  A lane
        //      finds the triangle of its row like a sample finds its span below, and fetches that triangle's bounds from
        //      the lane that owns it with shuffles.
        uint32_t nspans = 0;
        uint32_t rincl = (uint32_t)nrows;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, rincl, o);
            if (lane >= o) rincl += y;
        }
        const uint32_t rfirst = rincl - (uint32_t)nrows, R = __shfl_sync(FULL, rincl, 31);
        const uint32_t rfirst_s = nrows > 0 ? rfirst : 0x7fffffffu;
        // tile-local bbox corner + which side each edge bounds, in one word
        const uint32_t geo = (uint32_t)(X0 - tx0) | ((uint32_t)(X1 - tx0) << 4) | ((uint32_t)(Y0 - ty0) << 8) |
                             (E.m0 == INT_MAX ? 0x1000u : 0u) | (E.m1 == INT_MAX ? 0x2000u : 0u) | (E.m2 == INT_MAX ? 0x4000u : 0u);
        uint32_t rbefore = 0;
        for (uint32_t rb = 0; rb < R; rb += 32) {
            unsigned rstart;
            asm("shl.b32 %0, 1, %1;" : "=r"(rstart) : "r"(rfirst_s - rb));
            const unsigned rstarts = __reduce_or_sync(FULL, rstart);
            const int e = (int)rbefore + __popc(rstarts & lane_le) - 1;
            rbefore += (uint32_t)__popc(rstarts);
            TRB_CHECK(e >= 0 && e < 32);
            const bool ract = rb + lane < R;
            const int r = (int)(rb + lane - __shfl_sync(FULL, rfirst, e));
            SpanEdges Ee;
            Ee.ay = __shfl_sync(FULL, E.ay, e);
            Ee.a0 = __shfl_sync(FULL, E.a0, e); Ee.b0 = __shfl_sync(FULL, E.b0, e);
            Ee.a1 = __shfl_sync(FULL, E.a1, e); Ee.b1 = __shfl_sync(FULL, E.b1, e);
            Ee.a2 = __shfl_sync(FULL, E.a2, e); Ee.b2 = __shfl_sync(FULL, E.b2, e);
            const uint32_t ge = __shfl_sync(FULL, geo, e);
            Ee.m0 = (ge & 0x1000u) ? INT_MAX : INT_MIN;
            Ee.m1 = (ge & 0x2000u) ? INT_MAX : INT_MIN;
            Ee.m2 = (ge & 0x4000u) ? INT_MAX : INT_MIN;
            const int eX0 = tx0 + (int)(ge & 15u), eX1 = tx0 + (int)((ge >> 4) & 15u), ey = ty0 + (int)((ge >> 8) & 15u) + r;
            int xa = 0, xb = -1;
            if (ract) span_of_row(Ee, Ee.ay - pixel_centre(ey), eX0, eX1, xa, xb);
#if defined(TRB_DEBUG_CHECKS)
            if (ract) {           // the skipped columns must all fail the reference's test
                const double2* q = reinterpret_cast<const double2*>(&sm.recs[e]);
                TriSetup ts;
                ts.ax = q[0].x; ts.ay = q[0].y; ts.s00 = q[1].x; ts.s01 = q[1].y; ts.s10 = q[2].x; ts.s11 = q[2].y;
                ts.uz = q[3].x; ts.ruz = q[3].y; ts.z0 = q[4].x; ts.z1 = q[4].y; ts.z2 = q[5].x;
                ts.x0 = ts.y0 = ts.x1 = ts.y1 = 0;
                assert(r >= 0 && ey < ty0 + TILE);
                for (int x = eX0; x <= eX1; ++x) {
                    double bb[3], zz;
                    if (x < xa || x > xb) assert(!eval_sample(ts, x, ey, bb, zz));
                }
            }
#endif
            const bool ne = xa <= xb;
            const unsigned nb = __ballot_sync(FULL, ne);
            if (ne) {
                TRB_CHECK(xa >= eX0 && xb <= eX1 && nspans + __popc(nb & lane_lt) < (unsigned)RW_SPAN_CAP);
                sm.spans[nspans + __popc(nb & lane_lt)] = (uint32_t)e | ((uint32_t)(xa - tx0) << 5) |
                                                          ((uint32_t)(ey - ty0) << 9) | ((uint32_t)(xb - xa) << 13);
            }
            nspans += (uint32_t)__popc(nb);
        }
        __syncwarp();
        if (!tile_landed) {            // the tile has been travelling while the records were gathered and the spans listed
            mbar_wait(bar, 0);
            tile_landed = true;
        }
        // ---- the spans, 32 at a time: their samples are laid end to end and dealt out to the lanes
        for (uint32_t g = 0; g < nspans; g += 32) {
            const bool sv = g + lane < nspans;
            const uint32_t ent = sv ? sm.spans[g + lane] : 0u;
            const uint32_t len = sv ? (ent >> 13) + 1u : 0u;
            uint32_t incl = len;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += y;
            }
            const uint32_t first = incl - len, S = __shfl_sync(FULL, incl, 31);
            const uint32_t first_s = sv ? first : 0x7fffffffu;   // lanes without a span never start one
            uint32_t before = 0;                     // spans of the group that start before the current window
            for (uint32_t bs = 0; bs < S; bs += 32) {
                // which span does sample bs + lane belong to?  The spans' first samples inside this window of 32 form a
                // bit mask: one REDUX, one POPC.  shl.b32 yields 0 for shift counts >= 32: spans that started in an
                // earlier window (the difference wraps) or start in a later one contribute nothing
                unsigned startbit;
                asm("shl.b32 %0, 1, %1;" : "=r"(startbit) : "r"(first_s - bs));
                const unsigned starts = __reduce_or_sync(FULL, startbit);
                const int j = (int)before + __popc(starts & lane_le) - 1;
                before += (uint32_t)__popc(starts);
                TRB_CHECK(j >= 0 && j < 32);
                const uint32_t pj = __shfl_sync(FULL, ent | (first << 17), j);   // span (17 bits) and its first sample (<= 512) in one shuffle
                const uint32_t ent_j = pj & 0x1ffffu;
                const uint32_t l = bs + lane - (pj >> 17);
                const bool act = bs + lane < S;
                const int p = (int)(((ent_j >> 5) & 255u) + l);       // row * 16 + first column + l
                const double2* q = reinterpret_cast<const double2*>(&sm.recs[ent_j & 31u]);
                const double2 r5 = q[5];
                const unsigned long long gf = (unsigned long long)__double_as_longlong(r5.y);
                const uint32_t gid_e = (uint32_t)gf;
                bool frag = false;
                unsigned long long key = 0;
                if (act) {
                    TRB_CHECK(p >= 0 && p < TILE * TILE && l <= (ent_j >> 13));
                    const double2 r0 = q[0], r1 = q[1], r2 = q[2], r3 = q[3], r4 = q[4];
                    const double z = eval_sample_fast(reinterpret_cast<const double*>(q), r0.x, r0.y, r1.x, r1.y, r2.x, r2.y, r3.x,
                                                      r3.y, r4.x, r4.y, r5.x, (gf >> 32) != 0ull, tx0 + (p & 15), ty0 + (p >> 4));
                    if (finite_d(z)) {                                // our_gl.cpp:160 (NaN: not covered)
                        frag = true;
                        key = depth_key_dev(__dadd_rn(z, 0.0));      // fragment_key: -0.0 + 0.0 == +0.0, everything else unchanged
                        ++covered;
                    }
                }
                // apply: plain read-compare-write of (key, id).  Lanes of ONE triangle hit distinct pixels; lanes of different
                // triangles may meet in a pixel: they then take turns, so the outcome is the lexicographic minimum
                // whatever the order.
                bool clash = false;
                unsigned rank = 0;
                // the pixel's current (key, id): requested before the collision test so that the two latencies overlap
                const unsigned long long cur0 = sm.zk[frag ? p : 0];
                const uint32_t vid0 = sm.vid[frag ? p : 0];
                if (starts >> 1) {
                    // do two lanes of the window hit the same pixel?  Every lane with a fragment writes its number into the
                    // pixel's claim byte and reads it back: when two lanes share a pixel one of them finds the other's
                    // number.  (MATCH.ANY answers the same question, but the instruction behind it collected 16 % of the
                    // kernel's stall samples; it is kept for the rare window that does collide.)
                    if (frag) sm.claim[p] = (uint8_t)lane;
                    __syncwarp();
                    const bool lost = frag && sm.claim[p] != (uint8_t)lane;
                    if (__any_sync(FULL, lost)) {
                        const unsigned peers = __match_any_sync(FULL, frag ? (unsigned)p : 256u + (unsigned)lane);
                        rank = __popc(peers & lane_lt);
                        clash = true;
                    }
                }
                if (!clash) {
                    if (frag && (key < cur0 || (key == cur0 && gid_e < vid0))) {
                        sm.zk[p] = key;
                        sm.vid[p] = gid_e;
                        ++touched;
                        const unsigned kh = (unsigned)(key >> 32), kl = (unsigned)key;
                        if (kh < zmin_hi || (kh == zmin_hi && kl < zmin_lo)) { zmin_hi = kh; zmin_lo = kl; }
                    }
                } else {
                    const unsigned turns = __reduce_max_sync(FULL, rank);
                    for (unsigned r = 0; r <= turns; ++r) {
                        if (frag && rank == r) apply(p, key, gid_e);
                        __syncwarp();
                    }
                }
                __syncwarp();
            }
        }
        __syncwarp();        // the records and the span list are rewritten by the next batch
    }
    // A tile without a single covered sample is left alone; otherwise the whole tile is stored (no read-back of the
    // old values to find out what changed: the loads cost more than the stores of unchanged pixels save)
    covered = __reduce_add_sync(FULL, covered);
    if constexpr (SPLIT) {
        // deliver the private tile, then count this slice in; the last one to arrive folds the tile's slices into the frame
        __syncwarp();
        unsigned long long* sk = sp.keys + (size_t)item * (TILE * TILE);
        uint32_t* si = sp.ids + (size_t)item * (TILE * TILE);
        #pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int p = j * 32 + lane;
            __stcg(sk + p, sm.zk[p]);
            __stcg(si + p, sm.vid[p]);
        }
        __threadfence();
        __syncwarp();
        uint32_t arrived = 0;
        if (lane == 0) {
            arrived = atomicAdd(sp.done + tslot, 1u);
            if (covered) atomicAdd(&f.stats[view].frag_covered, (unsigned long long)covered);
        }
        arrived = __shfl_sync(FULL, arrived, 0);
        if (arrived != nslices - 1u) return;
        __threadfence();
        const size_t first_item = (size_t)(item - slice);
        uint32_t changed = 0;
        unsigned long long zmin = ~0ull;
        #pragma unroll 1
        for (int j = 0; j < 8; ++j) {
            const int p = j * 32 + lane, x = tx0 + (p & 15), y = ty0 + (p >> 4);
            if (x >= f.W || y >= f.H) continue;
            const size_t gp = (size_t)y * f.W + x;
            const unsigned long long k_in = gz[gp];
            const uint32_t id_in = gv[gp];
            unsigned long long bk = k_in;
            uint32_t bi = id_in;
            for (uint32_t s = 0; s < nslices; ++s) {
                const unsigned long long k = __ldcg(sp.keys + (first_item + s) * (TILE * TILE) + p);
                if (k > bk) continue;
                const uint32_t id = __ldcg(sp.ids + (first_item + s) * (TILE * TILE) + p);
                if (k < bk) { bk = k; bi = id; }
                else if (id < bi) bi = id;                 // ties: lowest id = first submitted
            }
            if (bk != k_in || bi != id_in) {
                gz[gp] = bk;
                gv[gp] = bi;
                ++changed;
                if (bk != k_in) zmin = min(zmin, bk);      // min_z of our_gl.cpp:197, as in raster_tile_cta
            }
        }
        changed = __reduce_add_sync(FULL, changed);
        for (int o = 16; o; o >>= 1) zmin = min(zmin, __shfl_xor_sync(FULL, zmin, o));
        if (lane == 0 && changed) {
            atomicAdd(&f.stats[view].touched, (unsigned long long)changed);
            atomicMin(&f.stats[view].zmin_key, zmin);
        }
        return;
    }
    if (covered == 0) return;
    touched = __reduce_add_sync(FULL, touched);
    if (touched) {
        if (use_tma) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the lanes' shared-memory writes -> the copy engine
            __syncwarp();
            if (lane == 0) {
                tma_store_3d(&maps.key, tx0, ty0, view, smem_addr(sm.zk));
                tma_store_3d(&maps.vis, tx0, ty0, view, smem_addr(sm.vid));
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the tile has left shared memory
            }
        } else {
            #pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int p = j * 32 + lane, x = tx0 + (p & 15), y = ty0 + (p >> 4);
                if (x < f.W && y < f.H) {
                    const size_t gp = (size_t)y * f.W + x;
                    gz[gp] = sm.zk[p];
                    gv[gp] = sm.vid[p];
                }
            }
        }
    }
    unsigned long long zmin = ((unsigned long long)zmin_hi << 32) | zmin_lo;
    for (int o = 16; o; o >>= 1) zmin = min(zmin, __shfl_xor_sync(FULL, zmin, o));
    if (lane == 0) {
        atomicAdd(&f.stats[view].frag_covered, (unsigned long long)covered);
        atomicMin(&f.stats[view].zmin_key, zmin);
        if (touched) atomicAdd(&f.stats[view].touched, (unsigned long long)touched);
    }
}

// the id pass of the unbinned fallback (see unbinned_pass); the depth pass runs inside k_raster
template <bool IDS>
__global__ void __launch_bounds__(TPB) k_unbinned(FrameDev f, uint32_t ntris, SlotIds ids,
                                                  const uint2* __restrict__ tribox, const TriRec* __restrict__ trirec,
                                                  const DrawCtl* __restrict__ ctl) {
    if (!ctl->overflow) return;
    const int view = blockIdx.y;
    unbinned_pass<IDS>(f.zkey, f.vis, f.stats, f.npix, f.W, f.H, view, view + 1, ntris, ids, tribox, trirec,
                       blockIdx.x * (TPB / 32) + (threadIdx.x >> 5), gridDim.x * (TPB / 32));
}

// ---------------------------------------------------------------------------------------------
// flush: the fragment() calls of our_gl.cpp:187-192, once per visible pixel
// ---------------------------------------------------------------------------------------------
constexpr int SHADE_MAX_SM_DRAWS = 32;
static_assert(sizeof(DrawDev) % 4 == 0, "DrawDev is staged word by word");
__device__ __forceinline__ void stage_draw_table(DrawDev* sm_draws, const DrawDev* __restrict__ draws, int ndraws) {
    const int words = min(ndraws, SHADE_MAX_SM_DRAWS) * (int)(sizeof(DrawDev) / 4);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(draws);
    uint32_t* dst = reinterpret_cast<uint32_t*>(sm_draws);
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
}
constexpr int SHADE_PX_PER_THREAD = 4;   // one 16-byte id load per thread: sparse frames (configs 4, 5) stay cheap

// PhongShader / EyeShader / shadow-mapped Phong for one pixel with fp32 lighting: `at` = the three vertices' raw
// attributes, pc = the perspective-correct barycentrics (fp64, exact).  Texture coordinates, the texel choice and the
// shadow test stay fp64 (fastshade.cuh explains why that keeps every channel within one code of the reference).
// Everything the pixel reads about its draw comes from the draw's LitF block (eight 16-byte loads).
__device__ __forceinline__ uint32_t texel_index32(uint32_t w, uint32_t h, double u, double v) {   // model.cpp:420-423
    const int x = clamp_i(x86_int(u * (double)(int)w), 0, (int)w - 1);
    const int y = clamp_i(x86_int(v * (double)(int)h), 0, (int)h - 1);
    return (uint32_t)x + (uint32_t)y * w;          // < 2^32: upload_texture caps w, h at 65536 and w * h below 2^32
}
// TGAColor(p, bpp) (tgaimage.h:47-51): the first min(bpp, 3) bytes, the rest 0
__device__ __forceinline__ void fetch3(const uint8_t* px, uint32_t ti, uint32_t bpp, int c[3]) {
    const uint8_t* q = px + (size_t)ti * bpp;
    c[0] = (int)__ldg(q);
    c[1] = bpp > 1 ? (int)__ldg(q + 1) : 0;
    c[2] = bpp > 2 ? (int)__ldg(q + 2) : 0;
}
template <bool C2>
__device__ __forceinline__ void lit_fast(const DrawDev& D, int view, const float (*at)[8], const double pc[3], uint8_t col[3]) {
    const bool shadowed = C2 && D.kind == 4 /*SHADOW_PHONG*/;
    const trbf::LitF* LP = reinterpret_cast<const trbf::LitF*>(D.litf) + view;
    const trbf::LitTex LF = trbf::load_lit_tex(LP);
    const double tu = (double)at[0][6] * pc[0] + (double)at[1][6] * pc[1] + (double)at[2][6] * pc[2];  // main.cpp:100-101
    const double tv = (double)at[0][7] * pc[0] + (double)at[1][7] * pc[1] + (double)at[2][7] * pc[2];
    // the maps are sampled at the same (u, v): one texel index serves both when they have one size
    int base[3] = {255, 255, 255}, nmc[3] = {0, 0, 0};
    uint32_t ti = 0;
    if (LF.dbpp) {
        ti = texel_index32(LF.dw, LF.dh, tu, tv);
        fetch3(LF.diffuse, ti, LF.dbpp, base);
    }
    const bool eye = D.kind == 2 /*EYE*/;
    const bool has_nm = !eye && LF.nbpp != 0;
    if (has_nm) {
        if (!LF.dbpp || LF.nw != LF.dw || LF.nh != LF.dh) ti = texel_index32(LF.nw, LF.nh, tu, tv);
        fetch3(LF.normal, ti, LF.nbpp, nmc);
    }
    float sf = 1.0f;
    if (shadowed) {
        const ShadowUniformsDev& SU = reinterpret_cast<const ShadowUniformsDev*>(D.uniforms)[view];
        double lc[3][4];
        for (int k = 0; k < 3; ++k)
            light_clip_from_position(SU.shadow, (double)at[k][0], (double)at[k][1], (double)at[k][2], lc[k]);
        sf = (float)shadow_factor(SU.shadow, lc, pc);
    }
    trbf::shade_lit_f32(eye, LP, at, pc, base, has_nm, nmc, sf, col);
}

// one visible pixel: p = x + y*W inside `view`, id = its winning triangle.  C2 = the frame has a
// config-2 shader (SHADOW_PHONG / GOURAUD); frames without one run the instantiation that does not
// carry their registers.
// base + i * STRIDE for a 32-bit element index as ONE multiply-add (the compiler's shift-and-add form costs two
// instructions per address, and a pixel forms nine of them)
template <unsigned STRIDE>
__device__ __forceinline__ const void* elem_addr(const void* base, uint32_t i) {
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(i), "r"(STRIDE), "l"((unsigned long long)base));
    return reinterpret_cast<const void*>(r);
}

// id -> (draw, triangle of the draw's mesh).  false: a winner of another rank's mesh that this context does not hold.
// `tab` is the draw table: the CTA's shared-memory copy when the frame has at most SHADE_MAX_SM_DRAWS draws (one level
// less in the dependent chain id -> draw -> indices -> records -> texels), else the table in global memory - never a
// per-access choice between the two, which would turn every field read into a generic load.
__device__ __forceinline__ bool resolve_winner(const DrawDev* tab, int ndraws, uint32_t id, int& draw, uint32_t& g0) {
    int lo = 0;                   // last draw with id_base < id (bases ascend)
    if (ndraws <= 8) {            // a frame loop's handful of draws: count instead of bisecting
        #pragma unroll 1
        for (int d = 1; d < ndraws; ++d) lo += tab[d].id_base < id ? 1 : 0;
    } else {
        int hi = ndraws - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (tab[mid].id_base < id) lo = mid; else hi = mid - 1;
        }
    }
    uint32_t t = id - tab[lo].id_base - 1u;              // triangle inside the range this context drew
    uint32_t first = tab[lo].first_tri;
    if (t >= tab[lo].ntris) {
        // a winner another rank rasterised (sort-last composite): find the draw whose MESH holds it
        int found = -1;
        for (int d = 0; d < ndraws && found < 0; ++d) {
            const long long g = (long long)id - tab[d].mesh_id_base - 1;
            if (g >= 0 && g < (long long)tab[d].mesh_ntris) found = d;
        }
        if (found < 0) return false;                     // not ours to shade
        lo = found;
        first = tab[found].first_tri;
        t = (uint32_t)((long long)id - tab[found].mesh_id_base - 1) - first;  // may wrap: first_tri + t is exact mod 2^32
    }
    draw = lo;
    g0 = first + t;                                      // triangle index in the mesh
    return true;
}
__device__ __forceinline__ void winner_vertices(const DrawDev* tab, int draw, uint32_t g0, uint32_t vi[3]) {
    const uint32_t* idx = tab[draw].idx;
    if (idx) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(elem_addr<12>(idx, g0));
        vi[0] = __ldg(q); vi[1] = __ldg(q + 1); vi[2] = __ldg(q + 2);
    } else {
        const uint32_t* inv = tab[draw].inv_perm;                        // ordered soup: the triangle's slot
        const uint32_t s = inv ? __ldg(inv + g0) : g0;
        vi[0] = s * 3u; vi[1] = s * 3u + 1u; vi[2] = s * 3u + 2u;        // implicit soup
    }
}

// one visible pixel (x, y) of `view`, p = x + y*W: its winner is triangle g0 (vertices i0, i1, i2) of draw `draw`
// vertex record of a winner's vertex.  LAZYV (the composite's shade pass): a vertex the rank's masked vertex stage skipped
// is transformed here - the same vrec_from_position on the same inputs as k_vertex_mesh, hence the same bits
template <bool LAZYV>
__device__ __forceinline__ VRec winner_vrec(const FrameDev& f, const DrawDev& D, int view, const VRec* vr, uint32_t i) {
    if (LAZYV && D.vmark && !__ldg(D.vmark + i)) {
        const float4 q = __ldg(D.pos4 + i);
        const double* m = D.mats + (size_t)view * 32;
        return vrec_from_position(m, m + 16, f.viewport, (double)q.x, (double)q.y, (double)q.z);
    }
    return load_vrec(reinterpret_cast<const VRec*>(elem_addr<32>(vr, i)));
}
template <bool C2, bool FAST, bool LAZYV>
__device__ __forceinline__ void shade_resolved(const FrameDev& f, const DrawDev* tab, int view, unsigned long long p, int x, int y,
                                               int draw, uint32_t g0, uint32_t i0, uint32_t i1, uint32_t i2) {
    const size_t gp = (size_t)view * f.npix + p;
    const DrawDev D = tab[draw];
    const VRec* vr = D.vrec + (size_t)view * D.nverts;
    const VRec va = winner_vrec<LAZYV>(f, D, view, vr, i0), vb = winner_vrec<LAZYV>(f, D, view, vr, i1),
               vc = winner_vrec<LAZYV>(f, D, view, vr, i2);
    TriSetup ts;
    setup_known_triangle(va, vb, vc, ts);   // a recorded winner passed every reject: no tests, no bbox
    double b[3], z, pc[3];
    eval_known_sample(ts, x, y, b, z);          // a recorded winner is covered: same arithmetic, no coverage tests
    {
        // exact bits of the reference's zbuffer[idx]: the stored key already is K(z) unless z is -0.0
        if (z == 0.0) f.zkey[gp] = depth_key(z);
        perspective_bary(b, va.iw, vb.iw, vc.iw, pc);
        uint8_t col[3];
        bool write = true;
        if (D.kind == 0 /*FLAT_BARY*/) {
            shade_flat_bary(pc, col);
        } else if (D.kind == 3 /*DEPTH*/) {
            write = false;
        } else if (FAST && !(C2 && D.kind == 5 /*GOURAUD*/)) {
            // fp32 lighting (fastshade.cuh): texture coordinates, texel choice and the shadow test stay fp64
            const uint32_t vi[3] = {i0, i1, i2};
            float at[3][8];
            #pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4* q = reinterpret_cast<const float4*>(elem_addr<32>(D.attr8, vi[k]));
                const float4 q0 = __ldg(q), q1 = __ldg(q + 1);
                at[k][0] = q0.x; at[k][1] = q0.y; at[k][2] = q0.z; at[k][3] = q0.w;
                at[k][4] = q1.x; at[k][5] = q1.y; at[k][6] = q1.z; at[k][7] = q1.w;
            }
            lit_fast<C2>(D, view, at, pc, col);
        } else if (!FAST || C2) {   // fp64 lighting; the fp32 kernels only carry it for GOURAUD
            const double* MV = D.mats + (size_t)view * 32;
            Varyings vy;
            if (!FAST && D.varyings) {
                const double* q = D.varyings + (size_t)g0 * 24;
                for (int k = 0; k < 3; ++k) {
                    vy.u[k] = q[k * 8]; vy.v[k] = q[k * 8 + 1];
                    vy.pos_eye[k] = D3{q[k * 8 + 2], q[k * 8 + 3], q[k * 8 + 4]};
                    vy.nrm_eye[k] = D3{q[k * 8 + 5], q[k * 8 + 6], q[k * 8 + 7]};
                }
            } else {
                const uint32_t vi[3] = {i0, i1, i2};
                for (int k = 0; k < 3; ++k) {
                    const float4* q = reinterpret_cast<const float4*>(D.attr8 + (size_t)vi[k] * 8);
                    float4 q0 = __ldg(q), q1 = __ldg(q + 1);
                    float at[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                    varyings_from_attr(MV, at, k, vy);
                }
            }
            if (C2 && D.kind == 5 /*GOURAUD*/) {
                shade_gouraud(reinterpret_cast<const LitUniforms*>(D.uniforms)[view], vy, pc, col);
            } else if (FAST) {
                write = false;   // unreachable: the host only runs the fp32 kernels when every lit draw has a LitF block
            } else if (C2 && D.kind == 4 /*SHADOW_PHONG*/) {
                const ShadowUniformsDev& SU = reinterpret_cast<const ShadowUniformsDev*>(D.uniforms)[view];
                double lc[3][4];
                const uint32_t vj[3] = {i0, i1, i2};
                for (int k = 0; k < 3; ++k) {
                    const float* q = D.attr8 + (size_t)vj[k] * 8;
                    light_clip_from_position(SU.shadow, (double)__ldg(q), (double)__ldg(q + 1), (double)__ldg(q + 2), lc[k]);
                }
                shade_lit(false, MV, SU.lit, vy, pc, col, shadow_factor(SU.shadow, lc, pc));
            } else {
                shade_lit(D.kind == 2 /*EYE*/, MV, reinterpret_cast<const LitUniforms*>(D.uniforms)[view], vy, pc, col);
            }
        }
        if (write) {
            uint8_t* c = f.color + gp * 3;
            c[0] = col[0]; c[1] = col[1]; c[2] = col[2];
        }
    }
}
template <bool C2, bool FAST, bool LAZYV>
__device__ __forceinline__ void shade_pixel_tab(const FrameDev& f, const DrawDev* tab, int ndraws, int view, unsigned long long p,
                                                int x, int y, uint32_t id) {
    int draw;
    uint32_t g0, vi[3];
    if (!resolve_winner(tab, ndraws, id, draw, g0)) return;
    winner_vertices(tab, draw, g0, vi);
    shade_resolved<C2, FAST, LAZYV>(f, tab, view, p, x, y, draw, g0, vi[0], vi[1], vi[2]);
}
template <bool C2, bool FAST, bool LAZYV = false>
__device__ __forceinline__ void shade_pixel(const FrameDev& f, const DrawDev* __restrict__ draws, int ndraws,
                                            const DrawDev* sm_draws, int view, unsigned long long p, int x, int y, uint32_t id) {
    // ONE table per frame: the shared-memory copy, or - frames with more draws than it holds (long immediate-mode
    // frames) - the table in global memory
    shade_pixel_tab<C2, FAST, LAZYV>(f, ndraws <= SHADE_MAX_SM_DRAWS ? sm_draws : draws, ndraws, view, p, x, y, id);
}

// The flush picks its kernel on the device (no host round trip): sparse frames (configs 4, 5) shade
// through a compacted pixel list so that warps stay full, dense frames take one thread per pixel.
__global__ void k_shade_decide(FrameDev f, unsigned long long rows_px) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < f.nviews; v += gridDim.x * blockDim.x) {
        DevStats* s = f.stats + v;
        s->shade_mode = (s->touched * 2 < rows_px) ? 1 : 0;
        s->touched = 0;
        s->list_len = 0;
    }
}

// Pass 1 of the flush: compact the pixels that have an unshaded winner into a dense list (one
// 16-byte id load per thread, one atomicAdd per warp), so that the expensive shade pass runs with
// full warps however sparse the frame is (config 5 touches ~10 % of its 67 M pixels).
__global__ void __launch_bounds__(TPB) k_shade_collect(FrameDev f, int row0, int row1, uint32_t* __restrict__ list) {
    const int view = blockIdx.y;
    if (!f.stats[view].shade_mode) return;
    unsigned long long* count = &f.stats[view].list_len;
    const unsigned long long first = (unsigned long long)row0 * f.W, last = (unsigned long long)row1 * f.W;
    const uint32_t* vis = f.vis + (size_t)view * f.npix;
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long per_block = (unsigned long long)TPB * SHADE_PX_PER_THREAD;
    // grid-stride over blocks of TPB * 4 pixels: dense views cost a handful of CTAs that exit at once
    for (unsigned long long b0 = first + blockIdx.x * per_block; b0 < last; b0 += gridDim.x * per_block) {
        const unsigned long long p0 = b0 + (unsigned long long)threadIdx.x * SHADE_PX_PER_THREAD;
        uint32_t ids[SHADE_PX_PER_THREAD];
        for (int j = 0; j < SHADE_PX_PER_THREAD; ++j) ids[j] = VIS_NONE;
        if (p0 < last) {
            if (p0 + SHADE_PX_PER_THREAD <= last && ((reinterpret_cast<uintptr_t>(vis + p0) & 15) == 0)) {
                const uint4 v = *reinterpret_cast<const uint4*>(vis + p0);
                ids[0] = v.x; ids[1] = v.y; ids[2] = v.z; ids[3] = v.w;
            } else {
                for (int j = 0; j < SHADE_PX_PER_THREAD; ++j)
                    if (p0 + j < last) ids[j] = vis[p0 + j];
            }
        }
        unsigned mine = 0;
        for (int j = 0; j < SHADE_PX_PER_THREAD; ++j) mine += (ids[j] != VIS_NONE && ids[j] != VIS_SHADED) ? 1u : 0u;
        // warp-aggregated append
        unsigned incl = mine;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += y;
        }
        const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        unsigned long long base = 0;
        if (lane == 31) base = atomicAdd(count, (unsigned long long)total);
        base = __shfl_sync(0xffffffffu, base, 31);
        uint32_t* out = list + (size_t)view * f.npix + base + (incl - mine);
        for (int j = 0; j < SHADE_PX_PER_THREAD; ++j)
            if (ids[j] != VIS_NONE && ids[j] != VIS_SHADED) *out++ = (uint32_t)(p0 + j);
    }
}

// The same list for a flush inside a tile-granular depth snapshot window (k_snap_save above): the snapshot flushed, so every
// pixel drawn since lies in a tile whose byte in `saved` is set - the list comes from those tiles instead of from a pass
// over the whole id plane (config 3: the eyes' flush read 265 MB of ids to find a few thousand pixels).
__global__ void __launch_bounds__(TPB) k_shade_collect_tiles(FrameDev f, uint32_t nslots, const uint8_t* __restrict__ saved,
                                                             int row0, int row1, uint32_t* __restrict__ list) {
    __shared__ uint32_t tiles[TPB];
    __shared__ uint32_t n_sh;
    const uint32_t slot = blockIdx.x * TPB + threadIdx.x;
    if (threadIdx.x == 0) n_sh = 0;
    __syncthreads();
    if (slot < nslots && saved[slot]) tiles[atomicAdd(&n_sh, 1u)] = slot;
    __syncthreads();
    const uint32_t n = n_sh;
    const unsigned lane = threadIdx.x & 31;
    constexpr int N = TILE * TILE / 32;
    for (uint32_t i = threadIdx.x >> 5; i < n; i += TPB / 32) {
        const uint32_t s = tiles[i], view = s / (uint32_t)f.ntiles, tile = s - view * (uint32_t)f.ntiles;
        if (!f.stats[view].shade_mode) continue;                       // a dense view: k_shade_dense takes it
        const int x = ((int)(tile % (uint32_t)f.tw) << TILE_SHIFT) + (int)(lane & (TILE - 1));
        const int y0 = ((int)(tile / (uint32_t)f.tw) << TILE_SHIFT) + (int)(lane >> TILE_SHIFT);
        const uint32_t* vis = f.vis + (size_t)view * f.npix;
        uint32_t ids[N];
        unsigned mine = 0;
#pragma unroll
        for (int it = 0; it < N; ++it) {
            const int y = y0 + 2 * it;
            ids[it] = (x < f.W && y >= row0 && y < row1) ? vis[(uint32_t)y * (uint32_t)f.W + (uint32_t)x] : VIS_NONE;
        }
#pragma unroll
        for (int it = 0; it < N; ++it) mine += (ids[it] != VIS_NONE && ids[it] != VIS_SHADED) ? 1u : 0u;
        unsigned incl = mine;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += up;
        }
        const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        unsigned long long base = 0;
        if (lane == 31) base = atomicAdd(&f.stats[view].list_len, (unsigned long long)total);
        base = __shfl_sync(0xffffffffu, base, 31);
        uint32_t* out = list + (size_t)view * f.npix + base + (incl - mine);
#pragma unroll
        for (int it = 0; it < N; ++it)
            if (ids[it] != VIS_NONE && ids[it] != VIS_SHADED) *out++ = (uint32_t)(y0 + 2 * it) * (uint32_t)f.W + (uint32_t)x;
    }
}

// Dense frames (most pixels have an unshaded winner): one thread per pixel, no list
#ifndef TRB_SHADE_MIN_BLOCKS
#define TRB_SHADE_MIN_BLOCKS 4
#endif
#ifndef TRB_SHADE_DENSE_PX
#define TRB_SHADE_DENSE_PX 8   // measured on config 3 (ms per step): 1 -> 1.975, 2 -> 1.81, 4 -> 1.70, 8 -> 1.665, 16 -> 1.654
#endif
constexpr int SHADE_DENSE_PX = TRB_SHADE_DENSE_PX;   // pixels per thread of k_shade_dense
// (8x4 pixel warp footprints instead of 32 pixels of a row were measured on config 3: 1.18 ms vs 1.10 ms - not kept)
template <bool C2, bool FAST>
__global__ void __launch_bounds__(TPB, TRB_SHADE_MIN_BLOCKS) k_shade_dense(FrameDev f, const DrawDev* __restrict__ draws, int ndraws,
                                                     int row0, int row1) {
    if (f.stats[blockIdx.y].shade_mode) return;
    __shared__ DrawDev sm_draws[SHADE_MAX_SM_DRAWS];
    stage_draw_table(sm_draws, draws, ndraws);
    const int view = blockIdx.y;
    // SHADE_DENSE_PX pixels per thread, TPB apart: their ids are requested together, so the first (cold: the id plane was
    // just written by the raster pass and is larger than L2) load of all of them is one wait instead of one each
    const unsigned long long first = (unsigned long long)row0 * f.W, last = (unsigned long long)row1 * f.W;
    const unsigned long long p0 = first + (unsigned long long)blockIdx.x * (TPB * SHADE_DENSE_PX) + threadIdx.x;
    uint32_t* vis = f.vis + (size_t)view * f.npix;
    uint32_t ids[SHADE_DENSE_PX];
    #pragma unroll
    for (int k = 0; k < SHADE_DENSE_PX; ++k) {
        const unsigned long long p = p0 + (unsigned long long)k * TPB;
        ids[k] = p < last ? vis[p] : VIS_NONE;
    }
    // (requesting the index triples of all the pixels up front as well was measured: 1.42 vs 1.30 ms - the kernel is bound
    // by instruction issue, and the extra local-memory traffic costs more than the overlapped wait saves)
    // pixel coordinates by one division per thread, then TPB columns further per pixel
    int y = (int)((uint32_t)p0 / (uint32_t)f.W);         // p < 2^32: a frame has fewer than 2^32 pixels
    int x = (int)((uint32_t)p0 - (uint32_t)y * (uint32_t)f.W);
    uint32_t* visp = vis + p0;
    #pragma unroll 1
    for (int k = 0; k < SHADE_DENSE_PX; ++k, x += TPB, visp += TPB) {
        while (x >= f.W) { x -= f.W; ++y; }
        const uint32_t id = ids[k];
        if (id == VIS_NONE || id == VIS_SHADED) continue;
        shade_pixel<C2, FAST>(f, draws, ndraws, sm_draws, view, p0 + (unsigned long long)k * TPB, x, y, id);
        *visp = VIS_SHADED;
    }
}

// Sparse frames, pass 2: persistent grid-stride loop over the compacted list
template <bool C2, bool FAST>
__global__ void __launch_bounds__(TPB, 3) k_shade(FrameDev f, const DrawDev* __restrict__ draws, int ndraws,
                                               const uint32_t* __restrict__ list) {
    if (!f.stats[blockIdx.y].shade_mode) return;
    __shared__ DrawDev sm_draws[SHADE_MAX_SM_DRAWS];
    stage_draw_table(sm_draws, draws, ndraws);
    const int view = blockIdx.y;
    const unsigned long long n = f.stats[view].list_len;
    const uint32_t* mylist = list + (size_t)view * f.npix;
    uint32_t* vis = f.vis + (size_t)view * f.npix;
    // four list entries per trip, their pixel numbers and ids requested before the first of them is shaded
    const unsigned long long stride = (unsigned long long)gridDim.x * TPB;
    for (unsigned long long i0 = (unsigned long long)blockIdx.x * TPB + threadIdx.x; i0 < n; i0 += 4 * stride) {
        uint32_t ps[4], ids[4];
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned long long i = i0 + (unsigned long long)k * stride;
            ps[k] = i < n ? mylist[i] : 0xffffffffu;
        }
        #pragma unroll
        for (int k = 0; k < 4; ++k) ids[k] = ps[k] != 0xffffffffu ? vis[ps[k]] : VIS_NONE;
        #pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            if (ps[k] == 0xffffffffu) break;
            const uint32_t yq = ps[k] / (uint32_t)f.W;
            shade_pixel<C2, FAST>(f, draws, ndraws, sm_draws, view, ps[k], (int)(ps[k] - yq * (uint32_t)f.W), (int)yq, ids[k]);
            vis[ps[k]] = VIS_SHADED;
        }
    }
}

// Fused sort-last composite + shade over NVLink peer memory (config 4): one thread per owned pixel
// loads the candidate (key, id) of every rank (peer pointers opened through CUDA IPC; loads of peer
// addresses travel over NVLink and bypass the local L2), keeps the exact (depth, id) minimum, makes
// it the local state and shades it - no intermediate all-reduced plane is ever written.
struct PeerPlanes {
    const unsigned long long* key[MAX_PEERS];
    const uint32_t* vis[MAX_PEERS];
    // one byte per COMPOSITE_CHUNK pixels of the rank's key plane: non-zero when the rank drew anything there
    // (k_chunk_touched, published with the rank's "drawn" counter); nullptr: unknown, read the rank's keys
    const uint8_t* touched[MAX_PEERS];
    int n;
};
constexpr int COMPOSITE_CHUNK = TPB;     // pixels per flag == pixels per CTA of k_composite_shade_p2p

// which chunks of the local key plane hold a fragment: one warp per chunk
__global__ void __launch_bounds__(TPB) k_chunk_touched(const unsigned long long* __restrict__ zkey, unsigned long long npix,
                                                       uint8_t* __restrict__ touched) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long chunk = (unsigned long long)blockIdx.x * (TPB / 32) + (threadIdx.x >> 5);
    const unsigned long long p0 = chunk * COMPOSITE_CHUNK;
    if (p0 >= npix) return;
    bool any = false;
    #pragma unroll
    for (int j = 0; j < COMPOSITE_CHUNK / 32; ++j) {
        const unsigned long long p = p0 + (unsigned)j * 32u + lane;
        any |= p < npix && zkey[p] != KEY_PLUS_INF;
    }
    const unsigned m = __ballot_sync(0xffffffffu, any);
    if (lane == 0) touched[chunk] = m ? 1 : 0;
}
template <bool C2, bool FAST>
__global__ void __launch_bounds__(TPB, TRB_SHADE_MIN_BLOCKS) k_composite_shade_p2p(FrameDev f, PeerPlanes peers,
                                                                                  const DrawDev* __restrict__ draws,
                                                                                  int ndraws, int row0, int row1) {
    __shared__ DrawDev sm_draws[SHADE_MAX_SM_DRAWS];
    __shared__ unsigned sm_ranks;
    stage_draw_table(sm_draws, draws, ndraws);
    const unsigned long long first = (unsigned long long)row0 * f.W, last = (unsigned long long)row1 * f.W;
    // a CTA takes one chunk of the peers' "touched" maps (chunks are aligned to the plane, not to the rows this rank owns)
    const unsigned long long chunk = first / COMPOSITE_CHUNK + blockIdx.x;
    if (threadIdx.x < 32) {
        bool t = false;
        if ((int)threadIdx.x < peers.n) {
            const uint8_t* q = peers.touched[threadIdx.x];
            t = q ? *reinterpret_cast<const volatile uint8_t*>(q + chunk) != 0 : true;
        }
        const unsigned m = __ballot_sync(0xffffffffu, t);
        if (threadIdx.x == 0) sm_ranks = m;
    }
    __syncthreads();
    if (sm_ranks == 0u) return;              // nobody drew here: the local planes already say so (cleared)
    const unsigned long long p = chunk * COMPOSITE_CHUNK + threadIdx.x;
    if (p < first || p >= last) return;
    // depth keys of the ranks that drew into this chunk (8 B each over NVLink), ids only from the ranks that hold the
    // minimum - usually one; a pixel nobody drew (key of +inf everywhere: fragments have finite depths) needs no id at all
    unsigned long long bk = KEY_PLUS_INF;
    unsigned holders = 0u;
    for (unsigned m = sm_ranks; m; m &= m - 1u) {
        const int r = __ffs(m) - 1;
        const unsigned long long kk = peers.key[r][p];
        if (kk < bk) { bk = kk; holders = 1u << r; }
        else if (kk == bk) holders |= 1u << r;
    }
    uint32_t bid = VIS_NONE;
    if (bk != KEY_PLUS_INF)
        for (unsigned m = holders; m; m &= m - 1u) bid = min(bid, peers.vis[__ffs(m) - 1][p]);
    f.zkey[p] = bk;
    if (bid == VIS_NONE || bid == VIS_SHADED) { f.vis[p] = bid; return; }
    const uint32_t yq = (uint32_t)p / (uint32_t)f.W;
    shade_pixel<C2, FAST, true>(f, draws, ndraws, sm_draws, 0, p, (int)((uint32_t)p - yq * (uint32_t)f.W), (int)yq, bid);
    f.vis[p] = VIS_SHADED;
}

// ---------------------------------------------------------------------------------------------
// Rank-to-rank synchronisation of the sort-last composite WITHOUT the host: every rank owns two counters in its
// own HBM (exported to its peers like the planes).  `drawn` = the last frame whose draws are complete, `done` = the
// last frame whose composite has finished reading the peers' planes.  A rank publishes a counter with a one-thread
// kernel behind the work it stands for (stream order) and waits for its peers' counters with another one-thread
// kernel in front of the work that needs them, so the streams of the ranks run in lock-step per frame while no host
// thread ever blocks.  The waiter occupies one thread of one SM and gives up after `timeout_ns` (a peer died): the
// verdict goes to mapped host memory and the next call on the context reports it.
// ---------------------------------------------------------------------------------------------
struct CommFlags {
    unsigned long long drawn, done, pad0_, pad1_;
};
struct CommWait {
    const unsigned long long* flag[MAX_PEERS_COMM];
    int n;
};
__global__ void k_comm_publish(unsigned long long* flag, unsigned long long value) {
    __threadfence_system();                        // the work queued before this kernel is visible to the peers first
    *reinterpret_cast<volatile unsigned long long*>(flag) = value;
    __threadfence_system();
}
__global__ void k_comm_wait(CommWait w, unsigned long long target, unsigned long long timeout_ns, uint32_t* __restrict__ timed_out) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int r = 0; r < w.n; ++r) {
        const volatile unsigned long long* f = w.flag[r];
        while (*f < target) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > timeout_ns) {
                *timed_out = 1u;
                __threadfence_system();
                return;
            }
            __nanosleep(100);
        }
    }
    __threadfence_system();
}

// ---------------------------------------------------------------------------------------------
// readback helpers and post passes
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_unmap_depth(const unsigned long long* __restrict__ zkey,
                                                     unsigned long long n, double* __restrict__ out) {
    unsigned long long i = (unsigned long long)blockIdx.x * TPB + threadIdx.x;
    if (i < n) out[i] = depth_from_key(zkey[i]);
}
__global__ void __launch_bounds__(TPB) k_count_finite(const unsigned long long* __restrict__ zkey,
                                                      unsigned long long n, unsigned long long* __restrict__ out) {
    __shared__ unsigned long long red[TPB / 32];
    unsigned long long s = 0, stride = (unsigned long long)gridDim.x * TPB;
    for (unsigned long long i = (unsigned long long)blockIdx.x * TPB + threadIdx.x; i < n; i += stride)
        s += finite_d(depth_from_key(zkey[i])) ? 1ull : 0ull;
    s = block_reduce_sum(s, red);
    if (threadIdx.x == 0 && s) atomicAdd(out, s);
}

// compute_ssao_at (main.cpp:324-362) per pixel; dir[] = cos/sin of main.cpp:333-334 evaluated by the
// host's libm so that the sample offsets are the reference's
struct SsaoDirs { double dx[8], dy[8]; };
__device__ __forceinline__ uint8_t ssao_at(const unsigned long long* zkey, int W, int H, int px, int py,
                                           const SsaoDirs& d) {
    double centre = depth_from_key(zkey[(size_t)px + (size_t)py * W]);
    double ao = 1.0;
    if (finite_d(centre)) {
        int occluded = 0, total = 0;
        for (int k = 0; k < 8; ++k)
            for (int s = 1; s <= 8; ++s) {
                double radius = (double)s / 8 * 16.0;                 // main.cpp:337
                int sx = (int)round(px + d.dx[k] * radius);
                int sy = (int)round(py + d.dy[k] * radius);
                if (sx < 0 || sx >= W || sy < 0 || sy >= H) continue;
                double sd = depth_from_key(zkey[(size_t)sx + (size_t)sy * W]);
                if (!finite_d(sd)) { total++; continue; }
                if (sd < centre - 1e-3) occluded++;
                total++;
            }
        if (total != 0) ao = 1.0 - ((double)occluded / (double)total) * 0.35;
    }
    return (uint8_t)(int)(255.0 * ao);                                  // main.cpp:760
}
__global__ void __launch_bounds__(TPB) k_ssao(const unsigned long long* __restrict__ zkey, int W, int H, SsaoDirs d,
                                              uint8_t* __restrict__ ao) {
    int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (x < W && y < H) ao[(size_t)x + (size_t)y * W] = ssao_at(zkey, W, H, x, y, d);
}
// final = phong * ao, main.cpp:768-783
__global__ void __launch_bounds__(TPB) k_composite_ao(const unsigned long long* __restrict__ zkey,
                                                      const uint8_t* __restrict__ color, int W, int H, SsaoDirs d,
                                                      uint8_t* __restrict__ out) {
    int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (x >= W || y >= H) return;
    size_t i = (size_t)x + (size_t)y * W;
    double fct = (double)ssao_at(zkey, W, H, x, y, d) / 255.0;
    for (int ch = 0; ch < 3; ++ch) out[3 * i + ch] = (uint8_t)(int)std_min(255.0, (double)color[3 * i + ch] * fct);
}
// save_zbuffer_image (main.cpp:269-314): pass 1 min/max over finite depths, pass 2 grey map
__global__ void __launch_bounds__(TPB) k_depth_range(const unsigned long long* __restrict__ zkey,
                                                     unsigned long long n, unsigned long long* __restrict__ lohi) {
    __shared__ unsigned long long red[TPB / 32];
    unsigned long long lo = ~0ull, hi = 0ull, stride = (unsigned long long)gridDim.x * TPB;
    for (unsigned long long i = (unsigned long long)blockIdx.x * TPB + threadIdx.x; i < n; i += stride) {
        unsigned long long k = zkey[i];
        if (finite_d(depth_from_key(k))) { lo = min(lo, k); hi = max(hi, k); }
    }
    lo = block_reduce_min64(lo, red);
    hi = ~block_reduce_min64(~hi, red);
    if (threadIdx.x == 0) { atomicMin(lohi, lo); atomicMax(lohi + 1, hi); }
}
__global__ void __launch_bounds__(TPB) k_depth_image(const unsigned long long* __restrict__ zkey,
                                                     unsigned long long n, const unsigned long long* __restrict__ lohi,
                                                     uint8_t* __restrict__ grey) {
    unsigned long long i = (unsigned long long)blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    // main.cpp:275: the search starts from the FINITE values 1e9 / -1e9
    double lo = std_min(1e9, lohi[0] == ~0ull ? 1e9 : depth_from_key(lohi[0]));
    double hi = std_max(-1e9, lohi[1] == 0ull ? -1e9 : depth_from_key(lohi[1]));
    if (hi - lo < 1e-7) hi = lo + 1e-7;                                 // main.cpp:294
    double z = depth_from_key(zkey[i]);
    uint8_t v = 255;
    if (finite_d(z)) {
        double t = (z - lo) / (hi - lo);
        v = (uint8_t)(int)(255.0 * (1.0 - t));                          // main.cpp:306
    }
    grey[i] = v;
}

// TGAColor(v, v, v) per pixel: the grey maps are written as 24-bit images (main.cpp:309, 761)
__global__ void __launch_bounds__(TPB) k_grey_to_bgr(const uint8_t* __restrict__ grey, unsigned long long n,
                                                     uint8_t* __restrict__ bgr) {
    unsigned long long i = (unsigned long long)blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    const uint8_t v = grey[i];
    bgr[3 * i] = v; bgr[3 * i + 1] = v; bgr[3 * i + 2] = v;
}

// ---------------------------------------------------------------------------------------------
// sort-last composite helpers (config 4): after the host all-reduced (min) the depth keys,
// drop local winners that lost; the id plane is then min-reduced too (lowest id == first submitted)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_composite_mask(const unsigned long long* __restrict__ local_key,
                                                        const unsigned long long* __restrict__ global_key,
                                                        uint32_t* __restrict__ vis, unsigned long long n) {
    unsigned long long i = (unsigned long long)blockIdx.x * TPB + threadIdx.x;
    if (i < n && (local_key[i] != global_key[i] || vis[i] == VIS_NONE)) vis[i] = 0x7FFFFFFFu;  // max int32
}
__global__ void __launch_bounds__(TPB) k_composite_finish(uint32_t* __restrict__ vis, unsigned long long n) {
    unsigned long long i = (unsigned long long)blockIdx.x * TPB + threadIdx.x;
    if (i < n && vis[i] == 0x7FFFFFFFu) vis[i] = VIS_NONE;
}
__global__ void __launch_bounds__(TPB) k_key_to_sortable_i64(unsigned long long* __restrict__ key, unsigned long long n) {
    // uint64 order -> int64 order (for collectives that only know signed types): flip the top bit
    unsigned long long i = (unsigned long long)blockIdx.x * TPB + threadIdx.x;
    if (i < n) key[i] ^= KEY_SIGN;
}

}  // namespace trbk
