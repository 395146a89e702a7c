// mesh_order.cu - coherent processing order of a large indexed mesh, built once at upload.
//
// The set-up kernel gathers three 32-byte vertex records per triangle.  When consecutive triangles of the index
// buffer are far apart in space (a recursively subdivided sphere emits the four children of every face in four
// separate blocks: neighbours in the array are on different faces of the icosahedron) every gather misses L2 and
// pulls a 64-byte DRAM atom: measured on config 4 (21 M triangles), k_setup_count moved 5.8 GB for 2.3 GB of
// algorithmic bytes at 77 % of the HBM peak.  Triangles processed in Morton order of their centroids share their
// vertices with the triangles processed just before them, so the records come from L2.
//
// Only the ORDER IN WHICH THE SET-UP AND BIN-FILL KERNELS VISIT the triangles changes: triangle ids (= submission
// order, the reference's tie rule our_gl.cpp:160-166) stay those of the index buffer, and the (depth, id) resolve is
// order independent, so every output bit is the same as without it.
//
// The sort itself is cub::DeviceRadixSort (CUDA toolkit) - a plain library sort at upload time, outside every timed
// region; the kernels around it are ours.  Separate translation unit so that the rasterization kernels do not
// recompile with it.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cstdint>

namespace trbmo {

constexpr int TPB = 256;

__device__ __forceinline__ unsigned ordered_bits(float v) {   // order-preserving map float -> unsigned (NaN sorts high)
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// bounding box of the finite positions: box[0..2] = min, box[3..5] = max as ordered bits
__global__ void __launch_bounds__(TPB) k_bbox(const float4* __restrict__ pos4, uint32_t nverts, unsigned* __restrict__ box) {
    unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
    for (uint32_t v = blockIdx.x * TPB + threadIdx.x; v < nverts; v += gridDim.x * TPB) {
        const float4 p = __ldg(pos4 + v);
        const float c[3] = {p.x, p.y, p.z};
        for (int k = 0; k < 3; ++k) {
            if (!(fabsf(c[k]) <= 3.0e38f)) continue;      // NaN / inf take no part
            const unsigned o = ordered_bits(c[k]);
            lo[k] = min(lo[k], o);
            hi[k] = max(hi[k], o);
        }
    }
    for (int k = 0; k < 3; ++k) {
        const unsigned l = __reduce_min_sync(0xffffffffu, lo[k]), h = __reduce_max_sync(0xffffffffu, hi[k]);
        if ((threadIdx.x & 31) == 0) {
            if (l != 0xffffffffu) atomicMin(box + k, l);
            if (h != 0u) atomicMax(box + 3 + k, h);
        }
    }
}

__device__ __forceinline__ unsigned spread10(unsigned v) {   // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// 30-bit Morton code of every triangle's centroid inside the mesh's bounding box; vals = identity
__global__ void __launch_bounds__(TPB) k_morton(const float4* __restrict__ pos4, const uint32_t* __restrict__ idx, uint32_t ntris,
                                                const unsigned* __restrict__ box, uint32_t* __restrict__ keys,
                                                uint32_t* __restrict__ vals) {
    const uint32_t t = blockIdx.x * TPB + threadIdx.x;
    if (t >= ntris) return;
    float lo[3], scale[3];
    for (int k = 0; k < 3; ++k) {
        const unsigned l = box[k], h = box[3 + k];
        lo[k] = l <= h ? from_ordered_bits(l) : 0.0f;
        const float ext = l <= h ? from_ordered_bits(h) - lo[k] : 0.0f;
        scale[k] = ext > 0.0f ? 1023.0f / ext : 0.0f;
    }
    // idx == nullptr: a triangle soup, triangle t = vertices 3t .. 3t + 2
    const float4 a = __ldg(pos4 + (idx ? __ldg(idx + 3 * (size_t)t) : 3u * t)), b = __ldg(pos4 + (idx ? __ldg(idx + 3 * (size_t)t + 1) : 3u * t + 1u)),
                 c = __ldg(pos4 + (idx ? __ldg(idx + 3 * (size_t)t + 2) : 3u * t + 2u));
    const float cen[3] = {(a.x + b.x + c.x) * (1.0f / 3.0f), (a.y + b.y + c.y) * (1.0f / 3.0f), (a.z + b.z + c.z) * (1.0f / 3.0f)};
    unsigned q[3];
    for (int k = 0; k < 3; ++k) {
        const float f = (cen[k] - lo[k]) * scale[k];
        q[k] = f >= 0.0f ? (unsigned)fminf(f, 1023.0f) : 0u;    // NaN -> 0
    }
    keys[t] = spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2);
    vals[t] = t;
}

// the index triples in processing order: idx_perm[3j + k] = idx[3 perm[j] + k]
__global__ void __launch_bounds__(TPB) k_gather_idx(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ perm, uint32_t ntris,
                                                    uint32_t* __restrict__ idx_perm) {
    const uint32_t j = blockIdx.x * TPB + threadIdx.x;
    if (j >= ntris) return;
    const uint32_t* q = idx + 3 * (size_t)__ldg(perm + j);
    uint32_t* o = idx_perm + 3 * (size_t)j;
    o[0] = __ldg(q); o[1] = __ldg(q + 1); o[2] = __ldg(q + 2);
}

// A soup shares no vertex, so its VERTEX ARRAYS are put into processing order themselves: slot j's vertices become
// 3j .. 3j + 2 (the set-up kernel then streams its vertex records instead of gathering them), inv[t] = slot of triangle t
// (the shade pass finds the vertices of a winning triangle id through it).
__global__ void __launch_bounds__(TPB) k_permute_soup(const float4* __restrict__ pos_in, const float4* __restrict__ attr_in,
                                                      const uint32_t* __restrict__ perm, uint32_t ntris, float4* __restrict__ pos_out,
                                                      float4* __restrict__ attr_out, uint32_t* __restrict__ inv) {
    const uint32_t v = blockIdx.x * TPB + threadIdx.x;          // a vertex of the ordered arrays
    if (v >= 3u * ntris) return;
    const uint32_t j = v / 3u, k = v - 3u * j, t = __ldg(perm + j);
    const size_t src = 3 * (size_t)t + k;
    pos_out[v] = __ldg(pos_in + src);
    attr_out[2 * (size_t)v] = __ldg(attr_in + 2 * src);
    attr_out[2 * (size_t)v + 1] = __ldg(attr_in + 2 * src + 1);
    if (k == 0) inv[t] = j;
}

// ---- vertex numbering of a large indexed mesh ------------------------------------------------------------------
// A subdivided mesh numbers its vertices by subdivision level: the three vertices of a triangle sit megabytes apart, and
// the vertices one rank's share of the mesh refers to (trb_draw_shard) are sprinkled evenly over the whole array.  The
// vertices are renumbered in Morton order of their positions (arrays permuted, index buffer rewritten): neighbouring
// triangles then gather neighbouring records, and a rank's vertices form runs that a masked vertex stage can skip over
// warp by warp.  Vertex numbers are internal (no entry point exposes them); triangle ids do not change.
__global__ void __launch_bounds__(TPB) k_vmorton(const float4* __restrict__ pos4, uint32_t nverts, const unsigned* __restrict__ box,
                                                 uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t v = blockIdx.x * TPB + threadIdx.x;
    if (v >= nverts) return;
    const float4 a = __ldg(pos4 + v);
    const float c[3] = {a.x, a.y, a.z};
    unsigned q[3];
    for (int k = 0; k < 3; ++k) {
        const unsigned l = box[k], h = box[3 + k];
        const float lo = l <= h ? from_ordered_bits(l) : 0.0f;
        const float ext = l <= h ? from_ordered_bits(h) - lo : 0.0f;
        const float f = (c[k] - lo) * (ext > 0.0f ? 1023.0f / ext : 0.0f);
        q[k] = f >= 0.0f ? (unsigned)fminf(f, 1023.0f) : 0u;    // NaN -> 0
    }
    keys[v] = spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2);
    vals[v] = v;
}
// vperm[new] = old: gather the vertex arrays into the new numbering and note where every old vertex went
__global__ void __launch_bounds__(TPB) k_permute_vertices(const float4* __restrict__ pos_in, const float4* __restrict__ attr_in,
                                                          const uint32_t* __restrict__ vperm, uint32_t nverts,
                                                          float4* __restrict__ pos_out, float4* __restrict__ attr_out,
                                                          uint32_t* __restrict__ inv) {
    const uint32_t n = blockIdx.x * TPB + threadIdx.x;
    if (n >= nverts) return;
    const uint32_t o = __ldg(vperm + n);
    pos_out[n] = __ldg(pos_in + o);
    attr_out[2 * (size_t)n] = __ldg(attr_in + 2 * (size_t)o);
    attr_out[2 * (size_t)n + 1] = __ldg(attr_in + 2 * (size_t)o + 1);
    inv[o] = n;
}
__global__ void __launch_bounds__(TPB) k_remap_indices(uint32_t* __restrict__ idx, uint64_t nidx, const uint32_t* __restrict__ inv) {
    for (uint64_t i = (uint64_t)blockIdx.x * TPB + threadIdx.x; i < nidx; i += (uint64_t)gridDim.x * TPB) idx[i] = __ldg(inv + idx[i]);
}

}  // namespace trbmo

// scratch of trb_vertex_order_apply: [keys_in | keys_out | vals_in | vperm | inv | box (256 B) | cub temp]
size_t trb_vertex_order_scratch_bytes(uint32_t nverts) {
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)nverts, 0, 30, (cudaStream_t)0);
    const size_t n4 = ((size_t)nverts * 4 + 255) & ~(size_t)255;
    return 5 * n4 + 256 + cub_bytes + 256;
}
// vertex arrays into Morton numbering (pos_out / attr_out), the index buffer rewritten in place; queued on `st`
cudaError_t trb_vertex_order_apply(const float4* pos_in, const float* attr_in, uint32_t nverts, uint32_t* idx, uint64_t nidx,
                                   float4* pos_out, float* attr_out, void* scratch, size_t scratch_bytes, int sms, cudaStream_t st) {
    using namespace trbmo;
    if (nverts == 0 || nverts > 0x7fffffffu) return cudaErrorInvalidValue;
    const size_t n4 = ((size_t)nverts * 4 + 255) & ~(size_t)255;
    char* s = (char*)scratch;
    uint32_t* keys_in = (uint32_t*)s;
    uint32_t* keys_out = (uint32_t*)(s + n4);
    uint32_t* vals_in = (uint32_t*)(s + 2 * n4);
    uint32_t* vperm = (uint32_t*)(s + 3 * n4);
    uint32_t* inv = (uint32_t*)(s + 4 * n4);
    unsigned* box = (unsigned*)(s + 5 * n4);
    void* cub_temp = s + 5 * n4 + 256;
    size_t cub_bytes = scratch_bytes - (5 * n4 + 256);
    const unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    cudaError_t e = cudaMemcpyAsync(box, init, sizeof(init), cudaMemcpyHostToDevice, st);   // pageable: staged before the call returns
    if (e != cudaSuccess) return e;
    const unsigned vb = (unsigned)((nverts + TPB - 1) / TPB);
    k_bbox<<<vb < (unsigned)sms * 8 ? vb : (unsigned)sms * 8, TPB, 0, st>>>(pos_in, nverts, box);
    k_vmorton<<<vb, TPB, 0, st>>>(pos_in, nverts, box, keys_in, vals_in);
    e = cub::DeviceRadixSort::SortPairs(cub_temp, cub_bytes, (const uint32_t*)keys_in, keys_out, (const uint32_t*)vals_in, vperm,
                                        (int)nverts, 0, 30, st);
    if (e != cudaSuccess) return e;
    k_permute_vertices<<<vb, TPB, 0, st>>>(pos_in, reinterpret_cast<const float4*>(attr_in), vperm, nverts, pos_out,
                                           reinterpret_cast<float4*>(attr_out), inv);
    k_remap_indices<<<(unsigned)sms * 8, TPB, 0, st>>>(idx, nidx, inv);
    return cudaGetLastError();
}

// vertex arrays of a soup in processing order (pos_out / attr_out hold nverts entries; vertices past 3 * ntris are copied
// as they are) and the inverse permutation; queued on `st`
cudaError_t trb_soup_order_apply(const float4* pos_in, const float* attr_in, uint32_t nverts, const uint32_t* perm, uint32_t ntris,
                                 float4* pos_out, float* attr_out, uint32_t* inv_perm, cudaStream_t st) {
    using namespace trbmo;
    if (ntris == 0 || ntris > 0x55555555u || 3ull * ntris > nverts) return cudaErrorInvalidValue;
    const uint32_t used = 3u * ntris;
    k_permute_soup<<<(used + TPB - 1) / TPB, TPB, 0, st>>>(pos_in, reinterpret_cast<const float4*>(attr_in), perm, ntris, pos_out,
                                                          reinterpret_cast<float4*>(attr_out), inv_perm);
    if (nverts > used) {
        cudaError_t e = cudaMemcpyAsync(pos_out + used, pos_in + used, (size_t)(nverts - used) * 16, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(attr_out + (size_t)used * 8, attr_in + (size_t)used * 8, (size_t)(nverts - used) * 32, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

// scratch the build needs besides its outputs: [keys_in | keys_out | vals_in | box (256 B) | cub temp]
size_t trb_mesh_order_scratch_bytes(uint32_t ntris) {
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)ntris, 0, 30, (cudaStream_t)0);
    const size_t n4 = ((size_t)ntris * 4 + 255) & ~(size_t)255;
    return 3 * n4 + 256 + cub_bytes + 256;
}

// perm_out[ntris], idx_perm_out[3 ntris]; everything is queued on `st`, nothing waits
cudaError_t trb_mesh_order_build(const float4* pos4, uint32_t nverts, const uint32_t* idx, uint32_t ntris, uint32_t* perm_out,
                                 uint32_t* idx_perm_out, void* scratch, size_t scratch_bytes, int sms, cudaStream_t st) {
    using namespace trbmo;
    if (ntris == 0 || ntris > 0x7fffffffu) return cudaErrorInvalidValue;
    const size_t n4 = ((size_t)ntris * 4 + 255) & ~(size_t)255;
    char* s = (char*)scratch;
    uint32_t* keys_in = (uint32_t*)s;
    uint32_t* keys_out = (uint32_t*)(s + n4);
    uint32_t* vals_in = (uint32_t*)(s + 2 * n4);
    unsigned* box = (unsigned*)(s + 3 * n4);
    void* cub_temp = s + 3 * n4 + 256;
    size_t cub_bytes = scratch_bytes - (3 * n4 + 256);
    const unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    cudaError_t e = cudaMemcpyAsync(box, init, sizeof(init), cudaMemcpyHostToDevice, st);   // pageable: staged before the call returns
    if (e != cudaSuccess) return e;
    const unsigned vb = (unsigned)((nverts + TPB - 1) / TPB);
    k_bbox<<<vb < (unsigned)sms * 8 ? vb : (unsigned)sms * 8, TPB, 0, st>>>(pos4, nverts, box);
    const unsigned tb = (unsigned)((ntris + TPB - 1) / TPB);
    k_morton<<<tb, TPB, 0, st>>>(pos4, idx, ntris, box, keys_in, vals_in);
    e = cub::DeviceRadixSort::SortPairs(cub_temp, cub_bytes, (const uint32_t*)keys_in, keys_out, (const uint32_t*)vals_in, perm_out,
                                        (int)ntris, 0, 30, st);
    if (e != cudaSuccess) return e;
    if (idx) k_gather_idx<<<tb, TPB, 0, st>>>(idx, perm_out, ntris, idx_perm_out);
    return cudaGetLastError();
}
