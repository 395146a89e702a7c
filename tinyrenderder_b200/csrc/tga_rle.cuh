// tga_rle.cuh - TGAImage::unload_rle_data (tgaimage.cpp:193-242) on the device, byte for byte.
//
// The reference walks the pixels of an image once, front to back (no restart at scanlines):
//   at `cur`: the run of pixels equal to pixel[cur], capped at 128;  run >= 2 -> RUN packet
//   (header 128 + run - 1, one pixel);  run == 1 -> RAW packet that grows while the next pixel
//   differs from ITS predecessor, capped at 128 (header n - 1, n pixels).
// A raw packet therefore ends up swallowing the FIRST pixel of the next run of equal pixels, which
// makes the packetisation depend on everything before it - a sequential automaton.  It parallelises
// because a maximal run of >= 2 equal pixels ("long run") forgets almost all of that history: the
// automaton reaches a long run either at a packet boundary (class P) or inside an open raw packet
// (class R, the run's first pixel is swallowed), and what it does from there to the next long run -
// run packets over the long run, raw packets of 128 over the single pixels that follow - depends on
// that one bit only.  So:
//   1. per pixel: eq[i] = pixel[i] == pixel[i-1];  a long run starts where !eq[i] && eq[i+1]
//   2. exclusive sum scan of the long-run starts  -> segment index of every pixel
//      (segment = one long run + the single pixels up to the next long run; segment 0 of an image =
//      the single pixels before its first long run)
//   3. per segment: the 2 -> 2 map "entry class -> entry class of the next segment"; an exclusive
//      scan under map composition gives every segment its entry class
//   4. per segment: bytes it emits;  exclusive sum scan -> where it writes
//   5. per pixel: which packet of its segment it belongs to -> header and pixel bytes written in place
// Every pass is a flat grid over all views of the batch (images never share a packet: eq is false
// across an image boundary and the map of a segment 0 ignores its input, which restarts the scan).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cassert>
#if defined(TRB_DEBUG_CHECKS)
#define TRB_RLE_CHECK(cond) assert(cond)
#else
#define TRB_RLE_CHECK(cond) ((void)0)
#endif

namespace trbr {

constexpr int RTPB = 256;
constexpr int RITEMS = 8;
constexpr int RBLOCK = RTPB * RITEMS;

struct AddU32 {
    typedef uint32_t T;
    __device__ static T identity() { return 0u; }
    __device__ static T op(T a, T b) { return a + b; }      // a comes first
};
// maps {P, R} -> {P, R} as two bits: bit 0 = image of P, bit 1 = image of R; op(a, b) = "a, then b"
struct MapCompose {
    typedef uint8_t T;
    __device__ static T identity() { return 2; }
    __device__ static T op(T a, T b) { return (T)(((b >> (a & 1)) & 1) | (((b >> ((a >> 1) & 1)) & 1) << 1)); }
};
struct LoadLongStart {      // scan input of pass 2: bit 1 of the per-pixel flags
    const uint8_t* flags;
    __device__ uint32_t operator()(size_t i) const { return (flags[i] >> 1) & 1u; }
};
template <class T>
struct LoadPlain {
    const T* p;
    __device__ T operator()(size_t i) const { return p[i]; }
};

// ordered block-wide exclusive scan of one value per thread (threads in index order)
template <class Op>
__device__ __forceinline__ typename Op::T block_scan_exclusive(typename Op::T v, typename Op::T* sh, typename Op::T& total) {
    typedef typename Op::T T;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T incl = v;
    for (int o = 1; o < 32; o <<= 1) {
        const T y = (T)__shfl_up_sync(0xffffffffu, (unsigned)incl, o);
        if (lane >= (unsigned)o) incl = Op::op(y, incl);
    }
    T excl = (T)__shfl_up_sync(0xffffffffu, (unsigned)incl, 1);
    if (lane == 0) excl = Op::identity();
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        T w = lane < RTPB / 32 ? sh[lane] : Op::identity();
        for (int o = 1; o < 32; o <<= 1) {
            const T y = (T)__shfl_up_sync(0xffffffffu, (unsigned)w, o);
            if (lane >= (unsigned)o) w = Op::op(y, w);
        }
        if (lane < RTPB / 32) sh[lane] = w;
    }
    __syncthreads();
    const T before = warp ? sh[warp - 1] : Op::identity();
    total = sh[RTPB / 32 - 1];
    __syncthreads();
    return Op::op(before, excl);
}

// three-kernel exclusive scan: per-block aggregates, one block scans the aggregates, per-block apply
template <class Op, class Load>
__global__ void __launch_bounds__(RTPB) k_rscan_partial(Load in, size_t n, typename Op::T* __restrict__ agg) {
    typedef typename Op::T T;
    __shared__ T sh[RTPB / 32];
    const size_t base = (size_t)blockIdx.x * RBLOCK + (size_t)threadIdx.x * RITEMS;
    T acc = Op::identity();
    for (int i = 0; i < RITEMS; ++i)
        if (base + i < n) acc = Op::op(acc, (T)in(base + i));
    T total;
    block_scan_exclusive<Op>(acc, sh, total);
    if (threadIdx.x == 0) agg[blockIdx.x] = total;
}
template <class Op>
__global__ void __launch_bounds__(RTPB) k_rscan_sums(typename Op::T* __restrict__ agg, uint32_t nblocks,
                                                     typename Op::T* __restrict__ total_out) {
    typedef typename Op::T T;
    __shared__ T sh[RTPB / 32];
    T carry = Op::identity();
    for (uint32_t base = 0; base < nblocks; base += RTPB) {
        const uint32_t e = base + threadIdx.x;
        const T v = e < nblocks ? agg[e] : Op::identity();
        T tot;
        const T ex = block_scan_exclusive<Op>(v, sh, tot);
        if (e < nblocks) agg[e] = Op::op(carry, ex);
        carry = Op::op(carry, tot);
    }
    if (threadIdx.x == 0) *total_out = carry;
}
template <class Op, class Load>
__global__ void __launch_bounds__(RTPB) k_rscan_final(Load in, size_t n, const typename Op::T* __restrict__ agg,
                                                      typename Op::T* __restrict__ out) {
    typedef typename Op::T T;
    __shared__ T sh[RTPB / 32];
    const size_t base = (size_t)blockIdx.x * RBLOCK + (size_t)threadIdx.x * RITEMS;
    T v[RITEMS];
    T acc = Op::identity();
    for (int i = 0; i < RITEMS; ++i) {
        v[i] = base + i < n ? (T)in(base + i) : Op::identity();
        acc = Op::op(acc, v[i]);
    }
    T tot;
    T ex = Op::op(agg[blockIdx.x], block_scan_exclusive<Op>(acc, sh, tot));
    for (int i = 0; i < RITEMS; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex = Op::op(ex, v[i]);
    }
}

template <int BPP>
__device__ __forceinline__ bool same_pixel(const uint8_t* px, size_t a, size_t b) {
    bool s = true;
    #pragma unroll
    for (int c = 0; c < BPP; ++c) s &= px[a * BPP + c] == px[b * BPP + c];
    return s;
}

// pass 1: bit 0 = equal to the previous pixel of the same image, bit 1 = first pixel of a long run
template <int BPP>
__global__ void __launch_bounds__(RTPB) k_rle_flags(const uint8_t* __restrict__ px, size_t npix, size_t total,
                                                    uint8_t* __restrict__ flags) {
    const size_t i = (size_t)blockIdx.x * RTPB + threadIdx.x;
    if (i >= total) return;
    const size_t r = i % npix;
    const bool eq = r != 0 && same_pixel<BPP>(px, i, i - 1);
    const bool eq_next = r + 1 != npix && same_pixel<BPP>(px, i + 1, i);
    flags[i] = (uint8_t)((eq ? 1 : 0) | ((!eq && eq_next) ? 2 : 0));
}

struct RleTables {
    uint32_t* seg_start;   // [NS + 1] first pixel (flat index) of the segment's long run; sentinel = total
    uint32_t* long_end;    // [NS + 1] one past the long run (== seg_start for a segment 0 and the sentinel)
    uint8_t* map;          // [NS] entry class -> next entry class
    uint8_t* prefix;       // [NS] composition of the maps before the segment
    uint32_t* bytes;       // [NS] bytes the segment emits
    uint32_t* base;        // [NS + 1] exclusive sum of bytes
};

// pass 2b: segment table.  ls_excl = exclusive count of long-run starts (flat); the segment of pixel i
// of view v is ls_excl[i] + (i starts a long run) + v, because every view adds its segment 0.
__global__ void __launch_bounds__(RTPB) k_rle_segments(const uint8_t* __restrict__ flags, const uint32_t* __restrict__ ls_excl,
                                                       const uint32_t* __restrict__ ls_total, size_t npix, size_t total,
                                                       uint32_t nviews, RleTables t) {
    const size_t i = (size_t)blockIdx.x * RTPB + threadIdx.x;
    if (i >= total) return;
    const uint32_t v = (uint32_t)(i / npix);
    const size_t r = i - (size_t)v * npix;
    const uint8_t f = flags[i];
    const uint32_t ex = ls_excl[i];
    if (f & 2) t.seg_start[ex + 1 + v] = (uint32_t)i;
    if (r == 0) {                                       // segment 0 of the view: no long run
        t.seg_start[ex + v] = (uint32_t)i;
        t.long_end[ex + v] = (uint32_t)i;
    } else if (!(f & 1) && (flags[i - 1] & 1)) {        // first pixel after a long run
        t.long_end[ex + v] = (uint32_t)i;
    }
    if (r + 1 == npix && (f & 1)) t.long_end[ex + v] = (uint32_t)(i + 1);   // long run that reaches the end of the image
    if (i == 0) {
        const uint32_t ns = *ls_total + nviews;
        t.seg_start[ns] = (uint32_t)total;
        t.long_end[ns] = (uint32_t)total;
    }
}

struct SegInfo {            // what a segment emits, given its entry class
    uint32_t p0, r0;        // first pixel of the run packets, first pixel of the raw packets
    uint32_t rle_px, rle_packets, nraw;
    bool exit_r;            // a raw packet is open when the next long run is reached
};
__device__ __forceinline__ SegInfo seg_info(uint32_t s, uint32_t e, uint32_t nxt, bool next_is_long, uint32_t entry_r) {
    SegInfo g;
    const uint32_t L = e - s;
    const uint32_t c = L ? entry_r : 0u;                 // a segment 0 always starts at a packet boundary
    const uint32_t X = L - c, m = X & 127u;
    g.p0 = s + c;
    g.rle_px = X - (m == 1u ? 1u : 0u);                  // a lone last pixel opens a raw packet instead
    g.rle_packets = (X >> 7) + (m >= 2u ? 1u : 0u);
    g.r0 = g.p0 + g.rle_px;
    const uint32_t cnt = nxt - g.r0;                     // (m == 1) + the single pixels
    g.exit_r = (cnt & 127u) != 0u;
    g.nraw = cnt + ((next_is_long && g.exit_r) ? 1u : 0u);   // ... + the swallowed first pixel of the next long run
    return g;
}

// pass 3a: the segment's map
__global__ void __launch_bounds__(RTPB) k_rle_maps(const uint32_t* __restrict__ ls_total, uint32_t nviews, RleTables t) {
    const uint32_t ns = *ls_total + nviews;
    const uint32_t k = blockIdx.x * RTPB + threadIdx.x;
    if (k >= ns) return;
    const uint32_t s = t.seg_start[k], e = t.long_end[k], nxt = t.seg_start[k + 1];
    const bool nl = t.long_end[k + 1] != nxt;
    const uint32_t m0 = seg_info(s, e, nxt, nl, 0).exit_r ? 1u : 0u, m1 = seg_info(s, e, nxt, nl, 1).exit_r ? 1u : 0u;
    t.map[k] = (uint8_t)(m0 | (m1 << 1));
}
// pass 4a: bytes per segment (prefix[k] applied to P = entry class)
template <int BPP>
__global__ void __launch_bounds__(RTPB) k_rle_sizes(const uint32_t* __restrict__ ls_total, uint32_t nviews, RleTables t) {
    const uint32_t ns = *ls_total + nviews;
    const uint32_t k = blockIdx.x * RTPB + threadIdx.x;
    if (k >= ns) return;
    const uint32_t s = t.seg_start[k], e = t.long_end[k], nxt = t.seg_start[k + 1];
    const SegInfo g = seg_info(s, e, nxt, t.long_end[k + 1] != nxt, t.prefix[k] & 1u);
    t.bytes[k] = g.rle_packets * (1 + BPP) + (g.nraw + 127u) / 128u + g.nraw * BPP;
}
// pass 5: every pixel writes itself (and the header of the packet it opens)
template <int BPP>
__global__ void __launch_bounds__(RTPB) k_rle_emit(const uint8_t* __restrict__ px, const uint8_t* __restrict__ flags,
                                                   const uint32_t* __restrict__ ls_excl, size_t npix, size_t total,
                                                   RleTables t, uint8_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * RTPB + threadIdx.x;
    if (i >= total) return;
    const uint32_t v = (uint32_t)(i / npix);
    uint32_t k = ls_excl[i] + ((flags[i] >> 1) & 1u) + v;
    uint32_t s = t.seg_start[k], e = t.long_end[k];
    if ((uint32_t)i == s && e != s && (t.prefix[k] & 1u)) {   // swallowed by the raw packet the previous segment left open
        --k;
        s = t.seg_start[k];
        e = t.long_end[k];
    }
    const uint32_t nxt = t.seg_start[k + 1];
    const SegInfo g = seg_info(s, e, nxt, t.long_end[k + 1] != nxt, t.prefix[k] & 1u);
    uint8_t* o = out + t.base[k];
    const uint32_t pi = (uint32_t)i;
    TRB_RLE_CHECK(pi >= g.p0 && pi < g.r0 + g.nraw && t.base[k + 1] - t.base[k] == t.bytes[k]);
    if (pi < g.r0) {
        const uint32_t d = pi - g.p0;
        if (d & 127u) return;                                  // only the first pixel of a run packet is stored
        const uint32_t q = d >> 7, len = min(128u, g.rle_px - (q << 7));
        o += q * (1 + BPP);
        TRB_RLE_CHECK(q < g.rle_packets && len >= 2u);
        o[0] = (uint8_t)(128u + len - 1u);
        #pragma unroll
        for (int c = 0; c < BPP; ++c) o[1 + c] = px[i * BPP + c];
    } else {
        const uint32_t d = pi - g.r0, q = d >> 7, w = d & 127u;
        o += g.rle_packets * (1 + BPP) + q * (1 + 128 * BPP) + 1 + w * BPP;
        TRB_RLE_CHECK((size_t)(o + BPP - out) <= (size_t)t.base[k + 1]);
        if (w == 0) o[-1] = (uint8_t)(min(128u, g.nraw - (q << 7)) - 1u);
        #pragma unroll
        for (int c = 0; c < BPP; ++c) o[c] = px[i * BPP + c];
    }
}
// per-view byte ranges: offsets[v] = base of the view's segment 0, offsets[nviews] = end
__global__ void k_rle_view_offsets(const uint32_t* __restrict__ ls_excl, const uint32_t* __restrict__ ls_total,
                                   size_t npix, uint32_t nviews, RleTables t, uint32_t* __restrict__ offsets) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nviews) offsets[v] = t.base[ls_excl[(size_t)v * npix] + v];
    if (v == nviews) offsets[v] = t.base[*ls_total + nviews];
}

}  // namespace trbr
