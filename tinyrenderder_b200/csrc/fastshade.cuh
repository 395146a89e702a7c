// fastshade.cuh - fp32 lighting for the deferred shade pass (the default; TRB_SHADE_EXACT=1 or
// trb_set_shade_exact() selects the all-fp64 restatement in exact.cuh instead).
//
// north_star asks for bit-exact coverage and depth and for shaded RGB "within 1 LSB per channel on
// at least 99.9 % of pixels".  Everything that DECIDES something stays in fp64 in the reference's
// operation order: which triangle wins the pixel, the barycentrics, the perspective-correct
// weights, the texture coordinates and therefore the texel that is fetched (model.cpp:420-423
// truncates, a one-ulp change of u could pick the neighbouring texel), the eye-pixel test of
// main.cpp:110-111 (on integers) and the shadow-map comparison.  What is left of
// PhongShader::fragment / EyeShader::fragment (main.cpp:92-170, 220-261) is a continuous function of
// its inputs up to the final (unsigned char) truncation, so evaluating it in fp32 (relative error
// ~1e-6, i.e. ~1e-3 LSB) moves a channel by at most one code, and only when the fp64 value sits
// within ~1e-3 of an integer.  B200 issues fp32 at twice the fp64 rate and sqrt / reciprocal are
// single MUFU instructions instead of ~20-instruction fp64 sequences.
#pragma once
#include "exact.cuh"

namespace trbf {
using namespace trbx;

struct F3 {
    float x, y, z;
};
__device__ __forceinline__ float dot3f(F3 a, F3 b) { return __fmaf_rn(a.z, b.z, __fmaf_rn(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ F3 scale3f(F3 v, float s) { return F3{v.x * s, v.y * s, v.z * s}; }
// normalized(), geometry.h:136-140 (a zero vector is returned unchanged)
__device__ __forceinline__ float rsqrt_fast(float x) {   // MUFU.RSQ, 2 ulp
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ F3 normalize3f(F3 v) {
    const float l2 = dot3f(v, v);
    return l2 >= 1.17549435e-38f ? scale3f(v, rsqrt_fast(l2)) : v;   // below FLT_MIN the ftz reciprocal root is inf
}
// rows 0..2 of a row-major 4x4 (12 floats) times (v, w)
__device__ __forceinline__ F3 mul_m34f(const float* M, F3 v, float w) {
    return F3{__fmaf_rn(M[3], w, __fmaf_rn(M[2], v.z, __fmaf_rn(M[1], v.y, M[0] * v.x))),
              __fmaf_rn(M[7], w, __fmaf_rn(M[6], v.z, __fmaf_rn(M[5], v.y, M[4] * v.x))),
              __fmaf_rn(M[11], w, __fmaf_rn(M[10], v.z, __fmaf_rn(M[9], v.y, M[8] * v.x)))};
}

struct alignas(16) LitF {   // what the fp32 lit pixel reads, per draw and view (built on the host): eight 16-byte loads
    float mv[12];           // ModelView rows 0..2
    F3 key, fill, rim;
    float normal_map_strength;
    uint32_t dbpp, nbpp;    // bytes per texel of the diffuse / normal map (0: map absent)
    uint32_t dw, dh, nw, nh;
    const uint8_t* diffuse; // TGAImage order, BGR(A) or grey
    const uint8_t* normal;
};
static_assert(sizeof(LitF) == 128, "LitF is loaded as eight float4");
// the block is read in 16-byte pieces where they are needed (all eight at once would cost 32 registers)
__device__ __forceinline__ float4 litf_q(const LitF* p, int i) { return __ldg(reinterpret_cast<const float4*>(p) + i); }
struct LitTex {             // pieces 5 (second half) .. 7: what the texel fetch needs
    uint32_t dbpp, nbpp, dw, dh, nw, nh;
    const uint8_t *diffuse, *normal;
};
__device__ __forceinline__ LitTex load_lit_tex(const LitF* p) {
    const float4 a = litf_q(p, 5), b = litf_q(p, 6), c = litf_q(p, 7);
    LitTex t;
    t.dbpp = __float_as_uint(a.z); t.nbpp = __float_as_uint(a.w);
    t.dw = __float_as_uint(b.x); t.dh = __float_as_uint(b.y); t.nw = __float_as_uint(b.z); t.nh = __float_as_uint(b.w);
    t.diffuse = reinterpret_cast<const uint8_t*>(((unsigned long long)__float_as_uint(c.y) << 32) | __float_as_uint(c.x));
    t.normal = reinterpret_cast<const uint8_t*>(((unsigned long long)__float_as_uint(c.w) << 32) | __float_as_uint(c.z));
    return t;
}

// One lit pixel.  attr = the three vertices' raw attributes (pos, nrm, uv as uploaded); pc = the
// perspective-correct barycentrics (fp64, exact); base / nmc = the texels already fetched at the exact fp64
// texture coordinates.  sf = shadow factor.
// The vertex stage is linear, so the attributes are interpolated first and transformed once:
// ModelView * (sum b_k p_k, 1) == sum b_k (ModelView * (p_k, 1)) because the weights sum to one; for the same
// reason the blend of main.cpp:121-125, g_eye * (1 - s) + (ModelView * nm) * s, is ModelView * (g * (1 - s) + nm * s).
// The specular MAP does not enter: Model::specular returns c[0] / 255.0f <= 1 (model.cpp:448-458), so
// std::max(1.0, specular) at main.cpp:107 / 246 is exactly 1.0 whatever the texel - the exponent is 1 (Phong) or
// 8 (eyes), and the eye-pixel test's `specular_power <= 5` (main.cpp:111) always holds.
__device__ __forceinline__ void shade_lit_f32(bool eye, const LitF* Lp, const float (*attr)[8], const double pc[3],
                                              const int base[3], bool has_nm, const int nmc[3], float sf,
                                              uint8_t out[3]) {
    struct { float mv[12]; F3 key, fill, rim; float normal_map_strength; } L;
    {
        const float4 q0 = litf_q(Lp, 0), q1 = litf_q(Lp, 1), q2 = litf_q(Lp, 2), q3 = litf_q(Lp, 3), q4 = litf_q(Lp, 4);
        const float4 q5 = litf_q(Lp, 5);
        L.mv[0] = q0.x; L.mv[1] = q0.y; L.mv[2] = q0.z; L.mv[3] = q0.w; L.mv[4] = q1.x; L.mv[5] = q1.y; L.mv[6] = q1.z; L.mv[7] = q1.w;
        L.mv[8] = q2.x; L.mv[9] = q2.y; L.mv[10] = q2.z; L.mv[11] = q2.w;
        L.key = F3{q3.x, q3.y, q3.z}; L.fill = F3{q3.w, q4.x, q4.y}; L.rim = F3{q4.z, q4.w, q5.x};
        L.normal_map_strength = q5.y;
    }
    const float b0 = (float)pc[0], b1 = (float)pc[1], b2 = (float)pc[2];
    F3 p, g;
    p.x = __fmaf_rn(attr[2][0], b2, __fmaf_rn(attr[1][0], b1, attr[0][0] * b0));
    p.y = __fmaf_rn(attr[2][1], b2, __fmaf_rn(attr[1][1], b1, attr[0][1] * b0));
    p.z = __fmaf_rn(attr[2][2], b2, __fmaf_rn(attr[1][2], b1, attr[0][2] * b0));
    g.x = __fmaf_rn(attr[2][3], b2, __fmaf_rn(attr[1][3], b1, attr[0][3] * b0));
    g.y = __fmaf_rn(attr[2][4], b2, __fmaf_rn(attr[1][4], b1, attr[0][4] * b0));
    g.z = __fmaf_rn(attr[2][5], b2, __fmaf_rn(attr[1][5], b1, attr[0][5] * b0));
    const F3 pos = mul_m34f(L.mv, p, b0 + b1 + b2);            // position_eye
    const F3 V = normalize3f(scale3f(pos, -1.0f));
    F3 N;
    float diff, spec_gain;
    if (!eye) {
        const bool eye_px = base[0] + base[1] + base[2] >= 651;                 // sum/765.0 >= 0.85, main.cpp:110-111
        F3 nm = F3{0.0f, 0.0f, 1.0f};                                           // model.cpp:429-431
        if (has_nm) {
            const float k = 2.0f / 255.0f;
            nm = normalize3f(F3{__fmaf_rn((float)nmc[2], k, -1.0f), __fmaf_rn((float)nmc[1], k, -1.0f),
                                __fmaf_rn((float)nmc[0], k, -1.0f)});           // model.cpp:440-442
        }
        const float s = eye_px ? 0.0f : L.normal_map_strength, s1 = 1.0f - s;   // main.cpp:116-125
        const F3 ne = mul_m34f(L.mv, F3{__fmaf_rn(nm.x, s, g.x * s1), __fmaf_rn(nm.y, s, g.y * s1), __fmaf_rn(nm.z, s, g.z * s1)}, 0.0f);
        const float l2 = dot3f(ne, ne);
        const float inv = (!eye_px && l2 >= 1.17549435e-38f) ? rsqrt_fast(l2) : 1.0f;   // main.cpp:123: eye pixels keep the raw normal
        N = scale3f(ne, inv);
        diff = fmaxf(0.0f, dot3f(N, L.key)) + fmaxf(0.0f, dot3f(N, L.fill)) * 0.35f + fmaxf(0.0f, dot3f(N, L.rim)) * 0.6f;
        spec_gain = 0.35f;
    } else {
        N = normalize3f(mul_m34f(L.mv, g, 0.0f));                               // main.cpp:225-227
        diff = fmaxf(0.0f, dot3f(N, L.key)) + fmaxf(0.0f, dot3f(N, L.rim)) * 0.6f;
        spec_gain = 1.5f;
    }
    const float nl2 = 2.0f * dot3f(N, L.key);
    const F3 R = normalize3f(F3{__fmaf_rn(N.x, nl2, -L.key.x), __fmaf_rn(N.y, nl2, -L.key.y), __fmaf_rn(N.z, nl2, -L.key.z)});
    const float rv = fmaxf(0.0f, dot3f(R, V));
    float spec = rv;                                                            // pow(rv, 1.0), main.cpp:150-152
    if (eye) { const float r2 = rv * rv, r4 = r2 * r2; spec = r4 * r4; }       // pow(rv, 8.0), main.cpp:246-250
    const float gain = 0.1f + diff * sf, add = 255.0f * (spec_gain * spec * sf);
    #pragma unroll
    for (int ch = 0; ch < 3; ++ch)
        out[ch] = (uint8_t)(int)fminf(255.0f, __fmaf_rn((float)base[ch], gain, add));  // main.cpp:164-165 / 255-256
}

}  // namespace trbf
