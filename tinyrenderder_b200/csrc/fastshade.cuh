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

struct LitF {               // fp32 copies of what the lighting reads (built on the host per draw and view)
    float mv[12];           // ModelView rows 0..2
    F3 key, fill, rim;
    float normal_map_strength;
};

// One lit pixel.  attr = the three vertices' raw attributes (pos, nrm, uv as uploaded); pc = the
// perspective-correct barycentrics (fp64, exact); base / nrm_texel / spec_c0 = the texels already
// fetched at the exact fp64 texture coordinates (nrm_texel < 0: no normal map).  sf = shadow factor.
// The vertex stage is linear, so the attributes are interpolated first and transformed once:
// ModelView * (sum b_k p_k, 1) == sum b_k (ModelView * (p_k, 1)) because the weights sum to one.
__device__ __forceinline__ void shade_lit_f32(bool eye, const LitF& L, const float (*attr)[8], const double pc[3],
                                              const int base[4], bool has_nm, const int nmc[4], float spec_f, float sf,
                                              uint8_t out[3]) {
    const float b0 = (float)pc[0], b1 = (float)pc[1], b2 = (float)pc[2];
    F3 p, g;
    p.x = __fmaf_rn(attr[2][0], b2, __fmaf_rn(attr[1][0], b1, attr[0][0] * b0));
    p.y = __fmaf_rn(attr[2][1], b2, __fmaf_rn(attr[1][1], b1, attr[0][1] * b0));
    p.z = __fmaf_rn(attr[2][2], b2, __fmaf_rn(attr[1][2], b1, attr[0][2] * b0));
    g.x = __fmaf_rn(attr[2][3], b2, __fmaf_rn(attr[1][3], b1, attr[0][3] * b0));
    g.y = __fmaf_rn(attr[2][4], b2, __fmaf_rn(attr[1][4], b1, attr[0][4] * b0));
    g.z = __fmaf_rn(attr[2][5], b2, __fmaf_rn(attr[1][5], b1, attr[0][5] * b0));
    const F3 pos = mul_m34f(L.mv, p, b0 + b1 + b2);            // position_eye
    const F3 gn = mul_m34f(L.mv, g, 0.0f);                     // normal_eye (main.cpp:84: ModelView, w = 0)
    const F3 V = normalize3f(scale3f(pos, -1.0f));
    F3 N;
    float diff, spec_pow, spec_gain;
    if (!eye) {
        spec_pow = fmaxf(1.0f, spec_f);                                         // main.cpp:107
        const bool eye_px = (base[0] + base[1] + base[2] >= 651) && spec_pow <= 5.0f;  // sum/765.0 >= 0.85, main.cpp:110-111
        if (eye_px) {
            N = gn;                                                             // main.cpp:123 (not normalised)
        } else {
            F3 nm = F3{0.0f, 0.0f, 1.0f};                                       // model.cpp:429-431
            if (has_nm) {
                const float k = 2.0f / 255.0f;
                nm = normalize3f(F3{__fmaf_rn((float)nmc[2], k, -1.0f), __fmaf_rn((float)nmc[1], k, -1.0f),
                                    __fmaf_rn((float)nmc[0], k, -1.0f)});       // model.cpp:440-442
            }
            const F3 nme = mul_m34f(L.mv, nm, 0.0f);                            // main.cpp:116-119
            const float s = L.normal_map_strength, s1 = 1.0f - s;
            N = normalize3f(F3{__fmaf_rn(nme.x, s, gn.x * s1), __fmaf_rn(nme.y, s, gn.y * s1), __fmaf_rn(nme.z, s, gn.z * s1)});
        }
        diff = fmaxf(0.0f, dot3f(N, L.key)) + fmaxf(0.0f, dot3f(N, L.fill)) * 0.35f + fmaxf(0.0f, dot3f(N, L.rim)) * 0.6f;
        spec_gain = 0.35f;
    } else {
        N = normalize3f(gn);                                                    // main.cpp:225-227
        diff = fmaxf(0.0f, dot3f(N, L.key)) + fmaxf(0.0f, dot3f(N, L.rim)) * 0.6f;
        spec_pow = fmaxf(1.0f, spec_f) * 8.0f;                                  // main.cpp:246
        spec_gain = 1.5f;
    }
    const float nl2 = 2.0f * dot3f(N, L.key);
    const F3 R = normalize3f(F3{__fmaf_rn(N.x, nl2, -L.key.x), __fmaf_rn(N.y, nl2, -L.key.y), __fmaf_rn(N.z, nl2, -L.key.z)});
    const float rv = fmaxf(0.0f, dot3f(R, V));
    float spec = 0.0f;
    if (rv > 0.0f) {
        if (spec_pow == 1.0f) spec = rv;
        else if (spec_pow == 8.0f) { const float r2 = rv * rv, r4 = r2 * r2; spec = r4 * r4; }
        else spec = powf(rv, spec_pow);
    }
    const float gain = 0.1f + diff * sf, add = 255.0f * (spec_gain * spec * sf);
    #pragma unroll
    for (int ch = 0; ch < 3; ++ch)
        out[ch] = (uint8_t)(int)fminf(255.0f, __fmaf_rn((float)base[ch], gain, add));  // main.cpp:164-165 / 255-256
}

}  // namespace trbf
