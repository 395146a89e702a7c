// tests/harness/exact_host.cpp - TEST ONLY.  Compiles tinyrenderder_b200/csrc/exact.cuh for the
// host (g++ -ffp-contract=off) and runs its functions in a trivial serial loop, so that the
// numerical specification the kernels use can be checked against the oracle without a GPU.
// The scheduling (binning, shared-memory resolve) is NOT exercised here - only the arithmetic
// and the order-independent (depth key, id) resolve rule.
#include "../../tinyrenderder_b200/csrc/exact.cuh"
#include <vector>
using namespace trbx;

extern "C" __attribute__((visibility("default")))
int hx_render(int W, int H, const double* viewport,
              int from_clip, const double* clip4_or_null, const float* attr8_or_null, unsigned nverts,
              const unsigned* idx_or_null, unsigned ntris, const double* MV, const double* PR,
              int kind, const double* lights9, double nms,
              const unsigned char* tex_d, int dw, int dh, int dbpp,
              const unsigned char* tex_n, int nw, int nh, int nbpp,
              const unsigned char* tex_s, int sw, int sh, int sbpp,
              unsigned id_base, unsigned long long* zkey /*in/out W*H*/, unsigned* vis /*in/out*/,
              unsigned char* bgr /*in/out*/, double* zout, unsigned long long* covered_out) {
    std::vector<VRec> vr(nverts);
    for (unsigned v = 0; v < nverts; ++v) {
        if (from_clip) vr[v] = vrec_from_clip(viewport, clip4_or_null[4*v], clip4_or_null[4*v+1], clip4_or_null[4*v+2], clip4_or_null[4*v+3]);
        else vr[v] = vrec_from_position(MV, PR, viewport, attr8_or_null[8*v], attr8_or_null[8*v+1], attr8_or_null[8*v+2]);
    }
    unsigned long long covered = 0;
    auto vi = [&](unsigned t, int k) { return idx_or_null ? idx_or_null[3*t+k] : 3*t+k; };
    for (unsigned t = 0; t < ntris; ++t) {
        TriSetup ts;
        if (setup_triangle(vr[vi(t,0)], vr[vi(t,1)], vr[vi(t,2)], W, H, ts) != SETUP_DRAW) continue;
        unsigned id = id_base + t + 1;
        for (int y = ts.y0; y <= ts.y1; ++y) for (int x = ts.x0; x <= ts.x1; ++x) {
            double b[3], z;
            if (!eval_sample(ts, x, y, b, z)) continue;
            ++covered;
            unsigned long long k = fragment_key(z);
            size_t p = (size_t)x + (size_t)y * W;
            if (k < zkey[p] || (k == zkey[p] && id < vis[p])) { if (k < zkey[p]) vis[p] = 0xFFFFFFFFu; zkey[p] = k; if (id < vis[p]) vis[p] = id; }
        }
    }
    // shade (flush)
    LitUniforms U{};
    if (lights9) { U.key = D3{lights9[0],lights9[1],lights9[2]}; U.fill = D3{lights9[3],lights9[4],lights9[5]}; U.rim = D3{lights9[6],lights9[7],lights9[8]}; }
    U.normal_map_strength = nms;
    U.diffuse = TexView{tex_d, dw, dh, dbpp}; U.normal = TexView{tex_n, nw, nh, nbpp}; U.specular = TexView{tex_s, sw, sh, sbpp};
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        size_t p = (size_t)x + (size_t)y * W;
        unsigned id = vis[p];
        if (id != 0xFFFFFFFFu && id != 0 && id > id_base && id <= id_base + ntris) {
            unsigned t = id - id_base - 1;
            const VRec &a = vr[vi(t,0)], &b_ = vr[vi(t,1)], &c = vr[vi(t,2)];
            TriSetup ts; setup_triangle(a, b_, c, W, H, ts);
            double b[3], z, pc[3];
            if (!eval_sample(ts, x, y, b, z)) return -2;
            zkey[p] = depth_key(z);
            perspective_bary(b, a.iw, b_.iw, c.iw, pc);
            unsigned char col[3];
            if (kind == 0) shade_flat_bary(pc, col);
            else {
                Varyings vy;
                for (int k = 0; k < 3; ++k) varyings_from_attr(MV, attr8_or_null + 8 * (size_t)vi(t,k), k, vy);
                shade_lit(kind == 2, MV, U, vy, pc, col);
            }
            bgr[3*p] = col[0]; bgr[3*p+1] = col[1]; bgr[3*p+2] = col[2];
            vis[p] = 0;
        }
        zout[p] = depth_from_key(zkey[p]);
    }
    if (covered_out) *covered_out = covered;
    return 0;
}
