// oracle/port_oracle.cpp - TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Self-contained CPU restatement of the reference's hot path, independent of the
// reference's headers, behind the C ABI of include/trb.h (prefix orc_).  It exists
// because oracle/_ref/libtrb_ref.so (the reference's real rasterize()) cannot report the
// order-independent counters and has process-wide static state; this port is pinned to
// it by tests/test_oracle_*.py (bit-equal z-buffers, equal colours, equal fragments_drawn
// on K1-K7 and on random scenes) and by tests/golden/*.json.
//
// Every function cites the reference lines it follows.  Arithmetic is IEEE double,
// compiled with -ffp-contract=off; the operation ORDER is the specification.
#define TRB_FN(name) orc_##name
#include "../include/trb.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace {

struct D3 { double x, y, z; };
struct M4 { double m[4][4]; };

// dot<n>, geometry.h:122-127: sum starts at +0.0, terms added in index order
inline double dot4(const double* a, const double* b) {
    double s = 0;
    for (int i = 0; i < 4; ++i) s += a[i] * b[i];
    return s;
}
inline double dot3(const D3& a, const D3& b) {
    double s = 0;
    s += a.x * b.x;
    s += a.y * b.y;
    s += a.z * b.z;
    return s;
}
// mat*vec, geometry.h:186-192
inline void mul_mv(const M4& M, const double v[4], double out[4]) {
    for (int i = 0; i < 4; ++i) out[i] = dot4(M.m[i], v);
}
// normalized, geometry.h:136-140: returns v unchanged when the length is exactly 0
inline D3 normalize3(const D3& v) {
    double len = std::sqrt(dot3(v, v));
    if (len == 0) return v;
    return D3{v.x / len, v.y / len, v.z / len};
}
inline D3 scale3(const D3& v, double s) { return D3{v.x * s, v.y * s, v.z * s}; }
inline D3 add3(const D3& a, const D3& b) { return D3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline D3 sub3(const D3& a, const D3& b) { return D3{a.x - b.x, a.y - b.y, a.z - b.z}; }

// (int)double on x86-64 is cvttsd2si: out-of-range and NaN give INT_MIN
inline int x86_int(double v) {
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
    return (int)v;
}
// std::min({a,b,c}) / std::max({a,b,c}) comparison order (min_element / max_element)
inline double min3(double a, double b, double c) {
    double m = a;
    if (b < m) m = b;
    if (c < m) m = c;
    return m;
}
inline double max3(double a, double b, double c) {
    double m = a;
    if (m < b) m = b;
    if (m < c) m = c;
    return m;
}
inline double max_d(double a, double b) { return (a < b) ? b : a; }  // std::max(a,b)
inline double min_d(double a, double b) { return (b < a) ? b : a; }  // std::min(a,b)

struct Tex {
    int w = 0, h = 0, bpp = 0;
    std::vector<uint8_t> px;
    // TGAImage::get (tgaimage.cpp:24-30) + TGAColor(p,bpp) (tgaimage.h:47-51): missing channels 0
    void fetch(const double uv[2], uint8_t out[4]) const {
        // model.cpp:420-423
        int x = std::clamp(x86_int(uv[0] * w), 0, w - 1);
        int y = std::clamp(x86_int(uv[1] * h), 0, h - 1);
        const uint8_t* p = &px[((size_t)x + (size_t)y * w) * bpp];
        for (int i = 0; i < 4; ++i) out[i] = i < bpp ? p[i] : 0;
    }
};

struct Mesh {
    std::vector<float> pos, nrm, uv;
    std::vector<uint32_t> idx;
};

struct Varyings {  // the per-triangle state PhongShader::vertex leaves behind, main.cpp:75-87
    double uv[3][2];
    D3 pos_eye[3];
    D3 nrm_eye[3];
    double light_clip[3][4];   // SHADOW_PHONG only
};

struct ShadowEnv {   // SHADOW_PHONG, authored in ref_harness.cpp (ShadowPhongAuthored)
    M4 lmv{}, lpr{}, lvp{};
    double bias = 0, darkening = 1;
    const std::vector<double>* map = nullptr;
    int w = 0, h = 0;
};

struct ShadeEnv {
    int kind = 0;
    M4 modelview{};
    ShadowEnv shadow;
    TrbPhongUniforms u{};
    const Tex* diffuse = nullptr;
    const Tex* normal = nullptr;
    const Tex* specular = nullptr;
};

// Model::diffuse / normal / specular, model.cpp:415-459
inline void sample_diffuse(const ShadeEnv& e, const double uv[2], uint8_t c[4]) {
    if (!e.diffuse) { c[0] = c[1] = c[2] = c[3] = 255; return; }
    e.diffuse->fetch(uv, c);
}
inline D3 sample_normal(const ShadeEnv& e, const double uv[2]) {
    if (!e.normal) return D3{0, 0, 1};
    uint8_t c[4];
    e.normal->fetch(uv, c);
    D3 n;
    n.x = (double)c[2] / 255.0 * 2.0 - 1.0;
    n.y = (double)c[1] / 255.0 * 2.0 - 1.0;
    n.z = (double)c[0] / 255.0 * 2.0 - 1.0;
    return normalize3(n);
}
inline float sample_specular(const ShadeEnv& e, const double uv[2]) {
    if (!e.specular) return 1.0f;
    uint8_t c[4];
    e.specular->fetch(uv, c);
    return c[0] / 255.0f;
}

inline D3 mix3(const D3 a[3], const double b[3]) {
    // vec*scalar then vec+vec, left to right (main.cpp:94-96)
    return add3(add3(scale3(a[0], b[0]), scale3(a[1], b[1])), scale3(a[2], b[2]));
}
inline D3 light(const double* p) { return D3{p[0], p[1], p[2]}; }

// fragment stage; returns the BGR bytes that framebuffer.set would write
void shade(const ShadeEnv& e, const Varyings& v, const double b[3], uint8_t out[3]) {
    if (e.kind == TRB_SHADER_FLAT_BARY || e.kind == TRB_SHADER_DEPTH) {
        for (int i = 0; i < 3; ++i) {
            double t = 255.0 * b[i];
            t = (t > 0.0) ? t : 0.0;
            t = (t < 255.0) ? t : 255.0;
            out[i] = (unsigned char)t;
        }
        return;
    }
    D3 pos = mix3(v.pos_eye, b);
    D3 gn = mix3(v.nrm_eye, b);
    double uv[2];
    for (int k = 0; k < 2; ++k) uv[k] = v.uv[0][k] * b[0] + v.uv[1][k] * b[1] + v.uv[2][k] * b[2];
    uint8_t base[4];
    sample_diffuse(e, uv, base);
    D3 key = light(e.u.key_dir_eye), fill = light(e.u.fill_dir_eye), rim = light(e.u.rim_dir_eye);

    if (e.kind == TRB_SHADER_GOURAUD) {  // GouraudAuthored in ref_harness.cpp
        double vi[3];
        for (int k = 0; k < 3; ++k) vi[k] = max_d(0.0, dot3(normalize3(v.nrm_eye[k]), key));
        double I = vi[0] * b[0] + vi[1] * b[1] + vi[2] * b[2];
        for (int ch = 0; ch < 3; ++ch) out[ch] = (unsigned char)min_d(255.0, (double)base[ch] * (0.1 + I));
        return;
    }
    double sf = 1.0;
    if (e.kind == TRB_SHADER_SHADOW_PHONG) {  // ShadowPhongAuthored::shadow_factor
        const ShadowEnv& S = e.shadow;
        double c[4];
        for (int k = 0; k < 4; ++k) c[k] = v.light_clip[0][k] * b[0] + v.light_clip[1][k] * b[1] + v.light_clip[2][k] * b[2];
        if (c[3] > 1e-12) {
            double ndc[4] = {c[0] / c[3], c[1] / c[3], c[2] / c[3], c[3] / c[3]};
            double sx = dot4(S.lvp.m[0], ndc), sy = dot4(S.lvp.m[1], ndc);
            if (sx >= 0.0 && sy >= 0.0) {
                int ix = x86_int(sx), iy = x86_int(sy);
                if (ix >= 0 && iy >= 0 && ix < S.w && iy < S.h) {
                    double zs = (*S.map)[(size_t)ix + (size_t)iy * S.w];
                    if (ndc[2] > zs + S.bias) sf = S.darkening;
                }
            }
        }
    }
    if (e.kind == TRB_SHADER_PHONG || e.kind == TRB_SHADER_SHADOW_PHONG) {  // main.cpp:92-170
        double spec_pow = max_d(1.0, (double)sample_specular(e, uv));
        double brightness = (base[0] + base[1] + base[2]) / (3.0 * 255.0);
        bool eye_px = (brightness >= 0.85) && (spec_pow <= 5.0);
        D3 nm = sample_normal(e, uv);
        double nm4[4] = {nm.x, nm.y, nm.z, 0.0}, nme[4];
        mul_mv(e.modelview, nm4, nme);
        D3 nm_eye{nme[0], nme[1], nme[2]};
        double s = e.u.normal_map_strength;
        D3 N = eye_px ? gn : normalize3(add3(scale3(gn, 1.0 - s), scale3(nm_eye, s)));
        D3 V = normalize3(scale3(pos, -1.0));  // operator- is v * -1.0, geometry.h:243-246
        double key_d = max_d(0.0, dot3(N, key)) * 1.0;
        D3 R = normalize3(sub3(scale3(N, 2.0 * dot3(N, key)), key));
        double rv = max_d(0.0, dot3(R, V));
        double key_s = (rv > 0.0 ? std::pow(rv, spec_pow) : 0.0) * 1.0;
        double fill_d = max_d(0.0, dot3(N, fill)) * 0.35;
        double rim_d = max_d(0.0, dot3(N, rim)) * 0.6;
        double diff = key_d + fill_d + rim_d;
        for (int ch = 0; ch < 3; ++ch) {
            double cv = base[ch];
            double val = cv * (0.10 + diff * sf) + 255.0 * ((0.35 * key_s) * sf);
            out[ch] = (unsigned char)min_d(255.0, val);
        }
    } else {  // EYE, main.cpp:220-261
        D3 N = normalize3(gn);
        D3 V = normalize3(scale3(pos, -1.0));
        double key_d = max_d(0.0, dot3(N, key)) * 1.0;
        double rim_d = max_d(0.0, dot3(N, rim)) * 0.6;
        double diff = key_d + rim_d;
        double spec_pow = max_d(1.0, (double)sample_specular(e, uv)) * 8.0;
        D3 R = normalize3(sub3(scale3(N, 2.0 * dot3(N, key)), key));
        double rv = max_d(0.0, dot3(R, V));
        double spec = (rv > 0.0 ? std::pow(rv, spec_pow) : 0.0);
        for (int ch = 0; ch < 3; ++ch) {
            double cv = base[ch];
            double val = cv * (0.1 + diff) + 255.0 * (1.5 * spec);
            out[ch] = (unsigned char)min_d(255.0, val);
        }
    }
}

struct View {
    std::vector<double> z, z_snap;
    std::vector<uint8_t> bgr;
    TrbStats st{};
};

void reset_stats(TrbStats& s) {
    std::memset(&s, 0, sizeof(s));
    s.bbox_min_x = s.bbox_min_y = INT_MAX;  // our_gl.cpp:20
    s.bbox_max_x = s.bbox_max_y = INT_MIN;
    s.z_min = std::numeric_limits<double>::infinity();       // our_gl.cpp:21
    s.z_max_covered = -std::numeric_limits<double>::infinity();
    s.z_max_ref = -std::numeric_limits<double>::infinity();  // our_gl.cpp:22
}

// rasterize(), our_gl.cpp:89-201, steps numbered as in SURVEY 8(a) R1
void rasterize_port(View& view, int W, int H, const M4& viewport, const double clip[3][4],
                    const ShadeEnv& env, const Varyings& vary, bool write_color) {
    TrbStats& st = view.st;
    ++st.triangles_submitted;                                            // (1) :90
    if (clip[0][3] <= 1e-12 || clip[1][3] <= 1e-12 || clip[2][3] <= 1e-12) return;  // (2) :94
    double ndc[3][4];
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 4; ++c) ndc[i][c] = clip[i][c] / clip[i][3];  // (3) :101
    bool out0 = ndc[0][2] < -1.0 || ndc[0][2] > 1.0;                       // (4) :103-106
    bool out1 = ndc[1][2] < -1.0 || ndc[1][2] > 1.0;
    bool out2 = ndc[2][2] < -1.0 || ndc[2][2] > 1.0;
    if (out0 && out1 && out2) return;
    for (int i = 0; i < 3; ++i)                                           // (5) :109-114
        for (int c = 0; c < 4; ++c)
            if (!std::isfinite(ndc[i][c])) return;
    double sx[3], sy[3];
    for (int i = 0; i < 3; ++i) {                                         // (6) :117-121
        sx[i] = dot4(viewport.m[0], ndc[i]);
        sy[i] = dot4(viewport.m[1], ndc[i]);
    }
    double e1x = sx[1] - sx[0], e1y = sy[1] - sy[0];                       // (7) :124-127
    double e2x = sx[2] - sx[0], e2y = sy[2] - sy[0];
    double cross = e1x * e2y - e1y * e2x;
    if (cross <= 0) return;
    // (8) :130-135
    int x0 = std::max(0, x86_int(std::floor(min3(sx[0], sx[1], sx[2]))));
    int x1 = std::min(W - 1, x86_int(std::ceil(max3(sx[0], sx[1], sx[2]))));
    int y0 = std::max(0, x86_int(std::floor(min3(sy[0], sy[1], sy[2]))));
    int y1 = std::min(H - 1, x86_int(std::ceil(max3(sy[0], sy[1], sy[2]))));
    if (x0 > x1 || y0 > y1) return;
    ++st.triangles_binned;
    st.bbox_min_x = std::min(st.bbox_min_x, x0);                          // (9) :138-141
    st.bbox_min_y = std::min(st.bbox_min_y, y0);
    st.bbox_max_x = std::max(st.bbox_max_x, x1);
    st.bbox_max_y = std::max(st.bbox_max_y, y1);
    double w0 = clip[0][3], w1 = clip[1][3], w2 = clip[2][3];

    // barycentric(), our_gl.cpp:77-86: the triangle-constant parts of s0, s1 and u.z
    double s00 = sx[2] - sx[0], s01 = sx[1] - sx[0];
    double s10 = sy[2] - sy[0], s11 = sy[1] - sy[0];
    for (int x = x0; x <= x1; ++x)                                        // (10) :147-149
        for (int y = y0; y <= y1; ++y) {
            double px = (double)x + 0.5, py = (double)y + 0.5;
            double s02 = sx[0] - px, s12 = sy[0] - py;
            double ux = s01 * s12 - s02 * s11;                            // cross(), geometry.h:143-149
            double uy = s02 * s10 - s00 * s12;
            double uz = s00 * s11 - s01 * s10;
            double b0, b1, b2;
            if (std::abs(uz) < 1e-12) {                                   // (11) :82-83
                b0 = -1; b1 = 1; b2 = 1;
            } else {
                b0 = 1.0 - (ux + uy) / uz;
                b1 = uy / uz;
                b2 = ux / uz;
            }
            if (b0 < 0 || b1 < 0 || b2 < 0) continue;                     // (12) :152
            double z = b0 * ndc[0][2] + b1 * ndc[1][2] + b2 * ndc[2][2];  // (13) :156-158
            if (!std::isfinite(z)) continue;                              // (14) :160
            ++st.fragments_covered;
            size_t idx = (size_t)x + (size_t)y * W;                       // (15) :162-165
            if (!(z < view.z[idx])) continue;
            double iw0 = (std::abs(w0) > 1e-12) ? (1.0 / w0) : 0.0;       // (16) :168-185
            double iw1 = (std::abs(w1) > 1e-12) ? (1.0 / w1) : 0.0;
            double iw2 = (std::abs(w2) > 1e-12) ? (1.0 / w2) : 0.0;
            double denom = b0 * iw0 + b1 * iw1 + b2 * iw2;
            double pc[3];
            if (std::abs(denom) < 1e-15) {
                pc[0] = b0; pc[1] = b1; pc[2] = b2;
            } else {
                pc[0] = (b0 * iw0) / denom;
                pc[1] = (b1 * iw1) / denom;
                pc[2] = (b2 * iw2) / denom;
            }
            uint8_t col[3];
            shade(env, vary, pc, col);                                    // (17) :187 (never discards)
            view.z[idx] = z;                                              // (18) :191-198
            if (write_color) std::memcpy(&view.bgr[idx * 3], col, 3);
            ++st.fragments_drawn_ref;
            st.z_min = std::min(st.z_min, z);
            st.z_max_ref = std::max(st.z_max_ref, z);
        }
}

}  // namespace

struct TrbCtx {
    std::string err;
    int w = 0, h = 0, nviews = 0;
    uint8_t clear[3] = {0, 0, 0};
    M4 viewport{};
    std::vector<View> views;
    std::vector<std::unique_ptr<Mesh>> meshes;
    std::vector<std::unique_ptr<Tex>> textures;
    struct ShadowMapCopy { std::vector<double> z; int w, h; };
    std::vector<ShadowMapCopy> shadow_maps;
    int threads = 1;
};

namespace {
int fail(TrbCtx* c, int code, const char* msg) {
    if (c) c->err = msg;
    return code;
}
M4 load_mat(const double* p) {
    M4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = p[i * 4 + j];
    return r;
}
const Tex* tex_of(TrbCtx* c, TrbTex t) {
    if (t == 0 || t > c->textures.size()) return nullptr;
    return c->textures[t - 1].get();
}

// PhongShader::vertex / EyeShader::vertex (main.cpp:71-90) for one face, then rasterize
void draw_face(TrbCtx* c, View& view, const Mesh& m, const M4& mv, const M4& pr, const ShadeEnv& env,
               uint64_t face, bool write_color) {
    double clip[3][4];
    Varyings vy;
    for (int k = 0; k < 3; ++k) {
        uint32_t vi = m.idx[face * 3 + k];  // Model::vert(iface,nthvert), model.cpp:396-400
        double p[4] = {m.pos[3 * vi], m.pos[3 * vi + 1], m.pos[3 * vi + 2], 1.0};
        double n[4] = {m.nrm[3 * vi], m.nrm[3 * vi + 1], m.nrm[3 * vi + 2], 0.0};
        vy.uv[k][0] = m.uv[2 * vi];
        vy.uv[k][1] = m.uv[2 * vi + 1];
        double pe[4], ne[4];
        mul_mv(mv, p, pe);
        vy.pos_eye[k] = D3{pe[0], pe[1], pe[2]};
        mul_mv(mv, n, ne);
        vy.nrm_eye[k] = D3{ne[0], ne[1], ne[2]};
        mul_mv(pr, pe, clip[k]);
        if (env.kind == TRB_SHADER_SHADOW_PHONG) {
            double le[4];
            mul_mv(env.shadow.lmv, p, le);
            mul_mv(env.shadow.lpr, le, vy.light_clip[k]);
        }
    }
    rasterize_port(view, c->w, c->h, c->viewport, clip, env, vy, write_color);
}
}  // namespace

extern "C" {

int orc_create(int, TrbCtx** out) {
    if (!out) return TRB_E_ARG;
    *out = new TrbCtx();
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) (*out)->viewport.m[i][j] = i == j;
    const char* t = getenv("TRB_ORACLE_THREADS");
    (*out)->threads = t ? std::max(1, atoi(t)) : 1;
    return TRB_OK;
}
int orc_destroy(TrbCtx* c) {
    delete c;
    return TRB_OK;
}
const char* orc_last_error(TrbCtx* c) { return c ? c->err.c_str() : "null context"; }
const char* orc_backend_name(void) { return "oracle-port"; }

int orc_upload_mesh(TrbCtx* c, const float* pos3, const float* nrm3, const float* uv2,
                    uint32_t nverts, const uint32_t* idx, uint64_t nidx, TrbMesh* out) {
    if (!c || !pos3 || !out || nidx % 3) return fail(c, TRB_E_ARG, "upload_mesh: bad argument");
    auto m = std::make_unique<Mesh>();
    m->pos.assign(pos3, pos3 + (size_t)nverts * 3);
    m->nrm.resize((size_t)nverts * 3);
    m->uv.resize((size_t)nverts * 2, 0.f);
    for (uint32_t i = 0; i < nverts; ++i)
        for (int k = 0; k < 3; ++k) m->nrm[3 * i + k] = nrm3 ? nrm3[3 * i + k] : (k == 2 ? 1.f : 0.f);
    if (uv2) m->uv.assign(uv2, uv2 + (size_t)nverts * 2);
    m->idx.resize(nidx);
    for (uint64_t i = 0; i < nidx; ++i) {
        uint32_t k = idx ? idx[i] : (uint32_t)i;
        if (k >= nverts) return fail(c, TRB_E_ARG, "upload_mesh: index out of range");
        m->idx[i] = k;
    }
    c->meshes.push_back(std::move(m));
    *out = c->meshes.size();
    return TRB_OK;
}
int orc_free_mesh(TrbCtx* c, TrbMesh m) {
    if (!c || m == 0 || m > c->meshes.size() || !c->meshes[m - 1]) return fail(c, TRB_E_ARG, "free_mesh");
    c->meshes[m - 1].reset();
    return TRB_OK;
}
int orc_upload_texture(TrbCtx* c, const uint8_t* texels, int w, int h, int bpp, TrbTex* out) {
    if (!c || !texels || !out || w <= 0 || h <= 0 || (bpp != 1 && bpp != 3 && bpp != 4))
        return fail(c, TRB_E_ARG, "upload_texture: bad argument");
    auto t = std::make_unique<Tex>();
    t->w = w;
    t->h = h;
    t->bpp = bpp;
    t->px.assign(texels, texels + (size_t)w * h * bpp);
    c->textures.push_back(std::move(t));
    *out = c->textures.size();
    return TRB_OK;
}
int orc_free_texture(TrbCtx* c, TrbTex t) {
    if (!c || t == 0 || t > c->textures.size() || !c->textures[t - 1]) return fail(c, TRB_E_ARG, "free_texture");
    c->textures[t - 1].reset();
    return TRB_OK;
}

int orc_begin_batch(TrbCtx* c, int w, int h, int nviews) {
    if (!c || w <= 0 || h <= 0 || nviews <= 0) return fail(c, TRB_E_ARG, "begin_batch: bad size");
    c->w = w;
    c->h = h;
    c->nviews = nviews;
    c->views.assign(nviews, View());
    for (auto& v : c->views) {
        v.z.assign((size_t)w * h, std::numeric_limits<double>::infinity());  // our_gl.cpp:72-74
        v.bgr.resize((size_t)w * h * 3);
        for (size_t i = 0; i < (size_t)w * h; ++i) std::memcpy(&v.bgr[3 * i], c->clear, 3);
        reset_stats(v.st);
    }
    return TRB_OK;
}
int orc_begin_frame(TrbCtx* c, int w, int h) { return orc_begin_batch(c, w, h, 1); }
int orc_set_clear_color(TrbCtx* c, uint8_t b, uint8_t g, uint8_t r) {
    if (!c) return TRB_E_ARG;
    c->clear[0] = b;
    c->clear[1] = g;
    c->clear[2] = r;
    return TRB_OK;
}
int orc_set_viewport(TrbCtx* c, const double* v) {
    if (!c || !v) return fail(c, TRB_E_ARG, "set_viewport");
    c->viewport = load_mat(v);
    return TRB_OK;
}

static int make_env(TrbCtx* c, int kind, const void* uniforms, size_t ubytes, int vi, const double* mv,
                    ShadeEnv& env) {
    env.kind = kind;
    if (mv) env.modelview = load_mat(mv);
    if (kind == TRB_SHADER_SHADOW_PHONG) {
        if (!uniforms || ubytes != sizeof(TrbShadowUniforms)) return fail(c, TRB_E_ARG, "draw: uniforms");
        const TrbShadowUniforms& su = ((const TrbShadowUniforms*)uniforms)[vi];
        if (su.shadow_map < 0 || (size_t)su.shadow_map >= c->shadow_maps.size()) return fail(c, TRB_E_ARG, "draw: shadow map");
        const auto& sm = c->shadow_maps[su.shadow_map];
        if (sm.w != su.shadow_w || sm.h != su.shadow_h) return fail(c, TRB_E_ARG, "draw: shadow map size");
        env.u = su.phong;
        env.diffuse = tex_of(c, env.u.diffuse);
        env.normal = tex_of(c, env.u.normal);
        env.specular = tex_of(c, env.u.specular);
        env.shadow.lmv = load_mat(su.light_modelview);
        env.shadow.lpr = load_mat(su.light_perspective);
        env.shadow.lvp = load_mat(su.light_viewport);
        env.shadow.bias = su.shadow_bias;
        env.shadow.darkening = su.shadow_darkening;
        env.shadow.map = &sm.z;
        env.shadow.w = sm.w;
        env.shadow.h = sm.h;
    } else if (kind == TRB_SHADER_PHONG || kind == TRB_SHADER_EYE || kind == TRB_SHADER_GOURAUD) {
        if (!uniforms || ubytes != sizeof(TrbPhongUniforms)) return fail(c, TRB_E_ARG, "draw: uniforms");
        env.u = ((const TrbPhongUniforms*)uniforms)[vi];
        env.diffuse = tex_of(c, env.u.diffuse);
        env.normal = tex_of(c, env.u.normal);
        env.specular = tex_of(c, env.u.specular);
    } else if (kind != TRB_SHADER_FLAT_BARY && kind != TRB_SHADER_DEPTH) {
        return fail(c, TRB_E_SHADER, "shader kind not available in oracle-port");
    }
    return TRB_OK;
}

int orc_draw_batch(TrbCtx* c, TrbMesh mesh, const double* mv, const double* pr, int kind,
                   const void* uniforms, size_t ubytes, uint64_t first, uint64_t ntris) {
    if (!c || c->views.empty()) return fail(c, TRB_E_ARG, "draw: no frame");
    if (mesh == 0 || mesh > c->meshes.size() || !c->meshes[mesh - 1]) return fail(c, TRB_E_ARG, "draw: bad mesh");
    const Mesh& m = *c->meshes[mesh - 1];
    if ((first + ntris) * 3 > m.idx.size()) return fail(c, TRB_E_ARG, "draw: triangle range");
    std::vector<ShadeEnv> envs(c->nviews);
    for (int vi = 0; vi < c->nviews; ++vi) {
        int rc = make_env(c, kind, uniforms, ubytes, vi, mv + 16 * vi, envs[vi]);
        if (rc) return rc;
    }
    auto work = [&](int vi) {
        M4 MV = load_mat(mv + 16 * vi), PR = load_mat(pr + 16 * vi);
        for (uint64_t f = first; f < first + ntris; ++f)
            draw_face(c, c->views[vi], m, MV, PR, envs[vi], f, kind != TRB_SHADER_DEPTH);
    };
    int nt = std::min(c->threads, c->nviews);
    if (nt <= 1) {
        for (int vi = 0; vi < c->nviews; ++vi) work(vi);
    } else {  // views are independent frames: one thread per view, striped
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; ++t)
            pool.emplace_back([&, t] {
                for (int vi = t; vi < c->nviews; vi += nt) work(vi);
            });
        for (auto& th : pool) th.join();
    }
    return TRB_OK;
}
int orc_draw(TrbCtx* c, TrbMesh mesh, const double* mv, const double* pr, int kind,
             const void* uniforms, size_t ubytes, uint64_t first, uint64_t ntris) {
    if (c && c->nviews != 1) return fail(c, TRB_E_ARG, "draw: batch frame needs draw_batch");
    return orc_draw_batch(c, mesh, mv, pr, kind, uniforms, ubytes, first, ntris);
}

int orc_submit_clip_triangles(TrbCtx* c, const double* clip12, const double* varyings, uint64_t n,
                              const double* mv, int kind, const void* uniforms, size_t ubytes) {
    if (!c || c->views.empty() || c->nviews != 1) return fail(c, TRB_E_ARG, "submit: needs a single-view frame");
    if (!clip12 && n) return fail(c, TRB_E_ARG, "submit: null clip");
    ShadeEnv env;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) env.modelview.m[i][j] = i == j;
    int rc = make_env(c, kind, uniforms, ubytes, 0, mv, env);
    if (rc) return rc;
    bool lit = kind == TRB_SHADER_PHONG || kind == TRB_SHADER_EYE;
    if (lit && !varyings) return fail(c, TRB_E_ARG, "submit: varyings required");
    for (uint64_t t = 0; t < n; ++t) {
        double clip[3][4];
        Varyings vy{};
        for (int v = 0; v < 3; ++v) {
            for (int k = 0; k < 4; ++k) clip[v][k] = clip12[t * 12 + v * 4 + k];
            if (lit) {
                const double* vr = varyings + t * 24 + v * 8;
                vy.uv[v][0] = vr[0];
                vy.uv[v][1] = vr[1];
                vy.pos_eye[v] = D3{vr[2], vr[3], vr[4]};
                vy.nrm_eye[v] = D3{vr[5], vr[6], vr[7]};
            }
        }
        rasterize_port(c->views[0], c->w, c->h, c->viewport, clip, env, vy, kind != TRB_SHADER_DEPTH);
    }
    return TRB_OK;
}

int orc_depth_snapshot(TrbCtx* c) {
    if (!c || c->views.empty()) return fail(c, TRB_E_ARG, "depth_snapshot");
    for (auto& v : c->views) v.z_snap = v.z;  // main.cpp:700
    return TRB_OK;
}
int orc_depth_restore(TrbCtx* c) {
    if (!c || c->views.empty()) return fail(c, TRB_E_ARG, "depth_restore");
    for (auto& v : c->views) {
        if (v.z_snap.size() != v.z.size()) return fail(c, TRB_E_ARG, "depth_restore: no snapshot");
        v.z = v.z_snap;  // main.cpp:730
    }
    return TRB_OK;
}
int orc_keep_depth_as_shadow_map(TrbCtx* c, int32_t* out) {
    if (!c || c->views.empty() || !out) return fail(c, TRB_E_ARG, "keep_depth_as_shadow_map");
    c->shadow_maps.push_back(TrbCtx::ShadowMapCopy{c->views[0].z, c->w, c->h});
    *out = (int32_t)c->shadow_maps.size() - 1;
    return TRB_OK;
}
int orc_release_shadow_maps(TrbCtx* c) {
    if (!c) return TRB_E_ARG;
    c->shadow_maps.clear();
    return TRB_OK;
}
int orc_flush(TrbCtx* c) { return c ? TRB_OK : TRB_E_ARG; }
int orc_end_frame(TrbCtx* c) { return c ? TRB_OK : TRB_E_ARG; }

static const double* orc_view_depth(TrbCtx* c, int view, int* w, int* h) {
    if (!c || view < 0 || view >= c->nviews) return nullptr;
    *w = c->w;
    *h = c->h;
    return c->views[view].z.data();
}
static const uint8_t* orc_view_color(TrbCtx* c, int view) {
    if (!c || view < 0 || view >= c->nviews) return nullptr;
    return c->views[view].bgr.data();
}
#include "post_restate.inc"

int orc_read_color(TrbCtx* c, int view, uint8_t* out) {
    if (!c || view < 0 || view >= c->nviews || !out) return fail(c, TRB_E_ARG, "read_color");
    std::memcpy(out, c->views[view].bgr.data(), c->views[view].bgr.size());
    return TRB_OK;
}
int orc_write_color(TrbCtx* c, int view, const uint8_t* bgr) {
    if (!c || view < 0 || view >= c->nviews || !bgr) return fail(c, TRB_E_ARG, "write_color");
    std::memcpy(c->views[view].bgr.data(), bgr, c->views[view].bgr.size());
    return TRB_OK;
}
int orc_read_depth(TrbCtx* c, int view, double* out) {
    if (!c || view < 0 || view >= c->nviews || !out) return fail(c, TRB_E_ARG, "read_depth");
    std::memcpy(out, c->views[view].z.data(), sizeof(double) * c->views[view].z.size());
    return TRB_OK;
}
int orc_readback_async(TrbCtx* c, uint8_t* const* color_out, double* const* depth_out) {
    if (!c || c->views.empty()) return TRB_E_ARG;
    for (int v = 0; v < c->nviews; ++v) {
        if (color_out && color_out[v]) orc_read_color(c, v, color_out[v]);
        if (depth_out && depth_out[v]) orc_read_depth(c, v, depth_out[v]);
    }
    return TRB_OK;
}
int orc_readback_wait(TrbCtx* c) { return c ? TRB_OK : TRB_E_ARG; }
int orc_read_visibility(TrbCtx* c, int, uint32_t*) { return fail(c, TRB_E_SHADER, "not in oracle"); }
int orc_get_stats(TrbCtx* c, int view, TrbStats* out) {
    if (!c || view < 0 || view >= c->nviews || !out) return fail(c, TRB_E_ARG, "get_stats");
    *out = c->views[view].st;
    uint64_t px = 0;
    double zmax = -std::numeric_limits<double>::infinity();
    for (double z : c->views[view].z)
        if (std::isfinite(z)) {
            ++px;
            zmax = std::max(zmax, z);
        }
    out->pixels_shaded = px;
    out->z_max_covered = zmax;
    return TRB_OK;
}
int orc_synchronize(TrbCtx* c) { return c ? TRB_OK : TRB_E_ARG; }
int orc_timer_start(TrbCtx* c) { return c ? TRB_OK : TRB_E_ARG; }
int orc_timer_stop_ms(TrbCtx* c, float* ms) {
    if (ms) *ms = 0.f;
    return c ? TRB_OK : TRB_E_ARG;
}
int orc_profile_enable(TrbCtx* c, int) { return c ? TRB_OK : TRB_E_ARG; }
int orc_profile_read(TrbCtx* c, TrbKernelTime*, int, int* n, int) {
    if (n) *n = 0;
    return c ? TRB_OK : TRB_E_ARG;
}
uint64_t orc_launch_count(TrbCtx*) { return 0; }
int orc_device_planes(TrbCtx* c, uint64_t*, uint64_t*, uint64_t*) { return fail(c, TRB_E_COMM, "oracle has no device"); }
int orc_set_triangle_id_base(TrbCtx* c, uint64_t) { return c ? TRB_OK : TRB_E_ARG; }
// trb_draw_shard on the CPU checker: rank r's share is a contiguous range of the index buffer (any partition of the
// triangles gives the same composited picture; the oracle keeps no ids)
int orc_draw_shard(TrbCtx* c, TrbMesh mesh, const double* mv, const double* pr, int kind, const void* uniforms,
                   size_t ubytes, int shard_rank, int shard_count) {
    if (!c || shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count) return fail(c, TRB_E_ARG, "draw_shard: bad rank / count");
    if (mesh == 0 || mesh > c->meshes.size() || !c->meshes[mesh - 1]) return fail(c, TRB_E_ARG, "draw: bad mesh");
    const uint64_t total = c->meshes[mesh - 1]->idx.size() / 3, n = (uint64_t)shard_count, r = (uint64_t)shard_rank;
    const uint64_t base = total / n, rem = total % n;
    return orc_draw(c, mesh, mv, pr, kind, uniforms, ubytes, r * base + (r < rem ? r : rem), base + (r < rem ? 1 : 0));
}
int orc_composite_save_local_depth(TrbCtx* c) { return fail(c, TRB_E_COMM, "oracle has no device"); }
int orc_composite_mask(TrbCtx* c) { return fail(c, TRB_E_COMM, "oracle has no device"); }
int orc_composite_finish(TrbCtx* c) { return fail(c, TRB_E_COMM, "oracle has no device"); }
int orc_ipc_export_planes(TrbCtx* c, void*, void*) { return fail(c, TRB_E_COMM, "oracle has no device"); }
int orc_ipc_open_peers(TrbCtx* c, const void*, const void*, int, int) { return fail(c, TRB_E_COMM, "oracle has no device"); }
int orc_open_peers_raw(TrbCtx* c, const uint64_t*, const uint64_t*, int, int) { return fail(c, TRB_E_COMM, "oracle has no device"); }
int orc_ipc_close_peers(TrbCtx* c) { return c ? TRB_OK : TRB_E_ARG; }
int orc_composite_shade_p2p(TrbCtx* c, int, int) { return fail(c, TRB_E_COMM, "oracle has no device"); }
int orc_comm_init(TrbCtx* const*, int) { return TRB_E_COMM; }
// frame recordings are CUDA graphs: the CPU checker renders every frame by plain calls
int orc_record_begin(TrbCtx* c) { return fail(c, TRB_E_ARG, "oracle has no device"); }
int orc_record_end(TrbCtx* c, TrbRecording*) { return fail(c, TRB_E_ARG, "oracle has no device"); }
int orc_replay(TrbCtx* c, TrbRecording, const TrbReplayDraw*, int) { return fail(c, TRB_E_ARG, "oracle has no device"); }
int orc_recording_free(TrbCtx* c, TrbRecording) { return fail(c, TRB_E_ARG, "oracle has no device"); }
int orc_comm_export(TrbCtx* c, void*, size_t) { return fail(c, TRB_E_COMM, "oracle has no device"); }
int orc_comm_open(TrbCtx* c, const void*, int, int) { return fail(c, TRB_E_COMM, "oracle has no device"); }
int orc_comm_close(TrbCtx* c) { return fail(c, TRB_E_COMM, "oracle has no device"); }
int orc_comm_shard(TrbCtx* c, uint64_t, uint64_t*, uint64_t*) { return fail(c, TRB_E_COMM, "oracle has no device"); }
int orc_comm_rows(TrbCtx* c, int*, int*) { return fail(c, TRB_E_COMM, "oracle has no device"); }
int orc_composite(TrbCtx* c) { return fail(c, TRB_E_COMM, "oracle has no device"); }
int orc_composite_group(TrbCtx* const*, int) { return TRB_E_COMM; }
int orc_set_shade_rows(TrbCtx* c, int, int) { return fail(c, TRB_E_COMM, "oracle has no device"); }

// ---- host helpers ------------------------------------------------------------------------
void orc_light_dir_eye(const double* mv, const double* dir, double* out) {
    // main.cpp:59-68: mat<3,3> * vec3 through dot<3>, then normalized
    D3 d{dir[0], dir[1], dir[2]};
    D3 r;
    r.x = dot3(D3{mv[0], mv[1], mv[2]}, d);
    r.y = dot3(D3{mv[4], mv[5], mv[6]}, d);
    r.z = dot3(D3{mv[8], mv[9], mv[10]}, d);
    r = normalize3(r);
    out[0] = r.x;
    out[1] = r.y;
    out[2] = r.z;
}
void orc_lookat(const double* eye, const double* center, const double* up, double* out) {
    // our_gl.cpp:25-41
    D3 e{eye[0], eye[1], eye[2]}, c{center[0], center[1], center[2]}, u{up[0], up[1], up[2]};
    D3 z = normalize3(sub3(e, c));
    D3 cx{u.y * z.z - u.z * z.y, u.z * z.x - u.x * z.z, u.x * z.y - u.y * z.x};  // cross(up,z)
    D3 x = normalize3(cx);
    D3 y{z.y * x.z - z.z * x.y, z.z * x.x - z.x * x.z, z.x * x.y - z.y * x.x};  // cross(z,x)
    for (int i = 0; i < 16; ++i) out[i] = (i % 5 == 0) ? 1.0 : 0.0;
    out[0] = x.x; out[1] = x.y; out[2] = x.z;
    out[4] = y.x; out[5] = y.y; out[6] = y.z;
    out[8] = z.x; out[9] = z.y; out[10] = z.z;
    out[3] = -dot3(x, e);
    out[7] = -dot3(y, e);
    out[11] = -dot3(z, e);
}
void orc_perspective(double fov_deg, double aspect, double zn, double zf, double* out) {
    // our_gl.cpp:44-56
    double fov_rad = fov_deg * M_PI / 180.0;
    double t = std::tan(fov_rad / 2.0);
    for (int i = 0; i < 16; ++i) out[i] = (i % 5 == 0) ? 1.0 : 0.0;
    out[0] = 1.0 / (aspect * t);
    out[5] = 1.0 / t;
    out[10] = (zf + zn) / (zn - zf);
    out[11] = (2.0 * zf * zn) / (zn - zf);
    out[14] = -1.0;
    out[15] = 0.0;
}
void orc_viewport(int x, int y, int w, int h, double* out) {
    // our_gl.cpp:59-69
    for (int i = 0; i < 16; ++i) out[i] = (i % 5 == 0) ? 1.0 : 0.0;
    out[0] = w / 2.0;
    out[5] = h / 2.0;
    out[3] = x + w / 2.0;
    out[7] = y + h / 2.0;
    out[10] = 1.0;
    out[11] = 0.0;
}
void orc_mat4_mul(const double* a, const double* b, double* out) {
    // geometry.h:195-205
    double r[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
            r[i * 4 + j] = s;
        }
    std::memcpy(out, r, sizeof(r));
}

// Frustum::createFromMatrix / intersects (our_gl.cpp:212-280), AABB::transform (geometry.h:297-327), restated
void orc_frustum_planes(const double* m, double* planes) {
    for (int p = 0; p < 6; ++p) {
        const int c = p / 2;
        const double s = (p & 1) ? -1.0 : 1.0;
        double nx = m[3] + s * m[c], ny = m[7] + s * m[4 + c], nz = m[11] + s * m[8 + c], d = m[15] + s * m[12 + c];
        double len = sqrt(dot3(D3{nx, ny, nz}, D3{nx, ny, nz}));
        if (len > 0.0) { nx = nx / len; ny = ny / len; nz = nz / len; d /= len; }
        planes[4 * p] = nx; planes[4 * p + 1] = ny; planes[4 * p + 2] = nz; planes[4 * p + 3] = d;
    }
}
int orc_frustum_intersects(const double* planes, const double* lo, const double* hi) {
    for (int p = 0; p < 6; ++p) {
        const double* pl = planes + 4 * p;
        D3 c{pl[0] >= 0 ? hi[0] : lo[0], pl[1] >= 0 ? hi[1] : lo[1], pl[2] >= 0 ? hi[2] : lo[2]};
        if (dot3(D3{pl[0], pl[1], pl[2]}, c) + pl[3] < 0) return 0;
    }
    return 1;
}
void orc_aabb_transform(const double* lo, const double* hi, const double* m, double* out_lo, double* out_hi) {
    double nlo[3] = {1e9, 1e9, 1e9}, nhi[3] = {-1e9, -1e9, -1e9};
    for (int i = 0; i < 8; ++i) {
        const double cx = (i & 1) ? hi[0] : lo[0], cy = (i & 2) ? hi[1] : lo[1], cz = (i & 4) ? hi[2] : lo[2];
        const double c4[4] = {cx, cy, cz, 1.0};
        double t[4];
        for (int r = 0; r < 4; ++r) t[r] = dot4(m + 4 * r, c4);
        for (int k = 0; k < 3; ++k) {
            const double v = t[k] / t[3];
            if (v < nlo[k]) nlo[k] = v;
            if (nhi[k] < v) nhi[k] = v;
        }
    }
    for (int k = 0; k < 3; ++k) { out_lo[k] = nlo[k]; out_hi[k] = nhi[k]; }
}
void orc_mat4_mul(const double* a, const double* b, double* out);
void orc_cull_batch(const double* perspective, const double* views, int n, const double* lo, const double* hi, uint8_t* out) {
    for (int v = 0; v < n; ++v) {
        double vp[16], planes[24];
        orc_mat4_mul(perspective, views + 16 * v, vp);
        orc_frustum_planes(vp, planes);
        out[v] = (uint8_t)orc_frustum_intersects(planes, lo, hi);
    }
}

void orc_mat4_mul_batch(const double* a, int n, const double* b, double* out) {
    for (int v = 0; v < n; ++v) orc_mat4_mul(a + 16 * v, b, out + 16 * v);
}
void orc_light_dir_eye_batch(const double* mvs, int n, const double* dir, double* out) {
    for (int v = 0; v < n; ++v) orc_light_dir_eye(mvs + 16 * v, dir, out + 3 * v);
}

}  // extern "C"
