// oracle/ref_harness.cpp - TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Wraps the reference's OWN rasterizer - /root/reference/our_gl.cpp and tgaimage.cpp,
// compiled where they lie by oracle/build_oracle.py into oracle/_ref/libtrb_ref.so - behind
// the C ABI of include/trb.h (prefix orc_), so the parity tests drive the real
// rasterize() (our_gl.cpp:89-201) exactly like main.cpp's draw loops do (main.cpp:660-666).
//
// What is the reference's code here: rasterize(), lookat(), init_perspective(),
// init_viewport(), init_zbuffer(), print_render_stats(), IShader, TGAImage/TGAColor,
// vec/mat.  What is restated (main.cpp and model.cpp need Assimp and cannot compile,
// SURVEY F4): PhongShader / EyeShader (main.cpp:39-262), the Model accessors
// (model.cpp:391-459), SSAO / z-image / AO composite (main.cpp:269-362, 756-786).
#define TRB_FN(name) orc_##name
#include "../include/trb.h"

#include "our_gl.h"  // the reference header, found through -I/root/reference

#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------
// Model stand-in: the accessor semantics of model.cpp:391-459 over caller-supplied arrays.
// ---------------------------------------------------------------------------------------
struct MeshData {
    std::vector<vec3> position, normal;
    std::vector<vec2> texcoord;
    std::vector<unsigned int> indices;
    bool alive = true;
};

struct StandInModel {
    const MeshData* mesh = nullptr;
    const TGAImage* diffuse_tex = nullptr;   // nullptr <=> materials[0].hasDiffuse()==false
    const TGAImage* normal_tex = nullptr;
    const TGAImage* specular_tex = nullptr;

    int nfaces() const { return (int)mesh->indices.size() / 3; }

    // model.cpp:396-412
    vec3 vert(int iface, int nthvert) const {
        int k = iface * 3 + nthvert;
        if (k < 0 || k >= (int)mesh->indices.size()) return vec3{0, 0, 0};
        return mesh->position[mesh->indices[k]];
    }
    vec3 normal(int iface, int nthvert) const {
        int k = iface * 3 + nthvert;
        if (k < 0 || k >= (int)mesh->indices.size()) return vec3{0, 0, 1};
        return mesh->normal[mesh->indices[k]];
    }
    vec2 uv(int iface, int nthvert) const {
        int k = iface * 3 + nthvert;
        if (k < 0 || k >= (int)mesh->indices.size()) return vec2{0, 0};
        return mesh->texcoord[mesh->indices[k]];
    }
    static void texel_of(const TGAImage& t, const vec2& uv, int& x, int& y) {
        // model.cpp:420-423: int() truncation then clamp
        x = std::clamp(int(uv.x * t.width()), 0, t.width() - 1);
        y = std::clamp(int(uv.y * t.height()), 0, t.height() - 1);
    }
    // model.cpp:415-425
    TGAColor diffuse(const vec2& uv) const {
        if (!diffuse_tex) return TGAColor(255, 255, 255, 255);
        int x, y;
        texel_of(*diffuse_tex, uv, x, y);
        return diffuse_tex->get(x, y);
    }
    // model.cpp:428-444
    vec3 normal(const vec2& uv) const {
        if (!normal_tex) return vec3{0, 0, 1};
        int x, y;
        texel_of(*normal_tex, uv, x, y);
        TGAColor c = normal_tex->get(x, y);
        vec3 n;
        n.x = (double)c[2] / 255.0 * 2.0 - 1.0;
        n.y = (double)c[1] / 255.0 * 2.0 - 1.0;
        n.z = (double)c[0] / 255.0 * 2.0 - 1.0;
        return normalized(n);
    }
    // model.cpp:446-459
    float specular(const vec2& uv) const {
        if (!specular_tex) return 1.0f;
        int x, y;
        texel_of(*specular_tex, uv, x, y);
        TGAColor c = specular_tex->get(x, y);
        return c[0] / 255.0f;
    }
};

// ---------------------------------------------------------------------------------------
// Shaders.  Vertex stage shared by Phong and Eye (main.cpp:71-90 == 199-218).
// ---------------------------------------------------------------------------------------
struct LitShaderBase : public IShader {
    const StandInModel* model = nullptr;
    vec2 varying_uv[3];
    vec3 varying_position_eye[3];
    vec3 varying_normal_eye[3];

    vec4 vertex(int face, int nth) override {
        vec3 p = model->vert(face, nth);
        vec3 n = model->normal(face, nth);
        varying_uv[nth] = model->uv(face, nth);
        vec4 pe = ModelView * make_vec4(p[0], p[1], p[2], 1.0);
        varying_position_eye[nth] = pe.xyz();
        vec4 ne = ModelView * make_vec4(n[0], n[1], n[2], 0.0);
        varying_normal_eye[nth] = ne.xyz();
        return Perspective * pe;
    }
    template <class V>
    static V mix3(const V* a, const vec3& b) {
        return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    }
};

struct PhongRestated : public LitShaderBase {
    vec3 key, fill, rim;
    double normal_map_strength = 1.0;

    // main.cpp:92-170
    std::pair<bool, TGAColor> fragment(const vec3 bar) const override {
        vec3 pos = mix3(varying_position_eye, bar);
        vec3 gn = mix3(varying_normal_eye, bar);
        vec2 uv = mix3(varying_uv, bar);

        TGAColor base = model->diffuse(uv);
        double spec_pow = std::max(1.0, (double)model->specular(uv));
        double brightness = (base[0] + base[1] + base[2]) / (3.0 * 255.0);
        bool eye_px = (brightness >= 0.85) && (spec_pow <= 5.0);

        vec3 nm = model->normal(uv);
        vec3 nm_eye = (ModelView * make_vec4(nm[0], nm[1], nm[2], 0.0)).xyz();
        vec3 N = eye_px ? gn
                        : normalized(gn * (1.0 - normal_map_strength) + nm_eye * normal_map_strength);
        vec3 V = normalized(-pos);

        double key_d = std::max(0.0, dot(N, key)) * 1.0;
        vec3 R = normalized(N * (2.0 * dot(N, key)) - key);
        double rv = std::max(0.0, dot(R, V));
        double key_s = (rv > 0.0 ? std::pow(rv, spec_pow) : 0.0) * 1.0;
        double fill_d = std::max(0.0, dot(N, fill)) * 0.35;
        double rim_d = std::max(0.0, dot(N, rim)) * 0.6;
        double diff = key_d + fill_d + rim_d;
        double ambient = 0.10;

        TGAColor out = base;
        for (int ch = 0; ch < 3; ++ch) {
            double cv = base[ch];
            double v = cv * (ambient + diff) + 255.0 * (0.35 * key_s);
            out[ch] = (unsigned char)std::min(255.0, v);
        }
        return {false, out};
    }
};

struct EyeRestated : public LitShaderBase {
    vec3 key, rim;

    // main.cpp:220-261
    std::pair<bool, TGAColor> fragment(const vec3 bar) const override {
        vec3 pos = mix3(varying_position_eye, bar);
        vec3 N = normalized(mix3(varying_normal_eye, bar));
        vec2 uv = mix3(varying_uv, bar);

        TGAColor base = model->diffuse(uv);
        vec3 V = normalized(-pos);
        double key_d = std::max(0.0, dot(N, key)) * 1.0;
        double rim_d = std::max(0.0, dot(N, rim)) * 0.6;
        double diff = key_d + rim_d;
        double spec_pow = std::max(1.0, (double)model->specular(uv)) * 8.0;
        vec3 R = normalized(N * (2.0 * dot(N, key)) - key);
        double rv = std::max(0.0, dot(R, V));
        double spec = (rv > 0.0 ? std::pow(rv, spec_pow) : 0.0);

        TGAColor out = base;
        for (int ch = 0; ch < 3; ++ch) {
            double cv = base[ch];
            double v = cv * (0.1 + diff) + 255.0 * (1.5 * spec);
            out[ch] = (unsigned char)std::min(255.0, v);
        }
        return {false, out};
    }
};

// ---- config 2 shaders.  The fork has no shadow-mapping or Gouraud shader, so they are
// AUTHORED here, once, as IShader subclasses that the reference's own rasterize() runs; the port
// oracle and the CUDA backend restate exactly this arithmetic.
// SHADOW_PHONG = PhongShader whose diffuse and specular terms are scaled by `darkening` when the
// fragment lies behind the depth pass' z-buffer; the light-space clip position is a varying because
// geometry.h has no matrix inverse.
struct ShadowPhongAuthored : public PhongRestated {
    mat<4, 4> light_mv, light_pr, light_vp;
    double bias = 0, darkening = 1;
    const std::vector<double>* shadow = nullptr;
    int sw = 0, sh = 0;
    vec4 varying_light_clip[3];

    vec4 vertex(int face, int nth) override {
        vec3 p = model->vert(face, nth);
        varying_light_clip[nth] = light_pr * (light_mv * make_vec4(p[0], p[1], p[2], 1.0));
        return PhongRestated::vertex(face, nth);
    }
    double shadow_factor(const vec3& bar) const {
        vec4 c = varying_light_clip[0] * bar[0] + varying_light_clip[1] * bar[1] + varying_light_clip[2] * bar[2];
        if (!(c[3] > 1e-12)) return 1.0;
        vec4 ndc = c / c[3];
        vec4 s = light_vp * ndc;
        if (!(s[0] >= 0.0 && s[1] >= 0.0)) return 1.0;
        int ix = int(s[0]), iy = int(s[1]);
        if (ix < 0 || iy < 0 || ix >= sw || iy >= sh) return 1.0;
        double zs = (*shadow)[(size_t)ix + (size_t)iy * sw];
        return (ndc[2] > zs + bias) ? darkening : 1.0;
    }
    std::pair<bool, TGAColor> fragment(const vec3 bar) const override {
        vec3 pos = mix3(varying_position_eye, bar);
        vec3 gn = mix3(varying_normal_eye, bar);
        vec2 uv = mix3(varying_uv, bar);
        TGAColor base = model->diffuse(uv);
        double spec_pow = std::max(1.0, (double)model->specular(uv));
        double brightness = (base[0] + base[1] + base[2]) / (3.0 * 255.0);
        bool eye_px = (brightness >= 0.85) && (spec_pow <= 5.0);
        vec3 nm = model->normal(uv);
        vec3 nm_eye = (ModelView * make_vec4(nm[0], nm[1], nm[2], 0.0)).xyz();
        vec3 N = eye_px ? gn : normalized(gn * (1.0 - normal_map_strength) + nm_eye * normal_map_strength);
        vec3 V = normalized(-pos);
        double key_d = std::max(0.0, dot(N, key)) * 1.0;
        vec3 R = normalized(N * (2.0 * dot(N, key)) - key);
        double rv = std::max(0.0, dot(R, V));
        double key_s = (rv > 0.0 ? std::pow(rv, spec_pow) : 0.0) * 1.0;
        double fill_d = std::max(0.0, dot(N, fill)) * 0.35;
        double rim_d = std::max(0.0, dot(N, rim)) * 0.6;
        double diff = key_d + fill_d + rim_d;
        double sf = shadow_factor(bar);
        TGAColor out = base;
        for (int ch = 0; ch < 3; ++ch) {
            double cv = base[ch];
            double v = cv * (0.10 + diff * sf) + 255.0 * ((0.35 * key_s) * sf);
            out[ch] = (unsigned char)std::min(255.0, v);
        }
        return {false, out};
    }
};

// GOURAUD: intensity max(0, normalized(normal_eye) . key) per vertex, interpolated; colour = diffuse * (0.1 + I)
struct GouraudAuthored : public LitShaderBase {
    vec3 key;
    double varying_intensity[3];
    vec4 vertex(int face, int nth) override {
        vec4 clip = LitShaderBase::vertex(face, nth);
        varying_intensity[nth] = std::max(0.0, dot(normalized(varying_normal_eye[nth]), key));
        return clip;
    }
    std::pair<bool, TGAColor> fragment(const vec3 bar) const override {
        double I = varying_intensity[0] * bar[0] + varying_intensity[1] * bar[1] + varying_intensity[2] * bar[2];
        vec2 uv = mix3(varying_uv, bar);
        TGAColor base = model->diffuse(uv);
        TGAColor out = base;
        for (int ch = 0; ch < 3; ++ch) out[ch] = (unsigned char)std::min(255.0, (double)base[ch] * (0.1 + I));
        return {false, out};
    }
};

// test shader of SURVEY K1-K7 / config 5: channel i = 255 * perspective-correct bary i
struct FlatBaryShader : public IShader {
    vec4 clip[3];
    vec4 vertex(int, int nth) override { return clip[nth]; }
    std::pair<bool, TGAColor> fragment(const vec3 bar) const override {
        TGAColor c(0, 0, 0, 255);
        for (int i = 0; i < 3; ++i) {
            double v = 255.0 * bar[i];
            v = (v > 0.0) ? v : 0.0;
            v = (v < 255.0) ? v : 255.0;
            c[i] = (unsigned char)v;
        }
        return {false, c};
    }
};

// mesh-driven variant: same colour rule, vertex stage = Perspective*ModelView*p
struct FlatBaryMeshShader : public FlatBaryShader {
    const StandInModel* model = nullptr;
    vec4 vertex(int face, int nth) override {
        vec3 p = model->vert(face, nth);
        vec4 pe = ModelView * make_vec4(p[0], p[1], p[2], 1.0);
        return Perspective * pe;
    }
};

// ---------------------------------------------------------------------------------------
struct View {
    std::vector<double> depth;           // swapped with the global zbuffer while drawing
    std::vector<double> depth_snapshot;
    TGAImage color;
    uint64_t covered_px = 0;
};

struct RefStats {
    uint64_t triangles = 0, fragments = 0;
    int bx0 = 0, by0 = 0, bx1 = 0, by1 = 0;
};

RefStats parse_reference_stats() {
    // print_render_stats (our_gl.cpp:204-210) is the only window on the static counters
    std::ostringstream cap;
    std::streambuf* old = std::cerr.rdbuf(cap.rdbuf());
    print_render_stats();
    std::cerr.rdbuf(old);
    RefStats s;
    unsigned long long t = 0, f = 0;
    std::string str = cap.str();
    std::sscanf(str.c_str(), "DEBUG: triangles=%llu fragments_drawn=%llu bbox=[%d,%d] - [%d,%d]",
                &t, &f, &s.bx0, &s.by0, &s.bx1, &s.by1);
    s.triangles = t;
    s.fragments = f;
    return s;
}

}  // namespace

struct TrbCtx {
    std::string err;
    int w = 0, h = 0, nviews = 0;
    uint8_t clear[3] = {0, 0, 0};
    mat<4, 4> viewport_m = mat<4, 4>::identity();
    std::vector<View> views;
    std::vector<std::unique_ptr<MeshData>> meshes;
    std::vector<std::unique_ptr<TGAImage>> textures;
    RefStats at_begin;
    TGAImage scratch;  // colour sink of DEPTH draws
    struct ShadowMapCopy { std::vector<double> z; int w, h; };
    std::vector<ShadowMapCopy> shadow_maps;
};

namespace {

mat<4, 4> load_mat(const double* m) {
    mat<4, 4> r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r[i][j] = m[i * 4 + j];
    return r;
}
void store_mat(const mat<4, 4>& m, double* out) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out[i * 4 + j] = m[i][j];
}
int fail(TrbCtx* c, int code, const char* msg) {
    if (c) c->err = msg;
    return code;
}
const TGAImage* tex_of(TrbCtx* c, TrbTex t) {
    if (t == 0 || t > c->textures.size()) return nullptr;
    return c->textures[t - 1].get();
}
vec3 v3(const double* p) { return vec3{p[0], p[1], p[2]}; }

// the per-face loop of main.cpp:660-666 on one view
template <class Shader>
void draw_faces(Shader& sh, View& view, TGAImage& target, uint64_t first, uint64_t n) {
    std::swap(zbuffer, view.depth);
    for (uint64_t f = first; f < first + n; ++f) {
        vec4 clip[3];
        for (int v = 0; v < 3; ++v) clip[v] = sh.vertex((int)f, v);
        rasterize(clip, sh, target);
    }
    std::swap(zbuffer, view.depth);
}

}  // namespace

extern "C" {

int orc_create(int, TrbCtx** out) {
    if (!out) return TRB_E_ARG;
    *out = new TrbCtx();
    return TRB_OK;
}
int orc_destroy(TrbCtx* c) {
    delete c;
    return TRB_OK;
}
const char* orc_last_error(TrbCtx* c) { return c ? c->err.c_str() : "null context"; }
const char* orc_backend_name(void) { return "oracle-ref"; }

int orc_upload_mesh(TrbCtx* c, const float* pos3, const float* nrm3, const float* uv2,
                    uint32_t nverts, const uint32_t* idx, uint64_t nidx, TrbMesh* out) {
    if (!c || !pos3 || !out || nidx % 3) return fail(c, TRB_E_ARG, "upload_mesh: bad argument");
    auto m = std::make_unique<MeshData>();
    m->position.resize(nverts);
    m->normal.resize(nverts);
    m->texcoord.resize(nverts);
    for (uint32_t i = 0; i < nverts; ++i) {
        // Assimp floats widened to double, model.cpp:160-185
        m->position[i] = vec3{pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]};
        m->normal[i] = nrm3 ? vec3{nrm3[3 * i], nrm3[3 * i + 1], nrm3[3 * i + 2]} : vec3{0, 0, 1};
        m->texcoord[i] = uv2 ? vec2{uv2[2 * i], uv2[2 * i + 1]} : vec2{0, 0};
    }
    m->indices.resize(nidx);
    for (uint64_t i = 0; i < nidx; ++i) {
        unsigned int k = idx ? idx[i] : (unsigned int)i;
        if (k >= nverts) return fail(c, TRB_E_ARG, "upload_mesh: index out of range");
        m->indices[i] = k;
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
    auto t = std::make_unique<TGAImage>(w, h, bpp);
    std::memcpy(t->buffer(), texels, (size_t)w * h * bpp);
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
    c->views.clear();
    c->views.resize(nviews);
    for (auto& v : c->views) {
        init_zbuffer(w, h);  // our_gl.cpp:72-74 fills the global; move it into the view
        v.depth.swap(zbuffer);
        v.color = TGAImage(w, h, TGAImage::RGB, TGAColor(c->clear[2], c->clear[1], c->clear[0]));
    }
    c->scratch = TGAImage(w, h, TGAImage::RGB);
    c->at_begin = parse_reference_stats();
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
    c->viewport_m = load_mat(v);
    return TRB_OK;
}

int orc_draw_batch(TrbCtx* c, TrbMesh mesh, const double* mv, const double* pr, int kind,
                   const void* uniforms, size_t ubytes, uint64_t first, uint64_t ntris) {
    if (!c || c->views.empty()) return fail(c, TRB_E_ARG, "draw: no frame");
    if (mesh == 0 || mesh > c->meshes.size() || !c->meshes[mesh - 1]) return fail(c, TRB_E_ARG, "draw: bad mesh");
    const MeshData* md = c->meshes[mesh - 1].get();
    if ((first + ntris) * 3 > md->indices.size()) return fail(c, TRB_E_ARG, "draw: triangle range");
    for (int vi = 0; vi < c->nviews; ++vi) {
        View& view = c->views[vi];
        ModelView = load_mat(mv + 16 * vi);
        Perspective = load_mat(pr + 16 * vi);
        Viewport = c->viewport_m;
        StandInModel model;
        model.mesh = md;
        if (kind == TRB_SHADER_PHONG || kind == TRB_SHADER_EYE) {
            if (!uniforms || ubytes != sizeof(TrbPhongUniforms)) return fail(c, TRB_E_ARG, "draw: uniforms");
            const TrbPhongUniforms& u = ((const TrbPhongUniforms*)uniforms)[vi];
            model.diffuse_tex = tex_of(c, u.diffuse);
            model.normal_tex = tex_of(c, u.normal);
            model.specular_tex = tex_of(c, u.specular);
            if (kind == TRB_SHADER_PHONG) {
                PhongRestated sh;
                sh.model = &model;
                sh.key = v3(u.key_dir_eye);
                sh.fill = v3(u.fill_dir_eye);
                sh.rim = v3(u.rim_dir_eye);
                sh.normal_map_strength = u.normal_map_strength;
                draw_faces(sh, view, view.color, first, ntris);
            } else {
                EyeRestated sh;
                sh.model = &model;
                sh.key = v3(u.key_dir_eye);
                sh.rim = v3(u.rim_dir_eye);
                draw_faces(sh, view, view.color, first, ntris);
            }
        } else if (kind == TRB_SHADER_FLAT_BARY || kind == TRB_SHADER_DEPTH) {
            FlatBaryMeshShader sh;
            sh.model = &model;
            draw_faces(sh, view, kind == TRB_SHADER_DEPTH ? c->scratch : view.color, first, ntris);
        } else if (kind == TRB_SHADER_GOURAUD) {
            if (!uniforms || ubytes != sizeof(TrbPhongUniforms)) return fail(c, TRB_E_ARG, "draw: uniforms");
            const TrbPhongUniforms& u = ((const TrbPhongUniforms*)uniforms)[vi];
            model.diffuse_tex = tex_of(c, u.diffuse);
            GouraudAuthored sh;
            sh.model = &model;
            sh.key = v3(u.key_dir_eye);
            draw_faces(sh, view, view.color, first, ntris);
        } else if (kind == TRB_SHADER_SHADOW_PHONG) {
            if (!uniforms || ubytes != sizeof(TrbShadowUniforms)) return fail(c, TRB_E_ARG, "draw: uniforms");
            const TrbShadowUniforms& su = ((const TrbShadowUniforms*)uniforms)[vi];
            if (su.shadow_map < 0 || (size_t)su.shadow_map >= c->shadow_maps.size()) return fail(c, TRB_E_ARG, "draw: shadow map");
            const auto& sm = c->shadow_maps[su.shadow_map];
            if (sm.w != su.shadow_w || sm.h != su.shadow_h) return fail(c, TRB_E_ARG, "draw: shadow map size");
            model.diffuse_tex = tex_of(c, su.phong.diffuse);
            model.normal_tex = tex_of(c, su.phong.normal);
            model.specular_tex = tex_of(c, su.phong.specular);
            ShadowPhongAuthored sh;
            sh.model = &model;
            sh.key = v3(su.phong.key_dir_eye);
            sh.fill = v3(su.phong.fill_dir_eye);
            sh.rim = v3(su.phong.rim_dir_eye);
            sh.normal_map_strength = su.phong.normal_map_strength;
            sh.light_mv = load_mat(su.light_modelview);
            sh.light_pr = load_mat(su.light_perspective);
            sh.light_vp = load_mat(su.light_viewport);
            sh.bias = su.shadow_bias;
            sh.darkening = su.shadow_darkening;
            sh.shadow = &sm.z;
            sh.sw = sm.w;
            sh.sh = sm.h;
            draw_faces(sh, view, view.color, first, ntris);
        } else {
            return fail(c, TRB_E_SHADER, "draw: shader kind not available in oracle-ref");
        }
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
    View& view = c->views[0];
    if (mv) ModelView = load_mat(mv);
    Viewport = c->viewport_m;
    std::swap(zbuffer, view.depth);
    int rc = TRB_OK;
    if (kind == TRB_SHADER_FLAT_BARY || kind == TRB_SHADER_DEPTH) {
        FlatBaryShader sh;
        TGAImage& target = kind == TRB_SHADER_DEPTH ? c->scratch : view.color;
        for (uint64_t t = 0; t < n; ++t) {
            vec4 clip[3];
            for (int v = 0; v < 3; ++v)
                clip[v] = make_vec4(clip12[t * 12 + v * 4], clip12[t * 12 + v * 4 + 1],
                                    clip12[t * 12 + v * 4 + 2], clip12[t * 12 + v * 4 + 3]);
            rasterize(clip, sh, target);
        }
    } else if ((kind == TRB_SHADER_PHONG || kind == TRB_SHADER_EYE) && varyings && uniforms &&
               ubytes == sizeof(TrbPhongUniforms)) {
        const TrbPhongUniforms& u = *(const TrbPhongUniforms*)uniforms;
        StandInModel model;
        MeshData empty;
        model.mesh = &empty;
        model.diffuse_tex = tex_of(c, u.diffuse);
        model.normal_tex = tex_of(c, u.normal);
        model.specular_tex = tex_of(c, u.specular);
        PhongRestated ph;
        EyeRestated ey;
        ph.model = ey.model = &model;
        ph.key = ey.key = v3(u.key_dir_eye);
        ph.fill = v3(u.fill_dir_eye);
        ph.rim = ey.rim = v3(u.rim_dir_eye);
        ph.normal_map_strength = u.normal_map_strength;
        LitShaderBase& sh = kind == TRB_SHADER_PHONG ? (LitShaderBase&)ph : (LitShaderBase&)ey;
        for (uint64_t t = 0; t < n; ++t) {
            vec4 clip[3];
            for (int v = 0; v < 3; ++v) {
                const double* cv = clip12 + t * 12 + v * 4;
                const double* vr = varyings + t * 24 + v * 8;
                clip[v] = make_vec4(cv[0], cv[1], cv[2], cv[3]);
                sh.varying_uv[v] = vec2{vr[0], vr[1]};
                sh.varying_position_eye[v] = vec3{vr[2], vr[3], vr[4]};
                sh.varying_normal_eye[v] = vec3{vr[5], vr[6], vr[7]};
            }
            rasterize(clip, sh, view.color);
        }
    } else {
        rc = fail(c, TRB_E_SHADER, "submit: shader kind / varyings");
    }
    std::swap(zbuffer, view.depth);
    return rc;
}

int orc_depth_snapshot(TrbCtx* c) {
    if (!c || c->views.empty()) return fail(c, TRB_E_ARG, "depth_snapshot");
    for (auto& v : c->views) v.depth_snapshot = v.depth;  // main.cpp:700
    return TRB_OK;
}
int orc_depth_restore(TrbCtx* c) {
    if (!c || c->views.empty()) return fail(c, TRB_E_ARG, "depth_restore");
    for (auto& v : c->views) {
        if (v.depth_snapshot.size() != v.depth.size()) return fail(c, TRB_E_ARG, "depth_restore: no snapshot");
        v.depth = v.depth_snapshot;  // main.cpp:730
    }
    return TRB_OK;
}
int orc_keep_depth_as_shadow_map(TrbCtx* c, int32_t* out) {
    if (!c || c->views.empty() || !out) return fail(c, TRB_E_ARG, "keep_depth_as_shadow_map");
    c->shadow_maps.push_back(TrbCtx::ShadowMapCopy{c->views[0].depth, c->w, c->h});
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
    return c->views[view].depth.data();
}
static const uint8_t* orc_view_color(TrbCtx* c, int view) {
    if (!c || view < 0 || view >= c->nviews) return nullptr;
    return c->views[view].color.buffer();
}
#define ORC_HAVE_REFERENCE_TGA 1   // post_restate.inc: encode through the reference's own TGAImage writer
#include "post_restate.inc"

int orc_read_color(TrbCtx* c, int view, uint8_t* out) {
    if (!c || view < 0 || view >= c->nviews || !out) return fail(c, TRB_E_ARG, "read_color");
    std::memcpy(out, c->views[view].color.buffer(), (size_t)c->w * c->h * 3);
    return TRB_OK;
}
int orc_write_color(TrbCtx* c, int view, const uint8_t* bgr) {
    if (!c || view < 0 || view >= c->nviews || !bgr) return fail(c, TRB_E_ARG, "write_color");
    std::memcpy(c->views[view].color.buffer(), bgr, (size_t)c->w * c->h * 3);
    return TRB_OK;
}
int orc_read_depth(TrbCtx* c, int view, double* out) {
    if (!c || view < 0 || view >= c->nviews || !out) return fail(c, TRB_E_ARG, "read_depth");
    std::memcpy(out, c->views[view].depth.data(), sizeof(double) * c->w * c->h);
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
int orc_read_visibility(TrbCtx* c, int, uint32_t*) { return fail(c, TRB_E_SHADER, "not in oracle-ref"); }
int orc_get_stats(TrbCtx* c, int view, TrbStats* out) {
    if (!c || view < 0 || view >= c->nviews || !out) return fail(c, TRB_E_ARG, "get_stats");
    std::memset(out, 0, sizeof(*out));
    RefStats now = parse_reference_stats();
    // the reference's counters are process-wide statics: deltas since begin, over ALL views
    // every view is sent the same draws, so the per-view count is the total / nviews
    out->triangles_submitted = (now.triangles - c->at_begin.triangles) / (uint64_t)c->nviews;
    out->fragments_drawn_ref = now.fragments - c->at_begin.fragments;  // all views together
    out->bbox_min_x = now.bx0;  // cumulative since the library was loaded
    out->bbox_min_y = now.by0;
    out->bbox_max_x = now.bx1;
    out->bbox_max_y = now.by1;
    uint64_t px = 0;
    double zmin = std::numeric_limits<double>::infinity();
    for (double z : c->views[view].depth)
        if (std::isfinite(z)) {
            ++px;
            zmin = std::min(zmin, z);
        }
    out->pixels_shaded = px;
    out->z_min = zmin;
    out->z_max_covered = std::numeric_limits<double>::quiet_NaN();
    out->z_max_ref = std::numeric_limits<double>::quiet_NaN();
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
    const uint64_t total = c->meshes[mesh - 1]->indices.size() / 3, n = (uint64_t)shard_count, r = (uint64_t)shard_rank;
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

// ---- host helpers: straight through the reference's own functions ----------------------
void orc_light_dir_eye(const double* mv, const double* dir, double* out) {
    // main.cpp:59-68
    mat<3, 3> nm;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) nm[i][j] = mv[i * 4 + j];
    vec3 r = normalized(nm * v3(dir));
    out[0] = r.x;
    out[1] = r.y;
    out[2] = r.z;
}
void orc_lookat(const double* eye, const double* center, const double* up, double* out) {
    lookat(v3(eye), v3(center), v3(up));  // our_gl.cpp:25-41
    store_mat(ModelView, out);
}
void orc_perspective(double fov, double aspect, double zn, double zf, double* out) {
    init_perspective(fov, aspect, zn, zf);  // our_gl.cpp:44-56
    store_mat(Perspective, out);
}
void orc_viewport(int x, int y, int w, int h, double* out) {
    init_viewport(x, y, w, h);  // our_gl.cpp:59-69
    store_mat(Viewport, out);
}
void orc_mat4_mul(const double* a, const double* b, double* out) {
    store_mat(load_mat(a) * load_mat(b), out);  // geometry.h:195-205
}

// the reference's own Frustum / AABB (our_gl.cpp:212-280, geometry.h:272-328)
void orc_frustum_planes(const double* m, double* planes) {
    Frustum f = Frustum::createFromMatrix(load_mat(m));
    for (int p = 0; p < 6; ++p) {
        planes[4 * p] = f.planes[p].normal.x; planes[4 * p + 1] = f.planes[p].normal.y;
        planes[4 * p + 2] = f.planes[p].normal.z; planes[4 * p + 3] = f.planes[p].d;
    }
}
int orc_frustum_intersects(const double* planes, const double* lo, const double* hi) {
    Frustum f;
    for (int p = 0; p < 6; ++p) {
        f.planes[p].normal = vec3{planes[4 * p], planes[4 * p + 1], planes[4 * p + 2]};
        f.planes[p].d = planes[4 * p + 3];
    }
    return f.intersects(AABB(v3(lo), v3(hi))) ? 1 : 0;
}
void orc_aabb_transform(const double* lo, const double* hi, const double* m, double* out_lo, double* out_hi) {
    AABB b = AABB(v3(lo), v3(hi)).transform(load_mat(m));
    for (int k = 0; k < 3; ++k) { out_lo[k] = b.min[k]; out_hi[k] = b.max[k]; }
}
void orc_cull_batch(const double* perspective, const double* views, int n, const double* lo, const double* hi, uint8_t* out) {
    for (int v = 0; v < n; ++v) {
        Frustum f = Frustum::createFromMatrix(load_mat(perspective) * load_mat(views + 16 * v));   // main.cpp:623-624
        out[v] = f.intersects(AABB(v3(lo), v3(hi))) ? 1 : 0;
    }
}

void orc_mat4_mul_batch(const double* a, int n, const double* b, double* out) {
    for (int v = 0; v < n; ++v) orc_mat4_mul(a + 16 * v, b, out + 16 * v);
}
void orc_light_dir_eye_batch(const double* mvs, int n, const double* dir, double* out) {
    for (int v = 0; v < n; ++v) orc_light_dir_eye(mvs + 16 * v, dir, out + 3 * v);
}

}  // extern "C"
