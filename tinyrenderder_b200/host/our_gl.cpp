// our_gl.cpp - the reference's pipeline API (our_gl.cpp:12-280) over the B200 backend's C ABI.
// Host code only: matrices and batching; every fragment is produced by libtrb.so.
#include <our_gl.h>
#include <model.h>

#include "../../include/trb.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>

mat<4, 4> ModelView = mat<4, 4>::identity();
mat<4, 4> Perspective = mat<4, 4>::identity();
mat<4, 4> Viewport = mat<4, 4>::identity();
std::vector<double> zbuffer;

namespace {

struct Backend {
    TrbCtx* ctx = nullptr;
    int w = 0, h = 0;
    bool frame = false;
    std::vector<double> views;   // gl_begin_views: row-major view matrices of the batch (empty: single frame)
    // pending immediate-mode batch (consecutive rasterize() calls with the same shader state)
    std::vector<double> clip, vary;
    TrbDeviceShader key;
    double key_mv[16];
    double key_vp[16];           // Viewport at the time the batch was opened: the reference applies it per rasterize() call
    bool have_key = false;
    std::vector<uint8_t> color_tmp;
    std::vector<uint32_t> vis_tmp;
    // several GPUs (gl_set_devices): context k of the table
    int device = -1;             // -1: $TRB_DEVICE or 0
    size_t view0 = 0;            // gl_begin_views: first global frame index of this context's block
    struct Handles { std::uint64_t mesh = 0, diffuse = 0, normal = 0, specular = 0; };
    std::map<const Model*, Handles> models;   // contexts 1..n-1 keep their copies of the models here (context 0: Model::dev_*)
    // frame recordings of a views batch (gl_record_views_begin .. gl_replay_views): per recording the library's handle and,
    // per gl_draw_model_views call, what is needed to compute the call's matrices / uniforms for other cameras
    struct RecDraw { int kind; double mm[16], pr[16], kw[3], fw[3], rw[3], nms; Handles h; };
    struct Rec { std::uint64_t handle = 0; size_t nviews = 0; std::vector<RecDraw> draws; };
    std::vector<Rec> recs;
    bool recording = false;
    Rec pending;
};
std::vector<Backend>& BS() {
    static std::vector<Backend> v(1);
    return v;
}
Backend& B() { return BS()[0]; }     // the primary context: everything single-GPU goes through it
bool g_comm_ready = false;           // composite group of the contexts formed (same frame size since)
std::uint64_t g_next_id = 0;         // multi-GPU frames: submission index of the next triangle, the same on every context

[[noreturn]] void die(const char* what, int rc) {
    std::string msg = std::string("tinyrenderder-b200: ") + what + " failed (rc=" + std::to_string(rc) + "):";
    for (Backend& b : BS())
        if (b.ctx && trb_last_error(b.ctx)[0]) msg += std::string(" ") + trb_last_error(b.ctx);
    std::cerr << msg << std::endl;
    throw std::runtime_error(msg);
}
#define CK(call)                         \
    do {                                 \
        int rc_ = (call);                \
        if (rc_ != TRB_OK) die(#call, rc_); \
    } while (0)

void flat(const mat<4, 4>& m, double* out) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out[i * 4 + j] = m[i][j];
}

TrbCtx* context_of(Backend& b) {
    if (!b.ctx) {
        const char* e = std::getenv("TRB_DEVICE");
        int rc = trb_create(b.device >= 0 ? b.device : (e ? std::atoi(e) : 0), &b.ctx);
        if (rc != TRB_OK) {
            b.ctx = nullptr;
            die("trb_create (no sm_100 device, and there is no CPU fallback)", rc);
        }
    }
    return b.ctx;
}
Backend::Handles upload_to(Backend& b, const Model& m) {
    TrbCtx* c = context_of(b);
    const auto& V = m.getVertices();
    const auto& I = m.getIndices();
    std::vector<float> pos(V.size() * 3), nrm(V.size() * 3), uv(V.size() * 2);
    for (size_t i = 0; i < V.size(); ++i) {
        for (int k = 0; k < 3; ++k) {
            pos[3 * i + k] = (float)V[i].position[k];   // values originate as fp32 (model.cpp:160-185)
            nrm[3 * i + k] = (float)V[i].normal[k];
        }
        uv[2 * i] = (float)V[i].texcoord.x;
        uv[2 * i + 1] = (float)V[i].texcoord.y;
    }
    std::vector<uint32_t> idx(I.begin(), I.end());
    Backend::Handles h;
    TrbMesh mh = 0;
    CK(trb_upload_mesh(c, pos.data(), nrm.data(), uv.data(), (uint32_t)V.size(), idx.data(), idx.size(), &mh));
    h.mesh = mh;
    auto tex = [&](const TGAImage& t, std::uint64_t& out) {
        out = 0;
        if (t.width() <= 0) return;
        TrbTex th = 0;
        CK(trb_upload_texture(c, t.buffer(), t.width(), t.height(), t.bytespp(), &th));
        out = th;
    };
    if (m.getMaterialCount() > 0) {  // only materials[0] is ever sampled (model.cpp:416-425)
        tex(m.getMaterial(0).diffuse, h.diffuse);
        tex(m.getMaterial(0).normal, h.normal);
        tex(m.getMaterial(0).specular, h.specular);
    }
    return h;
}
// the model's buffers in context `slot`, uploaded on first use
Backend::Handles handles_of(size_t slot, const Model& m) {
    Backend& b = BS()[slot];
    if (slot == 0) {
        if (!m.dev_uploaded) {
            Backend::Handles h = upload_to(b, m);
            m.dev_mesh = h.mesh; m.dev_diffuse = h.diffuse; m.dev_normal = h.normal; m.dev_specular = h.specular;
            m.dev_uploaded = true;
        }
        Backend::Handles h;
        h.mesh = m.dev_mesh; h.diffuse = m.dev_diffuse; h.normal = m.dev_normal; h.specular = m.dev_specular;
        return h;
    }
    auto it = b.models.find(&m);
    if (it == b.models.end()) it = b.models.emplace(&m, upload_to(b, m)).first;
    return it->second;
}
void upload_model(const Model& m) { (void)handles_of(0, m); }

void fill_uniforms(const TrbDeviceShader& d, TrbPhongUniforms& u) {
    std::memset(&u, 0, sizeof(u));
    for (int i = 0; i < 3; ++i) {
        u.key_dir_eye[i] = d.key_dir_eye[i];
        u.fill_dir_eye[i] = d.fill_dir_eye[i];
        u.rim_dir_eye[i] = d.rim_dir_eye[i];
    }
    u.normal_map_strength = d.normal_map_strength;
    if (d.model) {
        upload_model(*d.model);
        u.diffuse = d.model->dev_diffuse;
        u.normal = d.model->dev_normal;
        u.specular = d.model->dev_specular;
    }
}

void require_frame() {
    if (!B().frame) throw std::runtime_error("tinyrenderder-b200: call init_zbuffer(width, height) before drawing");
}

void submit_pending() {
    Backend& b = B();
    if (b.clip.empty()) return;
    TrbPhongUniforms u;
    fill_uniforms(b.key, u);
    CK(trb_set_viewport(b.ctx, b.key_vp));   // the viewport the queued triangles were submitted under
    const bool lit = b.key.kind == TRB_SHADER_PHONG || b.key.kind == TRB_SHADER_EYE;
    CK(trb_submit_clip_triangles(b.ctx, b.clip.data(), lit ? b.vary.data() : nullptr, b.clip.size() / 12, b.key_mv,
                                 b.key.kind, lit ? &u : nullptr, lit ? sizeof(u) : 0));
    b.clip.clear();
    b.vary.clear();
    b.have_key = false;
}

bool same_state(const TrbDeviceShader& a, const TrbDeviceShader& c, const double* mv_a, const double* mv_c,
                const double* vp_a, const double* vp_c) {
    return !std::memcmp(vp_a, vp_c, 128) && a.kind == c.kind && a.model == c.model && a.normal_map_strength == c.normal_map_strength &&
           !std::memcmp(a.key_dir_eye, c.key_dir_eye, 24) && !std::memcmp(a.fill_dir_eye, c.fill_dir_eye, 24) &&
           !std::memcmp(a.rim_dir_eye, c.rim_dir_eye, 24) && !std::memcmp(mv_a, mv_c, 128);
}

}  // namespace

void trb_host_release_model(const Model& m) {
    for (size_t k = 1; k < BS().size(); ++k) {        // the copies the other contexts hold
        Backend& o = BS()[k];
        auto it = o.models.find(&m);
        if (it == o.models.end() || !o.ctx) continue;
        if (it->second.mesh) trb_free_mesh(o.ctx, it->second.mesh);
        if (it->second.diffuse) trb_free_texture(o.ctx, it->second.diffuse);
        if (it->second.normal) trb_free_texture(o.ctx, it->second.normal);
        if (it->second.specular) trb_free_texture(o.ctx, it->second.specular);
        o.models.erase(it);
    }
    Backend& b = B();
    if (!b.ctx || !m.dev_uploaded) return;
    if (m.dev_mesh) trb_free_mesh(b.ctx, m.dev_mesh);
    if (m.dev_diffuse) trb_free_texture(b.ctx, m.dev_diffuse);
    if (m.dev_normal) trb_free_texture(b.ctx, m.dev_normal);
    if (m.dev_specular) trb_free_texture(b.ctx, m.dev_specular);
    m.dev_mesh = m.dev_diffuse = m.dev_normal = m.dev_specular = 0;
    m.dev_uploaded = false;
}

TrbCtx* gl_context() { return context_of(B()); }

// ---- several GPUs from one C++ program ------------------------------------------------------------
void gl_set_devices(const std::vector<int>& devices) {
    if (devices.empty()) throw std::runtime_error("tinyrenderder-b200: gl_set_devices needs at least one device");
    for (Backend& b : BS())
        if (b.ctx) throw std::runtime_error("tinyrenderder-b200: gl_set_devices must be called before the first frame");
    BS().assign(devices.size(), Backend());
    for (size_t k = 0; k < devices.size(); ++k) BS()[k].device = devices[k];
    g_comm_ready = false;
}
int gl_device_count() { return (int)BS().size(); }

namespace {
void shard(size_t total, size_t k, size_t n, size_t& first, size_t& count) {   // contiguous blocks, the first ones take the remainder
    const size_t base = total / n, rem = total % n;
    first = k * base + std::min(k, rem);
    count = base + (k < rem ? 1 : 0);
}
}  // namespace

// ---- setup functions: host math through the backend's reference-order helpers ---------------
void lookat(const vec3 eye, const vec3 center, const vec3 up) {
    double e[3] = {eye.x, eye.y, eye.z}, c[3] = {center.x, center.y, center.z}, u[3] = {up.x, up.y, up.z}, m[16];
    trb_lookat(e, c, u, m);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) ModelView[i][j] = m[i * 4 + j];
}
void init_perspective(double fov_deg, double aspect, double znear, double zfar) {
    double m[16];
    trb_perspective(fov_deg, aspect, znear, zfar, m);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) Perspective[i][j] = m[i * 4 + j];
}
void init_viewport(int x, int y, int w, int h) {
    double m[16];
    trb_viewport(x, y, w, h, m);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) Viewport[i][j] = m[i * 4 + j];
}
void init_zbuffer(int width, int height) {
    zbuffer.assign((size_t)width * height, std::numeric_limits<double>::infinity());
    if (BS().size() > 1 && g_comm_ready && (B().w != width || B().h != height)) {   // the group is tied to a frame size
        for (Backend& o : BS()) CK(trb_comm_close(o.ctx));
        g_comm_ready = false;
    }
    for (Backend& b : BS()) {    // with several GPUs every context takes part in the picture (sharded draws + composite)
        context_of(b);
        b.clip.clear();
        b.vary.clear();
        b.have_key = false;
        CK(trb_begin_frame(b.ctx, width, height));
        b.w = width;
        b.h = height;
        b.frame = true;
        b.views.clear();
        b.view0 = 0;
    }
    g_next_id = 0;
}

// ---- drawing -------------------------------------------------------------------------------------
void rasterize(const Triangle& clip, const IShader& shader, TGAImage& framebuffer) {
    (void)framebuffer;
    require_frame();
    if (BS().size() > 1)
        throw std::runtime_error("tinyrenderder-b200: with several GPUs a picture is drawn model by model (gl_draw_model "
                                 "shards the triangles); rasterize() one triangle at a time is a single-GPU path");
    Backend& b = B();
    TrbDeviceShader d;
    if (!shader.device_shader(d))
        throw std::runtime_error("tinyrenderder-b200: rasterize() got a shader without a device implementation "
                                 "(IShader::device_shader); a GPU cannot call a host fragment() and there is no CPU fallback");
    double mv[16], vp[16];
    flat(ModelView, mv);
    flat(Viewport, vp);
    if (b.have_key && !same_state(b.key, d, b.key_mv, mv, b.key_vp, vp)) submit_pending();
    if (!b.have_key) {
        b.key = d;
        std::memcpy(b.key_mv, mv, sizeof(mv));
        std::memcpy(b.key_vp, vp, sizeof(vp));
        b.have_key = true;
    }
    for (int v = 0; v < 3; ++v)
        for (int k = 0; k < 4; ++k) b.clip.push_back(clip[v][k]);
    if (d.varyings) b.vary.insert(b.vary.end(), d.varyings, d.varyings + 24);
    else b.vary.insert(b.vary.end(), 24, 0.0);
}

void triangle(const Triangle& clip_verts, const IShader& shader, TGAImage& image, std::vector<double>& zbuffer_arg) {
    if (&zbuffer_arg != &zbuffer)
        throw std::runtime_error("tinyrenderder-b200: triangle() depth-tests against the device-resident z-buffer behind the "
                                 "global `zbuffer`; pass that vector");
    rasterize(clip_verts, shader, image);
}

void gl_draw_model(const Model& model, const IShader& shader, TGAImage& framebuffer) {
    (void)framebuffer;
    require_frame();
    submit_pending();
    TrbDeviceShader d;
    if (!shader.device_shader(d))
        throw std::runtime_error("tinyrenderder-b200: gl_draw_model() got a shader without a device implementation");
    d.model = &model;
    double mv[16], pr[16], vp[16];
    flat(ModelView, mv);
    flat(Perspective, pr);
    flat(Viewport, vp);
    const bool lit = d.kind == TRB_SHADER_PHONG || d.kind == TRB_SHADER_EYE;
    const size_t n = BS().size(), ntris = (size_t)model.nfaces();
    if (n > 1 && !B().views.empty()) throw std::runtime_error("tinyrenderder-b200: gl_draw_model inside gl_begin_views; use gl_draw_model_views");
    // one GPU: the whole face loop is one draw.  Several GPUs: context k draws share k of the model with GLOBAL
    // submission indices, so that gl_composite resolves depth ties exactly like one sequential loop would
    for (size_t k = 0; k < n; ++k) {
        Backend& b = BS()[k];
        const Backend::Handles h = handles_of(k, model);
        TrbPhongUniforms u;
        fill_uniforms(d, u);
        u.diffuse = h.diffuse; u.normal = h.normal; u.specular = h.specular;
        CK(trb_set_viewport(b.ctx, vp));
        if (n > 1) {
            // the backend deals the triangles out (trb_draw_shard); ids are g_next_id + face + 1 on every context
            CK(trb_set_triangle_id_base(b.ctx, g_next_id));
            CK(trb_draw_shard(b.ctx, h.mesh, mv, pr, d.kind, lit ? &u : nullptr, lit ? sizeof(u) : 0, (int)k, (int)n));
        } else {
            CK(trb_draw(b.ctx, h.mesh, mv, pr, d.kind, lit ? &u : nullptr, lit ? sizeof(u) : 0, 0, ntris));
        }
    }
    g_next_id += ntris;
}

// several GPUs, one picture: the sort-last composite of the contexts' depth / id planes (trb_comm_init once per frame
// size, then trb_composite_group: the exact (depth, submission index) minimum per pixel, every context shades the rows
// it owns), rows gathered into `framebuffer` and the global zbuffer
void gl_composite(TGAImage& framebuffer) {
    require_frame();
    const size_t n = BS().size();
    if (n < 2) { gl_flush(framebuffer); return; }
    std::vector<TrbCtx*> ctxs;
    for (Backend& b : BS()) ctxs.push_back(b.ctx);
    if (!g_comm_ready) {
        CK(trb_comm_init(ctxs.data(), (int)n));
        g_comm_ready = true;
    }
    CK(trb_composite_group(ctxs.data(), (int)n));
    const int w = B().w, h = B().h;
    framebuffer = TGAImage(w, h, TGAImage::RGB);
    zbuffer.resize((size_t)w * h);
    std::vector<uint8_t> col((size_t)w * h * 3);
    std::vector<double> z((size_t)w * h);
    for (size_t k = 0; k < n; ++k) {
        int y0 = 0, y1 = 0;
        CK(trb_comm_rows(ctxs[k], &y0, &y1));
        CK(trb_read_color(ctxs[k], 0, col.data()));
        CK(trb_read_depth(ctxs[k], 0, z.data()));
        std::memcpy(framebuffer.buffer() + (size_t)y0 * w * 3, col.data() + (size_t)y0 * w * 3, (size_t)(y1 - y0) * w * 3);
        std::memcpy(zbuffer.data() + (size_t)y0 * w, z.data() + (size_t)y0 * w, (size_t)(y1 - y0) * w * sizeof(double));
    }
}

void gl_flush(TGAImage& framebuffer) {
    require_frame();
    if (BS().size() > 1 && B().views.empty()) { gl_composite(framebuffer); return; }
    Backend& b = B();
    submit_pending();
    CK(trb_flush(b.ctx));
    const size_t n = (size_t)b.w * b.h;
    b.color_tmp.resize(n * 3);
    b.vis_tmp.resize(n);
    CK(trb_read_color(b.ctx, 0, b.color_tmp.data()));
    CK(trb_read_visibility(b.ctx, 0, b.vis_tmp.data()));
    // only pixels some fragment was drawn to are copied: the caller's framebuffer keeps whatever it
    // held elsewhere, exactly like framebuffer.set() per fragment (our_gl.cpp:192)
    if (framebuffer.width() == b.w && framebuffer.height() == b.h)
        for (int y = 0; y < b.h; ++y)
            for (int x = 0; x < b.w; ++x) {
                size_t p = (size_t)x + (size_t)y * b.w;
                if (b.vis_tmp[p] != 0u) continue;  // 0 == shaded at least once since begin
                TGAColor c = framebuffer.get(x, y);
                c[0] = b.color_tmp[3 * p];
                c[1] = b.color_tmp[3 * p + 1];
                c[2] = b.color_tmp[3 * p + 2];
                framebuffer.set(x, y, c);
            }
    zbuffer.resize(n);
    CK(trb_read_depth(b.ctx, 0, zbuffer.data()));
}

void gl_zbuffer_snapshot() {
    require_frame();
    submit_pending();
    if (BS().size() > 1 && B().views.empty())
        throw std::runtime_error("tinyrenderder-b200: z-buffer snapshots of a picture that is split over several GPUs are not supported");
    for (Backend& b : BS())
        if (b.frame) CK(trb_depth_snapshot(b.ctx));
}
void gl_zbuffer_restore(TGAImage& framebuffer) {
    require_frame();
    submit_pending();
    for (Backend& b : BS())
        if (b.frame) CK(trb_depth_restore(b.ctx));
    if (B().views.empty()) gl_flush(framebuffer);     // a batch of frames is read with gl_read_view / gl_write_tga_files
}

static void grey_to_image(const std::vector<uint8_t>& g, int w, int h, TGAImage& img) {
    img = TGAImage(w, h, TGAImage::RGB);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            uint8_t v = g[(size_t)x + (size_t)y * w];
            img.set(x, y, TGAColor(v, v, v));
        }
}
void gl_ssao(TGAImage& ao_map) {
    require_frame();
    submit_pending();
    Backend& b = B();
    std::vector<uint8_t> g((size_t)b.w * b.h);
    CK(trb_ssao(b.ctx, 0, g.data()));
    grey_to_image(g, b.w, b.h, ao_map);
}
void gl_zbuffer_image(TGAImage& grey) {
    require_frame();
    submit_pending();
    Backend& b = B();
    std::vector<uint8_t> g((size_t)b.w * b.h);
    CK(trb_depth_image(b.ctx, 0, g.data()));
    grey_to_image(g, b.w, b.h, grey);
}
void gl_composite_ao(TGAImage& final_result) {
    require_frame();
    submit_pending();
    Backend& b = B();
    std::vector<uint8_t> c((size_t)b.w * b.h * 3);
    CK(trb_composite_ao(b.ctx, 0, c.data()));
    final_result = TGAImage(b.w, b.h, TGAImage::RGB);
    std::memcpy(final_result.buffer(), c.data(), c.size());
}

namespace {
// what one gl_draw_model_views call passes to trb_draw_batch for the cameras `views` (row-major, n x 16)
void view_params(const std::vector<double>& views, const Backend::RecDraw& d, std::vector<double>& mvs, std::vector<double>& prs,
                 std::vector<TrbPhongUniforms>& u) {
    const int n = (int)(views.size() / 16);
    mvs.resize((size_t)n * 16);
    prs.resize((size_t)n * 16);
    trb_mat4_mul_batch(views.data(), n, d.mm, mvs.data());               // ModelView_i = view_i * model (main.cpp:653)
    for (int v = 0; v < n; ++v) std::memcpy(&prs[(size_t)v * 16], d.pr, sizeof(d.pr));
    std::vector<double> ke((size_t)n * 3), fe((size_t)n * 3), re((size_t)n * 3);
    trb_light_dir_eye_batch(mvs.data(), n, d.kw, ke.data());             // initLightDirections per view
    trb_light_dir_eye_batch(mvs.data(), n, d.fw, fe.data());
    trb_light_dir_eye_batch(mvs.data(), n, d.rw, re.data());
    u.resize((size_t)n);
    for (int v = 0; v < n; ++v) {
        std::memset(&u[v], 0, sizeof(TrbPhongUniforms));
        for (int i = 0; i < 3; ++i) {
            u[v].key_dir_eye[i] = ke[(size_t)v * 3 + i];
            u[v].fill_dir_eye[i] = fe[(size_t)v * 3 + i];
            u[v].rim_dir_eye[i] = re[(size_t)v * 3 + i];
        }
        u[v].normal_map_strength = d.nms;
        u[v].diffuse = d.h.diffuse;
        u[v].normal = d.h.normal;
        u[v].specular = d.h.specular;
    }
}
}  // namespace

void gl_begin_views(const std::vector<mat<4, 4>>& views, int width, int height) {
    if (views.empty()) throw std::runtime_error("tinyrenderder-b200: gl_begin_views needs at least one view");
    zbuffer.assign((size_t)width * height, std::numeric_limits<double>::infinity());
    if (g_comm_ready) {                            // a batch of frames is not the picture the composite group was formed for
        for (Backend& o : BS()) CK(trb_comm_close(o.ctx));
        g_comm_ready = false;
    }
    // frames are independent: with several GPUs context k takes block k of the cameras, nothing is
    // ever exchanged between the contexts
    const size_t n = BS().size();
    for (size_t k = 0; k < n; ++k) {
        Backend& b = BS()[k];
        size_t first = 0, count = views.size();
        shard(views.size(), k, n, first, count);
        b.clip.clear();
        b.vary.clear();
        b.have_key = false;
        b.views.clear();
        b.frame = false;
        b.view0 = first;
        if (count == 0) continue;
        context_of(b);
        CK(trb_begin_batch(b.ctx, width, height, (int)count));
        b.w = width;
        b.h = height;
        b.frame = true;
        b.views.resize(count * 16);
        for (size_t v = 0; v < count; ++v) flat(views[first + v], &b.views[v * 16]);
    }
}

void gl_draw_model_views(const Model& model, int kind, const mat<4, 4>& model_matrix, const vec3& key_world,
                         const vec3& fill_world, const vec3& rim_world, double normal_map_strength) {
    require_frame();
    if (B().views.empty()) throw std::runtime_error("tinyrenderder-b200: gl_draw_model_views needs gl_begin_views");
    if (kind != TRB_SHADER_PHONG && kind != TRB_SHADER_EYE)
        throw std::runtime_error("tinyrenderder-b200: gl_draw_model_views draws PhongShader (1) or EyeShader (2)");
    double mm[16], pr[16], vp[16];
    flat(model_matrix, mm);
    flat(Perspective, pr);
    flat(Viewport, vp);
    const double kw[3] = {key_world.x, key_world.y, key_world.z}, fw[3] = {fill_world.x, fill_world.y, fill_world.z},
                 rw[3] = {rim_world.x, rim_world.y, rim_world.z};
    for (size_t k = 0; k < BS().size(); ++k) {
        Backend& b = BS()[k];
        if (!b.frame || b.views.empty()) continue;
        const Backend::Handles h = handles_of(k, model);
        std::vector<double> mvs, prs;
        std::vector<TrbPhongUniforms> u;
        Backend::RecDraw rd;
        rd.kind = kind;
        std::memcpy(rd.mm, mm, sizeof(mm)); std::memcpy(rd.pr, pr, sizeof(pr));
        std::memcpy(rd.kw, kw, sizeof(kw)); std::memcpy(rd.fw, fw, sizeof(fw)); std::memcpy(rd.rw, rw, sizeof(rw));
        rd.nms = normal_map_strength;
        rd.h = h;
        view_params(b.views, rd, mvs, prs, u);
        if (b.recording) b.pending.draws.push_back(rd);
        CK(trb_set_viewport(b.ctx, vp));
        CK(trb_draw_batch(b.ctx, h.mesh, mvs.data(), prs.data(), kind, u.data(), sizeof(TrbPhongUniforms), 0,
                          (uint64_t)model.nfaces()));
    }
}

// ---- frame recordings of a views batch: trb_record_begin / trb_record_end / trb_replay on every context ------------
void gl_record_views_begin() {
    for (Backend& b : BS()) {
        if (b.recording) throw std::runtime_error("tinyrenderder-b200: gl_record_views_begin: already recording");
        context_of(b);
        CK(trb_record_begin(b.ctx));
        b.recording = true;
        b.pending = Backend::Rec();
    }
}
int gl_record_views_end() {
    int handle = 0;
    for (Backend& b : BS()) {
        if (!b.recording) throw std::runtime_error("tinyrenderder-b200: gl_record_views_end: not recording");
        b.recording = false;
        TrbRecording r = 0;
        CK(trb_record_end(b.ctx, &r));
        b.pending.handle = r;
        b.pending.nviews = b.views.size() / 16;
        b.recs.push_back(b.pending);
        handle = (int)b.recs.size();
    }
    return handle;
}
void gl_replay_views(int recording, const std::vector<mat<4, 4>>& views) {
    const size_t n = BS().size();
    for (size_t k = 0; k < n; ++k) {
        Backend& b = BS()[k];
        if (recording < 1 || (size_t)recording > b.recs.size()) throw std::runtime_error("tinyrenderder-b200: gl_replay_views: no such recording");
        const Backend::Rec& rec = b.recs[(size_t)recording - 1];
        size_t first = 0, count = views.size();
        shard(views.size(), k, n, first, count);
        if (count != rec.nviews) throw std::runtime_error("tinyrenderder-b200: gl_replay_views needs as many cameras as were recorded");
        if (count == 0) continue;
        b.views.resize(count * 16);
        b.view0 = first;
        for (size_t v = 0; v < count; ++v) flat(views[first + v], &b.views[v * 16]);
        std::vector<std::vector<double>> mvs(rec.draws.size()), prs(rec.draws.size());
        std::vector<std::vector<TrbPhongUniforms>> us(rec.draws.size());
        std::vector<TrbReplayDraw> draws(rec.draws.size());
        for (size_t i = 0; i < rec.draws.size(); ++i) {
            view_params(b.views, rec.draws[i], mvs[i], prs[i], us[i]);
            draws[i] = TrbReplayDraw{mvs[i].data(), prs[i].data(), us[i].data(), sizeof(TrbPhongUniforms)};
        }
        CK(trb_replay(b.ctx, rec.handle, draws.data(), (int)draws.size()));
        b.frame = true;
    }
}

namespace {
// the context that renders global frame `view` of the current batch, and the frame's index inside that context
Backend& owner_of_view(int view, int& local) {
    for (Backend& b : BS()) {
        const size_t n = b.views.size() / 16;
        if (b.frame && (size_t)view >= b.view0 && (size_t)view < b.view0 + n) {
            local = (int)((size_t)view - b.view0);
            return b;
        }
    }
    throw std::runtime_error("tinyrenderder-b200: no such frame in the current batch");
}
}  // namespace

void gl_read_view(int view, TGAImage& framebuffer) {
    require_frame();
    int local = 0;
    Backend& b = B().views.empty() ? B() : owner_of_view(view, local);
    if (B().views.empty()) local = view;
    const size_t n = (size_t)b.w * b.h;
    framebuffer = TGAImage(b.w, b.h, TGAImage::RGB);
    CK(trb_read_color(b.ctx, local, framebuffer.buffer()));
    zbuffer.resize(n);
    CK(trb_read_depth(b.ctx, local, zbuffer.data()));
}

bool gl_write_tga_files(int image, const std::vector<std::string>& filenames) {
    require_frame();
    submit_pending();           // triangles queued by rasterize() belong to the picture (as in gl_write_tga_file)
    size_t total = 0;
    for (Backend& b : BS())
        if (b.frame) total += b.views.empty() ? (&b == &B() ? 1 : 0) : b.views.size() / 16;
    if (filenames.size() != total) throw std::runtime_error("tinyrenderder-b200: gl_write_tga_files needs one name per frame");
    const size_t cap = (size_t)B().w * B().h * 3 + (size_t)B().w * B().h / 2 + 64;
    std::vector<std::vector<uint8_t>> files(total, std::vector<uint8_t>(cap));
    std::vector<uint64_t> sizes(total);
    // every context packetises its own frames on its own GPU (asynchronously: the GPUs work side by side), then one wait
    std::vector<std::vector<uint8_t*>> outs(BS().size());
    size_t at = 0;
    for (size_t k = 0; k < BS().size(); ++k) {
        Backend& b = BS()[k];
        const size_t nv = !b.frame ? 0 : (b.views.empty() ? (k == 0 ? 1 : 0) : b.views.size() / 16);
        if (nv == 0) continue;
        for (size_t v = 0; v < nv; ++v) outs[k].push_back(files[at + v].data());
        CK(trb_encode_tga_async(b.ctx, image, outs[k].data(), cap, &sizes[at]));
        at += nv;
    }
    for (Backend& b : BS())
        if (b.frame && b.ctx) CK(trb_readback_wait(b.ctx));
    bool ok = true;
    for (size_t v = 0; v < total; ++v) {
        std::ofstream f(filenames[v], std::ios::binary);
        if (!f.is_open()) {
            std::cerr << "can't open " << filenames[v] << "\n";
            ok = false;
            continue;
        }
        f.write((const char*)files[v].data(), (std::streamsize)sizes[v]);
        ok = ok && f.good();
    }
    return ok;
}

bool gl_write_tga_file(int image, const std::string& filename) {
    require_frame();
    submit_pending();
    Backend& b = B();
    if (b.views.size() > 16) throw std::runtime_error("tinyrenderder-b200: a batch of frames is written with gl_write_tga_files");
    const size_t cap = (size_t)b.w * b.h * 3 + (size_t)b.w * b.h / 2 + 64;
    std::vector<uint8_t> file(cap);
    uint8_t* out[1] = {file.data()};
    uint64_t size = 0;
    CK(trb_encode_tga(b.ctx, image, out, cap, &size));
    std::ofstream f(filename, std::ios::binary);
    if (!f.is_open()) {
        std::cerr << "can't open " << filename << "\n";   // same message as tgaimage.cpp:163
        return false;
    }
    f.write((const char*)file.data(), (std::streamsize)size);
    return f.good();
}

void print_render_stats() {
    Backend& b = B();
    if (!b.frame) {
        std::cerr << "DEBUG: triangles=0 (no frame)\n";
        return;
    }
    submit_pending();
    TrbStats s;
    CK(trb_get_stats(b.ctx, 0, &s));
    std::cerr << "DEBUG: triangles=" << s.triangles_submitted << " fragments_covered=" << s.fragments_covered
              << " pixels_shaded=" << s.pixels_shaded << " bbox=[" << s.bbox_min_x << "," << s.bbox_min_y << "] - ["
              << s.bbox_max_x << "," << s.bbox_max_y << "]"
              << " z-range=[" << (std::isfinite(s.z_min) ? std::to_string(s.z_min) : "inf") << ","
              << (std::isfinite(s.z_max_covered) ? std::to_string(s.z_max_covered) : "-inf") << "]\n";
}

// ---- frustum (host, bug-for-bug with our_gl.cpp:212-280) --------------------------------------------
Frustum Frustum::createFromMatrix(const mat<4, 4>& m) {
    Frustum f;
    // the reference reads COLUMN 3 +- column k of each row, i.e. planes of the transposed matrix
    const int col[6] = {0, 0, 1, 1, 2, 2};
    const double sgn[6] = {1, -1, 1, -1, 1, -1};
    for (int p = 0; p < 6; ++p) {
        f.planes[p].normal.x = m[0][3] + sgn[p] * m[0][col[p]];
        f.planes[p].normal.y = m[1][3] + sgn[p] * m[1][col[p]];
        f.planes[p].normal.z = m[2][3] + sgn[p] * m[2][col[p]];
        f.planes[p].d = m[3][3] + sgn[p] * m[3][col[p]];
        double len = norm(f.planes[p].normal);
        if (len > 0.0) {
            f.planes[p].normal = f.planes[p].normal / len;
            f.planes[p].d /= len;
        }
    }
    return f;
}
bool Frustum::intersects(const AABB& box) const {
    for (int i = 0; i < 6; ++i) {
        const Plane& pl = planes[i];
        vec3 far_corner = box.min;  // the corner furthest along the normal
        if (pl.normal.x >= 0) far_corner.x = box.max.x;
        if (pl.normal.y >= 0) far_corner.y = box.max.y;
        if (pl.normal.z >= 0) far_corner.z = box.max.z;
        if (pl.distance(far_corner) < 0) return false;
    }
    return true;
}
