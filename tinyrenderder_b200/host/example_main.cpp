// example_main.cpp - the scene section of the reference's main() (main.cpp:469-806) without its
// debug prints, written once and compiled twice:
//   device build  (-I tinyrenderder_b200/host, links libtrb.so): every fragment comes from the B200;
//   oracle build  (-I /root/reference + the reference's our_gl.cpp / tgaimage.cpp, -DEXAMPLE_ORACLE):
//                 the reference's own rasterize() on the CPU.  TEST INFRASTRUCTURE ONLY.
// tests/test_host_example.py renders the same OBJ/TGA assets with both and compares the outputs.
// usage: example <head.obj> <eyes.obj> <sponza.obj> <width> <height> <outdir> [--immediate]
//                [--eye x y z] [--target x y z]   (camera of main.cpp:587-589 by default; other cameras
//                exercise the model-level frustum cull of main.cpp:647, 680, 706)
#include <our_gl.h>
#include <model.h>
#include <model_manager.h>
#include <shaders.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>

#ifdef EXAMPLE_ORACLE
#include <example_oracle_glue.h>  // test-only: CPU post passes for the oracle build (tests/harness/ref_include)
#endif

static mat<4, 4> scale_matrix(double s) {  // main.cpp:365-371
    mat<4, 4> m = mat<4, 4>::identity();
    m[0][0] = m[1][1] = m[2][2] = s;
    return m;
}
static mat<4, 4> translation_matrix(double x, double y, double z) {  // main.cpp:374-380
    mat<4, 4> m = mat<4, 4>::identity();
    m[0][3] = x; m[1][3] = y; m[2][3] = z;
    return m;
}
static mat<4, 4> rotation_y_matrix(double a) {  // main.cpp:408-420
    mat<4, 4> m = mat<4, 4>::identity();
    double c = cos(a), s = sin(a);
    m[0][0] = c; m[0][2] = s; m[2][0] = -s; m[2][2] = c;
    return m;
}

static bool g_immediate = false;

// the per-face loop of main.cpp:660-666; on the device build this is ONE draw call unless --immediate
template <class Shader>
static void draw(const Model& model, Shader& shader, TGAImage& framebuffer) {
#ifdef TRB_DEVICE_BACKEND
    if (!g_immediate) { gl_draw_model(model, shader, framebuffer); return; }
#endif
    for (int face = 0; face < model.nfaces(); ++face) {
        vec4 clip[3];
        for (int v = 0; v < 3; ++v) clip[v] = shader.vertex(face, v);
#ifdef TRB_DEVICE_BACKEND
        triangle(clip, shader, framebuffer, zbuffer);   // upstream spelling of rasterize() (our_gl.h aliases)
#else
        rasterize(clip, shader, framebuffer);
#endif
    }
}

static void grey_image(const unsigned char* g, int w, int h, TGAImage& img) {
    img = TGAImage(w, h, TGAImage::RGB);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) img.set(x, y, TGAColor(g[x + y * w], g[x + y * w], g[x + y * w]));
}

int main(int argc, char** argv) {
    if (argc < 7) { std::cerr << "usage: example head.obj eyes.obj sponza.obj W H outdir [--immediate]\n"; return 2; }
    const int WIDTH = atoi(argv[4]), HEIGHT = atoi(argv[5]);
    const std::string out = argv[6];
    double eye_p[3] = {-3.4019, 2.2001, 1.8026}, target_p[3] = {1.3555, 1.5116, -0.9686};       // main.cpp:587-589
    for (int a = 7; a < argc; ++a) {
        if (!strcmp(argv[a], "--immediate")) g_immediate = true;
        else if (!strcmp(argv[a], "--eye") && a + 3 < argc) { for (int k = 0; k < 3; ++k) eye_p[k] = atof(argv[++a]); }
        else if (!strcmp(argv[a], "--target") && a + 3 < argc) { for (int k = 0; k < 3; ++k) target_p[k] = atof(argv[++a]); }
        else { std::cerr << "unknown argument " << argv[a] << "\n"; return 2; }
    }

    auto& mm = ModelManager::getInstance();
    auto head_model = mm.loadModel(argv[1]);
    auto eye_model = mm.loadModel(argv[2]);
    auto sponza_model = mm.loadModel(argv[3]);
    if (!head_model || !eye_model || !sponza_model) { std::cerr << "ERROR: Failed to load one or more models!\n"; return 1; }

    mat<4, 4> sponzaM = scale_matrix(0.014);                                                        // main.cpp:506-507
    mat<4, 4> headM = translation_matrix(0.0, 1.6815, 0.0) * rotation_y_matrix(-112.82 * M_PI / 180.0);  // :509-511
    AABB sponzaBox = sponza_model->getWorldAABB(sponzaM), headBox = head_model->getWorldAABB(headM);

    TGAImage framebuffer(WIDTH, HEIGHT, TGAImage::RGB);
    init_zbuffer(WIDTH, HEIGHT);                                                                   // main.cpp:606-612
    lookat(make_vec3(eye_p[0], eye_p[1], eye_p[2]), make_vec3(target_p[0], target_p[1], target_p[2]), make_vec3(0, 1, 0));
#ifdef TRB_DEVICE_BACKEND
    projection(70.0, (double)WIDTH / HEIGHT, 0.05, 500.0);     // upstream spellings of init_perspective / init_viewport
    viewport(0, 0, WIDTH, HEIGHT);
#else
    init_perspective(70.0, (double)WIDTH / HEIGHT, 0.05, 500.0);
    init_viewport(0, 0, WIDTH, HEIGHT);
#endif
    vec3 key = normalized(make_vec3(1.0, 1.4, 1.0)), fill = normalized(make_vec3(-0.3, 0.5, 0.2)),
         rim = normalized(make_vec3(-1.0, 0.8, -1.5));                                            // main.cpp:615-617
    Frustum frustum = Frustum::createFromMatrix(Perspective * ModelView);                          // main.cpp:623-624
    int rendered = 0, culled = 0;

    if (frustum.intersects(sponzaBox)) {                                                           // main.cpp:647-674
        ++rendered;
        mat<4, 4> view = ModelView;
        ModelView = ModelView * sponzaM;
        PhongShader sh(sponza_model.get());
        sh.initLightDirections(key, fill, rim);
        sh.normal_map_strength = 0.5;
        draw(*sponza_model, sh, framebuffer);
        ModelView = view;
    } else ++culled;

    if (frustum.intersects(headBox)) {                                                             // main.cpp:680-736
        ++rendered;
        mat<4, 4> view = ModelView;
        ModelView = ModelView * headM;
        PhongShader sh(head_model.get());
        sh.initLightDirections(key, fill, rim);
        draw(*head_model, sh, framebuffer);
#ifdef TRB_DEVICE_BACKEND
        gl_zbuffer_snapshot();                                   // std::vector<double> zbuffer_before_eyes = zbuffer;
#else
        std::vector<double> zbuffer_before_eyes = zbuffer;
#endif
        if (frustum.intersects(headBox)) {                       // the reference tests the HEAD box here (main.cpp:706)
            ++rendered;
            EyeShader eye(eye_model.get());
            eye.initLightDirections(key, rim);
            draw(*eye_model, eye, framebuffer);
        } else ++culled;
        ModelView = view;
#ifdef TRB_DEVICE_BACKEND
        gl_zbuffer_restore(framebuffer);                         // zbuffer = zbuffer_before_eyes;
#else
        zbuffer = zbuffer_before_eyes;
#endif
    } else ++culled;

#ifdef TRB_DEVICE_BACKEND
    gl_flush(framebuffer);
#endif
    framebuffer.write_tga_file(out + "/phong.tga");                                                // main.cpp:743
    TGAImage zimg, ao, fin;
#ifdef TRB_DEVICE_BACKEND
    gl_zbuffer_image(zimg);
    gl_ssao(ao);
    gl_composite_ao(fin);
#else
    {
        g_w = WIDTH; g_h = HEIGHT; g_color = framebuffer.buffer();
        std::vector<unsigned char> g((size_t)WIDTH * HEIGHT), c((size_t)WIDTH * HEIGHT * 3);
        orc_depth_image(nullptr, 0, g.data());
        grey_image(g.data(), WIDTH, HEIGHT, zimg);
        orc_ssao(nullptr, 0, g.data());
        grey_image(g.data(), WIDTH, HEIGHT, ao);
        orc_composite_ao(nullptr, 0, c.data());
        fin = TGAImage(WIDTH, HEIGHT, TGAImage::RGB);
        memcpy(fin.buffer(), c.data(), c.size());
    }
#endif
#ifdef TRB_DEVICE_BACKEND
    // the same four files once more, packetised on the device (tests compare them byte for byte)
    gl_write_tga_file(0, out + "/phong_dev.tga");
    gl_write_tga_file(1, out + "/zbuffer_dev.tga");
    gl_write_tga_file(2, out + "/ao_dev.tga");
    gl_write_tga_file(3, out + "/final_dev.tga");
#endif
    zimg.write_tga_file(out + "/zbuffer.tga");                                                     // main.cpp:751
    ao.write_tga_file(out + "/ao.tga");                                                            // main.cpp:764
    fin.write_tga_file(out + "/final.tga");                                                        // main.cpp:784
    std::ofstream zf(out + "/zbuffer.bin", std::ios::binary);
    zf.write((const char*)zbuffer.data(), zbuffer.size() * sizeof(double));
    print_render_stats();                                                                          // main.cpp:792
    // main.cpp:794-804: which models the (transposed-plane) frustum kept
    std::cout << "frustum sponza " << frustum.intersects(sponzaBox) << " head " << frustum.intersects(headBox) << std::endl;
    std::cout << "models rendered " << rendered << " culled " << culled << " faces "
              << sponza_model->nfaces() + head_model->nfaces() + eye_model->nfaces() << std::endl;
    return 0;
}
