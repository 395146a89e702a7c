// shaders.h - PhongShader and EyeShader of the reference's main.cpp (39-171, 176-262) as a header,
// host side.  vertex()/fragment() are complete host implementations (they are what runs in the
// oracle build, through the reference's own rasterize()); device_shader() tells the B200 backend
// which device shader + uniform block stands for them.  Compiles against either header set.
#pragma once
#include <our_gl.h>
#include <model.h>

#include <algorithm>
#include <cmath>

struct LitShader : public IShader {
    const Model* model;
    vec3 key_light_dir_eye, fill_light_dir_eye, rim_light_dir_eye;
    vec2 varying_uv[3];
    vec3 varying_position_eye[3];
    vec3 varying_normal_eye[3];
    double normal_map_strength = 1.0;
#ifdef TRB_DEVICE_BACKEND
    mutable double device_varyings[24];
#endif
    explicit LitShader(const Model* m) : model(m) {}

    static vec3 to_eye(const vec3& dir_world) {  // main.cpp:59-68
        mat<3, 3> nm;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) nm[i][j] = ModelView[i][j];
        return normalized(nm * dir_world);
    }
    vec4 vertex(int face, int nth) override {     // main.cpp:71-90 / 199-218
        vec3 p = model->vert(face, nth);
        vec3 n = model->normal(face, nth);
        varying_uv[nth] = model->uv(face, nth);
        vec4 pe = ModelView * make_vec4(p[0], p[1], p[2], 1.0);
        varying_position_eye[nth] = pe.xyz();
        vec4 ne = ModelView * make_vec4(n[0], n[1], n[2], 0.0);
        varying_normal_eye[nth] = ne.xyz();
        return Perspective * pe;
    }
#ifdef TRB_DEVICE_BACKEND
    bool describe(int kind, TrbDeviceShader& d) const {
        d.kind = kind;
        for (int i = 0; i < 3; ++i) {
            d.key_dir_eye[i] = key_light_dir_eye[i];
            d.fill_dir_eye[i] = fill_light_dir_eye[i];
            d.rim_dir_eye[i] = rim_light_dir_eye[i];
            double* q = device_varyings + 8 * i;
            q[0] = varying_uv[i].x; q[1] = varying_uv[i].y;
            q[2] = varying_position_eye[i].x; q[3] = varying_position_eye[i].y; q[4] = varying_position_eye[i].z;
            q[5] = varying_normal_eye[i].x; q[6] = varying_normal_eye[i].y; q[7] = varying_normal_eye[i].z;
        }
        d.normal_map_strength = normal_map_strength;
        d.model = model;
        d.varyings = device_varyings;
        return true;
    }
#endif
};

struct PhongShader : public LitShader {
    explicit PhongShader(const Model* m) : LitShader(m) {}
    void initLightDirections(const vec3& key, const vec3& fill, const vec3& rim) {
        key_light_dir_eye = to_eye(key);
        fill_light_dir_eye = to_eye(fill);
        rim_light_dir_eye = to_eye(rim);
    }
    std::pair<bool, TGAColor> fragment(const vec3 bar) const override {  // main.cpp:92-170
        vec3 pos = varying_position_eye[0] * bar[0] + varying_position_eye[1] * bar[1] + varying_position_eye[2] * bar[2];
        vec3 gn = varying_normal_eye[0] * bar[0] + varying_normal_eye[1] * bar[1] + varying_normal_eye[2] * bar[2];
        vec2 uv = varying_uv[0] * bar[0] + varying_uv[1] * bar[1] + varying_uv[2] * bar[2];
        TGAColor base = model->diffuse(uv);
        double spec_pow = std::max(1.0, (double)model->specular(uv));
        double brightness = (base[0] + base[1] + base[2]) / (3.0 * 255.0);
        bool eye_px = (brightness >= 0.85) && (spec_pow <= 5.0);
        vec3 nm = model->normal(uv);
        vec3 nm_eye = (ModelView * make_vec4(nm[0], nm[1], nm[2], 0.0)).xyz();
        vec3 N = eye_px ? gn : normalized(gn * (1.0 - normal_map_strength) + nm_eye * normal_map_strength);
        vec3 V = normalized(-pos);
        double key_d = std::max(0.0, dot(N, key_light_dir_eye)) * 1.0;
        vec3 R = normalized(N * (2.0 * dot(N, key_light_dir_eye)) - key_light_dir_eye);
        double rv = std::max(0.0, dot(R, V));
        double key_s = (rv > 0.0 ? std::pow(rv, spec_pow) : 0.0) * 1.0;
        double fill_d = std::max(0.0, dot(N, fill_light_dir_eye)) * 0.35;
        double rim_d = std::max(0.0, dot(N, rim_light_dir_eye)) * 0.6;
        double diff = key_d + fill_d + rim_d;
        TGAColor out = base;
        for (int ch = 0; ch < 3; ++ch) {
            double v = (double)base[ch] * (0.10 + diff) + 255.0 * (0.35 * key_s);
            out[ch] = (unsigned char)std::min(255.0, v);
        }
        return {false, out};
    }
#ifdef TRB_DEVICE_BACKEND
    bool device_shader(TrbDeviceShader& d) const override { return describe(1 /*TRB_SHADER_PHONG*/, d); }
#endif
};

struct EyeShader : public LitShader {
    explicit EyeShader(const Model* m) : LitShader(m) {}
    void initLightDirections(const vec3& key, const vec3& rim) {
        key_light_dir_eye = to_eye(key);
        rim_light_dir_eye = to_eye(rim);
    }
    std::pair<bool, TGAColor> fragment(const vec3 bar) const override {  // main.cpp:220-261
        vec3 pos = varying_position_eye[0] * bar[0] + varying_position_eye[1] * bar[1] + varying_position_eye[2] * bar[2];
        vec3 N = normalized(varying_normal_eye[0] * bar[0] + varying_normal_eye[1] * bar[1] + varying_normal_eye[2] * bar[2]);
        vec2 uv = varying_uv[0] * bar[0] + varying_uv[1] * bar[1] + varying_uv[2] * bar[2];
        TGAColor base = model->diffuse(uv);
        vec3 V = normalized(-pos);
        double key_d = std::max(0.0, dot(N, key_light_dir_eye)) * 1.0;
        double rim_d = std::max(0.0, dot(N, rim_light_dir_eye)) * 0.6;
        double diff = key_d + rim_d;
        double spec_pow = std::max(1.0, (double)model->specular(uv)) * 8.0;
        vec3 R = normalized(N * (2.0 * dot(N, key_light_dir_eye)) - key_light_dir_eye);
        double rv = std::max(0.0, dot(R, V));
        double spec = (rv > 0.0 ? std::pow(rv, spec_pow) : 0.0);
        TGAColor out = base;
        for (int ch = 0; ch < 3; ++ch) {
            double v = (double)base[ch] * (0.1 + diff) + 255.0 * (1.5 * spec);
            out[ch] = (unsigned char)std::min(255.0, v);
        }
        return {false, out};
    }
#ifdef TRB_DEVICE_BACKEND
    bool device_shader(TrbDeviceShader& d) const override { return describe(2 /*TRB_SHADER_EYE*/, d); }
#endif
};
