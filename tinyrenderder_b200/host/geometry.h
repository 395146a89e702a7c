// geometry.h - host-side math types with the API surface of the reference's geometry.h
// (vec<n>, mat<r,c>, dot/cross/normalized, Plane, AABB; geometry.h:13-328), written from scratch.
// These types only carry values to the C ABI (row-major double[16] == mat<4,4>::rows); device code
// has its own arithmetic in csrc/exact.cuh.  The OPERATION ORDER of every function matches the
// reference (sum from +0.0 in index order, componentwise divide, normalized() returning v when the
// length is exactly 0) because host-side setup (lookat, light directions) feeds the device.
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <iostream>

template <int N> struct vec {
    double data[N] = {};
    double& operator[](int i) { assert(i >= 0 && i < N); return data[i]; }
    double operator[](int i) const { assert(i >= 0 && i < N); return data[i]; }
};
template <> struct vec<2> {
    double x = 0, y = 0;
    double& operator[](int i) { assert(i >= 0 && i < 2); return i ? y : x; }
    double operator[](int i) const { assert(i >= 0 && i < 2); return i ? y : x; }
};
template <> struct vec<3> {
    double x = 0, y = 0, z = 0;
    double& operator[](int i) { assert(i >= 0 && i < 3); return i == 0 ? x : (i == 1 ? y : z); }
    double operator[](int i) const { assert(i >= 0 && i < 3); return i == 0 ? x : (i == 1 ? y : z); }
};
template <> struct vec<4> {
    double data[4] = {0, 0, 0, 0};
    double& operator[](int i) { assert(i >= 0 && i < 4); return data[i]; }
    double operator[](int i) const { assert(i >= 0 && i < 4); return data[i]; }
    double x() const { return data[0]; }
    double y() const { return data[1]; }
    double z() const { return data[2]; }
    double w() const { return data[3]; }
    vec<2> xy() const { vec<2> r; r.x = data[0]; r.y = data[1]; return r; }
    vec<3> xyz() const { vec<3> r; r.x = data[0]; r.y = data[1]; r.z = data[2]; return r; }
};
typedef vec<2> vec2;
typedef vec<3> vec3;
typedef vec<4> vec4;

template <int N> vec<N> operator+(const vec<N>& a, const vec<N>& b) { vec<N> r; for (int i = 0; i < N; ++i) r[i] = a[i] + b[i]; return r; }
template <int N> vec<N> operator-(const vec<N>& a, const vec<N>& b) { vec<N> r; for (int i = 0; i < N; ++i) r[i] = a[i] - b[i]; return r; }
template <int N> vec<N> operator*(const vec<N>& a, double s) { vec<N> r; for (int i = 0; i < N; ++i) r[i] = a[i] * s; return r; }
template <int N> vec<N> operator*(double s, const vec<N>& a) { return a * s; }
template <int N> vec<N> operator/(const vec<N>& a, double s) { vec<N> r; for (int i = 0; i < N; ++i) r[i] = a[i] / s; return r; }
template <int N> vec<N> operator-(const vec<N>& a) { return a * -1.0; }
template <int N> double dot(const vec<N>& a, const vec<N>& b) { double s = 0; for (int i = 0; i < N; ++i) s += a[i] * b[i]; return s; }
template <int N> double norm(const vec<N>& a) { return std::sqrt(dot(a, a)); }
template <int N> std::ostream& operator<<(std::ostream& o, const vec<N>& v) { for (int i = 0; i < N; ++i) o << v[i] << " "; return o; }

inline vec3 normalized(const vec3& v) {
    double len = norm<3>(v);
    if (len == 0) return v;
    return v / len;
}
inline vec3 cross(const vec3& a, const vec3& b) {
    vec3 r;
    r.x = a[1] * b[2] - a[2] * b[1];
    r.y = a[2] * b[0] - a[0] * b[2];
    r.z = a[0] * b[1] - a[1] * b[0];
    return r;
}
inline vec2 make_vec2(double x, double y) { vec2 r; r.x = x; r.y = y; return r; }
inline vec3 make_vec3(double x, double y, double z) { vec3 r; r.x = x; r.y = y; r.z = z; return r; }
inline vec4 make_vec4(double x, double y, double z, double w) { vec4 r; r[0] = x; r[1] = y; r[2] = z; r[3] = w; return r; }

template <int R, int C> struct mat {
    vec<C> rows[R];
    vec<C>& operator[](int r) { assert(r >= 0 && r < R); return rows[r]; }
    const vec<C>& operator[](int r) const { assert(r >= 0 && r < R); return rows[r]; }
    static mat<R, C> identity() {
        mat<R, C> m;
        for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) m[r][c] = (r == c) ? 1.0 : 0.0;
        return m;
    }
    mat<C, R> transpose() const {
        mat<C, R> m;
        for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) m[c][r] = rows[r][c];
        return m;
    }
};
template <int R, int C> vec<R> operator*(const mat<R, C>& M, const vec<C>& v) {
    vec<R> r;
    for (int i = 0; i < R; ++i) r[i] = dot<C>(M[i], v);
    return r;
}
template <int R1, int C1, int C2> mat<R1, C2> operator*(const mat<R1, C1>& A, const mat<C1, C2>& B) {
    mat<R1, C2> m;
    for (int i = 0; i < R1; ++i)
        for (int j = 0; j < C2; ++j) {
            m[i][j] = 0;
            for (int k = 0; k < C1; ++k) m[i][j] += A[i][k] * B[k][j];
        }
    return m;
}
template <int R, int C> std::ostream& operator<<(std::ostream& o, const mat<R, C>& M) { for (int i = 0; i < R; ++i) o << M[i] << "\n"; return o; }

// Plane / AABB (geometry.h:253-328): model-level frustum culling helpers, host only
struct Plane {
    vec3 normal;
    double d;
    Plane() : d(0) { normal.z = 1; }
    Plane(const vec3& n, const vec3& point) { normal = normalized(n); d = -dot(normal, point); }
    double distance(const vec3& p) const { return dot(normal, p) + d; }
};
struct AABB {
    vec3 min, max;
    AABB() {}
    AABB(const vec3& lo, const vec3& hi) : min(lo), max(hi) {}
    vec3 getCenter() const { return (min + max) * 0.5; }
    vec3 getSize() const { return max - min; }
    vec3 getHalfSize() const { return getSize() * 0.5; }
    bool intersects(const AABB& o) const {
        return (min.x <= o.max.x && max.x >= o.min.x) && (min.y <= o.max.y && max.y >= o.min.y) &&
               (min.z <= o.max.z && max.z >= o.min.z);
    }
    // the 8 corners through the matrix with a perspective divide, min/max from +-1e9 (geometry.h:297-327)
    AABB transform(const mat<4, 4>& m) const {
        vec3 lo = make_vec3(1e9, 1e9, 1e9), hi = make_vec3(-1e9, -1e9, -1e9);
        for (int k = 0; k < 8; ++k) {
            vec4 p = m * make_vec4((k & 1) ? max.x : min.x, (k & 2) ? max.y : min.y, (k & 4) ? max.z : min.z, 1.0);
            vec3 q = p.xyz() / p.w();
            lo.x = std::min(lo.x, q.x); lo.y = std::min(lo.y, q.y); lo.z = std::min(lo.z, q.z);
            hi.x = std::max(hi.x, q.x); hi.y = std::max(hi.y, q.y); hi.z = std::max(hi.z, q.z);
        }
        return AABB(lo, hi);
    }
};
