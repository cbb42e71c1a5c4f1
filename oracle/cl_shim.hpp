// cl_shim.hpp -- TEST INFRASTRUCTURE.  Just enough of OpenCL C 1.2 in C++ to compile the reference's own device
// kernel, /root/reference/internal/ocl/tracer.cl, for the host CPU (oracle/build_ref.py; output oracle/_ref/).
//
// Why: neither Go nor an OpenCL runtime exists in this image, so the reference cannot be *run* as shipped.  Its
// arithmetic, however, lives entirely in tracer.cl, and OpenCL C is close enough to C++ that the file compiles once
// this header supplies the vector types, the handful of built-ins the kernel calls, and the address-space keywords.
// The compiled kernel is the real reference code: tests/test_oracle_vs_reference.py pins oracle/tracer_oracle.cpp (the
// restatement every GPU parity test uses) against it pixel by pixel.
//
// What this header decides (because OpenCL leaves it to the implementation), always the same way as the oracle:
//   * float sin() -- the one place the image is chaotic in (noise3D, tracer.cl:314-317) -- is canon_sinf
//     (oracle/canon_rng.h): the correctly rounded single-precision sine; no FMA contraction (-ffp-contract=off);
//   * dot() sums x,y,z,w left to right; normalize() divides by length(); max/min are fmax/fmin;
//   * read_imagef() is the OpenCL 1.2 (8.2) formula for NORMALIZED | REPEAT | LINEAR on CL_RGBA / CL_UNORM_INT8,
//     unfused float arithmetic in the order the specification writes it.
// Everything else -- every line of geometry, shading and control flow -- is the reference's.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "canon_rng.h"

// ---- address spaces, qualifiers ------------------------------------------------------------------------------------
#define __kernel
#define __global
#define __constant const
#define __local            // per work-item here: the kernel uses its __local variables privately
#define __private

// ---- vector types --------------------------------------------------------------------------------------------------
// Only doubles inside, so their natural layout is the packed layout of the reference's wire records (ref_driver.cpp
// static_asserts the sizes and offsets).  A scalar converts to a vector by broadcast, as in OpenCL C
// (`double2 out = (0,0);`, tracer.cl:379, is a comma expression); v[i] indexes components (tracer.cl:152, 577, 790).
struct double2 {
    double x, y;
    double2() = default;
    double2(double a) : x(a), y(a) {}
    double2(double a, double b) : x(a), y(b) {}
};
struct double4 {
    double x, y, z, w;
    double4() = default;
    double4(double a) : x(a), y(a), z(a), w(a) {}
    double4(double a, double b, double c, double d) : x(a), y(b), z(c), w(d) {}
    double& operator[](int i) { return (&x)[i]; }
    const double& operator[](int i) const { return (&x)[i]; }
};
struct float4 {
    float x, y, z, w;
    float4() = default;
    float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};
struct double16 {                                          // rows (tracer.cl:370-373) and elements (tracer.cl:790)
    double4 s0123, s4567, s89AB, sCDEF;
    double& operator[](int i) { return (&s0123.x)[i]; }
    const double& operator[](int i) const { return (&s0123.x)[i]; }
};
static_assert(sizeof(double4) == 32 && sizeof(double16) == 128 && sizeof(double2) == 16, "vector sizes");

static inline double2 mk_double2(double a, double b) { return {a, b}; }
static inline double4 mk_double4(double a, double b, double c, double d) { return {a, b, c, d}; }
static inline double4 mk_double4(double a) { return {a, a, a, a}; }
static inline float4 mk_float4(float a, float b, float c, float d) { return {a, b, c, d}; }

static inline double4 operator+(double4 a, double4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
static inline double4 operator-(double4 a, double4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
static inline double4 operator*(double4 a, double4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
static inline double4 operator/(double4 a, double4 b) { return {a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w}; }
static inline double4 operator*(double4 a, double s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
static inline double4 operator*(double s, double4 a) { return {s * a.x, s * a.y, s * a.z, s * a.w}; }
static inline double4 operator/(double4 a, double s) { return {a.x / s, a.y / s, a.z / s, a.w / s}; }
static inline double4 operator+(double4 a, double s) { return {a.x + s, a.y + s, a.z + s, a.w + s}; }
static inline double4 operator-(double4 a, double s) { return {a.x - s, a.y - s, a.z - s, a.w - s}; }
static inline double4 operator-(double4 a) { return {-a.x, -a.y, -a.z, -a.w}; }
static inline double4& operator+=(double4& a, double4 b) { a = a + b; return a; }
static inline double4& operator-=(double4& a, double4 b) { a = a - b; return a; }
static inline double4& operator*=(double4& a, double4 b) { a = a * b; return a; }
static inline double4& operator*=(double4& a, double s) { a = a * s; return a; }
static inline double4& operator/=(double4& a, double s) { a = a / s; return a; }

// ---- built-ins the kernel calls -----------------------------------------------------------------------------------
static inline double dot(double4 a, double4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
static inline double4 cross(double4 a, double4 b) {          // OpenCL cross(): w = 0
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x, 0.0};
}
static inline double length(double4 a) { return std::sqrt(dot(a, a)); }
static inline double4 normalize(double4 a) { const double l = length(a); return {a.x / l, a.y / l, a.z / l, a.w / l}; }
static inline double max(double a, double b) { return std::fmax(a, b); }
static inline double min(double a, double b) { return std::fmin(a, b); }
static inline double convert_double(int v) { return double(v); }
static inline double convert_double(unsigned v) { return double(v); }
// (float sin() is declared inside the kernel's namespace by ref_driver.cpp: <cmath> already owns ::sin(float))
static inline float fract(float x, float* ip) { const float f = std::floor(x); *ip = f; return std::fmin(x - f, 0x1.fffffep-1f); }

static thread_local int cl_shim_global_id = 0;
static inline int get_global_id(int) { return cl_shim_global_id; }

// ---- images -------------------------------------------------------------------------------------------------------
struct cl_shim_image { const uint8_t* data; int w, h, layers; };
typedef cl_shim_image image2d_array_t;
typedef int sampler_t;
enum { CLK_NORMALIZED_COORDS_TRUE = 1, CLK_ADDRESS_REPEAT = 2, CLK_FILTER_LINEAR = 4 };

static inline float4 cl_shim_texel(const cl_shim_image& t, int layer, int i, int j) {
    const uint8_t* p = t.data + ((size_t(layer) * t.h + j) * t.w + i) * 4;
    return {p[0] / 255.0f, p[1] / 255.0f, p[2] / 255.0f, p[3] / 255.0f};
}
static inline float4 read_imagef(const cl_shim_image& t, sampler_t, float4 c) {
    if (!t.data) return {0.f, 0.f, 0.f, 0.f};        // the reference binds a blank image when a class has no textures (ocltracer.go:248-251)
    const float u = (c.x - std::floor(c.x)) * float(t.w);
    const float v = (c.y - std::floor(c.y)) * float(t.h);
    int i0 = int(std::floor(u - 0.5f)), j0 = int(std::floor(v - 0.5f));
    int i1 = i0 + 1, j1 = j0 + 1;
    if (i0 < 0) i0 = t.w + i0;
    if (i1 > t.w - 1) i1 = i1 - t.w;
    if (j0 < 0) j0 = t.h + j0;
    if (j1 > t.h - 1) j1 = j1 - t.h;
    const float a = (u - 0.5f) - std::floor(u - 0.5f);
    const float b = (v - 0.5f) - std::floor(v - 0.5f);
    int layer = int(std::rint(c.z));
    if (layer < 0) layer = 0;
    if (layer > t.layers - 1) layer = t.layers - 1;
    const float4 t00 = cl_shim_texel(t, layer, i0, j0), t10 = cl_shim_texel(t, layer, i1, j0);
    const float4 t01 = cl_shim_texel(t, layer, i0, j1), t11 = cl_shim_texel(t, layer, i1, j1);
    const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
    float4 o;
    o.x = ((w00 * t00.x + w10 * t10.x) + w01 * t01.x) + w11 * t11.x;
    o.y = ((w00 * t00.y + w10 * t10.y) + w01 * t01.y) + w11 * t11.y;
    o.z = ((w00 * t00.z + w10 * t10.z) + w01 * t01.z) + w11 * t11.z;
    o.w = ((w00 * t00.w + w10 * t10.w) + w01 * t01.w) + w11 * t11.w;
    return o;
}
