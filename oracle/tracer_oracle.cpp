// tracer_oracle.cpp -- TEST INFRASTRUCTURE.  CPU restatement of the reference's device kernel
// internal/ocl/tracer.cl (`trace`, lines 831-1187, and every live helper it calls), driven the way
// internal/ocl/ocltracer.go:256-376 drives it (one seed per pixel, whole scene, RGBA doubles out).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library, and only as the checker / the timed CPU baseline.  The product (libptcuda)
// never links or calls it.
//
// PINNED on the reference kernel itself: tests/test_oracle_vs_reference.py renders every scene with this file AND with
// the reference's own kernel source (internal/ocl/tracer.cl compiled for the host CPU through cl_shim.hpp, see
// build_ref.py) on the same records and seeds and requires bit-identical pixels (fp64; both use the canonical float
// sin() of canon_rng.h -- OpenCL leaves sin's accuracy and FMA contraction to the implementation).  What cannot be
// pinned: the Go host side and a real OpenCL runtime (neither exists in this image), and the fp32 mode / fast RNG
// stream, which the reference does not have.  The reference's own unit-test vectors are replayed in
// tests/test_oracle_golden.py: the ray/AABB slab test (shapes/boundingbox_test.go:203-262), spherical UVs
// (shapes/sphericalmap_test.go:16-23), cube-face selection and cube-map lookups (shapes/cubemap_test.go:9-165).
//
// Structure follows the kernel line by line (same loop nest, same 4-wide vector arithmetic, same
// quirks); `Real` is double for the tracer.cl semantics and float for the "fp32 mode".  Variables
// the kernel declares __local but uses per work-item (tracer.cl:839,846,851,866,876) are private
// here -- the intended semantics.  Build with -ffp-contract=off.
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/ptwire.h"
#include "canon_rng.h"

namespace {

enum Counter {
    C_PATHS, C_SEGMENTS, C_MISSES, C_OBJ_PLANE, C_OBJ_SPHERE, C_OBJ_CYL, C_OBJ_CUBE, C_OBJ_GROUP, C_BOX_TESTS,
    C_TRI_DET, C_TRI_U, C_TRI_V, C_TRI_FULL, C_TRI_RECORDED, C_SHADED, C_DIFFUSE, C_MIRROR, C_REFRACT, C_THIN_PASS,
    C_TEX_LOOKUPS, C_NOISE, C_MAX_XS, C_XS_OVERFLOW, C_COUNT
};

template <typename Real>
struct V4 {
    Real x, y, z, w;
};
template <typename Real> inline V4<Real> operator+(V4<Real> a, V4<Real> b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
template <typename Real> inline V4<Real> operator-(V4<Real> a, V4<Real> b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
template <typename Real> inline V4<Real> operator*(V4<Real> a, V4<Real> b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
template <typename Real> inline V4<Real> operator*(V4<Real> a, Real s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
template <typename Real> inline V4<Real> neg(V4<Real> a) { return {-a.x, -a.y, -a.z, -a.w}; }
template <typename Real> inline Real dot(V4<Real> a, V4<Real> b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
template <typename Real> inline V4<Real> cross(V4<Real> a, V4<Real> b) {   // OpenCL cross(): w = 0
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x, Real(0)};
}
template <typename Real> inline Real length(V4<Real> a) { return std::sqrt(dot(a, a)); }
template <typename Real> inline V4<Real> normalize(V4<Real> a) { Real l = length(a); return {a.x / l, a.y / l, a.z / l, a.w / l}; }

template <typename Real> struct M16 { Real m[16]; };

// tracer.cl:369-376
template <typename Real> inline V4<Real> mul(const M16<Real>& a, V4<Real> v) {
    const Real* m = a.m;
    Real r0 = ((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3] * v.w;
    Real r1 = ((m[4] * v.x + m[5] * v.y) + m[6] * v.z) + m[7] * v.w;
    Real r2 = ((m[8] * v.x + m[9] * v.y) + m[10] * v.z) + m[11] * v.w;
    Real r3 = ((m[12] * v.x + m[13] * v.y) + m[14] * v.z) + m[15] * v.w;
    return {r0, r1, r2, r3};
}

template <typename Real> struct Obj {   // tracer.cl:37-63, converted once to Real
    M16<Real> transform;
    M16<Real> inverse, inverse_transpose;
    V4<Real> color, emission, bb_min, bb_max;
    Real refractive_index, min_y, max_y, reflectivity, tsx, tsy, tsx_nm, tsy_nm;
    long type;
    int child_count;
    int children[64];
    bool is_textured, is_textured_nm;
    unsigned char texture_index, texture_index_nm;
};
template <typename Real> struct Grp { V4<Real> bb_min, bb_max; int tri_offset, tri_count, children[2]; };  // tracer.cl:24-35
template <typename Real> struct Tri { V4<Real> p1, e1, e2, n1, n2, n3, color; };                       // tracer.cl:82-93
template <typename Real> struct Cam { int width, height; Real pixel_size, half_width, half_height, aperture, focal_length; M16<Real> inverse; };

struct Texture { const uint8_t* data; int w, h, layers; };

template <typename Real> struct Scene {
    std::vector<Obj<Real>> objects;
    std::vector<Grp<Real>> groups;
    std::vector<Tri<Real>> tris;
    Cam<Real> cam;
    Texture tex[3];
    unsigned samples;
    int rng_fast;   // 0: correctly rounded sine (parity stream); 1: the reproducible fast stream (canon_rng.h)
    int features;   // bit 0: next-event estimation (tracer.cl:1168 un-commented), bit 1: cylinder caps (tracer.cl:437-444 un-commented)
    Real PI, EPSILON;
};

template <typename Real> V4<Real> cv4(const double* p) { return {Real(p[0]), Real(p[1]), Real(p[2]), Real(p[3])}; }
template <typename Real> M16<Real> cm16(const double* p) { M16<Real> m; for (int i = 0; i < 16; ++i) m.m[i] = Real(p[i]); return m; }

// ---- texture sampling: OpenCL 1.2 (8.2) CLK_NORMALIZED_COORDS_TRUE | CLK_ADDRESS_REPEAT |
// CLK_FILTER_LINEAR on a CL_RGBA / CL_UNORM_INT8 image2d_array (tracer.cl:829, ocltracer.go:231-244).
// All arithmetic in float, unfused, in the order written.
struct F4 { float x, y, z, w; };
inline F4 texel(const Texture& t, int layer, int i, int j) {
    const uint8_t* p = t.data + ((size_t(layer) * t.h + j) * t.w + i) * 4;
    return {p[0] / 255.0f, p[1] / 255.0f, p[2] / 255.0f, p[3] / 255.0f};
}
inline F4 read_imagef(const Texture& t, float s, float tt, float layer_f) {
    if (!t.data) return {0.f, 0.f, 0.f, 0.f};     // the reference binds a blank 1024x1024 image (ocltracer.go:248-251)
    float u = (s - floorf(s)) * float(t.w);
    float v = (tt - floorf(tt)) * float(t.h);
    int i0 = int(floorf(u - 0.5f)), j0 = int(floorf(v - 0.5f));
    int i1 = i0 + 1, j1 = j0 + 1;
    if (i0 < 0) i0 = t.w + i0;
    if (i1 > t.w - 1) i1 = i1 - t.w;
    if (j0 < 0) j0 = t.h + j0;
    if (j1 > t.h - 1) j1 = j1 - t.h;
    float a = (u - 0.5f) - floorf(u - 0.5f);
    float b = (v - 0.5f) - floorf(v - 0.5f);
    int layer = int(rintf(layer_f));
    if (layer < 0) layer = 0;
    if (layer > t.layers - 1) layer = t.layers - 1;
    F4 t00 = texel(t, layer, i0, j0), t10 = texel(t, layer, i1, j0), t01 = texel(t, layer, i0, j1), t11 = texel(t, layer, i1, j1);
    float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
    F4 o;
    o.x = ((w00 * t00.x + w10 * t10.x) + w01 * t01.x) + w11 * t11.x;
    o.y = ((w00 * t00.y + w10 * t10.y) + w01 * t01.y) + w11 * t11.y;
    o.z = ((w00 * t00.z + w10 * t10.z) + w01 * t01.z) + w11 * t11.z;
    o.w = ((w00 * t00.w + w10 * t10.w) + w01 * t01.w) + w11 * t11.w;
    return o;
}

// ---- UV mappings -----------------------------------------------------------------------------
template <typename Real> inline Real maxX(Real a, Real b, Real c) { return std::fmax(std::fmax(a, b), c); }
template <typename Real> inline Real minX(Real a, Real b, Real c) { return std::fmin(std::fmin(a, b), c); }

// tracer.cl:113-175 (constants exactly as written there)
template <typename Real> void cube_uv(V4<Real> p, Real& ou, Real& ov) {
    Real ax = std::fabs(p.x), ay = std::fabs(p.y), az = std::fabs(p.z);
    Real coord = maxX(ax, ay, az);
    Real u, v;
    if (coord == p.x) {            // right
        u = std::fmod(Real(1.0) - p.z, Real(2)) / Real(2.0); v = std::fmod(p.y + Real(1.0), Real(2)) / Real(2.0);
        ou = Real(0.5) + u * Real(0.25); ov = Real(0.6666666) - v * Real(0.333333); return;
    }
    if (coord == -p.x) {           // left
        u = std::fmod(p.z + Real(1.0), Real(2)) / Real(2.0); v = std::fmod(p.y + Real(1.0), Real(2)) / Real(2.0);
        ou = u * Real(0.25); ov = Real(0.6666666) - v * Real(0.333333); return;
    }
    if (coord == p.y) {            // up
        u = std::fmod(p.x + Real(1.0), Real(2)) / Real(2.0); v = std::fmod(Real(1.0) - p.z, Real(2)) / Real(2.0);
        ou = Real(0.25) + u * Real(0.25); ov = Real(1.0) - v * Real(0.333333); return;
    }
    if (coord == -p.y) {           // down
        u = std::fmod(p.x + Real(1.0), Real(2)) / Real(2.0); v = std::fmod(p.z + Real(1.0), Real(2)) / Real(2.0);
        ou = Real(0.25) + u * Real(0.25); ov = v * Real(0.333333); return;
    }
    if (coord == p.z) {            // front
        u = std::fmod(p.x + Real(1.0), Real(2)) / Real(2.0); v = std::fmod(p.y + Real(1.0), Real(2)) / Real(2.0);
        ou = Real(0.25) + u * Real(0.25); ov = Real(0.6666666) - v * Real(0.333333); return;
    }
    u = std::fmod(Real(1.0) - p.x, Real(2)) / Real(2.0); v = std::fmod(p.y + Real(1.0), Real(2)) / Real(2.0);   // back
    ou = Real(0.75) + u * Real(0.25); ov = Real(0.6666666) - v * Real(0.333333);
}

// tracer.cl:178-213 (PI is the float-literal constant)
template <typename Real> void spherical_map(V4<Real> p, Real PI, Real& u, Real& v) {
    Real theta = std::atan2(p.x, p.z);
    V4<Real> vec = {p.x, p.y, p.z, Real(0)};
    Real radius = length(vec);
    Real phi = std::acos(p.y / radius);
    Real raw_u = theta / (Real(2.0) * PI);
    u = Real(1) - (raw_u + Real(0.5));
    v = Real(1) - phi / PI;
}

// tracer.cl:221-248
template <typename Real> void sunflower(int amount, Real alpha, int point_number, Real PI, Real& ox, Real& oy) {
    Real idx = Real(point_number);
    Real sqp = std::sqrt(Real(amount));
    Real b = std::round(alpha * sqp);
    Real phi = (std::sqrt(Real(5.0)) + Real(1.0)) / Real(2.0);
    Real r = Real(1.0);
    if (idx <= (Real(amount) - b)) r = std::sqrt(idx - Real(0.5)) / std::sqrt(Real(amount) - (b + Real(1.0)) / Real(2.0));
    Real theta = Real(2.0) * PI * idx / (phi * phi);
    ox = r * std::cos(theta);
    oy = r * std::sin(theta);
}

// tracer.cl:250-280
template <typename Real> inline void check_axis(Real origin, Real direction, Real lo, Real hi, Real eps, Real& tmin, Real& tmax) {
    Real a = lo - origin, b = hi - origin;
    if (std::fabs(direction) >= eps) { tmin = a / direction; tmax = b / direction; }
    else { tmin = a * Real(HUGE_VAL); tmax = b * Real(HUGE_VAL); }
    if (tmin > tmax) { Real t = tmin; tmin = tmax; tmax = t; }
}
template <typename Real> inline bool ray_box(V4<Real> o, V4<Real> d, V4<Real> lo, V4<Real> hi, Real eps) {
    Real x0, x1, y0, y1, z0, z1;
    check_axis(o.x, d.x, lo.x, hi.x, eps, x0, x1);
    check_axis(o.y, d.y, lo.y, hi.y, eps, y0, y1);
    check_axis(o.z, d.z, lo.z, hi.z, eps, z0, z1);
    return maxX(x0, y0, z0) < minX(x1, y1, z1);
}

// tracer.cl:485-505
template <typename Real> Real schlick(V4<Real> eye, V4<Real> n, Real n1, Real n2) {
    Real c = dot(eye, n);
    if (n1 > n2) {
        Real r = n1 / n2;
        Real sin2 = (r * r) * (Real(1.0) - (c * c));
        if (sin2 > Real(1.0)) return Real(1.0);
        c = std::sqrt(Real(1.0) - sin2);
    }
    Real t = (n1 - n2) / (n1 + n2);
    Real r0 = t * t;
    return r0 + (Real(1) - r0) * std::pow(Real(1) - c, Real(5));
}
// tracer.cl:507-533
template <typename Real> V4<Real> refracted(V4<Real> eye, V4<Real> n, Real n1, Real n2) {
    Real ratio = n1 / n2;
    Real cos_i = dot(eye, n);
    Real sin2 = (ratio * ratio) * (Real(1.0) - (cos_i * cos_i));
    if (sin2 > Real(1.0)) return {0, 0, 0, 0};
    Real cos_t = std::sqrt(Real(1.0) - sin2);
    return (n * ((ratio * cos_i) - cos_t)) - eye * ratio;
}

struct Ctx {   // tracer.cl:96-102, unbounded here (the kernel's 64 slots overflow silently)
    std::vector<double> t;
    std::vector<unsigned> obj;
    std::vector<int> tri;   // triangle index, -1 for analytic shapes
    std::vector<double> u, v;
    void clear() { t.clear(); obj.clear(); tri.clear(); u.clear(); v.clear(); }
    void push(double tt, unsigned o, int tr = -1, double uu = 0, double vv = 0) { t.push_back(tt); obj.push_back(o); tri.push_back(tr); u.push_back(uu); v.push_back(vv); }
};

template <typename Real>
struct Tracer {
    const Scene<Real>& sc;
    uint64_t cnt[C_COUNT];
    Ctx ctx;
    explicit Tracer(const Scene<Real>& s) : sc(s) { std::memset(cnt, 0, sizeof cnt); }

    float noise(float x, float y, float z) { cnt[C_NOISE]++; return canon_noise3d_mode(x, y, z, sc.rng_fast); }

    // tracer.cl:537-742.  Returns object index (-1 = none), t and the winning record's slot.
    int closest(V4<Real> ro, V4<Real> rd, Real& out_t, int& slot) {
        const Real EPS = sc.EPSILON;
        ctx.clear();
        for (unsigned j = 0; j < sc.objects.size(); ++j) {
            const Obj<Real>& ob = sc.objects[j];
            V4<Real> o = mul(ob.inverse, ro);
            V4<Real> d = mul(ob.inverse, rd);
            if (ob.type == 0) {                                   // plane, tracer.cl:478-483
                cnt[C_OBJ_PLANE]++;
                Real t = Real(0.0);
                if (std::fabs(d.y) > EPS) t = -o.y / d.y;
                if (t != Real(0.0)) ctx.push(t, j);
            } else if (ob.type == 1) {                            // sphere, tracer.cl:448-476
                cnt[C_OBJ_SPHERE]++;
                V4<Real> to_center = o - V4<Real>{0, 0, 0, Real(1.0)};
                Real a = dot(d, d);
                Real b = Real(2.0) * dot(d, to_center);
                Real c = dot(to_center, to_center) - Real(1.0);
                Real disc = (b * b) - Real(4) * a * c;
                Real t1 = 0, t2 = 0;
                if (disc > Real(0.0)) {
                    t1 = (-b - std::sqrt(disc)) / (Real(2) * a);
                    t2 = (-b + std::sqrt(disc)) / (Real(2) * a);
                }
                if (t1 != Real(0.0)) ctx.push(t1, j);
                if (t2 != Real(0.0)) ctx.push(t2, j);
            } else if (ob.type == 2) {                            // cylinder side, tracer.cl:396-446
                cnt[C_OBJ_CYL]++;
                Real out0 = 0, out1 = 0, out2 = 0, out3 = 0;
                Real rdx2 = d.x * d.x, rdz2 = d.z * d.z;
                Real a = rdx2 + rdz2;
                if (!(std::fabs(a) < EPS)) {
                    Real b = Real(2) * o.x * d.x + Real(2) * o.z * d.z;
                    Real rox2 = o.x * o.x, roz2 = o.z * o.z;
                    Real c1 = rox2 + roz2 - Real(1);
                    Real disc = b * b - Real(4) * a * c1;
                    if (!(disc < Real(0.0))) {
                        Real t0 = (-b - std::sqrt(disc)) / (Real(2) * a);
                        Real t1 = (-b + std::sqrt(disc)) / (Real(2) * a);
                        Real y0 = o.y + t0 * d.y;
                        if (y0 > ob.min_y && y0 < ob.max_y) out0 = t0;
                        Real y1 = o.y + t1 * d.y;
                        if (y1 > ob.min_y && y1 < ob.max_y) out1 = t1;
                        if (sc.features & 2) {                    // tracer.cl:437-444 enabled: intersectCaps, :290-310
                            Real cx = 0, cy = 0;
                            if (!(std::fabs(d.y) < EPS)) {
                                auto check_cap = [&](Real t) {   // tracer.cl:282-286
                                    Real x = o.x + t * d.x, z = o.z + t * d.z;
                                    return x * x + z * z <= Real(1.0);
                                };
                                Real tc1 = (ob.min_y - o.y) / d.y;
                                if (check_cap(tc1)) cx = tc1;
                                Real tc2 = (ob.max_y - o.y) / d.y;
                                if (check_cap(tc2)) cy = tc2;
                            }
                            if (cx > Real(0.0)) out2 = cx;
                            if (cy > Real(0.0)) out3 = cy;
                        }
                    }
                }
                if (out0 != Real(0)) ctx.push(out0, j);
                if (out1 != Real(0)) ctx.push(out1, j);
                if (out2 != Real(0)) ctx.push(out2, j);
                if (out3 != Real(0)) ctx.push(out3, j);
            } else if (ob.type == 3) {                            // cube, tracer.cl:378-394
                cnt[C_OBJ_CUBE]++;
                Real x0, x1, y0, y1, z0, z1;
                check_axis(o.x, d.x, Real(-1.0), Real(1.0), EPS, x0, x1);
                check_axis(o.y, d.y, Real(-1.0), Real(1.0), EPS, y0, y1);
                check_axis(o.z, d.z, Real(-1.0), Real(1.0), EPS, z0, z1);
                Real tmin = maxX(x0, y0, z0), tmax = minX(x1, y1, z1);
                Real a = 0, b = 0;
                if (!(tmin > tmax)) { a = tmin; b = tmax; }
                if (a != Real(0.0)) ctx.push(a, j);
                if (b != Real(0.0)) ctx.push(b, j);
            } else if (ob.type == 4) {                            // group, tracer.cl:598-720
                cnt[C_OBJ_GROUP]++;
                cnt[C_BOX_TESTS]++;
                if (!ray_box(o, d, ob.bb_min, ob.bb_max, EPS)) continue;
                for (int ci = 0; ci < ob.child_count; ++ci) {
                    int stack[64] = {0};
                    int sidx = 0;
                    int node = ob.children[ci];
                    bool have = true;     // `current != 0`
                    for (; have || sidx > -1;) {
                        for (; have;) {
                            const Grp<Real>& g = sc.groups[node];
                            cnt[C_BOX_TESTS]++;
                            if (!ray_box(o, d, g.bb_min, g.bb_max, EPS)) break;
                            for (int n = g.tri_offset; n < g.tri_offset + g.tri_count; ++n) {
                                const Tri<Real>& tr = sc.tris[n];
                                cnt[C_TRI_DET]++;
                                V4<Real> dir_x_e2 = cross(d, tr.e2);
                                Real det = dot(tr.e1, dir_x_e2);
                                if (std::fabs(det) < EPS) continue;
                                cnt[C_TRI_U]++;
                                Real f = Real(1.0) / det;
                                V4<Real> p1_to_o = o - tr.p1;
                                Real u = f * dot(p1_to_o, dir_x_e2);
                                if (u < 0 || u > 1) continue;
                                cnt[C_TRI_V]++;
                                V4<Real> o_x_e1 = cross(p1_to_o, tr.e1);
                                Real v = f * dot(d, o_x_e1);
                                if (v < 0 || (u + v) > 1) continue;
                                cnt[C_TRI_FULL]++;
                                Real t = f * dot(tr.e2, o_x_e1);
                                cnt[C_TRI_RECORDED]++;
                                ctx.push(t, j, n, u, v);
                            }
                            if (sidx < 64) stack[sidx] = node;
                            sidx++;
                            if (g.children[0] > 0) node = g.children[0];
                            else have = false;
                        }
                        sidx--;
                        if (sidx == -1) break;
                        const Grp<Real>& p = sc.groups[stack[sidx]];
                        if (p.children[1] > 0) { node = p.children[1]; have = true; }
                        else have = false;
                    }
                }
            }
        }
        if (ctx.t.size() > cnt[C_MAX_XS]) cnt[C_MAX_XS] = ctx.t.size();
        if (ctx.t.size() > 64) cnt[C_XS_OVERFLOW]++;
        out_t = Real(0.0); slot = -1;
        if (ctx.t.empty()) return -1;
        Real lowest = Real(1024.0);
        int idx = -1;
        for (size_t x = 0; x < ctx.t.size(); ++x) {
            Real t = Real(ctx.t[x]);
            if (t > EPS && t < lowest) { lowest = t; idx = int(ctx.obj[x]); slot = int(x); }
        }
        out_t = lowest;
        return idx;
    }

    // tracer.cl:745-779
    void ray_for_pixel(unsigned x, unsigned y, float rnd_x, float rnd_y, int sample, V4<Real>& origin, V4<Real>& direction) {
        const Cam<Real>& cam = sc.cam;
        V4<Real> in_view = {0, 0, Real(-1.0), Real(1.0)};
        V4<Real> origin_pt = {0, 0, 0, Real(1.0)};
        Real xo = cam.pixel_size * (Real(x) + Real(rnd_x));
        Real yo = cam.pixel_size * (Real(y) + Real(rnd_y));
        in_view.x = cam.half_width - xo;
        in_view.y = cam.half_height - yo;
        V4<Real> pixel = mul(cam.inverse, in_view);
        origin = mul(cam.inverse, origin_pt);
        direction = normalize(pixel - origin);
        if (cam.aperture != Real(0)) {
            V4<Real> pos = origin + direction * cam.focal_length;
            Real sx, sy;
            sunflower<Real>(int(sc.samples), Real(2), sample, sc.PI, sx, sy);
            V4<Real> no = {origin.x + (sy * cam.aperture), origin.y + (sx * cam.aperture), origin.z, Real(1.0)};
            direction = pos - no;
            origin = no;
        }
    }

    // tracer.cl:348-366
    V4<Real> hemisphere(V4<Real> n, Real x, Real y, Real z) {
        Real rand1 = Real(2.0) * sc.PI * Real(noise(float(x), float(y), float(z)));
        Real rand2 = Real(noise(float(y), float(z), float(x)));
        Real rand2s = std::sqrt(rand2);
        V4<Real> axis = (std::fabs(n.x) > Real(0.1)) ? V4<Real>{0, Real(1.0), 0, 0} : V4<Real>{Real(1.0), 0, 0, 0};
        V4<Real> u = normalize(cross(axis, n));
        V4<Real> v = cross(n, u);
        return u * std::cos(rand1) * rand2s + v * std::sin(rand1) * rand2s + n * std::sqrt(Real(1.0) - rand2);
    }

    struct Bounce { V4<Real> point, color, emission, normal; Real cos; bool is_refraction; };

    // tracer.cl:321-336 (sic: the latitude is shifted by 2*PI and y by PI/4 -- kept as written)
    V4<Real> random_point_on_sphere(Real r, Real u1, Real u2) {
        Real lat = std::acos(Real(2) * u1 - Real(1)) - sc.PI * Real(2);
        Real lon = Real(2) * sc.PI * u2;
        V4<Real> out = {0, 0, 0, Real(1.0)};
        out.x = std::cos(lat) * std::cos(lon) * r;
        out.y = (std::sin(lat) - sc.PI * Real(0.25)) * r;
        out.z = std::cos(lat) * std::sin(lon) * r;
        return out;
    }

    // tracer.cl:786-825, called from the shading loop when the feature is on (the call upstream keeps commented out, :1168)
    void next_event_estimation(const Bounce& b, float fgi_f, float fgi2_f, unsigned n_u, V4<Real> mask, unsigned x, V4<Real>& accum) {
        const Real EPS = sc.EPSILON;
        const Real fgi = Real(fgi_f), fgi2 = Real(fgi2_f), n = Real(n_u);       // the kernel's `double` parameters
        for (unsigned l = 0; l < sc.objects.size(); ++l) {
            const Obj<Real>& light = sc.objects[l];
            if (!(light.emission.x > Real(0.0))) continue;
            V4<Real> light_origin = {light.transform.m[3], light.transform.m[7], light.transform.m[11], Real(0.0)};
            Real scale_by = std::fmax(std::fmax(light.transform.m[0], light.transform.m[5]), light.transform.m[10]);
            V4<Real> light_scale = {scale_by, scale_by, scale_by, Real(1.0)};
            float r1 = noise(float(fgi), float(n + Real(x * l)), float(fgi2));
            float r2 = noise(float(fgi), float(fgi2), float(n + Real(x * x * l)));
            V4<Real> rpos = random_point_on_sphere(Real(1.0), Real(r1), Real(r2));
            V4<Real> light_position = light_origin + (rpos * light_scale);
            V4<Real> dir = normalize(light_position - b.point);
            V4<Real> origin = b.point + (dir * EPS);
            Real light_dot_normal = dot(dir, b.normal);
            if (light_dot_normal > Real(0.0)) {
                Real t; int slot;
                int hit = closest(origin, dir, t, slot);
                if (hit == int(l) && t > EPS) {
                    V4<Real> effective = b.color * light.emission;
                    Real attenuation = Real(1) - t / std::sqrt(t * t + light.transform.m[0] * light.transform.m[0]);
                    accum = accum + effective * light_dot_normal * mask * attenuation;
                }
            }
        }
    }

    // tracer.cl:831-1187, one pixel
    void pixel(unsigned x, unsigned y, double seed, double* out) {
        const Real EPS = sc.EPSILON;
        const unsigned samples = sc.samples;
        const unsigned num_objects = unsigned(sc.objects.size());
        Real color_weight = Real(1.0) / Real(samples);
        float fgi = float(seed / double(num_objects));
        float fgi2 = float(seed / double(samples));
        V4<Real> origin_pt = {0, 0, 0, Real(1.0)};
        V4<Real> colors = {0, 0, 0, 0};
        Bounce bounces[16];
        for (unsigned n = 0; n < samples; ++n) {
            cnt[C_PATHS]++;
            V4<Real> ro, rd;
            float jx = noise(fgi, float(n), fgi2);
            float jy = noise(fgi, fgi2, float(n));
            ray_for_pixel(x, y, jx, jy, int(n), ro, rd);
            unsigned actual = 0, effective = 0;
            bool entering = false, inside = false, exiting = false, reflecting = false;
            for (unsigned b = 0; b < 10 && effective < 4; ++b) {
                Real t; int slot;
                int hit = closest(ro, rd, t, slot);
                cnt[C_SEGMENTS]++;
                if (hit < 0) { cnt[C_MISSES]++; break; }   // the kernel re-traces the same ray until b==10; same result
                cnt[C_SHADED]++;
                const Obj<Real>& ob = sc.objects[hit];
                V4<Real> position = ro + rd * t;
                V4<Real> eye = neg(rd);
                V4<Real> on = {0, 0, 0, 0};
                if (ob.type == 0) {
                    if (ob.is_textured_nm) {
                        V4<Real> lp = mul(ob.inverse, position);
                        cnt[C_TEX_LOOKUPS]++;
                        F4 c = read_imagef(sc.tex[0], float(std::fabs(lp.x) * ob.tsx_nm), float(std::fabs(lp.z) * ob.tsy_nm), float(ob.texture_index_nm));
                        on = normalize(V4<Real>{Real(c.x), Real(c.y), Real(c.z), Real(0.0)});
                    } else on = {0, Real(1.0), 0, 0};
                } else if (ob.type == 1) {
                    V4<Real> lp = mul(ob.inverse, position);
                    on = lp - origin_pt;
                } else if (ob.type == 2) {
                    V4<Real> lp = mul(ob.inverse, position);
                    Real dist = lp.x * lp.x + lp.z * lp.z;       // pow(x,2)
                    if (dist < 1 && lp.y >= ob.max_y - EPS) on = {0, Real(1.0), 0, 0};
                    else if (dist < 1 && lp.y <= ob.min_y + EPS) on = {0, Real(-1.0), 0, 0};
                    else on = {lp.x, 0, lp.z, 0};
                } else if (ob.type == 3) {
                    V4<Real> lp = mul(ob.inverse, position);
                    Real maxc = maxX(std::fabs(lp.x), std::fabs(lp.y), std::fabs(lp.z));
                    if (maxc == std::fabs(lp.x)) on = {lp.x, 0, 0, 0};
                    else if (maxc == std::fabs(lp.y)) on = {0, lp.y, 0, 0};
                    else on = {0, 0, lp.z, 0};
                } else if (ob.type == 4) {
                    const Tri<Real>& tr = sc.tris[ctx.tri[slot]];
                    Real u = Real(ctx.u[slot]), v = Real(ctx.v[slot]);
                    on = tr.n2 * u + tr.n3 * v + tr.n1 * (Real(1.0) - u - v);   // tracer.cl:669
                }
                V4<Real> nv = mul(ob.inverse_transpose, on);
                nv.w = Real(0.0);
                nv = normalize(nv);
                if (dot(eye, nv) < Real(0.0)) nv = nv * Real(-1.0);
                V4<Real> over = position + nv * EPS;
                Real cosine = Real(1.0);
                entering = false; exiting = false; reflecting = false;
                auto reflect = [&]() {
                    Real ds = dot(rd, nv);
                    V4<Real> nn = (nv * Real(2.0)) * ds;
                    rd = rd - nn;
                    reflecting = true;
                    cnt[C_MIRROR]++;
                };
                if (ob.reflectivity != Real(0.0) && Real(noise(fgi, float(n), float(b))) < ob.reflectivity) {
                    reflect();
                } else if (ob.refractive_index == Real(-1.0)) {
                    if (schlick(eye, nv, Real(1.0), Real(1.5)) < Real(noise(fgi, float(n * n), float(b)))) {
                        over = position - nv * EPS;
                        cnt[C_THIN_PASS]++;
                    } else reflect();
                } else if (ob.refractive_index != Real(1.0)) {
                    if (!inside) {
                        Real sch = schlick(eye, nv, Real(1.0), ob.refractive_index);
                        Real rnd = Real(noise(fgi, float(n * n), float(b)));
                        if (sch < rnd) {
                            rd = refracted(eye, nv, Real(1.0), ob.refractive_index);
                            over = position - nv * EPS;
                            inside = true; entering = true; exiting = false;
                            cnt[C_REFRACT]++;
                        } else reflect();
                    } else {
                        Real sch = schlick(eye, nv, ob.refractive_index, Real(1.0));
                        if (sch < Real(noise(fgi, float(n * n), float(b)))) {
                            rd = refracted(eye, nv, ob.refractive_index, Real(1.0));
                            over = position - nv * EPS;
                            inside = false; entering = false; exiting = true;
                            cnt[C_REFRACT]++;
                        } else { reflect(); entering = false; exiting = false; }
                    }
                } else {
                    rd = hemisphere(nv, Real(fgi), Real(b), Real(n));
                    cosine = dot(rd, nv);
                    cnt[C_DIFFUSE]++;
                }
                ro = over;
                Bounce bn;
                bn.point = position;
                bn.normal = nv;
                bn.cos = cosine;
                bn.is_refraction = entering || exiting;
                if (ob.type == 4) {
                    bn.color = sc.tris[ctx.tri[slot]].color;
                    bn.emission = {0, 0, 0, 0};
                } else {
                    V4<Real> col = ob.color;
                    if (ob.is_textured) {
                        if (ob.type == 0) {
                            V4<Real> lp = mul(ob.inverse, position);
                            cnt[C_TEX_LOOKUPS]++;
                            F4 c = read_imagef(sc.tex[0], float(lp.x * ob.tsx), float(lp.z * ob.tsy), float(ob.texture_index));
                            col = {Real(c.x), Real(c.y), Real(c.z), Real(1.0)};
                        } else if (ob.type == 1) {
                            V4<Real> lp = mul(ob.inverse, position);
                            Real u, v;
                            spherical_map(lp, sc.PI, u, v);
                            cnt[C_TEX_LOOKUPS]++;
                            F4 c = read_imagef(sc.tex[1], float(u), float(Real(1.0) - v), float(ob.texture_index));
                            col = {Real(c.x), Real(c.y), Real(c.z), Real(1.0)};
                        } else if (ob.type == 3) {
                            V4<Real> lp = mul(ob.inverse, position);
                            Real u, v;
                            cube_uv(lp, u, v);
                            cnt[C_TEX_LOOKUPS]++;
                            F4 c = read_imagef(sc.tex[2], float(u), float(v), float(ob.texture_index));
                            col = {Real(c.x), Real(c.y), Real(c.z), Real(1.0)};
                        }
                    }
                    bn.color = col;
                    bn.emission = ob.emission;
                }
                bounces[actual] = bn;   // kernel indexes by b; identical because hits are contiguous from b=0
                if (!entering && !exiting && !reflecting) effective++;
                actual++;
                if (ob.emission.x > Real(0.0)) break;
            }
            // tracer.cl:1116-1179
            V4<Real> accum = {0, 0, 0, 0};
            V4<Real> mask = {Real(1.0), Real(1.0), Real(1.0), Real(1.0)};
            for (unsigned k = 0; k < actual; ++k) {
                const Bounce& bn = bounces[k];
                if (bn.is_refraction) continue;
                accum = accum + mask * bn.emission;
                if (bn.emission.x > Real(0.0)) {
                    if (k == 0) accum = bn.color;
                    break;
                }
                if (sc.features & 1) next_event_estimation(bn, fgi, fgi2, n, mask, k, accum);   // tracer.cl:1168
                mask = mask * bn.color;
                mask = mask * bn.cos;
            }
            colors = colors + accum;
        }
        out[0] = double(colors.x * color_weight);
        out[1] = double(colors.y * color_weight);
        out[2] = double(colors.z * color_weight);
        out[3] = 1.0;
    }
};

template <typename Real>
void build_scene(Scene<Real>& sc, const ptw_object* objs, int n_obj, const ptw_triangle* tris, int n_tri, const ptw_group* groups,
                 int n_grp, const ptw_camera* cam, const uint8_t* const* tex, const int32_t* tw, const int32_t* th,
                 const int32_t* tl, int samples, int rng_fast, int features) {
    sc.rng_fast = rng_fast;
    sc.features = features;
    sc.PI = Real(double(3.14159265359f));    // tracer.cl:1
    sc.EPSILON = Real(0.0001);               // tracer.cl:4
    sc.samples = unsigned(samples);
    for (int i = 0; i < n_obj; ++i) {
        const ptw_object& s = objs[i];
        Obj<Real> o;
        o.transform = cm16<Real>(s.transform);
        o.inverse = cm16<Real>(s.inverse);
        o.inverse_transpose = cm16<Real>(s.inverse_transpose);
        o.color = cv4<Real>(s.color); o.emission = cv4<Real>(s.emission);
        o.bb_min = cv4<Real>(s.bb_min); o.bb_max = cv4<Real>(s.bb_max);
        o.refractive_index = Real(s.refractive_index); o.min_y = Real(s.min_y); o.max_y = Real(s.max_y);
        o.reflectivity = Real(s.reflectivity);
        o.tsx = Real(s.texture_scale_x); o.tsy = Real(s.texture_scale_y);
        o.tsx_nm = Real(s.texture_scale_x_nm); o.tsy_nm = Real(s.texture_scale_y_nm);
        o.type = long(s.type); o.child_count = s.child_count;
        std::memcpy(o.children, s.children, sizeof o.children);
        o.is_textured = s.is_textured != 0; o.is_textured_nm = s.is_textured_nm != 0;
        o.texture_index = s.texture_index; o.texture_index_nm = s.texture_index_nm;
        sc.objects.push_back(o);
    }
    for (int i = 0; i < n_grp; ++i) {
        Grp<Real> g;
        g.bb_min = cv4<Real>(groups[i].bb_min); g.bb_max = cv4<Real>(groups[i].bb_max);
        g.tri_offset = groups[i].tri_offset; g.tri_count = groups[i].tri_count;
        g.children[0] = groups[i].children[0]; g.children[1] = groups[i].children[1];
        sc.groups.push_back(g);
    }
    for (int i = 0; i < n_tri; ++i) {
        Tri<Real> t;
        t.p1 = cv4<Real>(tris[i].p1); t.e1 = cv4<Real>(tris[i].e1); t.e2 = cv4<Real>(tris[i].e2);
        t.n1 = cv4<Real>(tris[i].n1); t.n2 = cv4<Real>(tris[i].n2); t.n3 = cv4<Real>(tris[i].n3);
        t.color = cv4<Real>(tris[i].color);
        sc.tris.push_back(t);
    }
    sc.cam.width = cam->width; sc.cam.height = cam->height;
    sc.cam.pixel_size = Real(cam->pixel_size); sc.cam.half_width = Real(cam->half_width); sc.cam.half_height = Real(cam->half_height);
    sc.cam.aperture = Real(cam->aperture); sc.cam.focal_length = Real(cam->focal_length);
    sc.cam.inverse = cm16<Real>(cam->inverse);
    for (int c = 0; c < 3; ++c) sc.tex[c] = Texture{tex ? tex[c] : nullptr, tw ? tw[c] : 0, th ? th[c] : 0, tl ? tl[c] : 0};
}

template <typename Real>
int run(const ptw_object* objs, int n_obj, const ptw_triangle* tris, int n_tri, const ptw_group* groups, int n_grp,
        const ptw_camera* cam, const uint8_t* const* tex, const int32_t* tw, const int32_t* th, const int32_t* tl,
        const double* seeds, int samples, int rng_fast, int features, int row0, int row1, int nthreads, double* out, uint64_t* counters) {
    Scene<Real> sc;
    build_scene(sc, objs, n_obj, tris, n_tri, groups, n_grp, cam, tex, tw, th, tl, samples, rng_fast, features);
    const int W = cam->width;
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next(row0);
    std::vector<std::vector<uint64_t>> cnts(size_t(nthreads), std::vector<uint64_t>(C_COUNT, 0));
    auto work = [&](int tid) {
        Tracer<Real> tr(sc);
        for (;;) {
            int y = next.fetch_add(1);
            if (y >= row1) break;
            for (int x = 0; x < W; ++x) tr.pixel(unsigned(x), unsigned(y), seeds[size_t(y) * W + x], out + (size_t(y - row0) * W + x) * 4);
        }
        for (int i = 0; i < C_COUNT; ++i) cnts[size_t(tid)][size_t(i)] = tr.cnt[i];
    };
    std::vector<std::thread> th_;
    for (int t = 1; t < nthreads; ++t) th_.emplace_back(work, t);
    work(0);
    for (auto& t : th_) t.join();
    if (counters) {
        for (int i = 0; i < C_COUNT; ++i) {
            uint64_t s = 0;
            for (int t = 0; t < nthreads; ++t) {
                if (i == C_MAX_XS) s = std::max(s, cnts[size_t(t)][size_t(i)]);
                else s += cnts[size_t(t)][size_t(i)];
            }
            counters[i] = s;
        }
    }
    return 0;
}

}  // namespace

extern "C" {

int oracle_counter_count(void) { return C_COUNT; }
const char* oracle_counter_name(int i) {
    static const char* names[C_COUNT] = {"paths", "segments", "misses", "obj_plane", "obj_sphere", "obj_cylinder", "obj_cube",
        "obj_group", "box_tests", "tri_det", "tri_u", "tri_v", "tri_full", "tri_recorded", "shaded", "diffuse", "mirror",
        "refract", "thin_pass", "tex_lookups", "noise", "max_intersections", "xs_overflow_segments"};
    return (i >= 0 && i < C_COUNT) ? names[i] : "";
}

// Renders rows [row0,row1) of the frame.  precision: 0 = float arithmetic, 1 = double (tracer.cl).
// rng_mode: 0 = parity stream, 1 = fast stream (both defined in canon_rng.h).
// seeds: width*height doubles (whole frame).  out: (row1-row0)*width*4 doubles.
// features: bit 0 next-event estimation, bit 1 cylinder caps -- the two code paths the reference ships switched off.
int oracle_trace3(const void* objects, int n_objects, const void* triangles, int n_triangles, const void* groups, int n_groups,
                  const void* camera, const uint8_t* const* tex, const int32_t* tex_w, const int32_t* tex_h,
                  const int32_t* tex_layers, const double* seeds, int samples, int precision, int rng_mode, int features, int row0, int row1,
                  int nthreads, double* out, uint64_t* counters) {
    if (!objects || n_objects < 1 || n_objects > PTW_MAX_OBJECTS || !camera || !seeds || !out || samples < 1) return 1;
    const ptw_camera* cam = static_cast<const ptw_camera*>(camera);
    if (row0 < 0 || row1 > cam->height || row0 > row1) return 2;
    auto* o = static_cast<const ptw_object*>(objects);
    auto* t = static_cast<const ptw_triangle*>(triangles);
    auto* g = static_cast<const ptw_group*>(groups);
    const int fast = rng_mode != 0;
    if (precision == 1) return run<double>(o, n_objects, t, n_triangles, g, n_groups, cam, tex, tex_w, tex_h, tex_layers, seeds, samples, fast, features, row0, row1, nthreads, out, counters);
    return run<float>(o, n_objects, t, n_triangles, g, n_groups, cam, tex, tex_w, tex_h, tex_layers, seeds, samples, fast, features, row0, row1, nthreads, out, counters);
}

int oracle_trace2(const void* objects, int n_objects, const void* triangles, int n_triangles, const void* groups, int n_groups,
                  const void* camera, const uint8_t* const* tex, const int32_t* tex_w, const int32_t* tex_h,
                  const int32_t* tex_layers, const double* seeds, int samples, int precision, int rng_mode, int row0, int row1,
                  int nthreads, double* out, uint64_t* counters) {
    return oracle_trace3(objects, n_objects, triangles, n_triangles, groups, n_groups, camera, tex, tex_w, tex_h, tex_layers, seeds,
                         samples, precision, rng_mode, 0, row0, row1, nthreads, out, counters);
}

int oracle_trace(const void* objects, int n_objects, const void* triangles, int n_triangles, const void* groups, int n_groups,
                 const void* camera, const uint8_t* const* tex, const int32_t* tex_w, const int32_t* tex_h,
                 const int32_t* tex_layers, const double* seeds, int samples, int precision, int row0, int row1, int nthreads,
                 double* out, uint64_t* counters) {
    return oracle_trace2(objects, n_objects, triangles, n_triangles, groups, n_groups, camera, tex, tex_w, tex_h, tex_layers, seeds,
                         samples, precision, 0, row0, row1, nthreads, out, counters);
}

// Unit hooks (golden-vector tests).
float oracle_noise3d(float x, float y, float z) { return canon_noise3d(x, y, z); }
float oracle_sinf(float x) { return canon_sinf(x); }
void oracle_noise3d_array(const float* xyz, int n, float* out) { for (int i = 0; i < n; ++i) out[i] = canon_noise3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); }
void oracle_noise3d_array_mode(const float* xyz, int n, int rng_mode, float* out) { for (int i = 0; i < n; ++i) out[i] = canon_noise3d_mode(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], rng_mode != 0); }
void oracle_sinf_array(const float* x, int n, float* out) { for (int i = 0; i < n; ++i) out[i] = canon_sinf(x[i]); }
int oracle_ray_box(const double* o, const double* d, const double* lo, const double* hi) {
    return ray_box<double>({o[0], o[1], o[2], o[3]}, {d[0], d[1], d[2], d[3]}, {lo[0], lo[1], lo[2], lo[3]}, {hi[0], hi[1], hi[2], hi[3]}, 0.0001) ? 1 : 0;
}
void oracle_spherical_map(const double* p, double* uv) { spherical_map<double>({p[0], p[1], p[2], p[3]}, double(3.14159265359f), uv[0], uv[1]); }
void oracle_cube_uv(const double* p, double* uv) { cube_uv<double>({p[0], p[1], p[2], p[3]}, uv[0], uv[1]); }
void oracle_sunflower(int amount, int index, double* xy) { sunflower<double>(amount, 2.0, index, double(3.14159265359f), xy[0], xy[1]); }
void oracle_read_imagef(const uint8_t* data, int w, int h, int layers, float s, float t, float layer, float* out4) {
    Texture tx{data, w, h, layers};
    F4 c = read_imagef(tx, s, t, layer);
    out4[0] = c.x; out4[1] = c.y; out4[2] = c.z; out4[3] = c.w;
}
double oracle_schlick(const double* eye, const double* n, double n1, double n2) { return schlick<double>({eye[0], eye[1], eye[2], eye[3]}, {n[0], n[1], n[2], n[3]}, n1, n2); }

}  // extern "C"
