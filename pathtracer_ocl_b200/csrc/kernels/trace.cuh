// trace.cuh -- the path-tracing kernel for sm_100a.
//
// Replaces the reference's OpenCL device kernel `trace` (reference internal/ocl/tracer.cl:831-1187)
// and every helper on its path.  Same observable behaviour -- camera rays with AA jitter and
// sunflower depth of field (tracer.cl:745-779, 221-248), plane / sphere / cylinder / cube / BVH
// triangle intersection (378-483, 598-720), reflect / thin-glass / refract / diffuse branching
// (982-1061), texture lookups (907-911, 1076-1093), mask-and-accumulate shading (1116-1179), the
// hash RNG (314-317) -- but organised for a Blackwell SM instead of mirroring the OpenCL source:
//
//   * one launch covers the whole frame (the reference enqueues H/4 four-scanline batches);
//   * a thread owns a pixel and a slice of its samples and runs ONE flat loop whose iteration is
//     a path segment: when a lane's path ends it regenerates the next camera ray in place, so lanes
//     of a warp stay busy instead of idling until the longest path of the warp finishes (the
//     reference nests samples x bounces per work-item); the slices of a pixel are the warps of a
//     block and the blocks of a thread-block cluster, and are summed through (distributed) shared
//     memory in the kernel's epilogue -- one launch, one store per pixel, optionally straight into
//     a frame on another GPU (the multi-GPU gather fused into the kernel);
//   * closest hit is a running minimum in registers (the reference zero-fills a 6.9 KB `context`
//     per bounce and scans it afterwards, tracer.cl:886, 727-741) and shading is fused into the
//     segment loop (the reference stores bounces and replays them, 1071-1096 -> 1116-1179);
//   * the per-object data the intersection code needs travels in the kernel parameter block, i.e.
//     the constant bank (the reference copies 16 KB of 1024-byte records into __local per
//     work-item, tracer.cl:846-849): planes and spheres as 16-byte records of UNROLLED slots at
//     fixed offsets -- no loop, no type dispatch, spheres intersected in world space -- the rest as
//     affine inverse + bounds for a short loop; material data is fetched per hit;
//   * mesh objects are re-indexed on the host by an 8-wide SAH BVH over the caller's own triangle
//     records, walked cooperatively -- eight lanes share one ray, four rays per warp -- and every
//     candidate hit is checked against the caller's BVH boxes so that exactly the triangles the
//     reference's stack walk (tracer.cl:624-714) would have tested can win (see "mesh objects"
//     below); nodes and triangles are 16-byte-vectorised records (48 B of test data per triangle
//     instead of a 512-byte stride); warps with few rays at the mesh put the walk off so that rays
//     share rounds, warps whose lanes all reach it walk one ray per lane, and the tiles that look
//     at a mesh are launched first (the host orders them by measured clocks);
//   * the depth-of-field lens points depend only on the sample index, so they come from a table
//     built once on the host instead of two sqrt, a divide and a sincos per path;
//   * everything is templated on the arithmetic type: float = "fp32 mode", double = tracer.cl's
//     precision (its formulas, not its operation order: agreement to the 1e-6 gate, not bit for bit).
//     The fp32 instantiation uses the SFU approximations (rcp / rsqrt / sqrt / sin / cos) for its
//     own arithmetic; the RNG and the texture filter stay exactly rounded in both modes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rng.cuh"

#ifndef PTK_KERNEL_VERSION
#define PTK_KERNEL_VERSION "r02v6"      // names the ncu captures under profiles/ that belong to this kernel source
#endif

namespace ptk {

// Resident-blocks hints for the register allocator, tuned by measurement on B200 (profiles/): the fp32 kernels run
// best at 8 blocks x 128 threads (64 registers; reference scene 10.90 / 10.91 / 11.25 / 11.26 Gpaths/s at 5 / 6 / 7 / 8
// blocks -- the analytic kernel fits 64 registers with 8 bytes of spill), fp64 needs more registers per thread
// (analytic kernel at 3 / 4 / 5 blocks: 4.64 / 4.69 / 5.03 Gpaths/s; mesh kernel, teapot: 0.78 / 0.88 / 0.79).
#ifndef PTK_MIN_BLOCKS
#define PTK_MIN_BLOCKS 8
#endif
#ifndef PTK_MIN_BLOCKS_F64
#define PTK_MIN_BLOCKS_F64 5
#endif
#ifndef PTK_MESH_MIN_BLOCKS_F64
#define PTK_MESH_MIN_BLOCKS_F64 4
#endif
#ifndef PTK_MESH_MIN_BLOCKS
#define PTK_MESH_MIN_BLOCKS 8
#endif
#ifndef PTK_BLOCK_THREADS
#define PTK_BLOCK_THREADS 128
#endif

constexpr int kMaxObjects = 16;
// Unrolled slots of the intersection loop (see Params::fast): three runs in scene order -- spheres, planes, spheres --
// which is how Cornell-box scenes are laid out (light first, walls, then the contents).
constexpr int kFastA = 2, kFastB = 8, kFastC = 6, kFastSlots = kFastA + kFastB + kFastC;
constexpr int kBlockThreads = PTK_BLOCK_THREADS;
constexpr int kBlockWarps = kBlockThreads / 32;
constexpr int kMaxCluster = 8;             // CTAs per cluster (portable limit): up to kMaxCluster * kBlockWarps sample slices per pixel
constexpr int kTileW = 8, kTileH = 4;   // pixels covered by one warp

template <typename R> struct alignas(16) V4 { R x, y, z, w; };
template <typename R> struct V3 { R x, y, z; };

// ---- scalar math, overloaded on the arithmetic type -------------------------------------------
__device__ __forceinline__ float m_rcp(float a) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ double m_rcp(double a) { return 1.0 / a; }
__device__ __forceinline__ float m_div(float a, float b) { return a * m_rcp(b); }
__device__ __forceinline__ double m_div(double a, double b) { return a / b; }
__device__ __forceinline__ float m_sqrt(float a) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ double m_sqrt(double a) { return sqrt(a); }
__device__ __forceinline__ float m_rsqrt(float a) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ double m_rsqrt(double a) { return 1.0 / sqrt(a); }
__device__ __forceinline__ float m_abs(float a) { return fabsf(a); }
__device__ __forceinline__ double m_abs(double a) { return fabs(a); }
__device__ __forceinline__ float m_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double m_min(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float m_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double m_max(double a, double b) { return fmax(a, b); }
// sin/cos of an angle in [0, 2*pi): the fp32 version shifts into [-pi, pi) where MUFU.SIN/COS have
// their best absolute accuracy (~5e-7) and flips the signs back
__device__ __forceinline__ void m_sincos_2pi(float a, float* s, float* c) {
    float sa, ca;
    __sincosf(a - 3.14159265358979f, &sa, &ca);
    *s = -sa; *c = -ca;
}
__device__ __forceinline__ void m_sincos_2pi(double a, double* s, double* c) { sincos(a, s, c); }
__device__ __forceinline__ void m_sincos(float a, float* s, float* c) { sincosf(a, s, c); }
__device__ __forceinline__ void m_sincos(double a, double* s, double* c) { sincos(a, s, c); }
__device__ __forceinline__ float m_acos(float a) { return acosf(a); }
__device__ __forceinline__ double m_acos(double a) { return acos(a); }
__device__ __forceinline__ float m_atan2(float a, float b) { return atan2f(a, b); }
__device__ __forceinline__ double m_atan2(double a, double b) { return atan2(a, b); }
__device__ __forceinline__ float m_fmod(float a, float b) { return fmodf(a, b); }
__device__ __forceinline__ double m_fmod(double a, double b) { return fmod(a, b); }
template <typename R> __device__ __forceinline__ R m_huge();
template <> __device__ __forceinline__ float m_huge<float>() { return __int_as_float(0x7f800000); }
template <> __device__ __forceinline__ double m_huge<double>() { return __longlong_as_double(0x7ff0000000000000LL); }

template <typename R> __device__ __forceinline__ V3<R> operator+(V3<R> a, V3<R> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename R> __device__ __forceinline__ V3<R> operator-(V3<R> a, V3<R> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename R> __device__ __forceinline__ V3<R> operator*(V3<R> a, R s) { return {a.x * s, a.y * s, a.z * s}; }
template <typename R> __device__ __forceinline__ V3<R> operator*(V3<R> a, V3<R> b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
template <typename R> __device__ __forceinline__ R dot(V3<R> a, V3<R> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename R> __device__ __forceinline__ V3<R> cross(V3<R> a, V3<R> b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename R> __device__ __forceinline__ V3<R> normalize(V3<R> a) { return a * m_rsqrt(dot(a, a)); }

// ---- device scene ------------------------------------------------------------------------------
// Scene objects are converted once on the host from the 1024-byte wire records (ptw_object).  All
// transforms on the path are affine (bottom row 0,0,0,1; points keep w=1, directions w=0), so only
// the top three rows of the inverse and the 3x3 of the inverse-transpose are kept.
//
// "Hot" half: what the intersection loop reads for EVERY object on EVERY segment; lives in the
// kernel parameter block (constant bank, uniform loads).
template <typename R> struct alignas(16) DObjHot {
    R inv[12];          // rows 0..2 of `inverse` (tracer.cl:547-548)
    R aux[6];           // cylinder: min_y, max_y; group: the object's own AABB bb_min.xyz, bb_max.xyz (tracer.cl:609)
    int type;           // 0 plane 1 sphere 2 cylinder 3 cube 4 group, anything else: never hit
    int node_begin, node_end;   // group: its range in the re-emitted reference nodes (node_lo / node_hi / node_parent)
    int pad;
};
// "Shade" half: read once per hit, indexed by the (lane-varying) hit object; lives in global memory.
// Ordered so that the common case (an untextured plane) touches only the first 64 bytes (fp32).
template <typename R> struct alignas(16) DObjShade {
    int type;
    int flags;          // bit0 textured, bit1 normal-mapped, bit2 needs the object-space hit point
    int tex_index, tex_index_nm;
    R reflectivity, refractive_index, min_y, max_y;
    R color[3];
    R emission[3];
    R plane_n[3];       // planes without a normal map: normalize(inverseTranspose * (0,1,0)), precomputed
    R tex_sx, tex_sy, tex_sx_nm, tex_sy_nm;
    R pad0;
    R inv[12];
    R invt[9];          // 3x3 of `inverseTranspose` (tracer.cl:953)
    R pad1[3];
};

// "Fast" objects: planes and spheres whose `inverse` is a similarity -- what Cornell-box scenes are made of.  They are
// tested by code that is unrolled per SLOT: slot K's record sits at a fixed offset of the kernel parameter block, so
// its coefficients reach the FFMAs through one uniform 128-bit constant load -- no per-lane load, no index arithmetic,
// no type dispatch (round 1's run loop spent ~17 % of all issue slots on exactly that).  A record is
//   plane:  (a, b, c, d) = row 1 of `inverse`: object-space y of a point is a*x + b*y + c*z + d (tracer.cl:478-483)
//   sphere: (cx, cy, cz, r^2): the world-space sphere |p - c|^2 = r^2 that the unit sphere maps to when `inverse` is a
//           similarity (scale s = 1/r): t solves |ro + t*rd - c|^2 = r^2, the reference's quadratic (tracer.cl:448-476)
//           divided through by s^2;
//           fast_kind 1, axis-aligned ellipsoid (`inverse` = diag(sx, sy, sz) * translate(-c), e.g. the flattened
//           ceiling light of the Cornell scenes): (cx, cy, cz, -) plus fast2 = (sx, sy, sz, -): the quadratic of
//           |S (ro + t*rd - c)|^2 = 1.
// Slots form three runs -- kFastA spheres, kFastB planes, kFastC spheres -- filled by the host with the longest
// subsequence of the scene's objects that fits "spheres, planes, spheres" in scene order; each run stops at its count.
// Fast objects therefore meet in scene order and ties resolve as upstream (first recorded wins, tracer.cl:731-739);
// everything else goes through the slow loop, which compares (t, object index) explicitly.
template <typename R> struct alignas(16) DFast { R a, b, c, d; };

template <typename R> struct DCam {
    R pixel_size, half_width, half_height, aperture, focal_length;
    R inv[12];          // rows 0..2 of the camera inverse
    int width, height;
};

struct DTex { const uchar4* data; int w, h, layers; };

// Next-event estimation (tracer.cl:786-825) samples every object whose emission.x > 0 as a sphere around the
// translation of its `transform`, scaled by the largest diagonal entry; transform[0] enters the attenuation.
template <typename R> struct alignas(16) DLight { R ox, oy, oz, scale, t0, er, eg, eb; int obj, pad[3]; };

// Per mesh (group) object: what the rebuilt BVH needs besides the object's hot record.
template <typename R> struct alignas(16) DMesh {
    R root_lo[3], root_hi[3];   // padded extent of the object's triangles: conservative pre-cull
    int bvh_root;               // index of the object's root node in bvh_*; -1 = no triangles
    int flags;                  // bit0: every reference node box contains its child boxes (checked on the host in R)
};

constexpr int kWide = 8;                       // children per node = lanes per ray in the cooperative walk
constexpr int kLeafTris = 8;                   // triangles per leaf (one per lane)
constexpr int kWideStack = 96;                 // deferred-child stack entries per ray (the host refuses trees that need more)
constexpr int kEmptyChild = (int)0x80000000;

template <typename R> struct Params {
    DFast<R> fast[kFastSlots];  // runs A (spheres) | B (planes) | C (spheres)
    DFast<R> fast2[kFastSlots]; // sphere slots of kind 1: the diagonal of `inverse`
    int fast_kind[kFastSlots];  // sphere slots: 0 = similarity, 1 = axis-aligned ellipsoid
    int fast_n[4];              // objects in run A, B, C
    int fast_obj[kFastSlots + 1];   // slot -> object index; entry kFastSlots ("no slot hit") maps to -1
    DObjHot<R> hot[kMaxObjects];    // read by the slow loop and the mesh walk only
    int slow_obj[kMaxObjects];  int n_slow;    // analytic objects outside the fast slots (general spheres, cylinders, cubes, overflow), scene order
    int slow_kind[kMaxObjects];                // 3: by its hot record; 0 plane / 1 similarity sphere / 2 ellipsoid: by slow_rec, with the slots' arithmetic
    DFast<R> slow_rec[kMaxObjects], slow_rec2[kMaxObjects];
    int mesh_obj[kMaxObjects];  int n_mesh;    // group objects with triangles, scene order
    const DObjShade<R>* shade;  int n_objects;
    // reference BVH (the caller's groups) re-emitted in the reference's visiting order: only the boxes and the
    // parent links are kept, for the "would the reference have tested this triangle" check
    const V4<R>* node_lo;       // (min.xyz, -)
    const V4<R>* node_hi;       // (max.xyz, -)
    const int* node_parent;     // -1 for a root child
    // rebuilt BVH: 8-wide, SAH, triangles in leaves only.  Node n, child c: wide[(n*8+c)*2] = (lo.xyz, child code),
    // wide[(n*8+c)*2+1] = (hi.xyz, -): padded box of the child.  Code >= 0: inner node; kEmptyChild: no child;
    // otherwise a leaf, ~code = (first slot << 4) | triangle count (<= 8).
    const V4<R>* wide;
    DMesh<R> mesh[kMaxObjects]; // per object (valid for groups)
    const int2* tri_info;       // per slot: (rank in the reference's recording order, reference node)
    const V4<R>* tri_test;      // 3 per slot: (p1.xyz,e1.x) (e1.yz,e2.xy) (e2.z,-,-,-)
    const V4<R>* tri_shade;     // 3 per slot: (n1.xyz,col.r) (n2.xyz,col.g) (n3.xyz,col.b)
    const R* lens;              // 2 per sample: sunflower(samples, 2, n), tracer.cl:235-248 (NULL without DoF)
    DCam<R> cam;
    DTex tex[3];
    const double* seeds;        // one per OWNED pixel: local rows, row-major
    const int* row_map;         // local row -> frame row
    const int* tile_order;      // launch position -> tile: expensive tiles first (see plan_tile_order / measured costs), or NULL
    unsigned* tile_cost;        // per tile: clocks / 256 the slowest warp of the tile took in this launch (feeds the next launch's order), or NULL
    void* out;                  // RGBA result, double4 (float4 when out_f32) per pixel, row-major, `width` pixels per row
    const int* out_row;         // local row -> row of `out`; NULL: rows are packed in local order.  With a map `out` may be a
                                // buffer shared by several devices (a peer device's memory, or another process's through
                                // CUDA IPC): the gather is then fused into this kernel's epilogue as NVLink stores
    double4* acc;               // progressive rendering: running per-pixel sums (local rows), or NULL
    int out_f32;
    int nee, caps;              // optional features the reference ships disabled (tracer.cl:1168, :437-444); 0 = upstream behaviour
    DLight<R> light[kMaxObjects];   // next-event estimation: the emissive objects, scene order
    int n_lights;
    int rows;                   // local rows rendered by this device
    int n_tiles;                // 8x4 pixel tiles of those rows
    int samples;                // total samples per pixel (enters the RNG seeding and the final weight)
    int sample_begin, sample_end;   // samples rendered by this launch: [begin, end) (0, samples for a full render)
    int slices;                 // sample slices per pixel: a power of two <= kBlockWarps * cluster size
    int slices_per_block;       // min(slices, kBlockWarps): the warps of a block are slices_per_block slices of
                                // kBlockWarps / slices_per_block tiles; a cluster holds all slices of its tiles
    int defer_below;
    int defer_max;              // mesh walk: iterations a warp with fewer than four rays at the mesh may put the walk off (0 = never)
    int lane_walk_min;          // mesh walk: rays per warp from which every lane walks its own ray (mesh_hit_lanes); 33 = never
    int stack_entries;          // mesh walk: entries of one 8-lane group's stack in dynamic shared memory (scene's worst case + 1)
    R pi;                       // (double)3.14159265359f, tracer.cl:1
    R eps;                      // 0.0001, tracer.cl:4
};

// ---- texture fetch: OpenCL 1.2 sampler NORMALIZED | REPEAT | LINEAR on RGBA8 (tracer.cl:829) -----
// Done by hand, unfused fp32 in the spec's order: CUDA's hardware filter interpolates with 8
// fractional bits, which would cost ~2e-3 of accuracy against the OpenCL result.
__device__ __forceinline__ float3 sample_rgba8(const DTex& t, float s, float tt, int layer) {
    if (t.data == nullptr) return make_float3(0.f, 0.f, 0.f);
    float u = __fmul_rn(__fsub_rn(s, floorf(s)), (float)t.w);
    float v = __fmul_rn(__fsub_rn(tt, floorf(tt)), (float)t.h);
    float um = __fsub_rn(u, 0.5f), vm = __fsub_rn(v, 0.5f);
    float fu = floorf(um), fv = floorf(vm);
    int i0 = (int)fu, j0 = (int)fv;
    int i1 = i0 + 1, j1 = j0 + 1;
    if (i0 < 0) i0 += t.w;
    if (i1 > t.w - 1) i1 -= t.w;
    if (j0 < 0) j0 += t.h;
    if (j1 > t.h - 1) j1 -= t.h;
    float a = __fsub_rn(um, fu), b = __fsub_rn(vm, fv);
    layer = max(0, min(layer, t.layers - 1));
    const uchar4* base = t.data + (size_t)layer * t.w * t.h;
    uchar4 c00 = __ldg(base + (size_t)j0 * t.w + i0), c10 = __ldg(base + (size_t)j0 * t.w + i1);
    uchar4 c01 = __ldg(base + (size_t)j1 * t.w + i0), c11 = __ldg(base + (size_t)j1 * t.w + i1);
    float w00 = __fmul_rn(__fsub_rn(1.0f, a), __fsub_rn(1.0f, b)), w10 = __fmul_rn(a, __fsub_rn(1.0f, b));
    float w01 = __fmul_rn(__fsub_rn(1.0f, a), b), w11 = __fmul_rn(a, b);
    // UNORM8 -> float is c / 255.0f correctly rounded.  One multiply by 1/255 plus one residual step
    // (e = c - q*255 exactly by FMA; q += e/255) gives that quotient bit for bit for all 256 inputs
    // (checked exhaustively in tests/test_oracle_golden.py) at 3 instructions instead of an IEEE divide.
    // (Measured and rejected: replacing the byte -> float conversion (I2F, XU pipe) by a byte permute into the mantissa
    // of 2^23 minus 2^23: two instructions for one, -7 % on the environment-map scene, -9 % on the textures scene.)
    auto unorm8 = [](unsigned char p) {
        const float c = (float)p, r = 0x1.010102p-8f;           // float(1/255)
        const float q = __fmul_rn(c, r);
        return __fmaf_rn(__fmaf_rn(-q, 255.0f, c), r, q);
    };
    auto mix = [&](unsigned char p00, unsigned char p10, unsigned char p01, unsigned char p11) {
        float r = __fmul_rn(w00, unorm8(p00));
        r = __fadd_rn(r, __fmul_rn(w10, unorm8(p10)));
        r = __fadd_rn(r, __fmul_rn(w01, unorm8(p01)));
        r = __fadd_rn(r, __fmul_rn(w11, unorm8(p11)));
        return r;
    };
    return make_float3(mix(c00.x, c10.x, c01.x, c11.x), mix(c00.y, c10.y, c01.y, c11.y), mix(c00.z, c10.z, c01.z, c11.z));
}

// 16/32-byte records through the read-only path
__device__ __forceinline__ V4<float> ldg4(const V4<float>* p) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    return {v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ V4<double> ldg4(const V4<double>* p) {
    double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    return {a.x, a.y, b.x, b.y};
}
__device__ __forceinline__ float ldg1(const float* p) { return __ldg(p); }
__device__ __forceinline__ double ldg1(const double* p) { return __ldg(p); }

// ---- geometry helpers ----------------------------------------------------------------------------
template <typename R> __device__ __forceinline__ V3<R> xf_point(const R* m, V3<R> p) {   // rows 0..2, w = 1
    return {m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
            m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]};
}
template <typename R> __device__ __forceinline__ V3<R> xf_dir(const R* m, V3<R> d) {     // rows 0..2, w = 0
    return {m[0] * d.x + m[1] * d.y + m[2] * d.z, m[4] * d.x + m[5] * d.y + m[6] * d.z, m[8] * d.x + m[9] * d.y + m[10] * d.z};
}

// Per-ray slab set-up for tracer.cl:250-268 checkAxis: t = (bound - origin) / direction when
// |direction| >= EPSILON, else (bound - origin) * HUGE_VAL.  Both cases are "(bound - origin) * k"
// with k = 1/direction or +inf, so k is computed once per ray and object instead of per box.
// (double keeps the true division for last-bit agreement with the reference arithmetic.)
template <typename R> struct Slab { R kx, ky, kz; bool dx, dy, dz; };
__device__ __forceinline__ Slab<float> make_slab(V3<float> d, float eps) {
    Slab<float> s;
    s.dx = m_abs(d.x) >= eps; s.dy = m_abs(d.y) >= eps; s.dz = m_abs(d.z) >= eps;
    s.kx = s.dx ? m_rcp(d.x) : m_huge<float>();
    s.ky = s.dy ? m_rcp(d.y) : m_huge<float>();
    s.kz = s.dz ? m_rcp(d.z) : m_huge<float>();
    return s;
}
__device__ __forceinline__ Slab<double> make_slab(V3<double> d, double eps) {
    Slab<double> s;
    s.dx = m_abs(d.x) >= eps; s.dy = m_abs(d.y) >= eps; s.dz = m_abs(d.z) >= eps;
    s.kx = s.ky = s.kz = m_huge<double>();
    return s;
}
__device__ __forceinline__ void axis_t(float o, float, float k, bool, float lo, float hi, float& tmin, float& tmax) {
    float t0 = (lo - o) * k, t1 = (hi - o) * k;
    bool sw = t0 > t1;
    tmin = sw ? t1 : t0; tmax = sw ? t0 : t1;
}
__device__ __forceinline__ void axis_t(double o, double d, double k, bool big, double lo, double hi, double& tmin, double& tmax) {
    double a = lo - o, b = hi - o;
    double t0 = big ? a / d : a * k, t1 = big ? b / d : b * k;
    bool sw = t0 > t1;
    tmin = sw ? t1 : t0; tmax = sw ? t0 : t1;
}
// tracer.cl:270-280 intersectRayWithBox
template <typename R>
__device__ __forceinline__ bool ray_box(V3<R> o, V3<R> d, const Slab<R>& s, R lx, R ly, R lz, R hx, R hy, R hz, R& tmin, R& tmax) {
    R x0, x1, y0, y1, z0, z1;
    axis_t(o.x, d.x, s.kx, s.dx, lx, hx, x0, x1);
    axis_t(o.y, d.y, s.ky, s.dy, ly, hy, y0, y1);
    axis_t(o.z, d.z, s.kz, s.dz, lz, hz, z0, z1);
    tmin = m_max(m_max(x0, y0), z0);
    tmax = m_min(m_min(x1, y1), z1);
    return tmin < tmax;
}

// tracer.cl:485-505
template <typename R> __device__ __forceinline__ R schlick(V3<R> eye, V3<R> n, R n1, R n2) {
    R c = dot(eye, n);
    if (n1 > n2) {
        R r = m_div(n1, n2);
        R sin2 = (r * r) * (R(1) - c * c);
        if (sin2 > R(1)) return R(1);
        c = m_sqrt(R(1) - sin2);
    }
    R t = m_div(n1 - n2, n1 + n2);
    R r0 = t * t;
    R k = R(1) - c;
    R k2 = k * k;
    return r0 + (R(1) - r0) * (k2 * k2 * k);
}
// tracer.cl:507-533
template <typename R> __device__ __forceinline__ V3<R> refracted(V3<R> eye, V3<R> n, R n1, R n2) {
    R ratio = m_div(n1, n2);
    R cos_i = dot(eye, n);
    R sin2 = (ratio * ratio) * (R(1) - cos_i * cos_i);
    if (sin2 > R(1)) return {R(0), R(0), R(0)};
    R cos_t = m_sqrt(R(1) - sin2);
    return n * (ratio * cos_i - cos_t) - eye * ratio;
}

// tracer.cl:113-175, constants as written there
template <typename R> __device__ __forceinline__ void cube_uv(V3<R> p, R& ou, R& ov) {
    R coord = m_max(m_max(m_abs(p.x), m_abs(p.y)), m_abs(p.z));
    const R third = R(0.333333), two3 = R(0.6666666);
    R u, v;
    if (coord == p.x) { u = m_fmod(R(1) - p.z, R(2)) / R(2); v = m_fmod(p.y + R(1), R(2)) / R(2); ou = R(0.5) + u * R(0.25); ov = two3 - v * third; }
    else if (coord == -p.x) { u = m_fmod(p.z + R(1), R(2)) / R(2); v = m_fmod(p.y + R(1), R(2)) / R(2); ou = u * R(0.25); ov = two3 - v * third; }
    else if (coord == p.y) { u = m_fmod(p.x + R(1), R(2)) / R(2); v = m_fmod(R(1) - p.z, R(2)) / R(2); ou = R(0.25) + u * R(0.25); ov = R(1) - v * third; }
    else if (coord == -p.y) { u = m_fmod(p.x + R(1), R(2)) / R(2); v = m_fmod(p.z + R(1), R(2)) / R(2); ou = R(0.25) + u * R(0.25); ov = v * third; }
    else if (coord == p.z) { u = m_fmod(p.x + R(1), R(2)) / R(2); v = m_fmod(p.y + R(1), R(2)) / R(2); ou = R(0.25) + u * R(0.25); ov = two3 - v * third; }
    else { u = m_fmod(R(1) - p.x, R(2)) / R(2); v = m_fmod(p.y + R(1), R(2)) / R(2); ou = R(0.75) + u * R(0.25); ov = two3 - v * third; }
}

// Closest-hit record kept in registers while scanning the scene.
template <typename R> struct Hit {
    R t;        // running minimum, starts at 1024 (tracer.cl:728)
    int obj;    // -1 = none
    int tri;    // winning triangle (groups)
    R u, v;     // its barycentrics (normal interpolation, tracer.cl:669)
};

// ---- mesh objects: rebuilt BVH, reference semantics ------------------------------------------------
constexpr unsigned kFullMask = 0xffffffffu;

// What the reference does for a group object (tracer.cl:598-720): object AABB, then for every root
// child an in-order walk of the caller's BVH in which EVERY triangle of EVERY node whose box chain the
// ray "hits" (tracer.cl:270-280: tmin < tmax, no t range, |d| < EPSILON -> +-HUGE_VAL) is tested and
// recorded; the winner is the smallest recorded t in (EPSILON, 1024), earliest recorded on ties.  The Go
// frontend's BVH keeps straddling triangles in inner nodes (46 % of the teapot's), so that walk tests
// hundreds of triangles per ray.
//
// Here the triangles of a group object are re-indexed by an 8-wide SAH BVH with leaves of <= 8
// triangles, built by the host layer over the SAME triangle records.  Two facts make the result the
// reference's:
//   (1) the rebuilt tree only ever culls: its boxes are padded supersets of their triangles and the
//       slab test keeps a box on any doubt (relative slack, NaN keeps), so every triangle whose
//       Moeller-Trumbore test could pass with EPSILON < t <= best-so-far is tested, with the reference's
//       arithmetic;
//   (2) a triangle that passes is accepted only if the reference would have tested it, i.e. if the ray
//       hits (reference rule, reference boxes) every node from the triangle's own reference node up to
//       its root child.  When parent boxes contain child boxes (checked on the host in the kernel's
//       arithmetic type) and no direction component is below EPSILON, the per-axis intervals of an
//       ancestor contain those of the node exactly -- every operation of the slab test is monotonic
//       under rounding -- so testing the triangle's own node box decides the whole chain; otherwise the
//       chain is walked through the parent links.
// Ties between equal t go to the lower rank (= position in the reference's recording order).
//
// Execution: only a few lanes of a warp have a ray that reaches a mesh on a given segment (ncu on the
// teapot scene: ~4 of 32), so a per-lane walk idles most of the warp (measured: 7 of 32 lanes active,
// profiles/r01_v8a_*).  Instead EIGHT LANES SHARE ONE RAY -- a warp walks up to four rays at a time:
// at an inner node lane c tests the box of child c; at a leaf lane c tests triangle c; the nearest child
// / closest candidate of a group is found with three shuffle-min steps; the other hit children go to a
// small per-ray stack in shared memory.  A walk is a handful of dependent steps instead of dozens, and
// the four groups of a warp run in lockstep, so all shuffles and ballots are warp-convergent.
template <typename R> struct IDir { R x, y, z; };
__device__ __forceinline__ IDir<float> inv_dir(V3<float> d) { return {m_rcp(d.x), m_rcp(d.y), m_rcp(d.z)}; }
__device__ __forceinline__ IDir<double> inv_dir(V3<double> d) { return {1.0 / d.x, 1.0 / d.y, 1.0 / d.z}; }
template <typename R> __device__ __forceinline__ R box_slack();
template <> __device__ __forceinline__ float box_slack<float>() { return 8e-6f; }
template <> __device__ __forceinline__ double box_slack<double>() { return 1e-13; }
__device__ __forceinline__ int child_code(float w) { return __float_as_int(w); }
__device__ __forceinline__ int child_code(double w) { return (int)w; }
// distance bound kept with a stacked child: a float that is <= the true value
__device__ __forceinline__ int stack_key(float tn) { return __float_as_int(tn); }
__device__ __forceinline__ int stack_key(double tn) { return __float_as_int(__double2float_rd(tn)); }

// Conservative slab interval of a padded box: [tn, tf] clipped to [0, limit]; "keep" unless provably empty.
// fmin/fmax drop a NaN operand (0 * inf on an axis the ray is parallel to), which only widens the interval.
template <typename R>
__device__ __forceinline__ bool keep_box(V3<R> o, IDir<R> k, R lx, R hx, R ly, R hy, R lz, R hz, R limit, R& tn) {
    const R x0 = (lx - o.x) * k.x, x1 = (hx - o.x) * k.x;
    const R y0 = (ly - o.y) * k.y, y1 = (hy - o.y) * k.y;
    const R z0 = (lz - o.z) * k.z, z1 = (hz - o.z) * k.z;
    tn = m_max(m_max(m_min(x0, x1), m_min(y0, y1)), m_max(m_min(z0, z1), R(0)));
    const R tf = m_min(m_min(m_max(x0, x1), m_max(y0, y1)), m_min(m_max(z0, z1), limit));
    return !(tn > tf + tf * box_slack<R>());
}

// (2) above.  `g` is the triangle's reference node.
template <typename R>
__device__ __forceinline__ bool reference_tests_node(const Params<R>& P, V3<R> o, V3<R> d, const Slab<R>& s, int g, bool whole_chain) {
    R t0, t1;
    do {
        const V4<R> lo = ldg4(&P.node_lo[g]), hi = ldg4(&P.node_hi[g]);
        if (!ray_box(o, d, s, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, t0, t1)) return false;
        g = whole_chain ? __ldg(&P.node_parent[g]) : -1;
    } while (g >= 0);
    return true;
}

template <typename R> __device__ __forceinline__ V3<R> shfl3(V3<R> v, int src) {
    return {__shfl_sync(kFullMask, v.x, src), __shfl_sync(kFullMask, v.y, src), __shfl_sync(kFullMask, v.z, src)};
}
// min over the 8 lanes of a group (xor offsets stay inside the group)
__device__ __forceinline__ float group_min(float v) {
    v = fminf(v, __shfl_xor_sync(kFullMask, v, 1)); v = fminf(v, __shfl_xor_sync(kFullMask, v, 2)); return fminf(v, __shfl_xor_sync(kFullMask, v, 4));
}
__device__ __forceinline__ double group_min(double v) {
    v = fmin(v, __shfl_xor_sync(kFullMask, v, 1)); v = fmin(v, __shfl_xor_sync(kFullMask, v, 2)); return fmin(v, __shfl_xor_sync(kFullMask, v, 4));
}
__device__ __forceinline__ int group_min(int v) {
    v = min(v, __shfl_xor_sync(kFullMask, v, 1)); v = min(v, __shfl_xor_sync(kFullMask, v, 2)); return min(v, __shfl_xor_sync(kFullMask, v, 4));
}

__device__ __forceinline__ unsigned group_min(unsigned v) {
    v = min(v, __shfl_xor_sync(kFullMask, v, 1)); v = min(v, __shfl_xor_sync(kFullMask, v, 2)); return min(v, __shfl_xor_sync(kFullMask, v, 4));
}
// The per-group traversal stacks live in dynamic shared memory and are addressed by their 32-bit shared-window
// address, computed once per thread (through a generic pointer the compiler re-derived the window base -- S2UR
// SR_CgaCtaId, MOV, LEA -- at every push and pop).
__device__ __forceinline__ void stack_store(unsigned addr, int code, int key) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" :: "r"(addr), "r"(code), "r"(key) : "memory");
}
__device__ __forceinline__ int2 stack_load(unsigned addr) {
    int2 e; asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y) : "r"(addr) : "memory"); return e;
}

#ifdef PTK_HIST
__device__ unsigned long long ptk_hist[40];      // tuning build only: histogram of rays per warp that enter a mesh walk
#endif
// One mesh object against the rays of the lanes with `want` set.  Called by all 32 lanes.  `ro`, `rd`:
// this lane's ray in world space.  `stk`: shared-window address of this lane's GROUP's stack.
//
// A step of the walk, per group: every lane does its share (inner node: lane c tests child c's box; leaf: lane c
// tests triangle c and, if it is a candidate, checks at once that the reference would have tested it), then
//   * the nearest hit child comes out of ONE unsigned min over packed keys -- the entry distance's float bits with
//     the lane number in the three low bits -- so the minimum names its lane (no second ballot, no find-first-set);
//     the other hit children are pushed with their key as the distance bound;
//   * accepted triangle candidates are reduced separately, in the kernel's arithmetic type and only on the (rare)
//     steps that have one: smallest t, then lowest rank (= the reference's recording order) with the lane packed
//     into the rank key.
template <typename R>
__device__ __forceinline__ void mesh_hit(const Params<R>& P, const DObjHot<R>& ob, const DMesh<R>& m, int j, V3<R> ro, V3<R> rd, bool want, int lane, Hit<R>& h, unsigned stk) {
    const R eps = P.eps;
    const int sub = lane & 7, gbase = lane & 24, grp = lane >> 3;
    const unsigned below = (1u << sub) - 1u;
    constexpr int kDone = 0x7fffffff;
    constexpr unsigned kNoKey = 0xffffffffu;
    unsigned todo = __ballot_sync(kFullMask, want);
#ifdef PTK_HIST
    if (lane == 0) atomicAdd(&ptk_hist[__popc(todo)], 1ull);
#endif
    while (todo) {                                                    // a round: up to four rays, one per group
        // owners of this round: the four lowest set bits of `todo`
        const unsigned t1 = todo & (todo - 1), t2 = t1 & (t1 - 1), t3 = t2 & (t2 - 1);
        const unsigned mine_bits = grp == 0 ? todo : grp == 1 ? t1 : grp == 2 ? t2 : t3;
        const bool active = mine_bits != 0u;
        const int owner = active ? __ffs(mine_bits) - 1 : lane;
        const unsigned round_bits = todo ^ (t3 & (t3 - 1));
        todo = t3 & (t3 - 1);
        // the owner's ray, taken from its world-space registers and moved to the object's space here (the matrix is a
        // uniform constant-bank operand): cheaper than keeping every lane's object-space ray alive across the rounds
        const V3<R> o = xf_point(ob.inv, shfl3(ro, owner)), d = xf_dir(ob.inv, shfl3(rd, owner));
        R ct = __shfl_sync(kFullMask, h.t, owner);
        const int cobj = __shfl_sync(kFullMask, h.obj, owner);
        const IDir<R> k = inv_dir(d);
        const Slab<R> s = make_slab(d, eps);
        const bool whole_chain = !(m.flags & 1) || !(s.dx && s.dy && s.dz);
        int crank = cobj > j ? 0x7fffffff : -1;       // equal t: an earlier object keeps the hit, a later one loses it to this mesh
        int cslot = -1;
        R cu = R(0), cv = R(0);
        int cur = active ? m.bvh_root : kDone, sp = 0;

        while (__any_sync(kFullMask, cur != kDone)) {
            unsigned key = kNoKey;                                    // inner: packed entry distance of this lane's child
            int code = 0;
            bool cand = false;                                        // leaf: this lane's triangle is an accepted candidate
            int rank = 0, slot = 0;
            R tt = R(0), tu = R(0), tv = R(0);
            if (cur >= 0 && cur != kDone) {                           // inner node: one child box per lane
                const V4<R> a = ldg4(&P.wide[(cur * kWide + sub) * 2]), b = ldg4(&P.wide[(cur * kWide + sub) * 2 + 1]);
                code = child_code(a.w);
                R tn;
                if (keep_box(o, k, a.x, b.x, a.y, b.y, a.z, b.z, ct * R(1.0001), tn) && code != kEmptyChild)
                    key = ((unsigned)stack_key(tn) & 0x7ffffff8u) | (unsigned)sub;
            } else if (cur < 0) {                                     // leaf: one triangle per lane
                const int lc = ~cur, count = lc & 15;
                slot = (lc >> 4) + sub;
                if (sub < count) {                                    // Moeller-Trumbore, tracer.cl:640-675
                    const V4<R> q0 = ldg4(&P.tri_test[3 * slot]), q1 = ldg4(&P.tri_test[3 * slot + 1]);
                    const V3<R> e2 = {q1.z, q1.w, ldg1(&P.tri_test[3 * slot + 2].x)};
                    const V3<R> e1 = {q0.w, q1.x, q1.y};
                    const V3<R> dxe2 = cross(d, e2);
                    const R det = dot(e1, dxe2);
                    const R f = m_rcp(det);
                    const V3<R> sv = {o.x - q0.x, o.y - q0.y, o.z - q0.z};
                    const R u = f * dot(sv, dxe2);
                    const V3<R> sxe1 = cross(sv, e1);
                    const R v = f * dot(d, sxe1);
                    const R t = f * dot(e2, sxe1);
                    const bool ok = !(m_abs(det) < eps) && !(u < R(0) || u > R(1)) && !(v < R(0) || (u + v) > R(1));
                    if (ok && t > eps && t <= ct) {
                        const int2 info = __ldg(&P.tri_info[slot]);   // (rank in the reference's recording order, reference node)
                        if ((t < ct || info.x < crank) && reference_tests_node(P, o, d, s, info.y, whole_chain)) {
                            cand = true; rank = info.x; tt = t; tu = u; tv = v;
                        }
                    }
                }
            }
            // nearest hit child of each group; the others go to the stack
            const unsigned kmin = group_min(key);
            const unsigned hit_bits = (__ballot_sync(kFullMask, key != kNoKey) >> gbase) & 0xffu;
            const int sel = (int)(kmin & 7u);
            const int next_code = __shfl_sync(kFullMask, code, gbase + sel);
            int next = kDone;                                          // kDone here = "pop"
            if (kmin != kNoKey) {
                next = next_code;
                const unsigned others = hit_bits & ~(1u << sel);
                if (key != kNoKey && sub != sel) stack_store(stk + 8u * (unsigned)(sp + __popc(others & below)), code, (int)(key & 0x7ffffff8u));
                sp += __popc(others);
            }
            // accepted triangle candidates: closest, then lowest rank
            if (__any_sync(kFullMask, cand)) {
                const R tbest = group_min(cand ? tt : m_huge<R>());
                const int rmin = group_min((cand && tt == tbest) ? (rank << 3) | sub : 0x7fffffff);
                const int src = gbase + (rmin & 7);
                const R wu = __shfl_sync(kFullMask, tu, src), wv = __shfl_sync(kFullMask, tv, src);
                const int wslot = __shfl_sync(kFullMask, slot, src);
                if (rmin != 0x7fffffff) { ct = tbest; crank = rmin >> 3; cslot = wslot; cu = wu; cv = wv; }
            }
            __syncwarp();                                              // stack writes visible to the group before any later pop
            if (cur != kDone && next == kDone) {                      // pop; children beyond the current best are dropped
                const R lim2 = ct * R(1.0001);
                while (sp > 0) {
                    --sp;
                    const int2 e = stack_load(stk + 8u * (unsigned)sp);
                    if (!(R(__int_as_float(e.y)) > lim2)) { next = e.x; break; }
                }
            }
            cur = next;
        }
        // hand the results back: owner number q of this round reads group q's registers
        const bool is_owner = ((round_bits >> lane) & 1u) != 0u;
        const int from = is_owner ? __popc(round_bits & ((1u << lane) - 1u)) * 8 : lane;
        const R rt = __shfl_sync(kFullMask, ct, from), ru = __shfl_sync(kFullMask, cu, from), rv = __shfl_sync(kFullMask, cv, from);
        const int rslot = __shfl_sync(kFullMask, cslot, from);
        if (is_owner && rslot >= 0) { h.t = rt; h.obj = j; h.tri = rslot; h.u = ru; h.v = rv; }
    }
}

// The same walk with ONE LANE PER RAY, for warps where many lanes reach the mesh at once (camera rays of a tile that
// looks at the mesh, a mirror in front of it).  Measured on B200 (histogram of rays per warp entering mesh_hit):
// cooperative rounds hold four rays, so a warp with 32 wanting lanes pays eight rounds -- 27 % (teapot), 42 % (gopher),
// 97 % (cube-map scene) of all rounds came from warps with nine or more rays.  Here every lane tests the (up to) eight
// children of its node itself, visits the nearest first and keeps the others on a private stack in local memory; leaves
// test their triangles one after the other with the same acceptance rule (closest t, then lowest rank, reference node
// chain).  No shuffles, no ballots; lanes diverge freely.  `o`, `d`: this lane's ray in the OBJECT's space.
// Threshold, measured (teapot / gopher / cube-map scene, Gpaths/s, at 8, 16, 24, 32 rays and "never"): 4.01 / 2.65 / 9.95,
// 4.13 / 2.84 / 9.91, 4.24 / 2.92 / 9.97, 4.26 / 2.94 / 9.71, 4.26 / 2.95 / 8.00 -- divergence and the local-memory
// stacks make the private walk pay only when (nearly) the whole warp wants the mesh.
constexpr int kLaneWalkMin = 30;
constexpr int kDeferMax = 3;                   // see closest_mesh
template <typename R>
__device__ __forceinline__ void mesh_hit_lanes(const Params<R>& P, const DMesh<R>& m, int j, V3<R> o, V3<R> d, const Slab<R>& s, Hit<R>& h) {
    const R eps = P.eps;
    constexpr int kDone = 0x7fffffff;
    const IDir<R> k = inv_dir(d);
    const bool whole_chain = !(m.flags & 1) || !(s.dx && s.dy && s.dz);
    R ct = h.t;
    int crank = h.obj > j ? 0x7fffffff : -1;      // equal t: an earlier object keeps the hit, a later one loses it to this mesh
    int cslot = -1;
    R cu = R(0), cv = R(0);
    int2 stk[kWideStack];
    int sp = 0, cur = m.bvh_root;
    while (cur != kDone) {
        int next = kDone;
        if (cur >= 0) {                                               // inner node: its children are packed at the front
            const R limit = ct * R(1.0001);
            int next_key = 0x7fffffff;
            for (int c = 0; c < kWide; ++c) {
                const V4<R> a = ldg4(&P.wide[(cur * kWide + c) * 2]), b = ldg4(&P.wide[(cur * kWide + c) * 2 + 1]);
                const int code = child_code(a.w);
                if (code == kEmptyChild) break;
                R tn;
                if (!keep_box(o, k, a.x, b.x, a.y, b.y, a.z, b.z, limit, tn)) continue;
                const int key = stack_key(tn) & 0x7fffffff;           // float bits of a lower bound of tn (tn >= 0)
                if (key < next_key) {                                 // nearer than the nearest so far: that one is deferred
                    if (next != kDone) stk[sp++] = make_int2(next, next_key);
                    next = code; next_key = key;
                } else stk[sp++] = make_int2(code, key);
            }
        } else {                                                      // leaf: its triangles in turn, Moeller-Trumbore (tracer.cl:640-675)
            const int lc = ~cur, count = lc & 15, first = lc >> 4;
            for (int i = 0; i < count; ++i) {
                const int slot = first + i;
                const V4<R> q0 = ldg4(&P.tri_test[3 * slot]), q1 = ldg4(&P.tri_test[3 * slot + 1]);
                const V3<R> e2 = {q1.z, q1.w, ldg1(&P.tri_test[3 * slot + 2].x)};
                const V3<R> e1 = {q0.w, q1.x, q1.y};
                const V3<R> dxe2 = cross(d, e2);
                const R det = dot(e1, dxe2);
                if (m_abs(det) < eps) continue;
                const R f = m_rcp(det);
                const V3<R> sv = {o.x - q0.x, o.y - q0.y, o.z - q0.z};
                const R u = f * dot(sv, dxe2);
                if (u < R(0) || u > R(1)) continue;
                const V3<R> sxe1 = cross(sv, e1);
                const R v = f * dot(d, sxe1);
                if (v < R(0) || (u + v) > R(1)) continue;
                const R t = f * dot(e2, sxe1);
                if (!(t > eps && t <= ct)) continue;
                const int2 info = __ldg(&P.tri_info[slot]);           // (rank in the reference's recording order, reference node)
                if ((t < ct || info.x < crank) && reference_tests_node(P, o, d, s, info.y, whole_chain)) {
                    ct = t; crank = info.x; cslot = slot; cu = u; cv = v;
                }
            }
        }
        if (next == kDone) {                                          // pop; children beyond the current best are dropped
            const R lim2 = ct * R(1.0001);
            while (sp > 0) {
                const int2 e = stk[--sp];
                if (!(R(__int_as_float(e.y)) > lim2)) { next = e.x; break; }
            }
        }
        cur = next;
    }
    if (cslot >= 0) { h.t = ct; h.obj = j; h.tri = cslot; h.u = cu; h.v = cv; }
}

// All mesh objects of the scene (tracer.cl:598-720), after the analytic objects.  Called by all 32 lanes.
//
// Deferral.  A cooperative round costs about as much as a whole analytic segment whether one or four of its groups
// have a ray, and most rounds are under-filled: on the teapot 52 % of all rounds came from warps with one to three rays
// at the mesh (gopher 35 %).  So a warp with fewer than four such rays may put the walk off: those lanes simply do
// nothing this iteration -- no shading, no state change -- and repeat the segment in the next one (the analytic scan runs
// for the other lanes anyway, and everything is a pure function of the path state), by which time more lanes have
// usually reached the mesh and the rays share a round.  `defer_age` (warp-uniform) bounds how long a warp waits; it does
// not wait when few other lanes are alive to make progress.  Returns true for a lane whose segment is put off.
template <typename R>
__device__ __forceinline__ bool closest_mesh(const Params<R>& P, V3<R> ro, V3<R> rd, bool live, int lane, Hit<R>& h, unsigned stk, int& defer_age) {
    for (int q = 0; q < P.n_mesh; ++q) {
        const int j = P.mesh_obj[q];
        const DObjHot<R>& ob = P.hot[j];
        const DMesh<R>& m = P.mesh[j];
        const V3<R> o = xf_point(ob.inv, ro), d = xf_dir(ob.inv, rd);
        const Slab<R> s = make_slab(d, P.eps);
        R t0, t1, tn;
        // object AABB under the reference rule (tracer.cl:609) AND the padded extent of the triangles within reach
        // of the closest hit so far; NaN / inf rays hit nothing upstream
        const bool finite = (o.x + o.y + o.z + d.x + d.y + d.z) * R(0) == R(0);
        const bool want = live && finite && ray_box(o, d, s, ob.aux[0], ob.aux[1], ob.aux[2], ob.aux[3], ob.aux[4], ob.aux[5], t0, t1) &&
                          keep_box(o, inv_dir(d), m.root_lo[0], m.root_hi[0], m.root_lo[1], m.root_hi[1], m.root_lo[2], m.root_hi[2], h.t * R(1.0001), tn);
        const int n_want = __popc(__ballot_sync(kFullMask, want));
        if (P.n_mesh == 1 && n_want > 0 && n_want < P.defer_below && defer_age < P.defer_max &&
            __popc(__ballot_sync(kFullMask, live && !want)) >= 8) {
            ++defer_age;
            return want;
        }
        defer_age = 0;
        // (fp32 only: in double the private stacks and the second walk cost the kernel more registers than they save time)
        if (sizeof(R) == 4 && n_want >= P.lane_walk_min) {
            if (want) mesh_hit_lanes<R>(P, m, j, o, d, s, h);
            __syncwarp();
        } else {
            mesh_hit<R>(P, ob, m, j, ro, rd, want, lane, h, stk);
        }
    }
    return false;
}

// A candidate of the slow loop: objects there are NOT visited in scene order relative to the fast slots, so the
// reference's "first recorded wins" (tracer.cl:731-739) is spelled out: on equal t the lower object index wins.
template <typename R> __device__ __forceinline__ void offer_ordered(Hit<R>& h, R t, int obj, R eps) {
    if (t > eps && (t < h.t || (t == h.t && obj < h.obj))) { h.t = t; h.obj = obj; }
}

// Fast objects (see DFast).  The arithmetic is spelled with explicit fused multiply-adds and shared by the unrolled
// slots and the slow loop's overflow entries, so an object gives bit-identical t on either path: coincident objects
// (two coplanar planes, a duplicated sphere) tie EXACTLY wherever they land, and the tie goes to the lower index as
// upstream.  `a` = dot(rd, rd) and `inv_a` = 1 / a are per-ray values shared by every similarity sphere.
__device__ __forceinline__ float m_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double m_fma(double a, double b, double c) { return fma(a, b, c); }
// sqrt(disc) for disc > 0, NaN otherwise (the reference records roots only when disc > 0, tracer.cl:465): in float
// disc * rsqrt(disc) does it without a compare -- rsqrt(0) = inf, inf * 0 = NaN, rsqrt(negative) = NaN
__device__ __forceinline__ float sqrt_pos(float disc) { return disc * m_rsqrt(disc); }
__device__ __forceinline__ double sqrt_pos(double disc) { return disc > 0.0 ? sqrt(disc) : __longlong_as_double(0x7ff8000000000000LL); }

// plane, tracer.cl:478-483: t = -o'.y / d'.y, valid when |d'.y| > EPSILON (`dy` returned for that test)
template <typename R> __device__ __forceinline__ R plane_t(const DFast<R>& f, V3<R> ro, V3<R> rd, R& dy) {
    const R oy = m_fma(f.a, ro.x, m_fma(f.b, ro.y, m_fma(f.c, ro.z, f.d)));
    dy = m_fma(f.a, rd.x, m_fma(f.b, rd.y, f.c * rd.z));
    return -oy * m_rcp(dy);
}
// sphere, tracer.cl:448-476 (both roots recorded when disc > 0): returns the first root beyond EPSILON -- the pair's
// winner since t0 <= t1 -- and the larger root in `t1`; without real roots both are NaN and fail every comparison
template <typename R> __device__ __forceinline__ R sphere_t(const DFast<R>& f, const DFast<R>& g, int kind, V3<R> ro, V3<R> rd, R a, R inv_a, R eps, R& t1) {
    V3<R> oc = {ro.x - f.a, ro.y - f.b, ro.z - f.c}, dd = rd;
    R c1 = f.d;
    if (kind != 0) {                                             // ellipsoid: scale the offset and the direction, own a and 1/a
        oc = {oc.x * g.a, oc.y * g.b, oc.z * g.c};
        dd = {rd.x * g.a, rd.y * g.b, rd.z * g.c};
        a = m_fma(dd.x, dd.x, m_fma(dd.y, dd.y, dd.z * dd.z)); inv_a = m_rcp(a); c1 = R(1);
    }
    const R hb = m_fma(dd.x, oc.x, m_fma(dd.y, oc.y, dd.z * oc.z));
    const R c = m_fma(oc.x, oc.x, m_fma(oc.y, oc.y, m_fma(oc.z, oc.z, -c1)));
    const R sq = sqrt_pos(m_fma(hb, hb, -(a * c)));
    const R t0 = (-hb - sq) * inv_a;
    t1 = (-hb + sq) * inv_a;
    return t0 > eps ? t0 : t1;
}

// Unrolled slots.  The winner is tracked as a SLOT number (an immediate); closest_analytic maps it to the object index once.
template <typename R, int SLOT>
__device__ __forceinline__ void fast_plane(const Params<R>& P, V3<R> ro, V3<R> rd, R eps, Hit<R>& h) {
    R dy;
    const R t = plane_t(P.fast[SLOT], ro, rd, dy);
    if (m_abs(dy) > eps && t > eps && t < h.t) { h.t = t; h.obj = SLOT; }
}
template <typename R, int SLOT>
__device__ __forceinline__ void fast_sphere(const Params<R>& P, V3<R> ro, V3<R> rd, R a, R inv_a, R eps, Hit<R>& h) {
    R t1;
    const R t = sphere_t(P.fast[SLOT], P.fast2[SLOT], P.fast_kind[SLOT], ro, rd, a, inv_a, eps, t1);
    if (t1 > eps && t < h.t) { h.t = t; h.obj = SLOT; }
}
template <typename R, int K, int END>
__device__ __forceinline__ void run_planes(const Params<R>& P, int n, V3<R> ro, V3<R> rd, R eps, Hit<R>& h) {
    if constexpr (K < END) {
        if (K - kFastA >= n) return;
        fast_plane<R, K>(P, ro, rd, eps, h);
        run_planes<R, K + 1, END>(P, n, ro, rd, eps, h);
    }
}
template <typename R, int K, int BEGIN, int END>
__device__ __forceinline__ void run_spheres(const Params<R>& P, int n, V3<R> ro, V3<R> rd, R a, R inv_a, R eps, Hit<R>& h) {
    if constexpr (K < END) {
        if (K - BEGIN >= n) return;
        fast_sphere<R, K>(P, ro, rd, a, inv_a, eps, h);
        run_spheres<R, K + 1, BEGIN, END>(P, n, ro, rd, a, inv_a, eps, h);
    }
}

// One analytic object of the slow loop against one ray (`type` is warp-uniform).
template <typename R>
__device__ __forceinline__ void test_object(const DObjHot<R>& ob, int j, int type, V3<R> ro, V3<R> rd, R eps, bool caps, Hit<R>& h) {
    if (type == 0) {                                             // plane, tracer.cl:478-483
        R oy = ob.inv[4] * ro.x + ob.inv[5] * ro.y + ob.inv[6] * ro.z + ob.inv[7];
        R dy = ob.inv[4] * rd.x + ob.inv[5] * rd.y + ob.inv[6] * rd.z;
        R t = m_div(-oy, dy);
        if (m_abs(dy) > eps) offer_ordered(h, t, j, eps);
    } else if (type == 1) {                                      // sphere, tracer.cl:448-476
        // Roots recorded only when the discriminant is strictly positive.  Half-b form: with hb = dot(d,o) the
        // reference's b = 2*hb, b*b - 4*a*c = 4*(hb*hb - a*c) and 2*a differ from these by exact powers of two only.
        V3<R> o = xf_point(ob.inv, ro), d = xf_dir(ob.inv, rd);
        const R a = dot(d, d), hb = dot(d, o), c = dot(o, o) - R(1);
        const R disc4 = hb * hb - a * c;
        if (disc4 > R(0)) {
            const R sq = m_sqrt(disc4), inv_a = m_rcp(a);
            offer_ordered(h, (-hb - sq) * inv_a, j, eps);
            offer_ordered(h, (-hb + sq) * inv_a, j, eps);
        }
    } else if (type == 2) {                                      // cylinder side, caps off, tracer.cl:396-446
        V3<R> o = xf_point(ob.inv, ro), d = xf_dir(ob.inv, rd);
        R a = d.x * d.x + d.z * d.z;
        if (!(m_abs(a) < eps)) {
            R b = R(2) * o.x * d.x + R(2) * o.z * d.z;
            R c = o.x * o.x + o.z * o.z - R(1);
            R disc = b * b - R(4) * a * c;
            if (!(disc < R(0))) {
                R sq = m_sqrt(disc), inv_den = m_rcp(R(2) * a);
                R t0 = (-b - sq) * inv_den, t1 = (-b + sq) * inv_den;
                R y0 = o.y + t0 * d.y, y1 = o.y + t1 * d.y;
                if (y0 > ob.aux[0] && y0 < ob.aux[1]) offer_ordered(h, t0, j, eps);
                if (y1 > ob.aux[0] && y1 < ob.aux[1]) offer_ordered(h, t1, j, eps);
                if (caps && !(m_abs(d.y) < eps)) {                  // end caps, tracer.cl:282-310 (upstream: disabled at :437-444)
                    const R tc0 = m_div(ob.aux[0] - o.y, d.y), tc1 = m_div(ob.aux[1] - o.y, d.y);
                    const R x0 = o.x + tc0 * d.x, z0 = o.z + tc0 * d.z, x1 = o.x + tc1 * d.x, z1 = o.z + tc1 * d.z;
                    if (x0 * x0 + z0 * z0 <= R(1) && tc0 > R(0)) offer_ordered(h, tc0, j, eps);
                    if (x1 * x1 + z1 * z1 <= R(1) && tc1 > R(0)) offer_ordered(h, tc1, j, eps);
                }
            }
        }
    } else if (type == 3) {                                      // cube, tracer.cl:378-394
        V3<R> o = xf_point(ob.inv, ro), d = xf_dir(ob.inv, rd);
        Slab<R> s = make_slab(d, eps);
        R tmin, tmax;
        ray_box(o, d, s, R(-1), R(-1), R(-1), R(1), R(1), R(1), tmin, tmax);
        if (!(tmin > tmax)) { offer_ordered(h, tmin, j, eps); offer_ordered(h, tmax, j, eps); }
    }
}

// Analytic objects against one ray: tracer.cl:537-597 of findClosestIntersection.  Planes and similarity-transformed
// spheres go through the unrolled fast slots in scene order; what remains (cylinders, cubes, stretched spheres, more
// than kFastSlots fast objects) through a loop over the constant-bank records.  Mesh objects: closest_mesh.
template <typename R>
__device__ __forceinline__ void closest_analytic(const Params<R>& P, V3<R> ro, V3<R> rd, Hit<R>& h) {
    const R eps = P.eps;
    h.t = R(1024); h.obj = kFastSlots; h.tri = -1; h.u = R(0); h.v = R(0);
    const R a = m_fma(rd.x, rd.x, m_fma(rd.y, rd.y, rd.z * rd.z)), inv_a = m_rcp(a);
    run_spheres<R, 0, 0, kFastA>(P, P.fast_n[0], ro, rd, a, inv_a, eps, h);
    run_planes<R, kFastA, kFastA + kFastB>(P, P.fast_n[1], ro, rd, eps, h);
    run_spheres<R, kFastA + kFastB, kFastA + kFastB, kFastSlots>(P, P.fast_n[2], ro, rd, a, inv_a, eps, h);
    h.obj = P.fast_obj[h.obj];                                   // slot -> object index (kFastSlots -> -1: nothing hit yet)
    for (int k = 0; k < P.n_slow; ++k) {
        const int j = P.slow_obj[k], kind = P.slow_kind[k];
        if (kind == 3) test_object<R>(P.hot[j], j, P.hot[j].type, ro, rd, eps, P.caps != 0, h);
        else if (kind == 0) {                                    // a plane beyond the unrolled slots: same arithmetic as there
            R dy;
            const R t = plane_t(P.slow_rec[k], ro, rd, dy);
            if (m_abs(dy) > eps) offer_ordered(h, t, j, eps);
        } else {                                                 // a fast-class sphere beyond the unrolled slots
            R t1;
            const R t = sphere_t(P.slow_rec[k], P.slow_rec2[k], kind - 1, ro, rd, a, inv_a, eps, t1);
            if (t1 > eps) offer_ordered(h, t, j, eps);
        }
    }
}

// Per-thread path state of the segment loop.
template <typename R> struct Path {
    V3<R> ro, rd;           // current ray
    V3<R> mask, accum;      // tracer.cl:1116
    unsigned n;             // sample index (tracer.cl:867)
    unsigned b, effective;  // bounce counters (tracer.cl:873-884)
    bool inside;
};

// rayForPixel, tracer.cl:745-779, for sample `gn` of pixel (px, py)
template <typename R, int RNG>
__device__ __forceinline__ void camera_ray(const Params<R>& P, R px, R py, float fgi, float fgi2, unsigned gn, V3<R> cam_origin, V3<R>& nxo, V3<R>& nxd) {
    float jx = noise3d<RNG>(fgi, (float)gn, fgi2);
    float jy = noise3d<RNG>(fgi, fgi2, (float)gn);
    R xo = P.cam.pixel_size * (px + R(jx));
    R yo = P.cam.pixel_size * (py + R(jy));
    V3<R> in_view = {P.cam.half_width - xo, P.cam.half_height - yo, R(-1)};
    V3<R> pixel = xf_point(P.cam.inv, in_view);
    nxo = cam_origin;
    nxd = normalize(pixel - nxo);
    if (P.lens != nullptr) {                                                          // aperture != 0
        V3<R> pos = nxo + nxd * P.cam.focal_length;
        R sx = ldg1(&P.lens[2 * gn]), sy = ldg1(&P.lens[2 * gn + 1]);                   // NaN at sample 0 when samples >= 3: kept
        V3<R> no = {nxo.x + sy * P.cam.aperture, nxo.y + sx * P.cam.aperture, nxo.z}; // x/y swap as upstream
        nxd = pos - no;                                                               // left unnormalised
        nxo = no;
    }
}

// One surface interaction: normal, material branch, next ray, fused mask/accumulate (tracer.cl:895-1107,
// 1116-1176).  Returns true when the path ends here.
// What next-event estimation needs from a shaded bounce (the reference's stored `bounce`, tracer.cl:72-80).
template <typename R> struct NeeInfo { V3<R> point, normal, color, mask; unsigned b; bool on; };

template <typename R, int RNG, bool NEE>
__device__ __forceinline__ bool shade_hit(const Params<R>& P, const Hit<R>& h, Path<R>& s, float fgi, NeeInfo<R>& ni) {
    const R eps = P.eps, pi = P.pi;
    V3<R>& ro = s.ro; V3<R>& rd = s.rd;
    const unsigned n = s.n;
    const DObjShade<R>& ob = P.shade[h.obj];
    const int type = ob.type;
    V3<R> position = ro + rd * h.t;
    V3<R> eye = {-rd.x, -rd.y, -rd.z};
    V3<R> lp = {R(0), R(0), R(0)};
    if (ob.flags & 4) lp = xf_point(ob.inv, position);                    // spheres, cylinders, cubes, textured planes
    V3<R> nv;
    V3<R> tri_color = {R(0), R(0), R(0)};
    if (type == 0 && !(ob.flags & 2)) {
        // plane without normal map: object normal (0,1,0) -> world normal is a constant of the
        // object, normalize(inverseTranspose * (0,1,0)), precomputed on the host (tracer.cl:913, 953-955)
        nv = {ob.plane_n[0], ob.plane_n[1], ob.plane_n[2]};
    } else {
        V3<R> on;
        if (type == 0) {                                                  // tracer.cl:906-911
            float3 c = sample_rgba8(P.tex[0], (float)(m_abs(lp.x) * ob.tex_sx_nm), (float)(m_abs(lp.z) * ob.tex_sy_nm), ob.tex_index_nm);
            on = normalize(V3<R>{R(c.x), R(c.y), R(c.z)});
        } else if (type == 1) {
            on = lp;                                                      // tracer.cl:919-920
        } else if (type == 2) {                                           // tracer.cl:924-932
            R dist = lp.x * lp.x + lp.z * lp.z;
            if (dist < R(1) && lp.y >= ob.max_y - eps) on = {R(0), R(1), R(0)};
            else if (dist < R(1) && lp.y <= ob.min_y + eps) on = {R(0), R(-1), R(0)};
            else on = {lp.x, R(0), lp.z};
        } else if (type == 3) {                                           // tracer.cl:938-946
            R ax = m_abs(lp.x), ay = m_abs(lp.y), az = m_abs(lp.z);
            R maxc = m_max(m_max(ax, ay), az);
            if (maxc == ax) on = {lp.x, R(0), R(0)};
            else if (maxc == ay) on = {R(0), lp.y, R(0)};
            else on = {R(0), R(0), lp.z};
        } else {                                                          // tracer.cl:669, 949
            const V4<R> s0 = ldg4(&P.tri_shade[3 * h.tri]), s1 = ldg4(&P.tri_shade[3 * h.tri + 1]), s2 = ldg4(&P.tri_shade[3 * h.tri + 2]);
            R w = R(1) - h.u - h.v;
            on = {s1.x * h.u + s2.x * h.v + s0.x * w, s1.y * h.u + s2.y * h.v + s0.y * w, s1.z * h.u + s2.z * h.v + s0.z * w};
            tri_color = {s0.w, s1.w, s2.w};
        }
        nv = {ob.invt[0] * on.x + ob.invt[1] * on.y + ob.invt[2] * on.z,
              ob.invt[3] * on.x + ob.invt[4] * on.y + ob.invt[5] * on.z,
              ob.invt[6] * on.x + ob.invt[7] * on.y + ob.invt[8] * on.z};   // tracer.cl:953-955
        nv = normalize(nv);
    }
    if (dot(eye, nv) < R(0)) nv = nv * R(-1);                                  // tracer.cl:962-964
    V3<R> over = position + nv * eps;
    const V3<R> under = position - nv * eps;

    // material branch, tracer.cl:975-1062
    R cosine = R(1);
    bool entering = false, exiting = false, reflecting = false;
    const R refl = ob.reflectivity, ri = ob.refractive_index;
    bool mirror = false;
    if (refl != R(0) && R(noise3d<RNG>(fgi, (float)n, (float)s.b)) < refl) {
        mirror = true;
    } else if (ri == R(-1)) {                                                 // thin glass
        if (schlick(eye, nv, R(1), R(1.5)) < R(noise3d<RNG>(fgi, (float)(n * n), (float)s.b))) over = under;
        else mirror = true;
    } else if (ri != R(1)) {
        R rnd = R(noise3d<RNG>(fgi, (float)(n * n), (float)s.b));
        if (!s.inside) {
            if (schlick(eye, nv, R(1), ri) < rnd) { rd = refracted(eye, nv, R(1), ri); over = under; s.inside = true; entering = true; }
            else mirror = true;
        } else {
            if (schlick(eye, nv, ri, R(1)) < rnd) { rd = refracted(eye, nv, ri, R(1)); over = under; s.inside = false; exiting = true; }
            else mirror = true;
        }
    } else {                                                                  // diffuse, tracer.cl:348-366
        R rand1 = R(2) * pi * R(noise3d<RNG>(fgi, (float)s.b, (float)n));
        R rand2 = R(noise3d<RNG>((float)s.b, (float)n, fgi));
        R rand2s = m_sqrt(rand2);
        // u = normalize(cross(axis, n)) with axis = (0,1,0) if |n.x| > 0.1 else (1,0,0); the cross
        // product with a unit axis is written out (same values: the other terms are exact zeros)
        const bool ay_axis = m_abs(nv.x) > R(0.1);
        V3<R> uu = ay_axis ? V3<R>{nv.z, R(0), -nv.x} : V3<R>{R(0), -nv.z, nv.y};
        uu = normalize(uu);
        V3<R> vv = cross(nv, uu);
        R s1, c1;
        m_sincos_2pi(rand1, &s1, &c1);
        rd = uu * (c1 * rand2s) + vv * (s1 * rand2s) + nv * m_sqrt(R(1) - rand2);
        cosine = dot(rd, nv);
    }
    if (mirror) {                                                             // tracer.cl:985-988
        R ds = dot(rd, nv);
        rd = rd - nv * (R(2) * ds);
        reflecting = true;
    }
    ro = over;

    // surface colour, tracer.cl:1071-1096
    V3<R> colr, emis;
    if (type == 4) { colr = tri_color; emis = {R(0), R(0), R(0)}; }
    else {
        colr = {ob.color[0], ob.color[1], ob.color[2]};
        emis = {ob.emission[0], ob.emission[1], ob.emission[2]};
        if (ob.flags & 1) {
            // texture coordinates per shape, then ONE bilinear fetch: a single inlined copy of the exact filter -- the
            // kernel's code size is what the textured scenes stall on (+27 % on the textures scene).  The fp64 kernels keep
            // a fetch per shape: merged, their register allocation came out worse (-7 % .. -17 % on scenes without textures).
            float tu = 0.f, tv = 0.f;
            int cls = -1;
            if (type == 0) { tu = (float)(lp.x * ob.tex_sx); tv = (float)(lp.z * ob.tex_sy); cls = 0; }
            else if (type == 1) {                                             // sphericalMap, tracer.cl:178-213
                R theta = m_atan2(lp.x, lp.z);
                R radius = sqrt(dot(lp, lp));
                R phi = m_acos(lp.y / radius);
                R su = R(1) - (theta / (R(2) * pi) + R(0.5));
                R sv = R(1) - phi / pi;
                tu = (float)su; tv = (float)(R(1) - sv); cls = 1;
            } else if (type == 3) {
                R cu, cv;
                cube_uv(lp, cu, cv);
                tu = (float)cu; tv = (float)cv; cls = 2;
            }
            if (sizeof(R) == 4) {
                if (cls >= 0) {
                    float3 c = sample_rgba8(P.tex[cls], tu, tv, ob.tex_index);
                    colr = {R(c.x), R(c.y), R(c.z)};
                }
            } else if (cls == 0) {
                float3 c = sample_rgba8(P.tex[0], tu, tv, ob.tex_index);
                colr = {R(c.x), R(c.y), R(c.z)};
            } else if (cls == 1) {
                float3 c = sample_rgba8(P.tex[1], tu, tv, ob.tex_index);
                colr = {R(c.x), R(c.y), R(c.z)};
            } else if (cls == 2) {
                float3 c = sample_rgba8(P.tex[2], tu, tv, ob.tex_index);
                colr = {R(c.x), R(c.y), R(c.z)};
            }
        }
    }

    // fused shading, tracer.cl:1116-1176: refraction bounces are skipped; an emitter adds
    // mask*emission (or, when hit directly by the camera ray, replaces accum by its colour)
    if (NEE) {          // tracer.cl:1168 runs for every stored bounce that is neither a refraction nor an emitter, with the mask BEFORE this bounce
        ni.on = !(entering || exiting) && !(emis.x > R(0));
        ni.point = position; ni.normal = nv; ni.color = colr; ni.mask = s.mask; ni.b = s.b;
    }
    if (!(entering || exiting)) {
        s.accum = s.accum + s.mask * emis;
        if (emis.x > R(0)) { if (s.b == 0) s.accum = colr; }
        s.mask = s.mask * colr;
        s.mask = s.mask * cosine;
    }
    if (!entering && !exiting && !reflecting) s.effective++;                    // tracer.cl:1099-1101
    s.b++;
    return (ob.emission[0] > R(0)) || !(s.b < 10u && s.effective < 4u);        // tracer.cl:1107, 884
}

// Thread-block cluster primitives, spelled in PTX on 32-bit shared-window addresses: going through cooperative_groups'
// generic pointers made the compiler address EVERY shared array of the kernel through the cluster window (an S2UR
// SR_CgaCtaId + LEA per access group in the hot loop).
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_map_shared(unsigned addr, unsigned rank) {
    unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ double cluster_ld_f64(unsigned addr) {
    double v; asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory"); return v;
}

// Frame geometry of a thread.  A warp covers an 8x4 pixel tile for one sample slice.  The warps of a block are
// slices_per_block slices of kBlockWarps / slices_per_block consecutive tiles; the blocks of a cluster hold the
// remaining slices of the same tiles (slice = cluster rank * slices_per_block + slice within the block).
struct PixelSlot { int lx, ly, gy, slice, tile; bool has_pixel; };
template <typename R> __device__ __forceinline__ PixelSlot pixel_slot(const Params<R>& P, unsigned cluster_rank, unsigned cluster_size) {
    const int W = P.cam.width;
    const int tiles_x = (W + kTileW - 1) / kTileW;
    const int spb = P.slices_per_block;                       // 1, 2 or 4 (power of two <= kBlockWarps)
    const int tiles_per_block = kBlockWarps / spb;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = (int)(blockIdx.x / cluster_size) * tiles_per_block + w / spb;
    const int tile = P.tile_order ? (slot < P.n_tiles ? P.tile_order[slot] : slot) : slot;
    PixelSlot s;
    s.tile = tile;
    s.slice = (int)cluster_rank * spb + (w % spb);
    s.lx = (tile % tiles_x) * kTileW + (lane & (kTileW - 1));
    s.ly = (tile / tiles_x) * kTileH + (lane / kTileW);
    s.has_pixel = s.lx < W && s.ly < P.rows;
    s.gy = s.has_pixel ? P.row_map[s.ly] : 0;
    return s;
}

// ---- the kernel ----------------------------------------------------------------------------------
// GROUPS = the scene contains mesh objects: only then is the cooperative BVH walk (and its shared-memory
// stacks) compiled in.
// NEE = next-event estimation compiled in (the reference ships it commented out; ptc_job.features enables it).
template <typename R, int RNG, bool GROUPS, bool NEE = false>
__global__ void __launch_bounds__(kBlockThreads, NEE ? (sizeof(R) == 8 ? 2 : 4) : (sizeof(R) == 8 ? (GROUPS ? PTK_MESH_MIN_BLOCKS_F64 : PTK_MIN_BLOCKS_F64) : (GROUPS ? PTK_MESH_MIN_BLOCKS : PTK_MIN_BLOCKS))) trace_kernel(const __grid_constant__ Params<R> P) {
    extern __shared__ int2 mesh_stacks[];       // GROUPS: one stack of P.stack_entries per 8-lane group (sized by the host from the scene's BVH)
    const unsigned cluster_size = cluster_nctarank(), cluster_rank = cluster_ctarank();
    const long long clock_begin = clock64();
    const PixelSlot px = pixel_slot(P, cluster_rank, cluster_size);
    const int lane = threadIdx.x & 31;
    const int W = P.cam.width;
    const int slice = px.slice;
    const unsigned samples = (unsigned)P.samples;
    const double seed = px.has_pixel ? P.seeds[(size_t)px.ly * W + px.lx] : 0.0;
    const float fgi = (float)(seed / (double)P.n_objects);     // tracer.cl:840
    const float fgi2 = (float)(seed / (double)samples);        // tracer.cl:841
    const R fx = R(px.lx), fy = R(px.gy);
    const V3<R> cam_origin = {P.cam.inv[3], P.cam.inv[7], P.cam.inv[11]};   // inverse * (0,0,0,1)

    // per-pixel radiance sums (tracer.cl:1179) live in shared memory: touched once per finished path, they would
    // otherwise hold six registers for the whole life of the thread
    __shared__ double col_sum[3][kBlockThreads];
    col_sum[0][threadIdx.x] = 0.0; col_sum[1][threadIdx.x] = 0.0; col_sum[2][threadIdx.x] = 0.0;
    Path<R> s;
    s.n = (unsigned)(P.sample_begin + slice);
    const unsigned n_end = (unsigned)P.sample_end;      // == samples unless a caller renders a sample range
    s.b = 0; s.effective = 0; s.inside = false;
    s.ro = {R(0), R(0), R(0)}; s.rd = {R(0), R(0), R(0)};
    s.mask = {R(1), R(1), R(1)}; s.accum = {R(0), R(0), R(0)};
    bool fresh = true;                 // need a new camera ray
    bool live = px.has_pixel;          // false once this lane has finished all its samples (lanes without a pixel
                                       // stay in the loop: the mesh walk is warp-cooperative and uses all 32 lanes)

    // Camera rays are generated one path AHEAD and parked in registers.  Generation runs only when
    // some lane needs a ray it does not have; at that moment every lane without a parked ray makes
    // its next one too, so the (RNG-heavy) generation code executes with most lanes active instead
    // of with the ~quarter of the lanes whose path happened to end on this iteration.  The RNG is a
    // pure function of (seed, sample, bounce), so evaluation order does not change any value.
    __shared__ R next_ray[6][kBlockThreads];          // the parked ray: written once and read once per path, so not in registers
    bool have_next = false;
    int defer_age = 0;
    unsigned stack_base = GROUPS ? (unsigned)__cvta_generic_to_shared(mesh_stacks) + (threadIdx.x / kWide) * (unsigned)P.stack_entries * 8u : 0u;
    asm volatile("" : "+r"(stack_base));      // opaque: keep it in a register instead of re-deriving the shared window base at every pop

    while (true) {
        if (fresh && s.n >= n_end) live = false;
        if (!__any_sync(kFullMask, live)) break;                 // the warp leaves together
        const bool starved = fresh && live && !have_next;
        if (__any_sync(kFullMask, starved)) {
            const unsigned gn = fresh ? s.n : s.n + (unsigned)P.slices;    // the sample this lane will start next
            if (live && !have_next && gn < n_end) {
                V3<R> nxo, nxd;
                camera_ray<R, RNG>(P, fx, fy, fgi, fgi2, gn, cam_origin, nxo, nxd);
                next_ray[0][threadIdx.x] = nxo.x; next_ray[1][threadIdx.x] = nxo.y; next_ray[2][threadIdx.x] = nxo.z;
                next_ray[3][threadIdx.x] = nxd.x; next_ray[4][threadIdx.x] = nxd.y; next_ray[5][threadIdx.x] = nxd.z;
                have_next = true;
            }
        }
        if (fresh && live) {
            s.ro = {next_ray[0][threadIdx.x], next_ray[1][threadIdx.x], next_ray[2][threadIdx.x]};
            s.rd = {next_ray[3][threadIdx.x], next_ray[4][threadIdx.x], next_ray[5][threadIdx.x]};
            have_next = false;
            s.b = 0; s.effective = 0; s.inside = false;
            s.mask = {R(1), R(1), R(1)}; s.accum = {R(0), R(0), R(0)};
            fresh = false;
        }

        Hit<R> h;
        closest_analytic<R>(P, s.ro, s.rd, h);
        bool deferred = false;                   // this lane's mesh walk was put off: it repeats the segment next iteration
        if (GROUPS) deferred = closest_mesh<R>(P, s.ro, s.rd, live, lane, h, stack_base, defer_age);

        bool done = live && !deferred;           // a miss ends the path (re-tracing it cannot hit either)
        NeeInfo<R> ni;
        ni.on = false;
        if (live && !deferred && h.obj >= 0) done = shade_hit<R, RNG, NEE>(P, h, s, fgi, ni);
        if constexpr (NEE) {
            // Next-event estimation, tracer.cl:786-825: one shadow ray per light towards a point of the light's bounding
            // sphere, for every lane whose bounce takes part; the shadow rays of a warp are traced together.
            ni.on = ni.on && live && !deferred && h.obj >= 0;
            for (int q = 0; q < P.n_lights; ++q) {
                const DLight<R>& L = P.light[q];
                const unsigned l = (unsigned)L.obj;
                const R nd = R(s.n);
                const float r1 = noise3d<RNG>(fgi, (float)(nd + R(ni.b * l)), fgi2);
                const float r2 = noise3d<RNG>(fgi, fgi2, (float)(nd + R(ni.b * ni.b * l)));
                const R lat = m_acos(R(2) * R(r1) - R(1)) - P.pi * R(2);        // randomPointOnSphere, tracer.cl:321-336, as written
                const R lon = R(2) * P.pi * R(r2);
                R slat, clat, slon, clon;
                m_sincos(lat, &slat, &clat); m_sincos(lon, &slon, &clon);
                const V3<R> lp = {L.ox + clat * clon * L.scale, L.oy + (slat - P.pi * R(0.25)) * L.scale, L.oz + clat * slon * L.scale};
                const V3<R> to = lp - ni.point;
                const V3<R> dir = to * m_div(R(1), m_sqrt(dot(to, to)));
                const V3<R> so = ni.point + dir * P.eps;
                const R ldn = dot(dir, ni.normal);
                const bool test = ni.on && ldn > R(0);
                if (__any_sync(kFullMask, test)) {
                    Hit<R> hs;
                    closest_analytic<R>(P, so, dir, hs);
                    int never = 0x7fffffff;                          // shadow rays are not deferred
                    if (GROUPS) closest_mesh<R>(P, so, dir, test, lane, hs, stack_base, never);
                    if (test && hs.obj == (int)l && hs.t > P.eps) {
                        const R att = R(1) - m_div(hs.t, m_sqrt(hs.t * hs.t + L.t0 * L.t0));
                        const V3<R> emis = {L.er, L.eg, L.eb};
                        s.accum = s.accum + ((ni.color * emis) * ldn) * ni.mask * att;
                    }
                }
            }
        }
        if (done) {
            col_sum[0][threadIdx.x] += (double)s.accum.x; col_sum[1][threadIdx.x] += (double)s.accum.y;   // tracer.cl:1179
            col_sum[2][threadIdx.x] += (double)s.accum.z;
            s.n += (unsigned)P.slices;
            fresh = true;
        }
    }

    // Epilogue: the slices of a pixel meet here -- the warps of this block through shared memory, the blocks of the
    // cluster through distributed shared memory -- and are summed in slice order (deterministic: the same order for
    // any grid shape), so no per-slice partial sums ever go to HBM.  The first slice's thread of the cluster's first
    // block owns the pixel's store; with an out_row map that store goes straight into the (possibly remote) frame.
    if (P.tile_cost && lane == 0 && px.tile < P.n_tiles) atomicMax(&P.tile_cost[px.tile], (unsigned)((clock64() - clock_begin) >> 8));
    if (cluster_size > 1) cluster_sync(); else __syncthreads();
    const int spb = P.slices_per_block;
    const int w = threadIdx.x >> 5;
    if (cluster_rank == 0 && (w % spb) == 0 && px.has_pixel) {
        double r = 0.0, g = 0.0, b = 0.0;
        const size_t lpix = (size_t)px.ly * W + px.lx;
        if (P.acc) { const double4 a = P.acc[lpix]; r = a.x; g = a.y; b = a.z; }
        const unsigned sums = (unsigned)__cvta_generic_to_shared(&col_sum[0][0]);
        for (unsigned cr = 0; cr < cluster_size; ++cr) {
            const unsigned remote = cluster_map_shared(sums, cr);       // block cr's col_sum (this block's own for cr == rank)
            for (int q = 0; q < spb; ++q) {
                const unsigned t = (threadIdx.x + 32 * q) * 8u;
                r += cluster_ld_f64(remote + t); g += cluster_ld_f64(remote + kBlockThreads * 8u + t); b += cluster_ld_f64(remote + 2u * kBlockThreads * 8u + t);
            }
        }
        if (P.acc) P.acc[lpix] = make_double4(r, g, b, 0.0);
        const double wgt = 1.0 / (double)P.samples;                                  // tracer.cl:837, 1184-1187
        const size_t opix = (size_t)(P.out_row ? P.out_row[px.ly] : px.ly) * W + px.lx;
        if (P.out_f32) reinterpret_cast<float4*>(P.out)[opix] = make_float4((float)(r * wgt), (float)(g * wgt), (float)(b * wgt), 1.0f);
        else reinterpret_cast<double4*>(P.out)[opix] = make_double4(r * wgt, g * wgt, b * wgt, 1.0);
    }
    if (cluster_size > 1) cluster_sync();          // remote shared memory stays alive until the first block has read it
}

// RGBA double -> float (the reference's frontend keeps float64 but its .raw writer and canvas store float32-range data,
// internal/app/raw/writer.go:11-35): halves the readback of a caller that only wants single precision.
__global__ void f32_kernel(const double4* __restrict__ in, float4* __restrict__ out, int pixels) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pixels) return;
    const double4 c = in[i];
    out[i] = make_float4((float)c.x, (float)c.y, (float)c.z, (float)c.w);
}

// Frontend tone step on the device (reference internal/app/tracer/pathtracer.go:42-59: no gamma,
// clamp(round(c*255)), alpha 255) so a caller that only wants the picture reads back 4 B/pixel instead of 32.
__global__ void rgba8_kernel(const double4* __restrict__ in, uchar4* __restrict__ out, int pixels) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pixels) return;
    const double4 c = in[i];
    auto q = [](double v) -> unsigned char {
        double r = round(v * 255.0);            // math.Round: half away from zero
        if (!(r >= 0.0)) r = 0.0;               // negatives and NaN
        if (r > 255.0) r = 255.0;
        return (unsigned char)r;
    };
    out[i] = make_uchar4(q(c.x), q(c.y), q(c.z), 255);
}

// Calibration of the FP32 issue roofline: 8 independent FFMA chains per thread, nothing else.
__global__ void __launch_bounds__(256) fma_peak_kernel(float* __restrict__ out, int iters) {
    float a0 = threadIdx.x * 1e-6f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float b = 0.999999f, c = 1e-7f * (blockIdx.x + 1);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// The same for the FP64 pipe (the fp64 mode's roofline): 8 independent DFMA chains per thread.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* __restrict__ out, int iters) {
    double a0 = threadIdx.x * 1e-6, a1 = a0 + 1., a2 = a0 + 2., a3 = a0 + 3., a4 = a0 + 4., a5 = a0 + 5., a6 = a0 + 6., a7 = a0 + 7.;
    const double b = 0.999999, c = 1e-7 * (blockIdx.x + 1);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// Test hook: evaluate the RNG on the device (parity tests compare it bit for bit with the oracle).
__global__ void noise3d_kernel(const float* __restrict__ xyz, float* __restrict__ out, int n, int mode) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    out[i] = mode == RNG_PARITY ? noise3d<RNG_PARITY>(x, y, z) : noise3d<RNG_FAST>(x, y, z);
}

}  // namespace ptk
