// ref_driver.cpp -- TEST INFRASTRUCTURE.  Host driver around the reference's own kernel (tracer.cl compiled for the
// CPU through cl_shim.hpp; see build_ref.py).  It plays the role of internal/ocl/ocltracer.go:256-376 computeBatch():
// the frame is cut into batches of 4 scanlines, each batch is one "NDRange" of width*rows work-items, work-item i of a
// batch reads seedX[i] of the batch's seeds and writes output[i*4 .. i*4+3].  Seeds come from the caller (one per
// pixel, row-major) instead of math/rand, exactly as in include/ptcuda.h.
//
// Only tests/ and bench.py's cpu_baseline / --impl reference legs may load the resulting library.
#include <algorithm>
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <thread>
#include <vector>

#include "cl_shim.hpp"

namespace refcl {
// sin() as the kernel sees it: the float overload -- the only one the image is chaotic in (noise3D, tracer.cl:314-317) --
// is the canonical correctly rounded sine of oracle/canon_rng.h; the double overload is libm's.
static inline float sin(float x) { return canon_sinf(x); }
static inline double sin(double x) { return std::sin(x); }
#include "tracer_cl.inc"   // the reference kernel, translated in memory by build_ref.py
}  // namespace refcl

static_assert(sizeof(refcl::object) == 1024 && sizeof(refcl::group) == 256 && sizeof(refcl::triangle) == 512 && sizeof(refcl::camera) == 256,
              "the reference's packed records must keep their wire sizes (ocltracer.go:25-96)");
static_assert(offsetof(refcl::object, inverse) == 128 && offsetof(refcl::object, color) == 384 && offsetof(refcl::object, type) == 456 &&
              offsetof(refcl::object, bbMin) == 520 && offsetof(refcl::object, childCount) == 584 && offsetof(refcl::object, isTextured) == 844,
              "object field offsets");
static_assert(offsetof(refcl::group, triOffset) == 128 && offsetof(refcl::group, children) == 140, "group field offsets");
static_assert(offsetof(refcl::triangle, e1) == 96 && offsetof(refcl::triangle, color) == 256, "triangle field offsets");
static_assert(offsetof(refcl::camera, pixelSize) == 16 && offsetof(refcl::camera, inverse) == 56, "camera field offsets");

extern "C" {

// Renders rows [row0, row1) of the frame; `out` receives (row1-row0)*width*4 doubles.  Returns 0.
int ref_trace(const void* objects, int n_objects, const void* triangles, int n_triangles, const void* groups, int n_groups, const void* camera,
              const uint8_t* const* tex, const int32_t* tex_w, const int32_t* tex_h, const int32_t* tex_layers, const double* seeds, int samples,
              int row0, int row1, int nthreads, double* out) {
    const refcl::camera* cam = static_cast<const refcl::camera*>(camera);
    const int W = cam->width;
    // the reference pads empty triangle / group slices with one zero record (ocltracer.go:106-120)
    static const refcl::triangle zero_tri = {};
    static const refcl::group zero_group = {};
    const refcl::triangle* tris = n_triangles > 0 ? static_cast<const refcl::triangle*>(triangles) : &zero_tri;
    const refcl::group* grps = n_groups > 0 ? static_cast<const refcl::group*>(groups) : &zero_group;
    cl_shim_image img[3];
    for (int k = 0; k < 3; ++k) img[k] = cl_shim_image{tex ? tex[k] : nullptr, tex_w ? tex_w[k] : 0, tex_h ? tex_h[k] : 0, tex_layers ? tex_layers[k] : 0};

    // work-items are independent: hand out pixels of the requested rows to threads; each call sets its own
    // get_global_id() and batch offset exactly as a batch of 4 scanlines would
    std::atomic<long> next{0};
    const long total = long(row1 - row0) * W;
    auto worker = [&] {
        for (;;) {
            const long begin = next.fetch_add(256);
            if (begin >= total) break;
            const long end = std::min(total, begin + 256);
            for (long p = begin; p < end; ++p) {
                const int y = row0 + int(p / W), x = int(p % W);
                const int batch_y0 = (y / 4) * 4;                              // ocltracer.go:212-223
                cl_shim_global_id = (y - batch_y0) * W + x;
                // the kernel indexes seedX / output by work-item id: give it views that start at this batch
                refcl::trace(static_cast<const refcl::object*>(objects), unsigned(n_objects), const_cast<refcl::triangle*>(tris),
                             const_cast<refcl::group*>(grps), out + long(batch_y0 - row0) * W * 4, seeds + long(batch_y0) * W, unsigned(samples),
                             const_cast<refcl::camera*>(cam), unsigned(batch_y0), img[0], img[1], img[2]);
            }
        }
    };
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    return 0;
}

// The reference's findClosestIntersection (tracer.cl:537-742) for `n_rays` world-space rays (6 doubles each: origin,
// direction).  Per ray, out[8] = t, object index (-1: none), the recorded normal xyz and colour rgb of the winning record
// (meaningful for triangles; tracer.cl:669-671).
int ref_closest(const void* objects, int n_objects, const void* triangles, int n_triangles, const void* groups, int n_groups,
                const double* rays, int n_rays, double* out) {
    static const refcl::triangle zero_tri = {};
    static const refcl::group zero_group = {};
    refcl::triangle* tris = const_cast<refcl::triangle*>(n_triangles > 0 ? static_cast<const refcl::triangle*>(triangles) : &zero_tri);
    refcl::group* grps = const_cast<refcl::group*>(n_groups > 0 ? static_cast<const refcl::group*>(groups) : &zero_group);
    std::vector<refcl::object> local(static_cast<const refcl::object*>(objects), static_cast<const refcl::object*>(objects) + n_objects);
    for (int r = 0; r < n_rays; ++r) {
        const double* q = rays + 6 * r;
        refcl::context ctx = {};
        const refcl::intersection ix = refcl::findClosestIntersection(local.data(), unsigned(n_objects), grps, tris, double4(q[0], q[1], q[2], 1.0),
                                                                      double4(q[3], q[4], q[5], 0.0), &ctx);
        double* o = out + 8 * r;
        o[0] = ix.t; o[1] = double(ix.lowestIntersectionIndex);
        const int k = ix.normalIndex >= 0 ? ix.normalIndex : 0;
        o[2] = ctx.xsTriangle[k].x; o[3] = ctx.xsTriangle[k].y; o[4] = ctx.xsTriangle[k].z;
        o[5] = ctx.xsTriangleColor[k].x; o[6] = ctx.xsTriangleColor[k].y; o[7] = ctx.xsTriangleColor[k].z;
    }
    return 0;
}

int ref_abi(void) { return 2; }

}  // extern "C"
