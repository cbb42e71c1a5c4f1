/*
 * canon_rng.h -- TEST INFRASTRUCTURE (oracle).  The canonical definition of the reference's
 * hash RNG, internal/ocl/tracer.cl:314-317:
 *
 *     noise3D(x,y,z) = fract(sin(x*112.9898f + y*179.233f + z*237.212f) * 43758.5453f)
 *
 * OpenCL leaves two things open that change the stream completely (a 1-ulp change of the sin
 * argument or result moves the output by up to ~1.0 / ~3e-3): whether the multiply-adds are fused,
 * and how accurate `sin` is.  Neither pocl nor any OpenCL device is available to pin them, so the
 * oracle fixes them ("parity unpinned" at this level):
 *
 *   - the three products and two sums are separately rounded float operations, left to right;
 *   - sin is evaluated by canon_sinf below: IEEE double arithmetic only (mul, fma, rint), a fixed
 *     two-term Cody-Waite reduction by pi and the degree-21 odd Taylor polynomial, rounded once to
 *     float.  The double result is within ~2e-16 of the true sine, so the float is the correctly
 *     rounded sinf for all but ~1e-9 of inputs -- i.e. what a high-quality OpenCL `sin` returns.
 *     Every step is a single correctly-rounded IEEE operation, so the CUDA kernel reproduces it
 *     bit for bit (pathtracer_ocl_b200/csrc/kernels/rng.cuh);
 *   - fract(v) = fmin(v - floor(v), 0x1.fffffep-1f) as in the OpenCL C spec.
 */
#ifndef ORACLE_CANON_RNG_H
#define ORACLE_CANON_RNG_H

#include <math.h>

static inline float canon_sinf(float x) {
    const double INV_PI = 0x1.45f306dc9c883p-2;
    const double PI_HI = 0x1.921fb54442d18p+1;
    const double PI_LO = 0x1.1a62633145c07p-53;
    double xd = (double)x;
    double q = rint(xd * INV_PI);
    double r = fma(-q, PI_HI, xd);
    r = fma(-q, PI_LO, r);
    double r2 = r * r;
    double p = 0x1.71b8ef6dcf572p-66;          /*  1/21! */
    p = fma(p, r2, -0x1.2f49b46814157p-57);    /* -1/19! */
    p = fma(p, r2, 0x1.952c77030ad4ap-49);     /*  1/17! */
    p = fma(p, r2, -0x1.ae7f3e733b81fp-41);    /* -1/15! */
    p = fma(p, r2, 0x1.6124613a86d09p-33);     /*  1/13! */
    p = fma(p, r2, -0x1.ae64567f544e4p-26);    /* -1/11! */
    p = fma(p, r2, 0x1.71de3a556c734p-19);     /*  1/9!  */
    p = fma(p, r2, -0x1.a01a01a01a01ap-13);    /* -1/7!  */
    p = fma(p, r2, 0x1.1111111111111p-7);      /*  1/5!  */
    p = fma(p, r2, -0x1.5555555555555p-3);     /* -1/3!  */
    double s = fma(r * r2, p, r);
    long long qi = (long long)q;
    if (qi & 1) s = -s;
    return (float)s;
}

/* Second reproducible stream ("fast"): the same exact double reduction by pi, then the degree-11 odd
 * Taylor polynomial in float with explicit fused multiply-adds.  About 1 ulp from the true sine,
 * i.e. still a conforming OpenCL `sin` (the spec allows 4 ulp), at a third of the cost on the GPU.
 * fmaf() is a single correctly rounded operation, so this too is bit-identical on CPU and GPU. */
static inline float canon_sinf_fast(float x) {
    const double INV_PI = 0x1.45f306dc9c883p-2;
    const double PI_HI = 0x1.921fb54442d18p+1;
    const double PI_LO = 0x1.1a62633145c07p-53;
    double xd = (double)x;
    double q = rint(xd * INV_PI);
    double rd = fma(-q, PI_LO, fma(-q, PI_HI, xd));
    float r = (float)rd;
    float r2 = r * r;
    float p = -2.5052108e-8f;
    p = fmaf(p, r2, 2.7557319e-6f);
    p = fmaf(p, r2, -1.9841270e-4f);
    p = fmaf(p, r2, 8.3333333e-3f);
    p = fmaf(p, r2, -1.6666667e-1f);
    float s = fmaf(r * r2, p, r);
    long long qi = (long long)q;
    return (qi & 1) ? -s : s;
}

static inline float canon_noise3d_mode(float x, float y, float z, int fast) {
    float a = x * 112.9898f;
    float b = y * 179.233f;
    float c = z * 237.212f;
    float arg = (a + b) + c;
    float v = (fast ? canon_sinf_fast(arg) : canon_sinf(arg)) * 43758.5453f;
    float f = v - floorf(v);
    return fminf(f, 0x1.fffffep-1f);
}

static inline float canon_noise3d(float x, float y, float z) { return canon_noise3d_mode(x, y, z, 0); }

#endif
