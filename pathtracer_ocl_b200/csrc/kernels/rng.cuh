// rng.cuh -- device implementation of the reference's hash RNG (reference internal/ocl/tracer.cl:314-317)
//
//     noise3D(x,y,z) = fract(sin(x*112.9898f + y*179.233f + z*237.212f) * 43758.5453f)
//
// PARITY mode evaluates the canonical stream: separately rounded float products/sums and a sine
// built only from correctly rounded IEEE double operations (mul, fma, round-to-nearest-even
// conversion), rounded once to float -- every step is exactly reproducible on any IEEE machine,
// which is what lets a 1-spp render be compared pixel by pixel with a CPU restatement.
// FAST mode keeps the exact double argument reduction (the argument reaches ~1e9, far outside what
// fp32 reduction can handle) but evaluates the polynomial in fp32: the stream is statistically the
// same hash but not bit-identical, so it is validated on converged images only.
#pragma once
#include <cuda_runtime.h>

namespace ptk {

enum { RNG_PARITY = 0, RNG_FAST = 1 };

// Constants of the canonical sine (oracle/canon_rng.h), in constant memory: a DFMA takes a 64-bit constant-bank
// operand directly, whereas a literal has to be built in a uniform register pair first (two UMOVs per coefficient,
// ~24 issue slots per pair of noise3D calls in round 1's SASS).
__constant__ double c_sin[16] = {
    0x1.45f306dc9c883p-2,      // 0: 1/pi
    0x1.921fb54442d18p+1,      // 1: pi, high part
    0x1.1a62633145c07p-53,     // 2: pi, low part
    0x1.71b8ef6dcf572p-66, -0x1.2f49b46814157p-57, 0x1.952c77030ad4ap-49, -0x1.ae7f3e733b81fp-41, 0x1.6124613a86d09p-33,
    -0x1.ae64567f544e4p-26, 0x1.71de3a556c734p-19, -0x1.a01a01a01a01ap-13, 0x1.1111111111111p-7, -0x1.5555555555555p-3,
    6755399441055744.0,        // 13: 1.5 * 2^52
    0.0, 0.0};

__device__ __forceinline__ float sin_parity(float x) {
    const double INV_PI = c_sin[0];
    const double PI_HI = c_sin[1];
    const double PI_LO = c_sin[2];
    double xd = (double)x;
    // q = rint(x/pi) by the add-and-subtract-1.5*2^52 trick (exact for |x/pi| < 2^51, ties to even like
    // rint); the parity of q is the low mantissa bit of the biased sum.  Avoids two 64-bit conversions.
    const double MAGIC = 6755399441055744.0;
    double biased = __dadd_rn(__dmul_rn(xd, INV_PI), MAGIC);
    const int qi = __double2loint(biased);
    double q = __dadd_rn(biased, -MAGIC);
    double r = __fma_rn(-q, PI_HI, xd);
    r = __fma_rn(-q, PI_LO, r);
    double r2 = __dmul_rn(r, r);
    double p = c_sin[3];
    p = __fma_rn(p, r2, c_sin[4]);
    p = __fma_rn(p, r2, c_sin[5]);
    p = __fma_rn(p, r2, c_sin[6]);
    p = __fma_rn(p, r2, c_sin[7]);
    p = __fma_rn(p, r2, c_sin[8]);
    p = __fma_rn(p, r2, c_sin[9]);
    p = __fma_rn(p, r2, c_sin[10]);
    p = __fma_rn(p, r2, c_sin[11]);
    p = __fma_rn(p, r2, c_sin[12]);
    double s = __fma_rn(__dmul_rn(r, r2), p, r);
    if (qi & 1) s = -s;
    return __double2float_rn(s);
}

__device__ __forceinline__ float sin_fast(float x) {
    const double INV_PI = c_sin[0];
    const double PI_HI = c_sin[1];
    const double PI_LO = c_sin[2];
    double xd = (double)x;
    const double MAGIC = 6755399441055744.0;
    double biased = __dadd_rn(__dmul_rn(xd, INV_PI), MAGIC);
    const int qi = __double2loint(biased);
    double q = __dadd_rn(biased, -MAGIC);
    double rd = __fma_rn(-q, PI_LO, __fma_rn(-q, PI_HI, xd));
    float r = __double2float_rn(rd);
    float r2 = r * r;
    float p = -2.5052108e-8f;
    p = fmaf(p, r2, 2.7557319e-6f);
    p = fmaf(p, r2, -1.9841270e-4f);
    p = fmaf(p, r2, 8.3333333e-3f);
    p = fmaf(p, r2, -1.6666667e-1f);
    float s = fmaf(r * r2, p, r);
    return (qi & 1) ? -s : s;
}

template <int MODE>
__device__ __forceinline__ float noise3d(float x, float y, float z) {
    float a = __fmul_rn(x, 112.9898f);
    float b = __fmul_rn(y, 179.233f);
    float c = __fmul_rn(z, 237.212f);
    float arg = __fadd_rn(__fadd_rn(a, b), c);
    float s = (MODE == RNG_PARITY) ? sin_parity(arg) : sin_fast(arg);
    float v = __fmul_rn(s, 43758.5453f);
    float f = __fsub_rn(v, floorf(v));
    return fminf(f, 0x1.fffffep-1f);
}

}  // namespace ptk
