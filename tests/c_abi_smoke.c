/* c_abi_smoke.c -- a plain C caller of libptcuda with exactly the argument shapes the cgo binding passes
 * (go/internal/cuda/cuda.go: every buffer a direct argument of ptc_render_flat2, Go slices as pointer + length,
 * the camera record by address, textures absent => NULL pointers and a zeroed dims array, err as a byte buffer).
 * The closest thing to compiling the Go binding in an image without a Go toolchain: it proves the header is valid C,
 * that the wire records have the sizes Go's structs have, and that the call works (or fails loudly without a GPU).
 *
 * Scene, built by hand from the wire layout (include/ptwire.h; ocltracer.go:25-96): a floor plane (y = 0), an emissive
 * sphere above it, and a diffuse similarity-transformed sphere resting on the floor; camera at (0, 1, -5) looking +z.
 *
 * Prints one line:  rc=<0|1> devices=<n> center=<mean R over the ball's upper half> corner=<mean R of a floor row> msg=<error text>
 * Exit status 0 when the outcome is the expected one for the machine: rendered and plausible with a CUDA device,
 * refused with "no usable CUDA device" without one.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ptcuda.h"
#include "ptwire.h"

_Static_assert(sizeof(ptw_object) == 1024 && sizeof(ptw_triangle) == 512 && sizeof(ptw_group) == 256 && sizeof(ptw_camera) == 256,
               "wire record sizes (ocltracer.go:25-96)");

static void identity(double *m) { memset(m, 0, 16 * sizeof(double)); m[0] = m[5] = m[10] = m[15] = 1.0; }

/* object with transform = translate(tx,ty,tz) * scale(s): inverse = scale(1/s) * translate(-t) */
static void place(ptw_object *o, long type, double tx, double ty, double tz, double s) {
    memset(o, 0, sizeof *o);
    identity(o->transform); identity(o->inverse); identity(o->inverse_transpose);
    o->transform[0] = o->transform[5] = o->transform[10] = s;
    o->transform[3] = tx; o->transform[7] = ty; o->transform[11] = tz;
    o->inverse[0] = o->inverse[5] = o->inverse[10] = 1.0 / s;
    o->inverse[3] = -tx / s; o->inverse[7] = -ty / s; o->inverse[11] = -tz / s;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) o->inverse_transpose[r * 4 + c] = o->inverse[c * 4 + r];
    o->type = type;
    o->refractive_index = 1.0;
    o->color[0] = o->color[1] = o->color[2] = 0.8; o->color[3] = 1.0;
    o->min_y = -1e9; o->max_y = 1e9;
    for (int k = 0; k < 64; ++k) o->children[k] = -1;
}

int main(void) {
    enum { W = 64, H = 48, SPP = 16 };
    ptw_object objs[3];
    place(&objs[0], 0, 0.0, 0.0, 0.0, 1.0);                     /* floor plane */
    place(&objs[1], 1, 0.0, 6.0, 0.0, 2.0);                     /* light */
    objs[1].emission[0] = objs[1].emission[1] = objs[1].emission[2] = 6.0;
    place(&objs[2], 1, 0.0, 1.0, 0.0, 1.0);                     /* diffuse ball */
    objs[2].color[0] = 0.9; objs[2].color[1] = 0.2; objs[2].color[2] = 0.2;

    ptw_camera cam;
    memset(&cam, 0, sizeof cam);
    cam.width = W; cam.height = H; cam.fov = 1.0471975512;
    const double half_view = tan(cam.fov / 2.0), aspect = (double)W / H;
    cam.half_width = half_view; cam.half_height = half_view / aspect;
    cam.pixel_size = cam.half_width * 2.0 / W;
    /* view transform of a camera at (0,1,-5) looking down +z with up +y is orientation(-x, y, -z) * translate(-from);
     * the kernel wants its inverse: translate(from) * orientation^-1 (the orientation is its own inverse) */
    identity(cam.inverse);
    cam.inverse[0] = -1.0; cam.inverse[10] = -1.0;
    cam.inverse[3] = 0.0; cam.inverse[7] = 1.0; cam.inverse[11] = -5.0;

    double *seeds = malloc(sizeof(double) * W * H), *out = malloc(sizeof(double) * W * H * 4);
    unsigned long long x = 0x5EED0001ULL;
    for (int i = 0; i < W * H; ++i) {                           /* splitmix64 -> [0,1) */
        unsigned long long z = (x += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
        seeds[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0);
    }
    int32_t tex_dims[9] = {0};
    int32_t devices[1] = {0};
    char err[512] = {0};
    const int n_dev = ptc_device_count();
    /* cgo: ptc_render_flat2(unsafe.Pointer(&objects[0]), len, nil, 0, nil, 0, unsafe.Pointer(&camera), nil, nil, nil,
     *                       &texDims[0], &seeds[0], samples, precision, rngMode, features, &devices[0], 1, &results[0], &errbuf[0], 512) */
    const int rc = ptc_render_flat2(&objs[0], 3, NULL, 0, NULL, 0, &cam, NULL, NULL, NULL, &tex_dims[0], &seeds[0], SPP, PTC_FP32,
                                    PTC_RNG_PARITY, 0, &devices[0], 1, &out[0], &err[0], (int)sizeof err);
    double center = -1.0, corner = -1.0;
    int ok;
    if (rc == 0) {
        double center_g = 0.0;                                  /* upper half of the red ball: R clearly above G */
        center = 0.0;
        for (int y = H / 2 - 8; y < H / 2 - 2; ++y)
            for (int xx = W / 2 - 4; xx < W / 2 + 4; ++xx) { center += out[(y * W + xx) * 4] / 48.0; center_g += out[(y * W + xx) * 4 + 1] / 48.0; }
        corner = 0.0;                                           /* floor in front of the ball, lit by the sphere light */
        for (int xx = 0; xx < W; ++xx) corner += out[((H - 2) * W + xx) * 4] / W;
        const double alpha = out[((H / 2) * W + W / 2) * 4 + 3];
        ok = n_dev > 0 && alpha == 1.0 && center > 0.0 && center > 2.0 * center_g && corner > 0.0 && isfinite(center) && isfinite(corner);
    } else {
        ok = n_dev == 0 && strstr(err, "no usable CUDA device") != NULL;
    }
    printf("rc=%d devices=%d center=%.6f corner=%.6f msg=%s\n", rc, n_dev, center, corner, err);
    free(seeds); free(out);
    return ok ? 0 : 2;
}
