/*
 * ptcuda.h -- C ABI of libptcuda, the B200 (sm_100a) replacement for the OpenCL render path of
 * eriklupander/pathtracer-ocl.
 *
 * What each entry point replaces in the reference (paths relative to the reference repo):
 *
 *   ptc_device_count / ptc_device_name
 *        cmd/pt/main.go:98-112 listDevices()  (cl.GetPlatforms()[0].GetDevices(All) + Name()).
 *   ptc_render
 *        internal/ocl/ocltracer.go:100-226 Trace() together with :228-254 prepareTextures() and
 *        :256-376 computeBatch(): device selection, texture packing, the per-4-scanline batch
 *        loop (buffer upload, SetArgs, NDRange, Finish, readback) and the device kernel
 *        internal/ocl/tracer.cl:831-1187 trace().  One call renders the whole frame.
 *   ptc_open / ptc_trace / ptc_read / ptc_close
 *        the same path split into its phases (scene upload, kernel, gather+readback) so a caller
 *        can keep a scene resident, time the phases separately, or render one shard of the
 *        frame per process (one process per GPU).  ptc_render == open + trace + read + close.
 *
 * Conventions (kept from the reference boundary):
 *   - plain pointers and sizes only; every pointer is borrowed for the duration of the call;
 *   - objects / triangles / groups / camera are the packed records of include/ptwire.h, i.e. the
 *     exact bytes the Go side passes to EnqueueWriteBuffer (ocltracer.go:312-338);
 *   - the result is W*H*4 doubles, RGBA, row-major, top row first, alpha = 1.0
 *     (tracer.cl:1184-1187);
 *   - the reference draws one random double per pixel on the host (ocltracer.go:260-263); here
 *     the caller passes them (`seeds`, row-major, one per pixel) so renders are reproducible;
 *   - errors: the reference aborts via logrus.Fatalf; these functions return non-zero and write
 *     a message into `err` instead, the Go wrapper turns that back into Fatalf;
 *   - there is no CPU fallback: without a usable CUDA device every compute entry point fails.
 *   - thread-compatible: distinct contexts may be used concurrently.  The only process-wide state is a
 *     mutex-protected table of per-device memory pools (driver handles) that keeps a bounded amount of device
 *     memory between calls; ptc_trim() returns it.
 *   - launch plan: a context may re-plan its launches after its first renders of a mesh scene (tile launch order and
 *     sample slices per pixel, from measured per-tile clocks).  Pixels are pure functions of (scene, seeds, sample), so
 *     only the order in which a pixel's samples are summed can change: fp64-mode results of a later render may differ
 *     from the first in the last bits.  PTC_SLICES=<n> in the environment pins the slices.
 *   - arithmetic: PTC_FP64 evaluates tracer.cl's formulas in double but not operation by operation (fused multiply-add,
 *     reciprocal multiplies, spheres intersected in world space), so it agrees with the reference to the 1e-6 gate, not
 *     bit for bit; decisions that sit exactly on a threshold can differ.  Only the CPU oracle is bit-exact.
 */
#ifndef PTCUDA_H
#define PTCUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTC_ABI_VERSION 1

/* precision */
#define PTC_FP32 0   /* fp32 arithmetic mode (north_star: "fp32 mode", tolerance 1e-3 at 1 spp)   */
#define PTC_FP64 1   /* fp64 arithmetic mode, tracer.cl's own precision   (tolerance 1e-6 at 1 spp) */

/* ptc_job.features: code paths the reference ships switched off (they change the image; default 0 = as upstream) */
#define PTC_FEATURE_NEE           1  /* next-event estimation, tracer.cl:786-825 (call commented out at :1168)    */
#define PTC_FEATURE_CYLINDER_CAPS 2  /* cylinder end caps, tracer.cl:282-310 (disabled at :437-444)              */

/* frame formats (ptc_frame_*) */
#define PTC_FRAME_F64 0  /* 4 doubles per pixel, what ptc_read returns */
#define PTC_FRAME_F32 1  /* 4 floats per pixel: half the NVLink / PCIe bytes */
#define PTC_FRAME_HANDLE_BYTES 64

/* rng_mode */
#define PTC_RNG_PARITY 0 /* canonical noise3D stream, bit-identical to the oracle (oracle/canon_rng.h)  */
#define PTC_RNG_FAST   1 /* same hash, hardware-approximated sin; statistically validated only       */

typedef struct ptc_job {
    int32_t abi_version;          /* must be PTC_ABI_VERSION */

    const void *objects;          /* n_objects  * 1024 B  (ptw_object)   */
    int32_t     n_objects;        /* 1..16 (tracer.cl:846)               */
    const void *triangles;        /* n_triangles * 512 B  (ptw_triangle), may be NULL when 0 */
    int32_t     n_triangles;
    const void *groups;           /* n_groups   * 256 B  (ptw_group), may be NULL when 0     */
    int32_t     n_groups;
    const void *camera;           /* 256 B (ptw_camera)                  */

    /* Texture classes in kernel-argument order (tracer.cl:833): 0 plane textures, 1 sphere maps,
     * 2 cube-cross maps.  RGBA8, row 0 = top, `layers` images of identical size packed back to
     * back (ocltracer.go:228-254).  tex[i] == NULL means "class unused". */
    const uint8_t *tex[3];
    int32_t        tex_w[3], tex_h[3], tex_layers[3];

    const double *seeds;          /* width*height doubles in [0,1), row-major */
    int32_t samples;              /* samples per pixel, >= 1 */
    int32_t precision;            /* PTC_FP32 | PTC_FP64 */
    int32_t rng_mode;             /* PTC_RNG_PARITY | PTC_RNG_FAST */

    /* Devices driven by THIS process.  NULL/0 => device 0 (the reference's --device-index
     * default).  With n_devices > 1 the frame is split across them as interleaved scanline
     * tiles; every device's kernel stores its pixels straight into one buffer on the first
     * device (NVLink peer stores), so the readback is a single D2H. */
    const int32_t *devices;
    int32_t        n_devices;

    /* Frame sharding across PROCESSES (one process per GPU).  Scanline tile k (rows
     * [k*rows_per_tile, (k+1)*rows_per_tile)) belongs to shard k % shard_count.  shard_count
     * <= 1 means "whole frame".  A sharded context renders only its own tiles; ptc_read then
     * returns just those rows, packed in increasing row order. */
    int32_t shard_index;
    int32_t shard_count;
    int32_t rows_per_tile;        /* 0 => 4, the reference's batch height (ocltracer.go:214) */

    int32_t features;             /* PTC_FEATURE_* bits; 0 = the reference's behaviour */
    int32_t reserved[7];          /* must be zero */
} ptc_job;

typedef struct ptc_stats {
    double  upload_ms;            /* host wall time of ptc_open (scene flatten + H2D)          */
    double  kernel_ms;            /* max over devices of CUDA-event time of the last ptc_trace */
    double  read_ms;              /* host wall time of the last ptc_read (gather + D2H)        */
    int64_t h2d_bytes;            /* bytes copied host->device by ptc_open (all devices)       */
    int64_t d2h_bytes;            /* bytes copied device->host by the last ptc_read            */
    int64_t p2p_bytes;            /* bytes moved device->device by the last ptc_read           */
    int64_t paths;                /* pixels_in_shard * samples                                 */
    int32_t kernel_launches;      /* trace-kernel launches in the last ptc_trace (all devices) */
    int32_t n_devices;
    int32_t rows;                 /* rows rendered by this context                             */
    int32_t reserved;
} ptc_stats;

typedef struct ptc_context ptc_context;
typedef struct ptc_frame ptc_frame;

/* --list-devices support.  ptc_device_count returns the number of CUDA devices (0 if none or if
 * the driver is unusable).  ptc_device_name returns 0 and a NUL-terminated name, non-zero on a
 * bad index. */
int ptc_device_count(void);
int ptc_device_name(int index, char *buf, int buflen);

/* One-shot drop-in for ocl.Trace().  out_rgba: rows_of_this_shard * width * 4 doubles. */
int ptc_render(const ptc_job *job, double *out_rgba, char *err, int errlen);

/* ptc_render with every buffer as a direct argument instead of through ptc_job: the form a cgo
 * caller needs (Go may pass Go pointers as call arguments but may not store them inside a struct
 * handed to C).  tex_dims = {w0,h0,layers0, w1,h1,layers1, w2,h2,layers2}; devices may be NULL. */
int ptc_render_flat(const void *objects, int32_t n_objects, const void *triangles, int32_t n_triangles,
                    const void *groups, int32_t n_groups, const void *camera,
                    const uint8_t *tex_plane, const uint8_t *tex_sphere, const uint8_t *tex_cube,
                    const int32_t *tex_dims, const double *seeds, int32_t samples, int32_t precision,
                    int32_t rng_mode, const int32_t *devices, int32_t n_devices, double *out_rgba,
                    char *err, int errlen);

/* ptc_render_flat with the ptc_job.features bits as one more argument. */
int ptc_render_flat2(const void *objects, int32_t n_objects, const void *triangles, int32_t n_triangles,
                     const void *groups, int32_t n_groups, const void *camera,
                     const uint8_t *tex_plane, const uint8_t *tex_sphere, const uint8_t *tex_cube,
                     const int32_t *tex_dims, const double *seeds, int32_t samples, int32_t precision,
                     int32_t rng_mode, int32_t features, const int32_t *devices, int32_t n_devices,
                     double *out_rgba, char *err, int errlen);

/* Phase API. */
int  ptc_open(const ptc_job *job, ptc_context **ctx, char *err, int errlen);
int  ptc_trace(ptc_context *ctx, char *err, int errlen);   /* launches + waits; result stays in HBM */
int  ptc_read(ptc_context *ctx, double *out_rgba, char *err, int errlen);
int  ptc_get_stats(const ptc_context *ctx, ptc_stats *stats);
void ptc_close(ptc_context *ctx);

/* Progressive rendering (replaces the reference's 4-scanline batching as the way to
 * bound one launch, ocltracer.go:212-223): render samples [sample_begin, sample_end) of every pixel and
 * ADD them to the context's accumulator.  After ranges that together cover [0, samples) exactly once
 * the result equals ptc_trace's; in between, ptc_read returns accumulated_sum / samples (a caller
 * previewing n of N samples scales by N/n).  ptc_reset zeroes the accumulator. */
int ptc_trace_range(ptc_context *ctx, int32_t sample_begin, int32_t sample_end, char *err, int errlen);
int ptc_reset(ptc_context *ctx, char *err, int errlen);

/* The frontend's tone step on the device (internal/app/tracer/pathtracer.go:42-59:
 * no gamma, clamp(round(c*255)), alpha 255): rows*width*4 bytes, 8x less readback than ptc_read. */
int ptc_read_rgba8(ptc_context *ctx, uint8_t *out_rgba8, char *err, int errlen);

/* The frame as 4 floats per pixel (SURVEY.md 8f-2 "RGBA8 + optional f32"; the reference's .raw writer and canvas
 * hold single-precision-range data, internal/app/raw/writer.go:11-35): rows*width*4 floats, half of ptc_read's bytes. */
int ptc_read_f32(ptc_context *ctx, float *out_rgba_f32, char *err, int errlen);

/* Replace the per-pixel seeds of an open context (width*height doubles, same layout as
 * ptc_job.seeds).  Mirrors the reference drawing fresh seeds for every batch. */
int ptc_set_seeds(ptc_context *ctx, const double *seeds, char *err, int errlen);

/* Device-resident result of local device `local_index` (0..n_devices-1): pointer to its packed
 * rows (rows * width * 4 doubles), for callers that gather on the device themselves (NCCL).  A context that
 * drives several peer-connected devices gathers inside its kernels: local device 0 then exposes ALL the
 * context's rows and the other devices report 0 doubles. */
int ptc_device_framebuffer(ptc_context *ctx, int local_index, void **dev_ptr, int64_t *n_doubles,
                           int32_t *cuda_device);

/* Frames: the final gather fused into the trace kernel.  A frame is a whole-image buffer
 * (width*height RGBA pixels, row-major, top row first) on ONE device.  A context with a frame attached stores each
 * finished pixel straight into it, addressed by frame row -- from the owning device locally, from any other device
 * as NVLink peer stores out of the kernel's epilogue -- so N sharded contexts (N GPUs of one process, or N processes
 * through ptc_frame_export / ptc_frame_import, CUDA IPC) leave the complete image on the frame's device with no copy
 * kernel, no staging buffer and no collective; the caller only has to order "all shards traced" before
 * ptc_frame_read (a barrier between processes).  Replaces the per-batch EnqueueReadBuffer of ocltracer.go:355-373.
 *   ptc_frame_create    allocate on `device` (zero-filled)
 *   ptc_frame_export    64-byte handle another process passes to ptc_frame_import (only the creator may export)
 *   ptc_frame_import    map another process's frame for stores from `device`
 *   ptc_set_frame       attach (or, with NULL, detach) -- ptc_read* then refuse, the pixels are in the frame
 *   ptc_frame_read      device -> host copy of the whole frame (width*height*4 doubles or floats)
 *   ptc_frame_device_pointer  the raw buffer, for callers that keep the image on the device */
int  ptc_frame_create(int device, int32_t width, int32_t height, int32_t format, ptc_frame **frame, char *err, int errlen);
int  ptc_frame_export(ptc_frame *frame, void *handle64, char *err, int errlen);
int  ptc_frame_import(int device, const void *handle64, int32_t width, int32_t height, int32_t format,
                      ptc_frame **frame, char *err, int errlen);
int  ptc_set_frame(ptc_context *ctx, ptc_frame *frame, char *err, int errlen);
int  ptc_frame_read(ptc_frame *frame, void *out, char *err, int errlen);
int  ptc_frame_device_pointer(ptc_frame *frame, void **dev_ptr, int64_t *bytes, int32_t *cuda_device);
void ptc_frame_destroy(ptc_frame *frame);

/* Rows owned by this context, in increasing order; returns the count, fills at most `cap`. */
int ptc_shard_rows(const ptc_context *ctx, int32_t *rows, int cap);

/* The same ownership rule as a pure host function (no device needed): rows of a `height`-row
 * frame that belong to shard `shard_index` of `shard_count` with `rows_per_tile`-row tiles
 * (0 => 4).  Returns the count, fills at most `cap`; -1 on bad arguments. */
int ptc_plan_rows(int32_t height, int32_t rows_per_tile, int32_t shard_index, int32_t shard_count,
                  int32_t *rows, int cap);

/* Single-GPU contexts draw device memory from a per-device pool that is kept between calls so a
 * render pays no cudaMalloc/cudaFree.  The pool retains at most 768 MB per device after a context closes;
 * ptc_trim returns that too. */
void ptc_trim(void);

const char *ptc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PTCUDA_H */
