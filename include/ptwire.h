/*
 * ptwire.h -- byte layouts of the scene records that cross the render boundary.
 *
 * These are the records the reference's Go frontend already builds and hands to its
 * OpenCL driver (reference: internal/ocl/ocltracer.go:25-96, mirrored on the device side by
 * internal/ocl/tracer.cl:6-93).  libptcuda consumes exactly these bytes, so the Go side
 * (internal/ocl/scene.go:14 BuildSceneBufferCL, internal/app/tracer/renderer.go:44-56) needs
 * no change to its data path.
 *
 * All records are packed, little-endian, row-major 4x4 matrices (translation in [3],[7],[11]).
 */
#ifndef PTWIRE_H
#define PTWIRE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#pragma pack(push, 1)

/* ocltracer.go:25-51 (CLObject) / tracer.cl:37-63 (object). 1024 bytes. */
typedef struct ptw_object {
    double  transform[16];        /*    0 */
    double  inverse[16];          /*  128 */
    double  inverse_transpose[16];/*  256 */
    double  color[4];             /*  384 */
    double  emission[4];          /*  416 */
    double  refractive_index;     /*  448 */
    int64_t type;                 /*  456  0 plane, 1 sphere, 2 cylinder, 3 cube, 4 group */
    double  min_y;                /*  464 */
    double  max_y;                /*  472 */
    double  reflectivity;         /*  480 */
    double  texture_scale_x;      /*  488 */
    double  texture_scale_y;      /*  496 */
    double  texture_scale_x_nm;   /*  504 */
    double  texture_scale_y_nm;   /*  512 */
    double  bb_min[4];            /*  520 */
    double  bb_max[4];            /*  552 */
    int32_t child_count;          /*  584 */
    int32_t children[64];         /*  588 */
    uint8_t is_textured;          /*  844 */
    uint8_t texture_index;        /*  845 */
    uint8_t is_textured_nm;       /*  846 */
    uint8_t texture_index_nm;     /*  847 */
    uint8_t is_env_map;           /*  848  (never read by the kernel, tracer.cl:60) */
    char    label[8];             /*  849 */
    uint8_t padding[167];         /*  857 */
} ptw_object;

/* ocltracer.go:53-64 (CLGroup) / tracer.cl:24-35 (group). 256 bytes. */
typedef struct ptw_group {
    double  bb_min[4];            /*   0 */
    double  bb_max[4];            /*  32 */
    double  color[4];             /*  64  unused by the kernel */
    double  emission[4];          /*  96  unused by the kernel */
    int32_t tri_offset;           /* 128 */
    int32_t tri_count;            /* 132 */
    int32_t child_group_count;    /* 136 */
    int32_t children[2];          /* 140 */
    uint8_t padding[108];         /* 148 */
} ptw_group;

/* ocltracer.go:66-78 (CLTriangle) / tracer.cl:82-93 (triangle). 512 bytes. */
typedef struct ptw_triangle {
    double  p1[4];                /*   0 */
    double  p2[4];                /*  32 */
    double  p3[4];                /*  64 */
    double  e1[4];                /*  96 */
    double  e2[4];                /* 128 */
    double  n1[4];                /* 160 */
    double  n2[4];                /* 192 */
    double  n3[4];                /* 224 */
    double  color[4];             /* 256 */
    uint8_t padding[224];         /* 288 */
} ptw_triangle;

/* ocltracer.go:85-96 (CLCamera) / tracer.cl:6-17 (camera). 256 bytes. */
typedef struct ptw_camera {
    int32_t width;                /*   0 */
    int32_t height;               /*   4 */
    double  fov;                  /*   8  unused by the kernel */
    double  pixel_size;           /*  16 */
    double  half_width;           /*  24 */
    double  half_height;          /*  32 */
    double  aperture;             /*  40 */
    double  focal_length;         /*  48 */
    double  inverse[16];          /*  56 */
    uint8_t padding[72];          /* 184 */
} ptw_camera;

#pragma pack(pop)

#ifdef __cplusplus
}
static_assert(sizeof(ptw_object) == 1024, "object record must be 1024 bytes");
static_assert(sizeof(ptw_group) == 256, "group record must be 256 bytes");
static_assert(sizeof(ptw_triangle) == 512, "triangle record must be 512 bytes");
static_assert(sizeof(ptw_camera) == 256, "camera record must be 256 bytes");
#endif

/* Kernel constants (tracer.cl:1-4). PI is a float literal widened to double on purpose. */
#define PTW_MAX_OBJECTS 16          /* tracer.cl:846  __local object objects[16] */
#define PTW_MAX_ROOT_CHILDREN 64    /* tracer.cl:55 */
#define PTW_BVH_STACK 64            /* tracer.cl:624 */

#endif /* PTWIRE_H */
