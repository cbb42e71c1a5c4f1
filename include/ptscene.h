/*
 * ptscene.h -- C ABI of libptscene, the host-side frontend that stands in for the reference's Go
 * packages above the render boundary (cmd/pt scene registry, internal/app/scenes, shapes, obj,
 * camera, geom and internal/ocl/scene.go BuildSceneBufferCL).  It produces the byte buffers
 * (include/ptwire.h) that libptcuda's ptc_render consumes.  It contains no rendering code.
 */
#ifndef PTSCENE_H
#define PTSCENE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pts_scene pts_scene;

/* cmd/pt/main.go:27-43, 92-96 (--list-scenes) */
int         pts_scene_count(void);
const char *pts_scene_name(int index);

/* Build a scene by registry name (unknown name -> "default", main.go:85-87) with the CLI's
 * width/height/aperture/focal-length (cmd/pt/main.go:47-52) and flatten it to wire buffers.
 * Returns NULL and fills `err` on failure (e.g. a missing asset; the reference panics). */
pts_scene *pts_scene_build(const char *name, int32_t width, int32_t height, double aperture,
                           double focal_length, const char *assets_dir, int32_t tex_scale,
                           char *err, int errlen);
void pts_scene_free(pts_scene *s);

void        pts_scene_counts(const pts_scene *s, int32_t *n_objects, int32_t *n_triangles, int32_t *n_groups);
const void *pts_scene_objects(const pts_scene *s);    /* n_objects   * 1024 B */
const void *pts_scene_triangles(const pts_scene *s);  /* n_triangles *  512 B (NULL if none) */
const void *pts_scene_groups(const pts_scene *s);     /* n_groups    *  256 B (NULL if none) */
const void *pts_scene_camera(const pts_scene *s);     /* 256 B */
/* texture class 0 plane, 1 sphere, 2 cube; returns layer count (0 = class unused) */
int32_t pts_scene_texture(const pts_scene *s, int32_t cls, const uint8_t **rgba, int32_t *width, int32_t *height);

/* Deterministic per-pixel seeds: splitmix64(seed) -> (x >> 11) * 2^-53, SURVEY.md 8d. */
void pts_fill_seeds(uint64_t seed, double *out, int64_t n);

/* Output writers of the frontend: internal/app/tracer/pathtracer.go:32-59 (PNG, clamp(round(c*255)),
 * no gamma) and internal/app/raw/writer.go:11-35 (big-endian header + f32 RGB). */
int pts_write_png(const char *path, const double *rgba, int32_t width, int32_t height);
int pts_write_raw(const char *path, const double *rgba, int32_t width, int32_t height);

/* Hooks that expose frontend helpers to the test-suite so the reference's own unit-test vectors
 * can be replayed against this restatement. */
void    pts_mat_multiply(const double *a16, const double *b16, double *out16);
void    pts_mat_inverse(const double *m16, double *out16);
void    pts_mat_transform(const char *kind, double x, double y, double z, double *out16); /* translate|scale|rotx|roty|rotz(x) */
void    pts_view_transform(const double *from4, const double *to4, const double *up4, double *out16);
int32_t pts_ray_box(const double *origin4, const double *dir4, const double *bbmin4, const double *bbmax4);
void    pts_spherical_map(const double *p4, double *uv2);
int32_t pts_cube_face(const double *p4);
/* BVH: split a box (bvh.go:9-49); out = leftmin4,leftmax4,rightmin4,rightmax4 */
void    pts_split_bounds(const double *bbmin4, const double *bbmax4, double *out16);
/* Parse an OBJ text, optionally compute vertex normals and Divide(threshold), flatten; returns a
 * scene holding a single group object (for objparser_test.go / bvh_test.go style checks). */
pts_scene *pts_scene_from_obj(const char *obj_text, const char *mtl_dir, int32_t vertex_normals,
                              int32_t divide_threshold, char *err, int errlen);
/* statistics of the parsed model behind a pts_scene_from_obj scene: vertices, normals, groups, triangles,
 * ignored lines */
void    pts_obj_stats(const pts_scene *s, int32_t *out5);

#ifdef __cplusplus
}
#endif
#endif
