// shapes.hpp -- host-side scene model: materials, primitives, groups, bounding boxes, the BVH
// splitter and the camera.  This is the C++ stand-in for the reference's Go frontend packages
// internal/app/{material,shapes,camera}; it exists so scenes (and therefore the byte buffers fed to
// libptcuda) can be built in an image without a Go toolchain.  Behavioural citations are per item.
#pragma once
#include <cstdint>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "geom.hpp"

namespace pt {

// internal/app/material/material.go:7-21
struct Material {
    Tuple4 color{1, 1, 1, 0};
    Tuple4 emission{0, 0, 0, 0};
    double refractive_index = 1.0;
    double reflectivity = 0.0;
    bool textured = false;
    uint8_t texture_id = 0;
    double texture_scale_x = 0, texture_scale_y = 0;
    bool textured_nm = false;
    uint8_t texture_id_nm = 0;
    double texture_scale_x_nm = 0, texture_scale_y_nm = 0;
    bool is_env_map = false;
};
Material default_material();                    // material.go:23-29
Material diffuse(double r, double g, double b); // material.go:31-37
Material glass();                               // material.go:38-45
Material mirror();                              // material.go:46-53
Material light_bulb();                          // material.go:54-60

// internal/app/shapes/boundingbox.go:8-21
struct BoundingBox {
    Tuple4 min{std::numeric_limits<double>::infinity(), std::numeric_limits<double>::infinity(),
               std::numeric_limits<double>::infinity(), 1.0};
    Tuple4 max{-std::numeric_limits<double>::infinity(), -std::numeric_limits<double>::infinity(),
               -std::numeric_limits<double>::infinity(), 1.0};
    bool contains_point(const Tuple4& p) const;   // boundingbox.go:26-29 (inclusive)
    bool contains_box(const BoundingBox& b) const;// boundingbox.go:31-33
    void add(const Tuple4& p);                    // boundingbox.go:40-60
    void merge(const BoundingBox& b);             // boundingbox.go:35-38
};

enum class Kind : int { Plane = 0, Sphere = 1, Cylinder = 2, Cube = 3, Group = 4, Triangle = 5 };

struct Shape;
using ShapePtr = std::shared_ptr<Shape>;

// One tagged record instead of the reference's interface + five structs.
struct Shape {
    Kind kind;
    std::string label;
    Mat4 transform = identity(), inverse = identity(), inverse_transpose = identity();
    Material material;
    // cylinder (shapes/cylinder.go:10-41)
    double min_y = -std::numeric_limits<double>::infinity();
    double max_y = std::numeric_limits<double>::infinity();
    bool closed = false;
    // triangle (shapes/triangle.go:11-95)
    Tuple4 p1{}, p2{}, p3{}, e1{}, e2{}, n{}, n1{}, n2{}, n3{};
    // group (shapes/group.go:9-51)
    std::vector<ShapePtr> children;
    BoundingBox bbox;

    explicit Shape(Kind k) : kind(k) {}
    // T = T * m; inverse; inverse-transpose  (e.g. shapes/plane.go:57-61, group.go:70-74)
    void set_transform(const Mat4& m);
    void set_material(const Material& m) { material = m; }
    // group only
    void add_child(const ShapePtr& c);            // group.go:118-127
    void recompute_bounds();                      // group.go:129-131
};

ShapePtr new_plane();      // shapes/plane.go:11-29   (colour 0,.5,1; refractive index 0)
ShapePtr new_sphere();     // shapes/sphere.go:14-29  (colour 1,.5,.5)
ShapePtr new_cylinder(double min_y, double max_y, bool closed);  // cylinder.go:35-41
ShapePtr new_cube();       // shapes/cube.go:9-23
ShapePtr new_group();      // shapes/group.go:29-51
// flat-shaded triangle: n = normalize(cross(e2, e1)) used for all three vertex normals (triangle.go:21-67)
ShapePtr new_triangle(const Tuple4& p1, const Tuple4& p2, const Tuple4& p3);
// triangle with explicit vertex normals (triangle.go:69-93)
ShapePtr new_triangle(const Tuple4& p1, const Tuple4& p2, const Tuple4& p3, const Tuple4& n1,
                      const Tuple4& n2, const Tuple4& n3);

BoundingBox bounds_of(const Shape& s);                          // boundingbox.go:92-134
BoundingBox parent_space_bounds(const Shape& s);                // boundingbox.go:62-65
BoundingBox transform_box(const BoundingBox& b, const Mat4& m); // boundingbox.go:67-90
void split_bounds(const BoundingBox& b, BoundingBox& left, BoundingBox& right);  // bvh.go:9-49
void divide(const ShapePtr& s, int threshold);                  // bvh.go:51-119

// Ray/AABB slab test of the host package (shapes IntersectRayWithBox, pinned by
// boundingbox_test.go:203-262); same arithmetic as the kernel's tracer.cl:250-280.
bool intersect_ray_with_box(const Tuple4& origin, const Tuple4& direction, const BoundingBox& b);
// UV helpers mirrored on the host by the reference (shapes/sphericalmap.go, shapes/cubemap.go).
void spherical_map(const Tuple4& p, double& u, double& v);
int cube_face_from_point(const Tuple4& p);  // 0 left, 1 right, 2 front, 3 back, 4 up, 5 down

// internal/app/camera/camera.go:8-81
struct Camera {
    int width = 0, height = 0;
    double fov = 0;
    Mat4 transform = identity(), inverse = identity();
    double pixel_size = 0, half_width = 0, half_height = 0;
    double aperture = 0, focal_length = 0;
};
Mat4 view_transform(const Tuple4& from, const Tuple4& to, const Tuple4& up);
Camera new_camera(int width, int height, double fov, const Tuple4& from, const Tuple4& look_at);

// internal/app/obj/objparser.go
struct ObjModel {
    std::vector<Tuple4> vertices, normals;            // index 0 is a placeholder (objparser.go:24-25)
    std::vector<std::pair<std::string, ShapePtr>> groups;  // first-appearance order (see to_group)
    int ignored_lines = 0;
    ShapePtr find(const std::string& name) const;
    // objparser.go:208-215.  The reference iterates a Go map (random order); this frontend uses
    // first-appearance order, which only affects tie-breaks between coincident hits.
    ShapePtr to_group() const;
};
struct Mtl {
    std::string name;
    Tuple4 ambient{}, diffuse{}, specular{};
    double shininess = 0, transparency = 0, refractive_index = 0;
};
// `mtl_dir`: where `mtllib` files are looked up (the reference opens them relative to the CWD,
// objparser.go:36; we resolve next to the .obj).
ObjModel parse_obj(const std::string& text, const std::string& mtl_dir);
std::vector<Mtl> parse_mtl(const std::string& text);
void compute_vertex_normals(std::vector<ShapePtr>& tris);      // objparser.go:137-178

}  // namespace pt
