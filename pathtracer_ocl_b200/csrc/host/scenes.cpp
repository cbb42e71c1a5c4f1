// scenes.cpp -- the fifteen scenes of the reference's cmd/pt registry (cmd/pt/main.go:27-43), restated
// from internal/app/scenes/*.go, and the Shape-graph -> wire-record flattening of
// internal/ocl/scene.go.  Compile with -ffp-contract=off.
#include "scenes.hpp"

#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace pt {

namespace {

const double kPi = 3.14159265358979323846;  // Go math.Pi

std::string read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("open " + path + ": no such file or directory");
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

Camera std_camera(const SceneConfig& c, const Tuple4& from = point(0, 0.1, -1.5), const Tuple4& to = point(0, 0.05, 0)) {
    Camera cam = new_camera(c.width, c.height, kPi / 3, from, to);
    cam.focal_length = c.focal_length;
    cam.aperture = c.aperture;
    return cam;
}

ShapePtr plane(const Mat4& a, const Material& m) { auto p = new_plane(); p->set_transform(a); p->set_material(m); return p; }
ShapePtr plane(const Mat4& a, const Mat4& b, const Material& m) { auto p = new_plane(); p->set_transform(a); p->set_transform(b); p->set_material(m); return p; }
ShapePtr sphere(const Mat4& a, const Mat4& b, const Material& m) { auto s = new_sphere(); s->set_transform(a); s->set_transform(b); s->set_material(m); return s; }

// The Cornell-box walls shared by most scenes (e.g. scenes/reference.go:22-56).
struct Box { ShapePtr left, right, floor, ceil, back, front; };
Box cornell(double back_z) {
    Box b;
    b.left = plane(translate(-.6, 0, 0), rotate_z(kPi / 2), diffuse(0.75, 0.25, 0.25));
    b.right = plane(translate(.6, 0, 0), rotate_z(kPi / 2), diffuse(0.25, 0.25, 0.75));
    b.floor = plane(translate(0, -.4, 0), diffuse(0.9, 0.8, 0.7));
    b.ceil = plane(translate(0, .4, 0), diffuse(0.9, 0.8, 0.7));
    b.back = plane(translate(0, 0, back_z), rotate_x(kPi / 2), diffuse(0.9, 0.8, 0.7));
    b.front = plane(translate(0, 0, -2), rotate_x(kPi / 2), diffuse(0.9, 0.8, 0.7));
    return b;
}
void label(const Box& b) {
    b.left->label = "leftwall"; b.right->label = "rghtwall"; b.floor->label = "floor   ";
    b.ceil->label = "ceiling "; b.back->label = "backwall"; b.front->label = "frntwall";
}

Material half_mirror() { Material m = mirror(); m.reflectivity = 0.8; m.color = color(0.97, 0.97, 0.843); return m; }
Material emitter(double r, double g, double b) { Material m = light_bulb(); m.emission = color(r, g, b); return m; }

ShapePtr ceiling_light(double y, double r, double g, double b) {  // flat ellipsoid under the ceiling
    auto l = new_sphere();
    l->set_transform(translate(0, y, 0));
    l->set_transform(scale(0.283, 0.01, 0.283));
    l->set_material(emitter(r, g, b));
    return l;
}

ObjModel load_obj(const SceneConfig& c, const std::string& file) {
    return parse_obj(read_file(c.assets_dir + "/" + file), c.assets_dir);
}

// scenes/teapot.go:83-108 and transparent_teapot.go:108-137: teapot.obj has no normals, so vertex
// normals are computed before the BVH split.
ShapePtr teapot_group(const SceneConfig& c, const std::vector<Mat4>& xf, const Material& m) {
    ObjModel model = load_obj(c, "teapot.obj");
    ShapePtr group = model.to_group();
    std::vector<ShapePtr> tris = group->children.at(0)->children;
    compute_vertex_normals(tris);
    group->recompute_bounds();
    for (const Mat4& t : xf) group->set_transform(t);
    group->set_material(m);
    divide(group, 50);
    group->recompute_bounds();
    return group;
}
// scenes/gopher.go:70-87
ShapePtr gopher_group(const SceneConfig& c, const std::vector<Mat4>& xf, const Material& m) {
    ObjModel model = load_obj(c, "gopher.obj");
    ShapePtr group = model.to_group();
    group->recompute_bounds();
    for (const Mat4& t : xf) group->set_transform(t);
    group->set_material(m);
    divide(group, 60);
    group->recompute_bounds();
    return group;
}
Material silver(double reflectivity) { Material m = diffuse(0.75, 0.75, 0.75); m.reflectivity = reflectivity; return m; }

// ---- scenes -----------------------------------------------------------------------------------

Scene scene_default(const SceneConfig& c) {  // scenes/ocl.go
    Scene s; s.camera = std_camera(c);
    Box b = cornell(.4);
    auto left_sphere = sphere(translate(-0.25, -0.24, 0.1), scale(0.16, 0.16, 0.16), diffuse(0.9, 0.8, 0.7));
    auto right_sphere = sphere(translate(0.25, -0.24, 0.1), scale(0.16, 0.16, 0.16), half_mirror());
    auto cyl = new_cylinder(0, 0.4, true);
    cyl->set_transform(translate(0.45, -0.5, -0.2));
    cyl->set_transform(scale(0.075, 1, 0.075));
    cyl->set_material(diffuse(0.92, 0.4, 0.8));
    auto cube = new_cube();
    cube->set_transform(translate(-0.3, -0.375, -0.3));
    cube->set_transform(scale(0.1, 0.05, 0.04));
    cube->set_transform(rotate_y(kPi / 4));
    cube->set_transform(rotate_z(kPi / 2));
    cube->set_material(diffuse(0.25, 0.25, 0.75));
    auto light = new_sphere();
    light->set_transform(translate(0, 1.36, 0));
    light->set_material(emitter(9, 8, 6));
    // Three triangles placed DIRECTLY under a top-level group: the flattener only descends into
    // child groups (scene.go:65-72), so these never reach the device -- kept for fidelity.
    auto group = new_group();
    group->set_material(diffuse(0.7, 0.4, 0.9));
    group->set_transform(translate(0.15, 0, -0.25));
    group->add_child(new_triangle(point(-0.2, -.4, 0), point(0.0, -.4, 0), point(0, -0.1, 0)));
    group->add_child(new_triangle(point(0, -.4, 0), point(0.2, -.4, 0), point(0, -0.1, 0)));
    group->add_child(new_triangle(point(0.1, -.4, -0.4), point(0, -0.1, 0), point(0, -.4, 0)));
    group->recompute_bounds();
    s.objects = {b.floor, b.ceil, b.left, b.right, b.back, left_sphere, right_sphere, cyl, cube, group, light};
    return s;
}

Scene scene_reference(const SceneConfig& c) {  // scenes/reference.go
    Scene s; s.camera = std_camera(c);
    Box b = cornell(.4);
    auto left_sphere = sphere(translate(-0.35, -0.28, -0.15), scale(0.12, 0.12, 0.12), diffuse(0.9, 0.8, 0.7));
    auto right_sphere = sphere(translate(0, -0.24, -0.30), scale(0.16, 0.16, 0.16), diffuse(0.9, 0.8, 0.7));
    s.objects = {ceiling_light(.399, 9, 9, 9), b.floor, b.ceil, b.left, b.right, b.back, left_sphere, right_sphere};
    return s;
}

Scene scene_teapot(const SceneConfig& c) {  // scenes/teapot.go
    Scene s; s.camera = std_camera(c);
    Box b = cornell(.4);
    auto left_sphere = sphere(translate(-0.35, -0.28, -0.15), scale(0.12, 0.12, 0.12), diffuse(0.9, 0.8, 0.7));
    auto group = teapot_group(c, {translate(0, -0.4, 0), scale(0.07, 0.07, 0.07)}, silver(0.2));
    auto light = new_sphere();
    light->set_transform(translate(0, .4, 0));
    light->set_transform(scale(0.3, 0.03, 0.3));
    light->set_material(emitter(9, 8, 6));
    s.objects = {light, b.floor, b.ceil, b.left, b.right, b.back, group, left_sphere};
    return s;
}

Scene scene_glass(const SceneConfig& c) {  // scenes/transparent_glass.go -- needs assets/glass.obj, absent upstream too
    (void)load_obj(c, "glass.obj");
    throw std::runtime_error("glass scene: assets/glass.obj is not shipped with the reference");
}

void add_gopher_and_bulb(Scene& s, const SceneConfig& c) {
    s.objects.push_back(gopher_group(c, {translate(-.4, -0.15, 0.2), rotate_z(-kPi / 2), rotate_x(-kPi / 4), scale(0.2, 0.2, 0.2)}, silver(0.2)));
    auto light = new_sphere();
    light->set_transform(translate(0, 1.36, 0));
    light->set_material(emitter(9, 8, 6));
    s.objects.push_back(light);
}

Scene scene_gopher(const SceneConfig& c) {  // scenes/gopher.go
    Scene s; s.camera = std_camera(c);
    Box b = cornell(1.4);
    auto right_sphere = sphere(translate(0.28, -0.24, 0.15), scale(0.16, 0.16, 0.16), half_mirror());
    s.objects = {b.floor, b.ceil, b.left, b.right, b.back, b.front, right_sphere};
    add_gopher_and_bulb(s, c);
    return s;
}

Scene scene_gopher_window(const SceneConfig& c) {  // scenes/gopher-with-window.go
    Scene s; s.camera = std_camera(c);
    Box b = cornell(1.4);
    auto box3 = [](const Mat4& t, const std::vector<Mat4>& rot, const Mat4& sc, const Material& m) {
        auto q = new_cube(); q->set_transform(t); for (auto& r : rot) q->set_transform(r); q->set_transform(sc); q->set_material(m); return q;
    };
    Material window = diffuse(0.75, 0.75, 1); window.emission = color(24, 24, 24);
    Material frame = diffuse(0.95, 0.95, 1);
    auto cube = box3(translate(0.6, .1, 0), {rotate_y(kPi / 2)}, scale(0.1, 0.16, 0.002), window);
    auto rborder = box3(translate(0.6, .1, -0.1), {rotate_y(kPi / 2)}, scale(0.01, 0.16, 0.02), frame);
    auto lborder = box3(translate(0.6, .1, 0.1), {rotate_y(kPi / 2)}, scale(0.01, 0.16, 0.02), frame);
    auto bborder = box3(translate(0.6, -.06, 0.0), {rotate_x(kPi / 2), rotate_y(kPi / 2)}, scale(0.01, 0.11, 0.04), frame);
    auto tborder = box3(translate(0.6, .26, 0.0), {rotate_x(kPi / 2), rotate_y(kPi / 2)}, scale(0.01, 0.11, 0.03), frame);
    auto center = sphere(translate(0, -0.28, -0.3), scale(0.12, 0.12, 0.12), diffuse(0.9, 0.8, 0.7));
    auto right_sphere = sphere(translate(0.28, -0.24, 0.15), scale(0.16, 0.16, 0.16), half_mirror());
    s.objects = {b.floor, b.ceil, b.left, b.right, b.back, cube, lborder, rborder, bborder, tborder, b.front, center, right_sphere};
    add_gopher_and_bulb(s, c);
    return s;
}

Scene scene_christian(const SceneConfig& c) {  // scenes/christian.go
    Scene s; s.camera = std_camera(c);
    Box b = cornell(.4);
    Material shiny = diffuse(0.9, 0.9, 0.9); shiny.reflectivity = 0.99;
    auto left_sphere = sphere(translate(-0.35, -0.28, -0.15), scale(0.12, 0.12, 0.12), shiny);
    auto group = teapot_group(c, {translate(0, -0.4, 0), scale(0.07, 0.07, 0.07)}, silver(0.2));
    Material bulb = emitter(90, 80, 60);
    Material cover_m = diffuse(0.8, 0.8, 0.8); cover_m.reflectivity = 0.95;
    auto lamp = [&](double x) { return sphere(translate(x, .3, 0), scale(0.03, 0.03, 0.03), bulb); };
    auto cover = [&](double x) {
        auto q = new_cylinder(0, 1, false);
        q->set_transform(translate(x, .295, 0));
        q->set_transform(scale(0.06, 0.4, 0.06));
        q->set_material(cover_m);
        return q;
    };
    s.objects = {lamp(-0.3), lamp(-0.1), lamp(0.1), lamp(0.3), cover(-0.3), cover(-0.1), cover(0.1), cover(0.3),
                 b.floor, b.ceil, b.left, b.right, b.back, group, left_sphere};
    return s;
}

Scene scene_textures(const SceneConfig& c) {  // scenes/texturedplanets.go
    Scene s; s.camera = std_camera(c);
    auto textured = [](Material m, uint8_t id, double sx, double sy) { m.textured = true; m.texture_id = id; m.texture_scale_x = sx; m.texture_scale_y = sy; return m; };
    auto with_nm = [](Material m, uint8_t id) { m.textured_nm = true; m.texture_id_nm = id; m.texture_scale_x_nm = 1.0; m.texture_scale_y_nm = 1.0; return m; };
    auto left = new_plane();
    left->set_transform(translate(-.6, 0, 0)); left->set_transform(rotate_x(kPi)); left->set_transform(rotate_z(kPi / 2)); left->set_transform(rotate_y(kPi / 2));
    left->set_material(with_nm(textured(diffuse(0.75, 0.25, 0.25), 0, 1, 1), 3));
    auto right = new_plane();
    right->set_transform(translate(.6, 0, 0)); right->set_transform(rotate_z(kPi / 2)); right->set_transform(rotate_y(kPi / 2));
    right->set_material(with_nm(textured(diffuse(0.25, 0.25, 0.75), 0, 1, 1), 3));
    auto floor = plane(translate(0, -.4, 0), textured(diffuse(0.9, 0.8, 0.7), 1, 0.25, 0.25));
    auto ceil = plane(translate(0, .4, 0), textured(diffuse(0.9, 0.8, 0.7), 2, 1, 1));
    auto back = plane(translate(0, 0, .4), rotate_x(kPi / 2), with_nm(textured(diffuse(0.9, 0.8, 0.7), 0, 1, 1), 3));
    // Textured spheres keep texture scale 0/0 as upstream (only planes use the scale).
    Material lsm = diffuse(0.9, 0.8, 0.7); lsm.textured = true; lsm.texture_id = 1;
    auto left_sphere = sphere(translate(-0.3, -0.1, -0.25), scale(0.2, 0.2, 0.2), lsm);
    Material rsm = diffuse(0.9, 0.8, 0.7); rsm.textured = true; rsm.texture_id = 0;
    auto right_sphere = new_sphere();
    right_sphere->set_transform(translate(0.2, 0, -0.3)); right_sphere->set_transform(rotate_y(kPi)); right_sphere->set_transform(scale(0.25, 0.25, 0.25));
    right_sphere->set_material(rsm);
    Material lm = emitter(10, 10, 10);
    auto l1 = sphere(translate(0, .395, -.9), scale(0.283, 0.01, 0.283), lm);
    auto l2 = sphere(translate(0, 0, -1.7), scale(0.283, 0.283, 0.01), lm);
    s.objects = {l1, l2, floor, ceil, left, right, back, left_sphere, right_sphere};
    int d = c.tex_scale < 1 ? 1 : c.tex_scale;
    for (uint32_t i = 0; i < 4; ++i) s.textures.push_back(make_texture(2048 / d, 2048 / d, 100 + i));   // squares, cobblestone, boards, normal map
    for (uint32_t i = 0; i < 2; ++i) s.sphere_textures.push_back(make_texture(4096 / d, 2048 / d, 200 + i));  // planet, jupiter
    return s;
}

Scene scene_envmap(const SceneConfig& c) {  // scenes/envmap.go
    Scene s; s.camera = std_camera(c, point(0, 0.1, -1.5), point(0, 0.15, 0));
    auto right_sphere = sphere(translate(0, -0.14, -0.30), scale(0.16, 0.16, 0.16), mirror());
    auto sky = new_sphere();
    sky->set_transform(scale(5, 5, 5));
    Material m = default_material();
    m.textured = true; m.texture_id = 0; m.texture_scale_x = 1.0; m.texture_scale_y = 1.0; m.emission = color(1, 1, 1);
    sky->material = m;
    s.objects = {right_sphere, sky};
    int d = c.tex_scale < 1 ? 1 : c.tex_scale;
    s.sphere_textures.push_back(make_texture(4096 / d, 2048 / d, 300));
    return s;
}

Scene scene_cubemap(const SceneConfig& c) {  // scenes/cubemap.go
    Scene s; s.camera = std_camera(c, point(0, 0.3, -2.7), point(0, 0.45, 0));
    auto right_sphere = sphere(translate(.2, 1, 2), scale(0.26, 0.26, 0.26), mirror());
    auto light = sphere(translate(1.1, 1, -4), scale(0.7, 0.7, 0.7), emitter(19.5, 19.5, 19.5));
    auto sky = new_cube();
    sky->set_transform(translate(0, 0, 0));
    sky->set_transform(scale(5, 5, 5));
    Material m = default_material();
    m.textured = true; m.texture_id = 0; m.texture_scale_x = 1.0; m.texture_scale_y = 1.0; m.emission = color(1, 1, 1); m.is_env_map = true;
    sky->material = m;
    auto group = gopher_group(c, {translate(-.7, -0.15, 0.2), rotate_z(-kPi / 2), rotate_x(-kPi / 4), scale(0.4, 0.4, 0.4)}, silver(0.0));
    s.objects = {light, right_sphere, sky, group};
    int d = c.tex_scale < 1 ? 1 : c.tex_scale;
    s.cube_textures.push_back(make_texture(4096 / d, 3072 / d, 400));
    return s;
}

Scene scene_reflection(const SceneConfig& c) {  // scenes/reflections.go
    Scene s; s.camera = std_camera(c);
    Box b = cornell(.4);
    auto left_sphere = sphere(translate(-0.35, -0.28, -0.15), scale(0.12, 0.12, 0.12), mirror());
    auto right_sphere = sphere(translate(0, -0.24, -0.30), scale(0.16, 0.16, 0.16), diffuse(0.9, 0.8, 0.7));
    s.objects = {ceiling_light(.399, 9, 9, 9), b.floor, b.ceil, b.left, b.right, b.back, left_sphere, right_sphere};
    return s;
}

Material dense_glass() { Material m = diffuse(0.9, 0.8, 0.7); m.refractive_index = 1.57; return m; }

Scene scene_transparency(const SceneConfig& c) {  // scenes/transparency.go
    Scene s; s.camera = std_camera(c);
    Box b = cornell(.6); label(b);
    auto left_sphere = sphere(translate(-0.25, -0.28, 0.25), scale(0.12, 0.12, 0.12), glass()); left_sphere->label = "left_spr";
    auto middle = sphere(translate(0, -0.24, -0.30), scale(0.16, 0.16, 0.16), dense_glass()); middle->label = "mddl_spr";
    auto right_sphere = sphere(translate(0.25, -0.28, 0.25), scale(0.12, 0.12, 0.12), mirror()); right_sphere->label = "right_spr";
    Material lm = emitter(9, 9, 9); lm.color = color(1, 1, 1);
    auto light = sphere(translate(0, .399, 0), scale(0.283, 0.01, 0.283), lm); light->label = "light   ";
    s.objects = {light, b.floor, b.ceil, b.left, b.right, b.back, left_sphere, middle, right_sphere};
    return s;
}

Scene scene_quad_lights(const SceneConfig& c) {  // scenes/transparency_quadlights.go
    Scene s; s.camera = std_camera(c);
    Box b = cornell(.6);
    auto left_sphere = sphere(translate(-0.25, -0.18, 0.25), scale(0.14, 0.14, 0.14), glass());
    auto middle = sphere(translate(0, -0.24, -0.30), scale(0.16, 0.16, 0.16), dense_glass());
    auto right_sphere = sphere(translate(0.35, -0.23, 0.2), scale(0.17, 0.17, 0.17), mirror());
    s.objects = {b.floor, b.ceil, b.left, b.right, b.back, left_sphere, middle, right_sphere};
    Material lm = emitter(9, 9, 9); lm.color = color(1, 1, 1);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
            auto l = new_cube();
            l->set_transform(translate(-0.25 + double(i) * 0.5, .399, -0.25 + double(j) * 0.5));
            l->set_transform(scale(0.15, 0.01, 0.15));
            l->set_material(lm);
            s.objects.push_back(l);
        }
    return s;
}

Scene scene_f_light(const SceneConfig& c) {  // scenes/transparency_f_light.go
    Scene s; s.camera = std_camera(c);
    Box b = cornell(.6);
    auto left_sphere = sphere(translate(-0.25, -0.18, 0.25), scale(0.14, 0.14, 0.14), glass());
    auto middle = sphere(translate(0, -0.24, -0.30), scale(0.16, 0.16, 0.16), dense_glass());
    auto right_sphere = sphere(translate(0.35, -0.23, 0.2), scale(0.17, 0.17, 0.17), mirror());
    Material lm = emitter(9, 9, 9); lm.color = color(1, 1, 1);
    auto bar = [&](const Mat4& t, const Mat4& sc) { auto l = new_cube(); l->set_transform(t); l->set_transform(sc); l->set_material(lm); return l; };
    s.objects = {b.floor, b.ceil, b.left, b.right, b.back, left_sphere, middle, right_sphere,
                 bar(translate(-0.125, .3999, 0.05), scale(0.05, 0.01, 0.45)),
                 bar(translate(-0.02, .3999, -0.35), scale(0.075, 0.01, 0.05)),
                 bar(translate(-0.05, .3999, 0), scale(0.075, 0.01, 0.05))};
    return s;
}

Scene scene_transparent_teapot(const SceneConfig& c) {  // scenes/transparent_teapot.go
    Scene s; s.camera = std_camera(c);
    Box b = cornell(.6); label(b);
    auto left_sphere = sphere(translate(-0.25, -0.28, 0.25), scale(0.12, 0.12, 0.12), diffuse(0.9, 0.8, 0.7)); left_sphere->label = "left_spr";
    auto right_sphere = sphere(translate(0.25, -0.28, 0.25), scale(0.12, 0.12, 0.12), glass()); right_sphere->label = "right_spr";
    Material thin = glass(); thin.refractive_index = -1.0; thin.reflectivity = 0.2;   // "glass without thickness"
    auto teapot = teapot_group(c, {translate(0, -0.38, -0.2), rotate_y(kPi / 12), scale(0.1, 0.1, 0.1)}, thin);
    teapot->label = "teapot  ";
    auto light = ceiling_light(.399, 9, 9, 9); light->label = "light   ";
    s.objects = {light, b.floor, b.ceil, b.left, b.right, b.back, left_sphere, right_sphere, teapot};
    return s;
}

struct Entry { const char* name; Scene (*fn)(const SceneConfig&); };
const Entry kScenes[] = {
    {"reference", scene_reference}, {"teapot", scene_teapot}, {"glass", scene_glass}, {"gopher", scene_gopher},
    {"gopher-window", scene_gopher_window}, {"christian", scene_christian}, {"textures", scene_textures},
    {"envmap", scene_envmap}, {"cubemap", scene_cubemap}, {"reflection", scene_reflection},
    {"transparency", scene_transparency}, {"transparency_quad_lights", scene_quad_lights},
    {"transparency_f_light", scene_f_light}, {"transparent_teapot", scene_transparent_teapot},
    {"default", scene_default},
};

}  // namespace

const std::vector<std::string>& scene_names() {
    static const std::vector<std::string> names = [] {
        std::vector<std::string> v;
        for (const Entry& e : kScenes) v.push_back(e.name);
        return v;
    }();
    return names;
}

Scene build_scene(const std::string& name, const SceneConfig& cfg) {
    for (const Entry& e : kScenes)
        if (name == e.name) return e.fn(cfg);
    return scene_default(cfg);
}

// ---- procedural textures ----------------------------------------------------------------------
namespace {
inline uint32_t hash3(uint32_t x, uint32_t y, uint32_t s) {
    uint32_t h = x * 0x9E3779B1u ^ (y * 0x85EBCA77u) ^ (s * 0xC2B2AE3Du);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}
// periodic value noise, `cells` lattice cells across the image in each direction
inline double value_noise(int x, int y, int w, int h, int cells_x, int cells_y, uint32_t salt) {
    double fx = double(x) * cells_x / w, fy = double(y) * cells_y / h;
    int ix = int(fx), iy = int(fy);
    double tx = fx - ix, ty = fy - iy;
    tx = tx * tx * (3 - 2 * tx); ty = ty * ty * (3 - 2 * ty);
    auto v = [&](int cx, int cy) { return (hash3(uint32_t(cx % cells_x), uint32_t(cy % cells_y), salt) & 0xFFFF) / 65535.0; };
    double a = v(ix, iy), b = v(ix + 1, iy), c = v(ix, iy + 1), d = v(ix + 1, iy + 1);
    return (a + (b - a) * tx) + ((c + (d - c) * tx) - (a + (b - a) * tx)) * ty;
}
}  // namespace

Image make_texture(int width, int height, uint32_t salt) {
    Image img;
    img.width = width < 4 ? 4 : width;
    img.height = height < 4 ? 4 : height;
    img.rgba.resize(size_t(img.width) * img.height * 4);
    const int w = img.width, h = img.height;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            uint8_t* px = &img.rgba[(size_t(y) * w + x) * 4];
            for (int ch = 0; ch < 3; ++ch) {
                uint32_t s = salt * 16 + ch;
                double v = 0.55 * value_noise(x, y, w, h, 8, 4, s) + 0.30 * value_noise(x, y, w, h, 32, 16, s + 4) +
                           0.15 * value_noise(x, y, w, h, 128, 64, s + 8);
                int q = int(v * 255.0 + 0.5);
                px[ch] = uint8_t(q < 0 ? 0 : (q > 255 ? 255 : q));
            }
            px[3] = 255;
        }
    return img;
}

// ---- flattening (internal/ocl/scene.go) -------------------------------------------------------
namespace {
void copy4(double* dst, const Tuple4& t) { for (int i = 0; i < 4; ++i) dst[i] = t[i]; }
void copy16(double* dst, const Mat4& m) { for (int i = 0; i < 16; ++i) dst[i] = m[i]; }

// scene.go:96-155 BuildCLGroup: node, then its own triangles contiguously, then <=2 child groups,
// depth-first.  Returns the node's index in `out.groups`.
int32_t flatten_group(const Shape& g, SceneBuffers& out) {
    const int32_t id = static_cast<int32_t>(out.groups.size());
    out.groups.emplace_back();
    std::memset(&out.groups[id], 0, sizeof(ptw_group));
    copy4(out.groups[id].bb_min, g.bbox.min);
    copy4(out.groups[id].bb_max, g.bbox.max);
    for (size_t i = 0; i < g.label.size() && i < sizeof(out.groups[id].padding); ++i) out.groups[id].padding[i] = uint8_t(g.label[i]);
    out.groups[id].tri_offset = static_cast<int32_t>(out.triangles.size());
    int32_t added = 0;
    for (const ShapePtr& c : g.children) {
        if (c->kind != Kind::Triangle) continue;
        ptw_triangle t;
        std::memset(&t, 0, sizeof t);
        copy4(t.p1, c->p1); copy4(t.p2, c->p2); copy4(t.p3, c->p3);
        copy4(t.e1, c->e1); copy4(t.e2, c->e2);
        copy4(t.n1, c->n1); copy4(t.n2, c->n2); copy4(t.n3, c->n3);
        copy4(t.color, c->material.color);
        out.triangles.push_back(t);
        ++added;
    }
    out.groups[id].tri_count = added;
    int32_t n_sub = 0;
    for (const ShapePtr& c : g.children) {
        if (c->kind != Kind::Group) continue;
        if (n_sub >= 2) throw std::runtime_error("BVH node with more than two child groups (scene.go:143 would panic)");
        int32_t child = flatten_group(*c, out);   // may reallocate out.groups
        out.groups[id].children[n_sub++] = child;
    }
    out.groups[id].child_group_count = n_sub > 0 ? n_sub : -1;
    return id;
}
}  // namespace

SceneBuffers build_scene_buffers(const Scene& scene) {
    SceneBuffers out;
    for (const ShapePtr& sp : scene.objects) {
        const Shape& s = *sp;
        ptw_object o;
        std::memset(&o, 0, sizeof o);
        std::memcpy(o.label, s.label.data(), s.label.size() < 8 ? s.label.size() : 8);
        copy16(o.transform, s.transform);
        copy16(o.inverse, s.inverse);
        copy16(o.inverse_transpose, s.inverse_transpose);
        copy4(o.color, s.material.color);
        copy4(o.emission, s.material.emission);
        o.refractive_index = s.material.refractive_index;
        for (int i = 0; i < 64; ++i) o.children[i] = -1;
        if (s.material.textured) {
            o.is_textured = 1; o.texture_index = s.material.texture_id;
            o.texture_scale_x = s.material.texture_scale_x; o.texture_scale_y = s.material.texture_scale_y;
        }
        if (s.material.textured_nm) {
            o.is_textured_nm = 1; o.texture_index_nm = s.material.texture_id_nm;
            o.texture_scale_x_nm = s.material.texture_scale_x_nm; o.texture_scale_y_nm = s.material.texture_scale_y_nm;
        }
        o.is_env_map = s.material.is_env_map ? 1 : 0;
        switch (s.kind) {
            case Kind::Plane: o.type = 0; break;
            case Kind::Sphere: o.type = 1; break;
            case Kind::Cylinder: o.type = 2; o.min_y = s.min_y; o.max_y = s.max_y; break;
            case Kind::Cube: o.type = 3; break;
            case Kind::Group: {
                o.type = 4;
                copy4(o.bb_min, s.bbox.min);
                copy4(o.bb_max, s.bbox.max);
                int idx = 0;
                for (const ShapePtr& c : s.children) {
                    if (c->kind != Kind::Group) continue;   // direct triangle children are dropped (scene.go:65-72)
                    if (idx >= 64) throw std::runtime_error("more than 64 root child groups");
                    o.children[idx++] = flatten_group(*c, out);
                    o.child_count++;
                }
                break;
            }
            default: o.type = 999; break;
        }
        o.reflectivity = s.material.reflectivity;
        out.objects.push_back(o);
    }
    std::memset(&out.camera, 0, sizeof out.camera);
    const Camera& c = scene.camera;
    out.camera.width = c.width; out.camera.height = c.height; out.camera.fov = c.fov;
    out.camera.pixel_size = c.pixel_size; out.camera.half_width = c.half_width; out.camera.half_height = c.half_height;
    out.camera.aperture = c.aperture; out.camera.focal_length = c.focal_length;
    copy16(out.camera.inverse, c.inverse);
    return out;
}

}  // namespace pt
