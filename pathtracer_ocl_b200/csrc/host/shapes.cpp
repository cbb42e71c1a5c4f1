// shapes.cpp -- see shapes.hpp.  Compile with -ffp-contract=off (Go/amd64 never fuses).
#include "shapes.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace pt {

// ---------------------------------------------------------------- materials
Material default_material() { Material m; m.color = rgb0(1, 1, 1); m.refractive_index = 1.0; return m; }
Material diffuse(double r, double g, double b) { Material m; m.color = rgb0(r, g, b); m.refractive_index = 1.0; return m; }
Material glass() { Material m; m.color = rgb0(1, 1, 1); m.refractive_index = 1.52; m.reflectivity = 0.05; return m; }
Material mirror() { Material m; m.color = rgb0(1, 1, 1); m.refractive_index = 1.0; m.reflectivity = 1.0; return m; }
Material light_bulb() { Material m; m.color = rgb0(1, 1, 1); m.emission = rgb0(8, 8, 8); m.refractive_index = 1.0; return m; }

// ---------------------------------------------------------------- bounding boxes
bool BoundingBox::contains_point(const Tuple4& p) const {
    return min[0] <= p[0] && min[1] <= p[1] && min[2] <= p[2] && max[0] >= p[0] && max[1] >= p[1] && max[2] >= p[2];
}
bool BoundingBox::contains_box(const BoundingBox& b) const { return contains_point(b.min) && contains_point(b.max); }
void BoundingBox::add(const Tuple4& p) {
    for (int i = 0; i < 3; ++i) {
        if (min[i] > p[i]) min[i] = p[i];
        if (max[i] < p[i]) max[i] = p[i];
    }
}
// Merging with an EMPTY box (min=+inf,max=-inf) widens this box to (-inf,+inf): add(+inf) raises
// max, add(-inf) lowers min.  That is what the reference does (boundingbox.go:35-60) and what the
// gopher scene's empty "DefaultGroup" triggers, so it is kept.
void BoundingBox::merge(const BoundingBox& b) { add(b.min); add(b.max); }

BoundingBox transform_box(const BoundingBox& b, const Mat4& m) {
    const Tuple4 corners[8] = {
        b.min,
        point(b.min[0], b.min[1], b.max[2]),
        point(b.min[0], b.max[1], b.min[2]),
        point(b.min[0], b.max[1], b.max[2]),
        point(b.max[0], b.min[1], b.min[2]),
        point(b.max[0], b.min[1], b.max[2]),
        point(b.max[0], b.max[1], b.min[2]),
        b.max,
    };
    BoundingBox out;
    for (const Tuple4& c : corners) out.add(multiply(m, c));
    return out;
}

BoundingBox bounds_of(const Shape& s) {
    BoundingBox box;
    if (s.kind == Kind::Group) {
        for (const ShapePtr& c : s.children) box.merge(parent_space_bounds(*c));
        return box;
    }
    if (s.kind == Kind::Triangle) {
        box.add(s.p1); box.add(s.p2); box.add(s.p3);
        return box;
    }
    box.min = point(-1, -1, -1);
    box.max = point(1, 1, 1);
    return box;
}
BoundingBox parent_space_bounds(const Shape& s) { return transform_box(bounds_of(s), s.transform); }

// ---------------------------------------------------------------- shapes
void Shape::set_transform(const Mat4& m) {
    if (kind == Kind::Triangle) throw std::logic_error("triangles have no transform");
    transform = multiply(transform, m);
    inverse = pt::inverse(transform);
    inverse_transpose = transpose(inverse);
}
void Shape::add_child(const ShapePtr& c) {
    children.push_back(c);
    bbox.merge(bounds_of(*c));
}
void Shape::recompute_bounds() { bbox = bounds_of(*this); }

ShapePtr new_plane() {
    auto s = std::make_shared<Shape>(Kind::Plane);
    s->material.color = rgb0(0, .5, 1);
    s->material.refractive_index = 0;
    return s;
}
ShapePtr new_sphere() {
    auto s = std::make_shared<Shape>(Kind::Sphere);
    s->material.color = rgb0(1, .5, .5);
    s->material.refractive_index = 1;
    return s;
}
ShapePtr new_cylinder(double min_y, double max_y, bool closed) {
    auto s = std::make_shared<Shape>(Kind::Cylinder);
    s->material = default_material();
    s->min_y = min_y; s->max_y = max_y; s->closed = closed;
    return s;
}
ShapePtr new_cube() {
    auto s = std::make_shared<Shape>(Kind::Cube);
    s->material = default_material();
    return s;
}
ShapePtr new_group() {
    auto s = std::make_shared<Shape>(Kind::Group);
    s->material = Material{};
    s->material.color = rgb0(0, 0, 0);
    s->material.refractive_index = 0;   // zero-valued Go struct (group.go:29-51 sets no material)
    return s;
}
static ShapePtr triangle_base(const Tuple4& p1, const Tuple4& p2, const Tuple4& p3) {
    auto s = std::make_shared<Shape>(Kind::Triangle);
    s->p1 = p1; s->p2 = p2; s->p3 = p3;
    s->e1 = sub(p2, p1);
    s->e2 = sub(p3, p1);
    s->n = normalize(cross(s->e2, s->e1));
    s->material = default_material();
    return s;
}
ShapePtr new_triangle(const Tuple4& p1, const Tuple4& p2, const Tuple4& p3) {
    ShapePtr s = triangle_base(p1, p2, p3);
    s->n1 = s->n2 = s->n3 = s->n;
    s->label = "Triangle3P";
    return s;
}
ShapePtr new_triangle(const Tuple4& p1, const Tuple4& p2, const Tuple4& p3, const Tuple4& n1, const Tuple4& n2,
                      const Tuple4& n3) {
    ShapePtr s = triangle_base(p1, p2, p3);
    s->n1 = n1; s->n2 = n2; s->n3 = n3;
    s->label = "Triangle";
    return s;
}

// ---------------------------------------------------------------- BVH
void split_bounds(const BoundingBox& b, BoundingBox& left, BoundingBox& right) {
    double dx = b.max[0] - b.min[0], dy = b.max[1] - b.min[1], dz = b.max[2] - b.min[2];
    double greatest = dx;
    if (dy > greatest) greatest = dy;
    if (dz > greatest) greatest = dz;
    double x0 = b.min[0], y0 = b.min[1], z0 = b.min[2];
    double x1 = b.max[0], y1 = b.max[1], z1 = b.max[2];
    if (greatest == dx) { x0 = x0 + dx / 2.0; x1 = x0; }
    else if (greatest == dy) { y0 = y0 + dy / 2.0; y1 = y0; }
    else { z0 = z0 + dz / 2.0; z1 = z0; }
    left.min = b.min;  left.max = point(x1, y1, z1);
    right.min = point(x0, y0, z0); right.max = b.max;
}

static int g_subgroup_counter = 0;

static void make_subgroup(Shape& g, const std::vector<ShapePtr>& members) {
    ShapePtr sub = new_group();
    sub->material = g.material;
    sub->label = "Subgroup " + std::to_string(++g_subgroup_counter);
    for (const ShapePtr& m : members) sub->add_child(m);
    g.add_child(sub);
}

void divide(const ShapePtr& s, int threshold) {
    if (s->kind != Kind::Group) return;
    Shape& g = *s;
    if (threshold <= static_cast<int>(g.children.size())) {
        BoundingBox lb, rb;
        split_bounds(bounds_of(g), lb, rb);
        std::vector<ShapePtr> left, right, remain;
        for (const ShapePtr& c : g.children) {
            BoundingBox cb = parent_space_bounds(*c);
            if (lb.contains_box(cb)) left.push_back(c);
            else if (rb.contains_box(cb)) right.push_back(c);
            else remain.push_back(c);
        }
        g.children = remain;
        g.recompute_bounds();
        if (!left.empty()) make_subgroup(g, left);
        if (!right.empty()) make_subgroup(g, right);
    }
    // Children appended above are visited too (Go re-evaluates g.Children[i] each iteration but the
    // range length is fixed at loop entry, after the subgroups were added).
    for (size_t i = 0; i < g.children.size(); ++i) divide(g.children[i], threshold);
}

// ---------------------------------------------------------------- host mirrors of kernel helpers
static void check_axis_host(double origin, double direction, double lo, double hi, double eps, double& tmin, double& tmax) {
    double a = lo - origin, b = hi - origin;
    if (std::fabs(direction) >= eps) { tmin = a / direction; tmax = b / direction; }
    else { tmin = a * HUGE_VAL; tmax = b * HUGE_VAL; }
    if (tmin > tmax) std::swap(tmin, tmax);
}
bool intersect_ray_with_box(const Tuple4& o, const Tuple4& d, const BoundingBox& b) {
    const double eps = 0.01;  // geom.Epsilon on the host (geom/types.go:5); the kernel uses 1e-4
    double x0, x1, y0, y1, z0, z1;
    check_axis_host(o[0], d[0], b.min[0], b.max[0], eps, x0, x1);
    check_axis_host(o[1], d[1], b.min[1], b.max[1], eps, y0, y1);
    check_axis_host(o[2], d[2], b.min[2], b.max[2], eps, z0, z1);
    double tmin = std::max(std::max(x0, y0), z0), tmax = std::min(std::min(x1, y1), z1);
    return tmin < tmax;
}
void spherical_map(const Tuple4& p, double& u, double& v) {
    double theta = std::atan2(p[0], p[2]);
    double radius = magnitude(vector(p[0], p[1], p[2]));
    double phi = std::acos(p[1] / radius);
    double raw_u = theta / (2 * M_PI);
    u = 1 - (raw_u + 0.5);
    v = 1 - phi / M_PI;
}
int cube_face_from_point(const Tuple4& p) {
    double c = std::max(std::max(std::fabs(p[0]), std::fabs(p[1])), std::fabs(p[2]));
    if (c == p[0]) return 0;
    if (c == -p[0]) return 1;
    if (c == p[1]) return 2;
    if (c == -p[1]) return 3;
    if (c == p[2]) return 4;
    return 5;
}

// ---------------------------------------------------------------- camera
Mat4 view_transform(const Tuple4& from, const Tuple4& to, const Tuple4& up) {
    Mat4 vt = identity();
    Tuple4 forward = normalize(sub(to, from));
    Tuple4 upn = normalize(up);
    Tuple4 left = cross(forward, upn);
    Tuple4 true_up = cross(left, forward);
    vt[0] = left[0]; vt[1] = left[1]; vt[2] = left[2];
    vt[4] = true_up[0]; vt[5] = true_up[1]; vt[6] = true_up[2];
    vt[8] = -forward[0]; vt[9] = -forward[1]; vt[10] = -forward[2];
    return multiply(vt, translate(-from[0], -from[1], -from[2]));
}
Camera new_camera(int width, int height, double fov, const Tuple4& from, const Tuple4& look_at) {
    Camera c;
    double half_view = std::tan(fov / 2);
    double aspect = static_cast<double>(width) / static_cast<double>(height);
    if (aspect >= 1.0) { c.half_width = half_view; c.half_height = half_view / aspect; }
    else { c.half_width = half_view * aspect; c.half_height = half_view; }
    c.pixel_size = (c.half_width * 2) / static_cast<double>(width);
    c.width = width; c.height = height; c.fov = fov;
    c.transform = view_transform(from, look_at, vector(0, 1, 0));
    c.inverse = inverse(c.transform);
    return c;
}

// ---------------------------------------------------------------- OBJ / MTL
static std::vector<std::string> fields(const std::string& line) {
    std::vector<std::string> out;
    std::istringstream is(line);
    std::string tok;
    while (is >> tok) out.push_back(tok);
    return out;
}
static std::vector<std::string> split(const std::string& s, char sep) {
    std::vector<std::string> out;
    size_t start = 0;
    while (true) {
        size_t p = s.find(sep, start);
        if (p == std::string::npos) { out.push_back(s.substr(start)); break; }
        out.push_back(s.substr(start, p - start));
        start = p + 1;
    }
    return out;
}
static double to_f(const std::string& s) { return std::strtod(s.c_str(), nullptr); }
static int to_i(const std::string& s) { return s.empty() ? 0 : static_cast<int>(std::strtol(s.c_str(), nullptr, 10)); }

std::vector<Mtl> parse_mtl(const std::string& text) {
    std::vector<Mtl> out;
    std::istringstream is(text);
    std::string row;
    while (std::getline(is, row)) {
        auto parts = fields(row);
        if (parts.empty()) continue;
        const std::string& k = parts[0];
        if (k == "newmtl") { Mtl m; m.name = parts[1]; out.push_back(m); continue; }
        if (out.empty()) continue;
        Mtl& cur = out.back();
        if (k == "Ns") cur.shininess = to_f(parts[1]);
        else if (k == "Ka") cur.ambient = color(to_f(parts[1]), to_f(parts[2]), to_f(parts[3]));
        else if (k == "Kd") cur.diffuse = color(to_f(parts[1]), to_f(parts[2]), to_f(parts[3]));
        else if (k == "Ks") cur.specular = color(to_f(parts[1]), to_f(parts[2]), to_f(parts[3]));
        else if (k == "Ni") cur.refractive_index = to_f(parts[1]);
        else if (k == "d") cur.transparency = 1 - to_f(parts[1]);
    }
    return out;
}

// objparser.go:181-196: colour = Ka + Kd + Ks, refractive index from Ni, everything else zero.
static Material mtl_to_material(const Mtl& m) {
    Material out;
    out.color = color(m.ambient[0] + m.diffuse[0] + m.specular[0], m.ambient[1] + m.diffuse[1] + m.specular[1],
                      m.ambient[2] + m.diffuse[2] + m.specular[2]);
    out.emission = {0, 0, 0, 0};
    out.refractive_index = m.refractive_index;
    return out;
}

ShapePtr ObjModel::find(const std::string& name) const {
    for (auto& kv : groups)
        if (kv.first == name) return kv.second;
    return nullptr;
}
ShapePtr ObjModel::to_group() const {
    ShapePtr g = new_group();
    g->label = "ROOT";
    for (auto& kv : groups) g->add_child(kv.second);
    return g;
}

ObjModel parse_obj(const std::string& text, const std::string& mtl_dir) {
    ObjModel out;
    out.vertices.push_back(point(0, 0, 0));
    out.normals.push_back(vector(0, 0, 0));
    std::vector<Mtl> mats;
    std::string current = "DefaultGroup";
    Material current_material = default_material();
    {
        ShapePtr g = new_group();
        g->label = current;
        out.groups.emplace_back(current, g);
    }
    std::istringstream is(text);
    std::string row;
    while (std::getline(is, row)) {
        auto parts = fields(row);
        if (parts.empty()) { out.ignored_lines++; continue; }
        const std::string& k = parts[0];
        if (k == "mtllib") {
            std::string path = mtl_dir.empty() ? parts[1] : mtl_dir + "/" + parts[1];
            std::ifstream f(path);
            if (!f) throw std::runtime_error("open " + path + ": no such file or directory");
            std::stringstream ss; ss << f.rdbuf();
            mats = parse_mtl(ss.str());
        } else if (k == "usemtl") {
            const Mtl* m = nullptr;
            for (const Mtl& c : mats) if (c.name == parts[1]) m = &c;
            if (!m) throw std::runtime_error("usemtl: unknown material " + parts[1]);
            current_material = mtl_to_material(*m);
            out.find(current)->set_material(current_material);
        } else if (k == "v") {
            out.vertices.push_back(point(to_f(parts[1]), to_f(parts[2]), to_f(parts[3])));
        } else if (k == "vn") {
            out.normals.push_back(vector(to_f(parts[1]), to_f(parts[2]), to_f(parts[3])));
        } else if (k == "f") {
            ShapePtr grp = out.find(current);
            if (row.find('/') == std::string::npos) {
                // vertex-only faces: fan triangulation, flat normal, default (white) material
                for (size_t i = 2; i + 1 < parts.size(); ++i) {
                    grp->add_child(new_triangle(out.vertices.at(to_i(parts[1])), out.vertices.at(to_i(parts[i])),
                                                out.vertices.at(to_i(parts[i + 1]))));
                }
            } else {
                for (size_t i = 2; i + 1 < parts.size(); ++i) {
                    auto s1 = split(parts[1], '/'), s2 = split(parts[i], '/'), s3 = split(parts[i + 1], '/');
                    int n1 = 0, n2 = 0, n3 = 0;
                    if (s1.size() == 3) { n1 = to_i(s1[2]); n2 = to_i(s2.size() > 2 ? s2[2] : ""); n3 = to_i(s3.size() > 2 ? s3[2] : ""); }
                    ShapePtr tri = new_triangle(out.vertices.at(to_i(s1[0])), out.vertices.at(to_i(s2[0])),
                                                out.vertices.at(to_i(s3[0])), out.normals.at(n1), out.normals.at(n2),
                                                out.normals.at(n3));
                    tri->material = current_material;
                    grp->add_child(tri);
                }
            }
        } else if (k == "g" || k == "o") {
            current = parts.size() > 1 ? parts[1] : std::string();
            if (!out.find(current)) {
                ShapePtr g = new_group();
                g->label = current;
                out.groups.emplace_back(current, g);
            }
        } else {
            out.ignored_lines++;
        }
    }
    return out;
}

void compute_vertex_normals(std::vector<ShapePtr>& tris) {
    // Same O(n^2) neighbourhood rule as the reference (match = all four coordinates within 0.01),
    // evaluated through a uniform grid so only nearby triangles are compared.  The accumulation
    // order (increasing j) is preserved, so sums are bit-identical to the brute-force loop.
    const size_t n = tris.size();
    struct Key { long x, y, z; bool operator<(const Key& o) const { return x != o.x ? x < o.x : (y != o.y ? y < o.y : z < o.z); } };
    const double cell = 0.02;
    auto key_of = [&](const Tuple4& p) { return Key{(long)std::floor(p[0] / cell), (long)std::floor(p[1] / cell), (long)std::floor(p[2] / cell)}; };
    std::vector<std::pair<Key, int>> entries;  // (cell, triangle) for every vertex
    entries.reserve(n * 3);
    for (size_t j = 0; j < n; ++j) {
        entries.emplace_back(key_of(tris[j]->p1), (int)j);
        entries.emplace_back(key_of(tris[j]->p2), (int)j);
        entries.emplace_back(key_of(tris[j]->p3), (int)j);
    }
    std::sort(entries.begin(), entries.end(), [](auto& a, auto& b) { return a.first < b.first || (!(b.first < a.first) && a.second < b.second); });
    auto candidates = [&](const Tuple4& p, std::vector<int>& out) {
        out.clear();
        Key k = key_of(p);
        for (long dx = -1; dx <= 1; ++dx) for (long dy = -1; dy <= 1; ++dy) for (long dz = -1; dz <= 1; ++dz) {
            Key q{k.x + dx, k.y + dy, k.z + dz};
            auto lo = std::lower_bound(entries.begin(), entries.end(), std::make_pair(q, -1),
                                       [](auto& a, auto& b) { return a.first < b.first || (!(b.first < a.first) && a.second < b.second); });
            for (; lo != entries.end() && !(q < lo->first) && !(lo->first < q); ++lo) out.push_back(lo->second);
        }
        std::sort(out.begin(), out.end());
        out.erase(std::unique(out.begin(), out.end()), out.end());
    };
    std::vector<Tuple4> r1(n), r2(n), r3(n);
    std::vector<int> cand;
    for (size_t i = 0; i < n; ++i) {
        const Shape& t = *tris[i];
        auto accumulate = [&](const Tuple4& p) {
            Tuple4 acc = t.n;
            candidates(p, cand);
            for (int j : cand) {
                if ((size_t)j == i) continue;
                const Shape& o = *tris[j];
                if (tuple_equals(p, o.p1) || tuple_equals(p, o.p2) || tuple_equals(p, o.p3)) acc = add(acc, o.n);
            }
            return normalize(acc);
        };
        r1[i] = accumulate(t.p1);
        r2[i] = accumulate(t.p2);
        r3[i] = accumulate(t.p3);
    }
    for (size_t i = 0; i < n; ++i) { tris[i]->n1 = r1[i]; tris[i]->n2 = r2[i]; tris[i]->n3 = r3[i]; }
}

}  // namespace pt
