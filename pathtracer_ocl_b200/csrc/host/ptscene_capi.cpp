// ptscene_capi.cpp -- C ABI of libptscene (include/ptscene.h) plus the frontend's image writers.
#include <cstdio>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/ptscene.h"
#include "scenes.hpp"

using namespace pt;

struct pts_scene {
    SceneBuffers buf;
    std::vector<uint8_t> tex[3];
    int32_t tex_w[3] = {0, 0, 0}, tex_h[3] = {0, 0, 0}, tex_layers[3] = {0, 0, 0};
    int32_t obj_stats[5] = {0, 0, 0, 0, 0};
};

static void set_err(char* err, int errlen, const std::string& msg) {
    if (err && errlen > 0) { std::snprintf(err, size_t(errlen), "%s", msg.c_str()); }
}

static void pack_textures(pts_scene& out, int cls, const std::vector<Image>& imgs) {
    if (imgs.empty()) return;
    out.tex_w[cls] = imgs[0].width; out.tex_h[cls] = imgs[0].height; out.tex_layers[cls] = int32_t(imgs.size());
    for (const Image& im : imgs) {
        if (im.width != imgs[0].width || im.height != imgs[0].height) throw std::runtime_error("textures of one class must share a size");
        out.tex[cls].insert(out.tex[cls].end(), im.rgba.begin(), im.rgba.end());
    }
}

extern "C" {

int pts_scene_count(void) { return int(scene_names().size()); }
const char* pts_scene_name(int index) {
    if (index < 0 || index >= int(scene_names().size())) return nullptr;
    return scene_names()[size_t(index)].c_str();
}

pts_scene* pts_scene_build(const char* name, int32_t width, int32_t height, double aperture, double focal_length,
                           const char* assets_dir, int32_t tex_scale, char* err, int errlen) {
    try {
        if (width <= 0 || height <= 0) throw std::runtime_error("width and height must be positive");
        SceneConfig cfg;
        cfg.width = width; cfg.height = height; cfg.aperture = aperture; cfg.focal_length = focal_length;
        if (assets_dir && *assets_dir) cfg.assets_dir = assets_dir;
        cfg.tex_scale = tex_scale < 1 ? 1 : tex_scale;
        Scene sc = build_scene(name ? name : "default", cfg);
        auto* out = new pts_scene;
        try {
            out->buf = build_scene_buffers(sc);
            pack_textures(*out, 0, sc.textures);
            pack_textures(*out, 1, sc.sphere_textures);
            pack_textures(*out, 2, sc.cube_textures);
        } catch (...) { delete out; throw; }
        return out;
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return nullptr;
    }
}

void pts_scene_free(pts_scene* s) { delete s; }

void pts_scene_counts(const pts_scene* s, int32_t* no, int32_t* nt, int32_t* ng) {
    if (no) *no = int32_t(s->buf.objects.size());
    if (nt) *nt = int32_t(s->buf.triangles.size());
    if (ng) *ng = int32_t(s->buf.groups.size());
}
const void* pts_scene_objects(const pts_scene* s) { return s->buf.objects.empty() ? nullptr : s->buf.objects.data(); }
const void* pts_scene_triangles(const pts_scene* s) { return s->buf.triangles.empty() ? nullptr : s->buf.triangles.data(); }
const void* pts_scene_groups(const pts_scene* s) { return s->buf.groups.empty() ? nullptr : s->buf.groups.data(); }
const void* pts_scene_camera(const pts_scene* s) { return &s->buf.camera; }
int32_t pts_scene_texture(const pts_scene* s, int32_t cls, const uint8_t** rgba, int32_t* w, int32_t* h) {
    if (cls < 0 || cls > 2 || s->tex_layers[cls] == 0) { if (rgba) *rgba = nullptr; if (w) *w = 0; if (h) *h = 0; return 0; }
    if (rgba) *rgba = s->tex[cls].data();
    if (w) *w = s->tex_w[cls];
    if (h) *h = s->tex_h[cls];
    return s->tex_layers[cls];
}

void pts_fill_seeds(uint64_t seed, double* out, int64_t n) {
    uint64_t state = seed;
    for (int64_t i = 0; i < n; ++i) {
        state += 0x9E3779B97F4A7C15ull;
        uint64_t z = state;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z = z ^ (z >> 31);
        out[i] = double(z >> 11) * (1.0 / 9007199254740992.0);
    }
}

// ---- writers ------------------------------------------------------------------------------
static uint8_t clamp8(double c) {   // pathtracer.go:50-59
    double r = std::round(c * 255.0);
    if (r > 255.0) r = 255.0; else if (r < 0.0) r = 0.0;
    if (r != r) r = 0.0;
    return uint8_t(r);
}
static uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
        init = true;
    }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
    return crc;
}
static void be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(uint8_t(x >> 24)); v.push_back(uint8_t(x >> 16)); v.push_back(uint8_t(x >> 8)); v.push_back(uint8_t(x)); }
static void chunk(std::vector<uint8_t>& png, const char* tag, const std::vector<uint8_t>& data) {
    be32(png, uint32_t(data.size()));
    std::vector<uint8_t> body(tag, tag + 4);
    body.insert(body.end(), data.begin(), data.end());
    png.insert(png.end(), body.begin(), body.end());
    be32(png, crc32_update(0xFFFFFFFFu, body.data(), body.size()) ^ 0xFFFFFFFFu);
}

int pts_write_png(const char* path, const double* rgba, int32_t width, int32_t height) {
    // 8-bit RGBA, alpha 255, zlib "stored" blocks (no compression dependency).
    std::vector<uint8_t> raw;
    raw.reserve(size_t(height) * (size_t(width) * 4 + 1));
    for (int y = 0; y < height; ++y) {
        raw.push_back(0);
        for (int x = 0; x < width; ++x) {
            const double* p = rgba + (size_t(y) * width + x) * 4;
            raw.push_back(clamp8(p[0])); raw.push_back(clamp8(p[1])); raw.push_back(clamp8(p[2])); raw.push_back(255);
        }
    }
    std::vector<uint8_t> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    size_t pos = 0;
    while (pos < raw.size() || raw.empty()) {
        size_t n = raw.size() - pos; if (n > 65535) n = 65535;
        bool last = pos + n >= raw.size();
        z.push_back(last ? 1 : 0);
        z.push_back(uint8_t(n & 0xFF)); z.push_back(uint8_t(n >> 8));
        z.push_back(uint8_t(~n & 0xFF)); z.push_back(uint8_t((~n >> 8) & 0xFF));
        for (size_t i = 0; i < n; ++i) { a = (a + raw[pos + i]) % 65521; b = (b + a) % 65521; }
        z.insert(z.end(), raw.begin() + long(pos), raw.begin() + long(pos + n));
        pos += n;
        if (last) break;
    }
    be32(z, (b << 16) | a);
    std::vector<uint8_t> png = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<uint8_t> ihdr;
    be32(ihdr, uint32_t(width)); be32(ihdr, uint32_t(height));
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(png, "IHDR", ihdr);
    chunk(png, "IDAT", z);
    chunk(png, "IEND", {});
    FILE* f = std::fopen(path, "wb");
    if (!f) return 1;
    size_t w = std::fwrite(png.data(), 1, png.size(), f);
    std::fclose(f);
    return w == png.size() ? 0 : 1;
}

int pts_write_raw(const char* path, const double* rgba, int32_t width, int32_t height) {
    std::vector<uint8_t> out;
    be32(out, 1); be32(out, 0); be32(out, uint32_t(width)); be32(out, uint32_t(height));
    for (size_t i = 0; i < size_t(width) * height; ++i)
        for (int c = 0; c < 3; ++c) {
            float f = float(rgba[i * 4 + c]);
            uint32_t bits; std::memcpy(&bits, &f, 4);
            be32(out, bits);
        }
    FILE* f = std::fopen(path, "wb");
    if (!f) return 1;
    size_t w = std::fwrite(out.data(), 1, out.size(), f);
    std::fclose(f);
    return w == out.size() ? 0 : 1;
}

// ---- test hooks -----------------------------------------------------------------------------
static Mat4 m16(const double* p) { Mat4 m; for (int i = 0; i < 16; ++i) m[i] = p[i]; return m; }
static Tuple4 t4(const double* p) { return {p[0], p[1], p[2], p[3]}; }
static void put16(const Mat4& m, double* o) { for (int i = 0; i < 16; ++i) o[i] = m[i]; }

void pts_mat_multiply(const double* a, const double* b, double* out) { put16(multiply(m16(a), m16(b)), out); }
void pts_mat_inverse(const double* m, double* out) { put16(inverse(m16(m)), out); }
void pts_mat_transform(const char* kind, double x, double y, double z, double* out) {
    std::string k = kind ? kind : "";
    if (k == "translate") put16(translate(x, y, z), out);
    else if (k == "scale") put16(scale(x, y, z), out);
    else if (k == "rotx") put16(rotate_x(x), out);
    else if (k == "roty") put16(rotate_y(x), out);
    else if (k == "rotz") put16(rotate_z(x), out);
    else put16(identity(), out);
}
void pts_view_transform(const double* from, const double* to, const double* up, double* out) { put16(view_transform(t4(from), t4(to), t4(up)), out); }
int32_t pts_ray_box(const double* o, const double* d, const double* mn, const double* mx) {
    BoundingBox b; b.min = t4(mn); b.max = t4(mx);
    return intersect_ray_with_box(t4(o), t4(d), b) ? 1 : 0;
}
void pts_spherical_map(const double* p, double* uv) { spherical_map(t4(p), uv[0], uv[1]); }
int32_t pts_cube_face(const double* p) { return cube_face_from_point(t4(p)); }
void pts_split_bounds(const double* mn, const double* mx, double* out) {
    BoundingBox b, l, r; b.min = t4(mn); b.max = t4(mx);
    split_bounds(b, l, r);
    for (int i = 0; i < 4; ++i) { out[i] = l.min[i]; out[4 + i] = l.max[i]; out[8 + i] = r.min[i]; out[12 + i] = r.max[i]; }
}

pts_scene* pts_scene_from_obj(const char* obj_text, const char* mtl_dir, int32_t vertex_normals, int32_t divide_threshold,
                              char* err, int errlen) {
    try {
        ObjModel model = parse_obj(obj_text ? obj_text : "", mtl_dir ? mtl_dir : "");
        ShapePtr group = model.to_group();
        int tris = 0;
        for (auto& kv : model.groups) tris += int(kv.second->children.size());
        if (vertex_normals) {
            std::vector<ShapePtr> all;
            for (auto& kv : model.groups) for (auto& c : kv.second->children) if (c->kind == Kind::Triangle) all.push_back(c);
            compute_vertex_normals(all);
        }
        group->recompute_bounds();
        if (divide_threshold > 0) { divide(group, divide_threshold); group->recompute_bounds(); }
        Scene sc;
        sc.camera = new_camera(4, 4, 1.0, point(0, 0, -5), point(0, 0, 0));
        sc.objects = {group};
        auto* out = new pts_scene;
        out->buf = build_scene_buffers(sc);
        out->obj_stats[0] = int32_t(model.vertices.size()); out->obj_stats[1] = int32_t(model.normals.size());
        out->obj_stats[2] = int32_t(model.groups.size()); out->obj_stats[3] = tris; out->obj_stats[4] = model.ignored_lines;
        return out;
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        return nullptr;
    }
}
void pts_obj_stats(const pts_scene* s, int32_t* out5) { for (int i = 0; i < 5; ++i) out5[i] = s->obj_stats[i]; }

}  // extern "C"
