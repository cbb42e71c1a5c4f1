// scenes.hpp -- scene registry and the scene -> wire-buffer flattening of the frontend.
#pragma once
#include <string>
#include <vector>

#include "../../../include/ptwire.h"
#include "shapes.hpp"

namespace pt {

// RGBA8 image, row 0 = top (Go image.NRGBA.Pix order, ocltracer.go:228-254).
struct Image {
    int width = 0, height = 0;
    std::vector<uint8_t> rgba;
};

// internal/app/scenes/scene.go:16-28
struct Scene {
    Camera camera;
    std::vector<ShapePtr> objects;
    std::vector<Image> textures, sphere_textures, cube_textures;
};

// cmd/configuration.go:5-16 -- the part of cmd.Cfg the scene factories read.
struct SceneConfig {
    int width = 640, height = 480;
    double aperture = 0.0, focal_length = 0.0;
    std::string assets_dir = "assets";
    // Texture images are absent from the reference repo; textured scenes get
    // deterministic procedural stand-ins.  tex_scale divides their nominal resolution (1 = full
    // size; tests use 8 or 16 to keep buffers small).
    int tex_scale = 1;
};

// cmd/pt/main.go:27-43, same names and order.
const std::vector<std::string>& scene_names();
// Unknown names fall back to "default" (main.go:85-87).
Scene build_scene(const std::string& name, const SceneConfig& cfg);

// Wire buffers: what BuildSceneBufferCL (internal/ocl/scene.go:14-155) returns plus the camera record
// assembled in internal/app/tracer/renderer.go:44-56.
struct SceneBuffers {
    std::vector<ptw_object> objects;
    std::vector<ptw_triangle> triangles;
    std::vector<ptw_group> groups;
    ptw_camera camera;
};
SceneBuffers build_scene_buffers(const Scene& scene);

// Procedural textures (deterministic; smooth enough that bilinear filtering matters).
Image make_texture(int width, int height, uint32_t salt);

}  // namespace pt
