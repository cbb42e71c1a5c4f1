// pt -- command-line frontend, the C++ stand-in for the reference's cmd/pt (cmd/pt/main.go:45-112)
// in an image without a Go toolchain.  Same flags, same defaults, same outputs:
//
//   --width 640 --height 480 --samples 1 --aperture 0 --focal-length 0 --scene gopher
//   --device-index 0 --list-devices --list-scenes
//
// Render() (internal/app/tracer/pathtracer.go:19-30): build the scene, flatten it
// (BuildSceneBufferCL), trace, write experiment.raw and out-<samples>-<W>x<H>.png.  The trace goes
// through the C ABI of libptcuda; extra flags select what the reference cannot: --precision
// fp32|fp64, --rng parity|fast, --devices 0,1,.. (multi-GPU in one process), --seed N, --nee, --cylinder-caps.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/ptcuda.h"
#include "../../../include/ptscene.h"

namespace {

struct Config {   // cmd/configuration.go:5-16
    int width = 640, height = 480, samples = 1, device_index = 0;
    double aperture = 0.0, focal_length = 0.0;
    bool list_devices = false, list_scenes = false;
    std::string scene = "gopher";
    // extensions
    int precision = PTC_FP32, rng_mode = PTC_RNG_PARITY, features = 0;
    unsigned long long seed = 0;
    bool seed_given = false;
    std::vector<int32_t> devices;
    std::string assets = "assets";
    int tex_scale = 1;
};

const char* env_or(const char* key, const char* fallback) {   // viper.AutomaticEnv(), main.go:64
    const char* v = std::getenv(key);
    return (v && *v) ? v : fallback;
}

bool take(int argc, char** argv, int& i, const char* name, std::string& out) {
    std::string a = argv[i];
    std::string flag = std::string("--") + name;
    if (a == flag) {
        if (i + 1 >= argc) { std::fprintf(stderr, "flag needs an argument: %s\n", flag.c_str()); std::exit(2); }
        out = argv[++i];
        return true;
    }
    if (a.rfind(flag + "=", 0) == 0) { out = a.substr(flag.size() + 1); return true; }
    return false;
}

}  // namespace

int main(int argc, char** argv) {
    Config cfg;
    cfg.width = std::atoi(env_or("WIDTH", "640"));
    cfg.height = std::atoi(env_or("HEIGHT", "480"));
    cfg.samples = std::atoi(env_or("SAMPLES", "1"));
    cfg.aperture = std::atof(env_or("APERTURE", "0"));
    cfg.scene = env_or("SCENE", "gopher");
    for (int i = 1; i < argc; ++i) {
        std::string v;
        std::string a = argv[i];
        if (take(argc, argv, i, "width", v)) cfg.width = std::atoi(v.c_str());
        else if (take(argc, argv, i, "height", v)) cfg.height = std::atoi(v.c_str());
        else if (take(argc, argv, i, "samples", v)) cfg.samples = std::atoi(v.c_str());
        else if (take(argc, argv, i, "aperture", v)) cfg.aperture = std::atof(v.c_str());
        else if (take(argc, argv, i, "focal-length", v)) cfg.focal_length = std::atof(v.c_str());
        else if (take(argc, argv, i, "scene", v)) cfg.scene = v;
        else if (take(argc, argv, i, "device-index", v)) cfg.device_index = std::atoi(v.c_str());
        else if (a == "--list-devices") cfg.list_devices = true;
        else if (a == "--list-scenes") cfg.list_scenes = true;
        else if (a == "--nee") cfg.features |= PTC_FEATURE_NEE;                       // tracer.cl:1168, commented out upstream
        else if (a == "--cylinder-caps") cfg.features |= PTC_FEATURE_CYLINDER_CAPS;   // tracer.cl:437-444, disabled upstream
        else if (take(argc, argv, i, "precision", v)) cfg.precision = (v == "fp64") ? PTC_FP64 : PTC_FP32;
        else if (take(argc, argv, i, "rng", v)) cfg.rng_mode = (v == "fast") ? PTC_RNG_FAST : PTC_RNG_PARITY;
        else if (take(argc, argv, i, "seed", v)) { cfg.seed = std::strtoull(v.c_str(), nullptr, 0); cfg.seed_given = true; }
        else if (take(argc, argv, i, "assets", v)) cfg.assets = v;
        else if (take(argc, argv, i, "tex-scale", v)) cfg.tex_scale = std::atoi(v.c_str());
        else if (take(argc, argv, i, "devices", v)) {
            size_t p = 0;
            while (p < v.size()) { cfg.devices.push_back(std::atoi(v.c_str() + p)); size_t q = v.find(',', p); if (q == std::string::npos) break; p = q + 1; }
        } else { std::fprintf(stderr, "unknown flag: %s\n", a.c_str()); return 2; }
    }

    if (cfg.list_devices) {   // main.go:98-112
        int n = ptc_device_count();
        for (int i = 0; i < n; ++i) {
            char name[256];
            if (ptc_device_name(i, name, sizeof name) == 0) std::printf("Index: %d Type: GPU Name: %s\n", i, name);
        }
        return 0;
    }
    if (cfg.list_scenes) {    // main.go:92-96
        for (int i = 0; i < pts_scene_count(); ++i) std::printf("%s\n", pts_scene_name(i));
        return 0;
    }

    auto t0 = std::chrono::steady_clock::now();
    char err[512] = {0};
    pts_scene* sc = pts_scene_build(cfg.scene.c_str(), cfg.width, cfg.height, cfg.aperture, cfg.focal_length, cfg.assets.c_str(),
                                    cfg.tex_scale, err, sizeof err);
    if (!sc) { std::fprintf(stderr, "FATA scene: %s\n", err); return 1; }

    const size_t px = size_t(cfg.width) * cfg.height;
    std::vector<double> seeds(px), out(px * 4);
    if (!cfg.seed_given) cfg.seed = (unsigned long long)std::chrono::system_clock::now().time_since_epoch().count();   // main.go:19
    pts_fill_seeds(cfg.seed, seeds.data(), (int64_t)px);

    ptc_job job;
    std::memset(&job, 0, sizeof job);
    job.abi_version = PTC_ABI_VERSION;
    int32_t no = 0, nt = 0, ng = 0;
    pts_scene_counts(sc, &no, &nt, &ng);
    job.objects = pts_scene_objects(sc); job.n_objects = no;
    job.triangles = pts_scene_triangles(sc); job.n_triangles = nt;
    job.groups = pts_scene_groups(sc); job.n_groups = ng;
    job.camera = pts_scene_camera(sc);
    for (int c = 0; c < 3; ++c) job.tex_layers[c] = pts_scene_texture(sc, c, &job.tex[c], &job.tex_w[c], &job.tex_h[c]);
    job.seeds = seeds.data();
    job.samples = cfg.samples;
    job.precision = cfg.precision;
    job.rng_mode = cfg.rng_mode;
    job.features = cfg.features;
    int32_t one = cfg.device_index;
    if (cfg.devices.empty()) { job.devices = &one; job.n_devices = 1; }
    else { job.devices = cfg.devices.data(); job.n_devices = (int32_t)cfg.devices.size(); }
    std::fprintf(stderr, "INFO trace with %d objects %dx%d\n", no, cfg.width, cfg.height);   // ocltracer.go:102

    if (ptc_render(&job, out.data(), err, sizeof err) != 0) {
        std::fprintf(stderr, "FATA %s\n", err);   // the reference aborts through logrus.Fatalf
        pts_scene_free(sc);
        return 1;
    }
    pts_scene_free(sc);

    if (pts_write_raw("experiment.raw", out.data(), cfg.width, cfg.height) != 0) std::fprintf(stderr, "ERRO error writing .raw file to disk\n");
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::fprintf(stderr, "INFO Finished in %.3fs (%.1f Mpaths/s end to end)\n", secs, double(px) * cfg.samples / secs / 1e6);
    char name[128];
    std::snprintf(name, sizeof name, "out-%d-%dx%d.png", cfg.samples, cfg.width, cfg.height);
    std::fprintf(stderr, "INFO writing output to file %s\n", name);
    return pts_write_png(name, out.data(), cfg.width, cfg.height);
}
