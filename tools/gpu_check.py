"""Development check run on the GPU box: parity of the CUDA path against the oracle on several
scenes, plus rough timings.  Not part of the test-suite (tests/ holds the real parity tests)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathtracer_ocl_b200 import scene as S, trace as T  # noqa: E402
from oracle import oracle as O  # noqa: E402


def parity(name, W, H, spp, ap=0.0, fl=0.0, tex_scale=8, rng=0):
    sc = S.build_scene(name, W, H, ap, fl, tex_scale=tex_scale)
    seeds = S.make_seeds(0x5EED0000 + spp, W * H)
    ref, cnt = O.trace(sc, seeds, spp, precision=1, rng_mode=rng)
    out = {}
    for prec, tol in ((T.FP64, 1e-6), (T.FP32, 1e-3)):
        img = T.render_scene(sc, spp, seeds, precision=prec, rng_mode=rng)
        err = np.abs(img[..., :3] - ref[..., :3]).max(axis=-1)
        out["fp64" if prec else "fp32"] = dict(frac=float((err <= tol).mean()), max=float(err.max()),
                                               nan=int(np.isnan(img).sum()))
    print(f"{name:24s} {W}x{H}@{spp} " + json.dumps(out), flush=True)
    return out


def timing(name, W, H, spp, ap=0.0, fl=0.0, prec=T.FP32, rng=T.RNG_PARITY, reps=2):
    sc = S.build_scene(name, W, H, ap, fl, tex_scale=1)
    seeds = S.make_seeds(0x5EED0002, W * H)
    with T.open_scene(sc, spp, seeds, precision=prec, rng_mode=rng) as ctx:
        for _ in range(reps):
            ctx.trace()
            st = ctx.stats()
        t0 = time.time()
        ctx.read()
        rd = time.time() - t0
    print(f"time {name:12s} {W}x{H}@{spp} prec={'fp64' if prec else 'fp32'} rng={rng}: kernel {st['kernel_ms']:.1f} ms "
          f"-> {st['paths'] / st['kernel_ms'] / 1e3:.1f} Mpaths/s; upload {st['upload_ms']:.1f} ms read {rd * 1e3:.1f} ms", flush=True)


if __name__ == "__main__":
    print(T.list_devices())
    x = np.random.default_rng(1).random((100000, 3)).astype(np.float32) * np.array([1, 2048, 2048], dtype=np.float32)
    ref = np.empty(len(x), np.float32)
    O.lib().oracle_noise3d_array(np.ascontiguousarray(x).ctypes.data, len(x), ref.ctypes.data)
    dev = T.debug_noise3d(x)
    print("noise3d bit-exact:", bool((ref.view(np.uint32) == dev.view(np.uint32)).all()), "mismatches", int((ref != dev).sum()))
    parity("default", 320, 240, 1)
    parity("default", 320, 240, 1, rng=1)
    parity("teapot", 160, 120, 1, rng=1)
    parity("reference", 320, 240, 1, 0.15, 1.6)
    parity("reference", 160, 120, 16, 0.15, 1.6)
    parity("transparency", 320, 240, 2)
    parity("teapot", 160, 120, 1)
    parity("gopher", 160, 120, 1)
    parity("transparent_teapot", 160, 120, 2)
    parity("textures", 160, 120, 2)
    parity("envmap", 160, 120, 2)
    parity("cubemap", 160, 120, 2)
    timing("reference", 1280, 960, 64, 0.15, 1.6)
    timing("reference", 1280, 960, 64, 0.15, 1.6, rng=T.RNG_FAST)
    timing("reference", 1280, 960, 16, 0.15, 1.6, prec=T.FP64)
    timing("teapot", 1280, 960, 4)
    timing("gopher", 1280, 960, 4)
