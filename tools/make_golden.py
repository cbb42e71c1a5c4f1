"""Generates tests/golden/*.npz: small renders by the CPU oracle (fp64) on fixed seeds.

The oracle is held bit-identical to the reference's own kernel (tracer.cl compiled for the CPU, oracle/_ref), and
tests/test_oracle_vs_reference.py checks that the compiled reference kernel reproduces every one of these files bit for
bit -- so they are outputs of the reference itself, usable where /root/reference does not exist (the GPU box).  The GPU
parity tests compare the CUDA path with both the live oracle and these files.  Re-run only if the reference changes:

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathtracer_ocl_b200 import scene as S  # noqa: E402
from oracle import oracle as O  # noqa: E402

# name, scene, width, height, spp, aperture, focal length, seed
CASES = [
    ("default_64x48_1", "default", 64, 48, 1, 0.0, 0.0, 0x5EED0001),
    ("reference_dof_64x48_4", "reference", 64, 48, 4, 0.15, 1.6, 0x5EED0002),
    ("teapot_48x36_1", "teapot", 48, 36, 1, 0.0, 0.0, 0x5EED0003),
    ("gopher_48x36_1", "gopher", 48, 36, 1, 0.0, 0.0, 0x5EED0004),
    ("transparency_64x48_2", "transparency", 64, 48, 2, 0.0, 0.0, 0x5EED0005),
    ("transparent_teapot_48x36_2", "transparent_teapot", 48, 36, 2, 0.0, 0.0, 0x5EED0006),
    ("textures_48x36_2", "textures", 48, 36, 2, 0.0, 0.0, 0x5EED0007),
    ("envmap_48x36_1", "envmap", 48, 36, 1, 0.0, 0.0, 0x5EED0008),
    ("cubemap_48x36_1", "cubemap", 48, 36, 1, 0.0, 0.0, 0x5EED0009),
]
TEX_SCALE = 16


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, scene, w, h, spp, ap, fl, seed in CASES:
        sc = S.build_scene(scene, w, h, ap, fl, tex_scale=TEX_SCALE)
        seeds = S.make_seeds(seed, w * h)
        img, cnt = O.trace(sc, seeds, spp, precision=1)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), rgba=img, scene=scene, width=w, height=h, spp=spp,
                            aperture=ap, focal_length=fl, seed=seed, tex_scale=TEX_SCALE,
                            segments=cnt["segments"], paths=cnt["paths"])
        print(name, img[..., :3].mean(axis=(0, 1)).round(4), cnt["segments"])


if __name__ == "__main__":
    main()
