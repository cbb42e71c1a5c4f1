"""Slice-count sweep on the full-size reference frame and on a 1/8-frame shard (what one of 8 GPUs renders)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathtracer_ocl_b200 import scene as S, trace as T
W, H, spp = 1280, 960, 2048
sc = S.build_scene("reference", W, H, 0.15, 1.6)
seeds = S.make_seeds(0x5EED0002, W * H)
for shards in (1, 8):
    for sl in [int(x) for x in os.environ.get("SLICES", "2,4,8,16,32").split(",")]:
        os.environ["PTC_SLICES"] = str(sl)
        with T.open_scene(sc, spp, seeds, shard_index=0, shard_count=shards) as ctx:
            best = 1e30
            for _ in range(3):
                ctx.trace(); best = min(best, ctx.stats()["kernel_ms"])
        print(f"shards {shards} slices {sl:3d}: {best:8.2f} ms  ({W * H * spp / shards / best / 1e3:8.1f} Mpaths/s)", flush=True)
