"""Kernel time of ONE shard of a config on one GPU (shard k of n), for tuning the slice / cluster plan of multi-GPU runs
without an n-GPU lease.   python tools/shard_time.py [ref|teapot|gopher|cube|tex ...]   env: SHARDS=8 PTC_SLICES=.."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathtracer_ocl_b200 import scene as S, trace as T
from tools.r2_time import CASES

n = int(os.environ.get("SHARDS", "8"))
out = []
for name in sys.argv[1:] or ["ref"]:
    scene, W, H, spp, ap, fl, prec = CASES[name]
    sc = S.build_scene(scene, W, H, ap, fl)
    seeds = S.make_seeds(0x5EED0002, W * H)
    worst = 0.0
    for k in (0, n // 2):
        with T.open_scene(sc, spp, seeds, precision=prec, shard_index=k, shard_count=n) as ctx:
            best = 1e30
            for _ in range(3):
                ctx.trace()
                best = min(best, ctx.stats()["kernel_ms"])
        worst = max(worst, best)
    out.append(f"{name} 1/{n} shard {worst:7.2f} ms -> {W * H * spp / n / worst / 1e3 * n:9.1f} Mpaths/s x{n}")
print(os.environ.get("PTC_SLICES", "auto"), "|", "  ".join(out), flush=True)
