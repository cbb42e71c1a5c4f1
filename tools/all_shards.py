"""Kernel time of every shard of a config on one GPU: first trace (includes the cost probe) and the best later one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathtracer_ocl_b200 import scene as S, trace as T
n = int(os.environ.get("SHARDS", "8"))
for name in sys.argv[1:] or ["teapot"]:
    sc = S.build_scene(name, 1280, 960); seeds = S.make_seeds(0x5EED0002, 1280 * 960)
    first, later = [], []
    for k in range(n):
        with T.open_scene(sc, 2048, seeds, shard_index=k, shard_count=n) as ctx:
            ts = []
            for _ in range(4):
                ctx.trace(); ts.append(ctx.stats()["kernel_ms"])
        first.append(round(ts[0], 1)); later.append(round(min(ts[1:]), 1))
    print(name, "first", first, "max", max(first), "| later", later, "max", max(later), "sum", round(sum(later), 1), flush=True)
