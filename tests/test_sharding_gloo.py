"""N>1 host logic on CPU: two processes over gloo shard a frame into interleaved scanline tiles,
each produces its own rows, and rank 0 gathers them into the full frame (the multi-GPU path of
bench.py with the render step replaced by the CPU oracle, which is allowed in tests)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, rpt, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from pathtracer_ocl_b200 import distributed as D, scene as S, trace as T
        sc = S.build_scene("default", W, H)
        seeds = S.make_seeds(0xABCD, W * H)
        rows = T.plan_rows(H, rank, world, rpt)
        # render only the rows this rank owns (contiguous runs of the interleaved tiles)
        parts = []
        for k in range(0, len(rows), rpt):
            r0, r1 = int(rows[k]), int(rows[min(k + rpt, len(rows)) - 1]) + 1
            img, _ = O.trace(sc, seeds, 1, 1, rows=(r0, r1), nthreads=1)
            parts.append(img.reshape(-1))
        local = torch.from_numpy(np.concatenate(parts) if parts else np.zeros(0))
        frame = D.gather_frame(local, H, W, rpt, dst=0)
        if rank == 0:
            np.save(out_path, frame.numpy())
        else:
            assert frame is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("H,rpt", [(22, 4), (16, 4), (9, 2)])
def test_two_rank_shard_and_gather_matches_full_frame(tmp_path, H, rpt):
    W = 24
    out_path = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(2, _free_port(), W, H, rpt, out_path), nprocs=2, join=True)
    from oracle import oracle as O
    from pathtracer_ocl_b200 import scene as S
    sc = S.build_scene("default", W, H)
    full, _ = O.trace(sc, S.make_seeds(0xABCD, W * H), 1, 1, nthreads=2)
    got = np.load(out_path)
    assert got.shape == (H, W, 4)
    assert np.array_equal(got, full)


def test_single_process_gather_is_identity():
    from pathtracer_ocl_b200 import distributed as D
    t = torch.arange(5 * 3 * 4, dtype=torch.float64)
    f = D.gather_frame(t, 5, 3)
    assert f.shape == (5, 3, 4) and torch.equal(f.reshape(-1), t)
