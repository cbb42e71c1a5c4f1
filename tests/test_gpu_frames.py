"""Round-2 additions to the C ABI on the device: frames (the gather fused into the trace kernel), the float32
readback, and the fast-slot intersection path's tie and overflow semantics."""
import os
import socket
import sys

import numpy as np
import pytest

from oracle import oracle as O
from pathtracer_ocl_b200 import scene as S, trace as T

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_contexts_render_into_one_frame():
    """Three shards of one frame, each its own context, all attached to one frame: the frame ends up holding the
    whole image, bit-identical to a single render; ptc_read refuses while a frame is attached."""
    W, H, spp = 96, 50, 3
    sc = S.build_scene("transparency", W, H)
    seeds = S.make_seeds(77, W * H)
    full = T.render_scene(sc, spp, seeds, precision=T.FP64)
    frame = T.Frame(0, W, H)
    assert np.all(frame.read() == 0.0)
    for r in range(3):
        with T.open_scene(sc, spp, seeds, precision=T.FP64, shard_index=r, shard_count=3) as ctx:
            ctx.set_frame(frame)
            ctx.trace()
            with pytest.raises(T.PtcError, match="attached frame"):
                ctx.read()
            ctx.set_frame(None)
            ctx.trace()
            assert np.array_equal(ctx.read().reshape(len(ctx.rows), W, 4), full[ctx.rows])
    assert np.array_equal(frame.read(), full)
    with pytest.raises(T.PtcError, match="frame is"):
        with T.open_scene(S.build_scene("default", 32, 24), 1, S.make_seeds(1, 32 * 24)) as ctx:
            ctx.set_frame(frame)
    frame.close()


def test_float32_frames_and_readback():
    W, H, spp = 64, 48, 2
    sc = S.build_scene("default", W, H)
    seeds = S.make_seeds(78, W * H)
    with T.open_scene(sc, spp, seeds, precision=T.FP32) as ctx:
        ctx.trace()
        f64 = ctx.read().reshape(H, W, 4).copy()
        f32 = ctx.read_f32()
        assert ctx.stats()["d2h_bytes"] == W * H * 16
        frame = T.Frame(0, W, H, T.FRAME_F32)
        ctx.set_frame(frame)
        ctx.trace()
        got = frame.read()
    assert f32.dtype == np.float32 and np.array_equal(f32, f64.astype(np.float32))
    assert got.dtype == np.float32 and np.array_equal(got, f64.astype(np.float32))
    frame.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _exchange_worker(rank, world, port, W, H, spp, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pathtracer_ocl_b200 import distributed as D, scene as S2, trace as T2
        n_dev = T2.lib().ptc_device_count()
        dev = rank % n_dev
        sc = S2.build_scene("teapot", W, H)
        seeds = S2.make_seeds(0xF00D, W * H)
        ex = D.FrameExchange(W, H, dev, dst=0)
        with T2.open_scene(sc, spp, seeds, devices=[dev], shard_index=rank, shard_count=world) as ctx:
            ctx.set_frame(ex.frame)
            ctx.trace()
        ex.barrier()
        if rank == 0:
            np.save(out_path, ex.frame.read())
        ex.barrier()
        ex.close()
    finally:
        dist.destroy_process_group()


def test_two_processes_render_into_one_frame_through_cuda_ipc(tmp_path):
    """The process-per-GPU path of bench.py in miniature: rank 0 owns the frame, rank 1 maps it (CUDA IPC) and both
    ranks' kernels store their rows into it.  Works on one GPU (both processes on device 0) and on several."""
    import torch.multiprocessing as mp
    W, H, spp = 96, 72, 2
    out_path = str(tmp_path / "frame.npy")
    mp.spawn(_exchange_worker, args=(2, _free_port(), W, H, spp, out_path), nprocs=2, join=True)
    sc = S.build_scene("teapot", W, H)
    full = T.render_scene(sc, spp, S.make_seeds(0xF00D, W * H))
    assert np.array_equal(np.load(out_path), full)


# ---- fast slots: ties, overflow into the slow loop, ellipsoids ---------------------------------------------------------
def _records(sc):
    """The scene's objects as raw 1024-byte rows (editable through .view(S.OBJECT_DTYPE))."""
    return sc.objects.reshape(-1, 1024).copy()


def _fields(rows):
    return rows.view(S.OBJECT_DTYPE).reshape(-1)


def _with_objects(sc, rows):
    return S.SceneBuffers(sc.name, sc.width, sc.height, np.ascontiguousarray(rows).reshape(-1), sc.triangles, sc.groups, sc.camera,
                          sc.textures)


def _check(sc, seeds, spp=2):
    ref, _ = O.trace(sc, seeds, spp, precision=1)
    for precision, tol in ((T.FP64, 1e-6), (T.FP32, 1e-3)):
        img = T.Trace(sc.objects, sc.triangles if sc.n_triangles else None, sc.groups if sc.n_groups else None, 0, spp, sc.camera,
                      seeds=seeds, precision=precision).reshape(sc.height, sc.width, 4)
        err = np.abs(img[..., :3] - ref[..., :3]).max(axis=-1)
        assert (err <= tol).mean() >= 0.999, (precision, float((err <= tol).mean()), float(err.max()))


def _set_transform(row, m):
    f = _fields(row)
    f["transform"][0] = m.reshape(-1)
    f["inverse"][0] = np.linalg.inv(m).reshape(-1)
    f["inverse_transpose"][0] = np.linalg.inv(m).T.reshape(-1)


def test_coincident_objects_resolve_to_the_lower_index_like_the_reference():
    """Two coplanar planes / two identical spheres with different colours: exactly equal t, the reference keeps the
    first recorded hit (tracer.cl:731-739).  Both orders, so a wrong tie rule cannot pass by luck."""
    W, H = 96, 72
    base = S.build_scene("reference", W, H)
    seeds = S.make_seeds(404, W * H)
    o = _records(base)
    f = _fields(o)
    planes = [i for i in range(len(o)) if f["type"][i] == 0]
    spheres = [i for i in range(len(o)) if f["type"][i] == 1 and f["emission"][i][0] == 0.0]
    for src in (planes[0], spheres[0]):
        dup = o[src:src + 1].copy()
        _fields(dup)["color"][0][:3] = (0.1, 0.9, 0.1)
        after = np.concatenate([o, dup])                              # duplicate last: the original keeps the hits
        before = np.concatenate([o[:src], dup, o[src:]])             # duplicate first: the duplicate takes them
        _check(_with_objects(base, after), seeds)
        _check(_with_objects(base, before), seeds)
        a = T.Trace(_with_objects(base, after).objects, None, None, 0, 1, base.camera, seeds=seeds)
        b = T.Trace(_with_objects(base, before).objects, None, None, 0, 1, base.camera, seeds=seeds)
        assert not np.array_equal(a, b)                               # the order is visible in the picture


def test_more_planes_and_spheres_than_fast_slots_and_out_of_pattern_orders():
    """14 objects: 9 planes (one more than the plane run holds) and 5 spheres, first in scene order and then shuffled so
    the "spheres, planes, spheres" pattern breaks and objects overflow into the slow loop; plus a rotated, stretched
    ellipsoid (slow loop) next to the scene's own axis-aligned one (fast kind 1)."""
    W, H = 80, 60
    base = S.build_scene("reference", W, H)
    seeds = S.make_seeds(405, W * H)
    o = _records(base)
    f = _fields(o)
    planes = o[f["type"] == 0]
    spheres = o[f["type"] == 1]
    extra = []
    for k in range(4):                                                 # extra planes: tilted copies of the floor, pushed back
        p = planes[0:1].copy()
        ang = 0.2 + 0.15 * k
        rot = np.array([[1, 0, 0, 0], [0, np.cos(ang), -np.sin(ang), 0], [0, np.sin(ang), np.cos(ang), 0], [0, 0, 0, 1.0]])
        m = _fields(p)["transform"][0].reshape(4, 4) @ rot
        m[2, 3] += 0.3 * k
        _set_transform(p, m)
        _fields(p)["color"][0][:3] = (0.2 + 0.2 * k, 0.5, 0.9 - 0.2 * k)
        extra.append(p)
    ell = spheres[-1:].copy()                                          # a rotated, stretched sphere: slow loop
    ang = 0.7
    rot = np.array([[np.cos(ang), -np.sin(ang), 0, 0], [np.sin(ang), np.cos(ang), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
    m = _fields(ell)["transform"][0].reshape(4, 4) @ rot @ np.diag([1.0, 0.4, 1.6, 1.0])
    m[0, 3] += 0.25
    _set_transform(ell, m)
    many = np.concatenate([o] + extra + [ell, spheres[-2:-1]])
    assert (_fields(many)["type"] == 0).sum() == 9 and len(many) == 14
    _check(_with_objects(base, many), seeds)
    rng = np.random.default_rng(5)
    for _ in range(3):
        _check(_with_objects(base, many[rng.permutation(len(many))]), seeds)
