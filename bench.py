#!/usr/bin/env python
"""bench.py -- Mpaths/s of the path-tracing hot path on BASELINE.json's configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1|2|3|4|5|5e|5c] [--impl reference]
                    [--precision fp32|fp64] [--rng parity|fast] [--no-extras] [--no-cpu-baseline]
                    [--scene NAME --width .. --height .. --samples .. --aperture .. --focal-length ..]

A "step" is one full render of the workload (every pixel, every sample) by the CUDA path.  The default
workload is BASELINE.json configs[1] (`--config 2`): the README reference scene, 1280x960 at 2048 spp,
aperture 0.15, focal length 1.6.  N > 1 (torchrun, one process per GPU) shards the same frame into
interleaved 4-scanline tiles -- total work fixed, so "scaling": "strong" -- and every rank's trace kernel
stores its finished pixels straight into a frame on rank 0's GPU (CUDA IPC mapping, NVLink peer stores from
the kernel epilogue): the gather is part of the timed kernel, there is no collective on the data path.

`value`    paths of the whole job / device time (CUDA events on the launching stream around the kernel,
           summed over the K steps, max over ranks); scene and seeds resident in HBM.
`e2e`      the same metric through the C ABI with HOST buffers, K steps: N = 1: the one-shot ptc_render
           (scene flattening, allocation, H2D of scene + seeds, kernel, D2H of the frame); N > 1: per rank
           ptc_open (H2D of the scene and of the seeds of the owned rows) + ptc_set_frame + ptc_trace, a
           barrier, then rank 0's ptc_frame_read into pinned host memory.
`roofline` SM FP32-issue roofline of the trace kernel (FP64 pipe in fp64 mode): model flops
           per launch from the oracle's event counters x fixed weights, over the kernel's CUDA-event time.
           `executed` / `traffic` are copied from the committed ncu summaries of the SAME kernel version.
`fp64`, `other_configs` (N = 1 default run): the same measurement for the fp64 mode of the headline
           config and for configs 3 / 4 / 5 (config 5 at a bounded sample count, stated), so one line
           carries every number the north star names.
`cpu_baseline` / `--impl reference`: the reference's own kernel source compiled for the host CPU
           (oracle/_ref), falling back to its CPU restatement (oracle, "port"); bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# BASELINE.json configs: name -> (scene, width, height, samples, aperture, focal length, label)
CONFIGS = {
    "1": ("default", 640, 480, 1, 0.0, 0.0, "BASELINE.json configs[0]"),
    "2": ("reference", 1280, 960, 2048, 0.15, 1.6, "BASELINE.json configs[1]"),
    "3": ("teapot", 1280, 960, 2048, 0.0, 0.0, "BASELINE.json configs[2]"),
    "4": ("gopher", 1280, 960, 2048, 0.0, 0.0, "BASELINE.json configs[3]"),
    "5": ("textures", 3840, 2160, 4096, 0.0, 0.0, "BASELINE.json configs[4], texture-mapped primitives"),
    "5e": ("envmap", 3840, 2160, 4096, 0.0, 0.0, "BASELINE.json configs[4], environment sphere"),
    "5c": ("cubemap", 3840, 2160, 4096, 0.0, 0.0, "BASELINE.json configs[4], environment cube + mesh"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="2", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--rng", default="parity", choices=["parity", "fast"])
    ap.add_argument("--scene", default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--samples", type=int, default=None)
    ap.add_argument("--aperture", type=float, default=None)
    ap.add_argument("--focal-length", type=float, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the fp64 / other_configs sub-records of the default N=1 run")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    a = ap.parse_args()
    scene, w, h, spp, apert, fl, label = CONFIGS[a.config]
    overridden = any(v is not None for v in (a.scene, a.width, a.height, a.samples, a.aperture, a.focal_length))
    a.scene = a.scene or scene
    a.width = a.width or w
    a.height = a.height or h
    a.samples = a.samples or spp
    a.aperture = apert if a.aperture is None else a.aperture
    a.focal_length = fl if a.focal_length is None else a.focal_length
    a.label = None if overridden else label
    return a


def workload_name(scene, w, h, spp, ap, fl, label=None):
    s = f"{scene} scene {w}x{h}@{spp}spp aperture={ap:g} focal={fl:g}"
    return f"{s} ({label})" if label else s


class quiet_stdout:
    """The reference kernel prints debug lines for one hard-coded pixel (tracer.cl:1131-1134, printf from device code;
    the CPU build inherits them): keep them out of this program's stdout, which carries the JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)

    def __exit__(self, *exc):
        os.dup2(self._saved, 1)
        os.close(self._null)
        os.close(self._saved)


# ---- CPU arm (oracle/ is allowed here as the timed baseline only) ----------------------------------------
# Two CPU implementations of the path exist: oracle/_ref -- the reference's OWN kernel source (tracer.cl) compiled for
# the host CPU through oracle/cl_shim.hpp, kind "reference" -- and the oracle, a restatement of it (kind "port") that also
# counts events for the cost model.  The tests hold them bit-identical (tests/test_oracle_vs_reference.py).
def cpu_kernel(scene, seeds, threads, spp_probe, counters=None):
    """(trace function, kind, note) of the CPU implementation to time: the compiled reference kernel when it is there --
    and when the scene stays inside the kernel's fixed 64-entry intersection arrays (tracer.cl:96-102), which the
    reference overflows silently and a CPU build would turn into memory corruption.  The check runs the oracle (which
    counts intersections per ray) on the first `spp_probe` samples of the seeds that will be timed and keeps a margin
    of 8 entries; a scene that comes closer is timed with the oracle instead."""
    from oracle import oracle as O
    if counters is None:
        counters = O.trace(scene, seeds, max(1, spp_probe), precision=1, nthreads=threads)[1]
    fits = counters["max_intersections"] <= 56
    if fits and O.ref_lib() is not None:
        def run_ref(scene, seeds, spp, threads):
            with quiet_stdout():
                return O.ref_trace(scene, seeds, spp, nthreads=threads)
        return (run_ref, "reference",
                "the reference's own kernel source (internal/ocl/tracer.cl) compiled for the host CPU through oracle/cl_shim.hpp "
                "(scalar g++ -O2 -ffp-contract=off, canonical double-precision sin, the kernel's debug printf live), "
                "work-items spread over all host threads; no OpenCL runtime exists in this image")
    return (lambda scene, seeds, spp, threads: O.trace(scene, seeds, spp, precision=1, nthreads=threads), "port",
            "CPU restatement of tracer.cl (oracle); the compiled reference kernel (oracle/_ref) is absent or the scene "
            "overflows its fixed 64-entry intersection arrays")


def cpu_sample(a, scene, seeds, target_s, threads):
    """Time the CPU implementation on the workload's frame at a reduced sample count (~target_s of CPU work); the event
    counters of the cost model come from the oracle on the same sample."""
    from oracle import oracle as O
    px = a.width * a.height
    t0 = time.perf_counter()
    _, cnt1 = O.trace(scene, seeds, 1, precision=1, nthreads=threads)
    t1 = time.perf_counter() - t0
    spp = int(max(2, min(a.samples, target_s / max(t1, 1e-3))))
    _, cnt = O.trace(scene, seeds, max(2, spp // 4), precision=1, nthreads=threads)     # counters + the overflow guard
    run, kind, note = cpu_kernel(scene, seeds, threads, max(2, spp // 4), counters=cnt)
    t0 = time.perf_counter()
    run(scene, seeds, spp, threads)
    dt = time.perf_counter() - t0
    flops = O.model_flops(cnt, dof=a.aperture != 0.0)
    return dict(spp=spp, seconds=dt, mpaths=px * spp / dt / 1e6, flops_per_path=flops / cnt["paths"],
                segments_per_path=cnt["segments"] / cnt["paths"], kind=kind, note=note)


def model_flops_small(scene_name, ap, fl, threads):
    """Model flops per path of a config from the oracle's event counters on a reduced frame (the per-path event mix
    depends on the scene and the camera, not on the resolution): 192x144 at 4 spp."""
    from oracle import oracle as O
    from pathtracer_ocl_b200 import scene as S
    sc = S.build_scene(scene_name, 192, 144, ap, fl, tex_scale=16)
    _, cnt = O.trace(sc, S.make_seeds(0x5EED0002, 192 * 144), 4, precision=1, nthreads=threads)
    return O.model_flops(cnt, dof=ap != 0.0) / cnt["paths"]


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from pathtracer_ocl_b200 import scene as S
    threads = os.cpu_count() or 1
    scene = S.build_scene(a.scene, a.width, a.height, a.aperture, a.focal_length)
    seeds = S.make_seeds(0x5EED0002, a.width * a.height)
    px = a.width * a.height
    budget = min(10.0, 150.0 / max(1, a.steps + a.warmup))
    from oracle import oracle as O
    t0 = time.perf_counter()
    O.trace(scene, seeds, 1, precision=1, nthreads=threads)
    t1 = time.perf_counter() - t0
    spp = int(max(1, min(a.samples, budget / max(t1, 1e-3))))
    run, kind, note = cpu_kernel(scene, seeds, threads, min(spp, 4))
    for _ in range(a.warmup):
        run(scene, seeds, spp, threads)
    times = []
    for _ in range(a.steps):
        t0 = time.perf_counter()
        run(scene, seeds, spp, threads)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    value = px * spp * a.steps / total / 1e6
    sample = (f"{a.width}x{a.height} frame at {spp} of {a.samples} spp per step: a rate metric, cost is linear in spp "
              f"(the b200 arm renders all {a.samples})")
    line = {
        "impl": "reference", "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": total / a.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a.scene, a.width, a.height, a.samples, a.aperture, a.focal_length, a.label),
                   "sample": sample, "rng": "parity", "seeds": "splitmix64(0x5EED0002), one per pixel",
                   "sharding": "none: host threads over work-items"},
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": threads, "kind": kind, "sample": sample, "note": note},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "extrapolated_full_config_wall_s": px * a.samples / (value * 1e6),
    }
    print(json.dumps(line), flush=True)
    return 0


# ---- clocks -----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while self.nv and not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- one measured workload ----------------------------------------------------------------------------------
class Env:
    """Process-wide handles of a bench run."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)      # > 126 MB L2
        self.peaks = {}
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.sm_count = torch.cuda.get_device_properties(self.dev).multi_processor_count
        self.sm_mhz_peak = float(self.peaks.get("sm_max_mhz", 1965.0))
        self._fma = {}

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def allsum(self, val):
        t = self.torch.tensor([val], dtype=self.torch.int64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t[0])

    def pipe_peak(self, fp64):
        """(nominal peak TFLOP/s, measured micro-benchmark TFLOP/s or None, description) of the issue roofline."""
        from pathtracer_ocl_b200 import trace as T
        lanes = 64 if fp64 else 128                     # FP64 FMA lanes per SM are half the FP32 ones on B200
        nominal = self.sm_count * lanes * 2 * self.sm_mhz_peak * 1e6 / 1e12
        if fp64 not in self._fma:
            try:
                self._fma[fp64] = T.debug_dfma_peak(self.local_rank) if fp64 else T.debug_fma_peak(self.local_rank)
            except Exception:
                self._fma[fp64] = None
        return nominal, self._fma[fp64], f"{self.sm_count} SMs x {lanes} lanes x 2 x {self.sm_mhz_peak:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz)"


def ncu_evidence(kind, scene, precision):
    """`roofline.traffic` / `roofline.executed` from the committed ncu summaries of THIS kernel version
    (profiles/traffic.json, profiles/executed.json, keyed "<kernel version>/<scene>_<precision>"); None when the
    committed captures belong to another kernel source."""
    from pathtracer_ocl_b200 import trace as T
    version = T.lib().ptc_version().decode().split("kernel ")[-1]
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", f"{kind}.json")))
    except Exception:
        return None
    return table.get(f"{version}/{scene}_{precision}")


def measure(env, scene_name, W, H, spp, ap, fl, precision_name, rng_name, steps, warmup, label=None, with_e2e=True,
            clocks=False, scene=None, seeds=None):
    """Device-resident and end-to-end throughput of one workload on env.world GPUs.  Returns a dict (all ranks)."""
    torch = env.torch
    from pathtracer_ocl_b200 import distributed as D, scene as S, trace as T
    precision = T.FP64 if precision_name == "fp64" else T.FP32
    rng_mode = T.RNG_FAST if rng_name == "fast" else T.RNG_PARITY
    world, rank, lr = env.world, env.rank, env.local_rank
    if scene is None:
        scene = S.build_scene(scene_name, W, H, ap, fl)
    if seeds is None:
        seeds = S.make_seeds(0x5EED0002, W * H)
    total_paths = W * H * spp

    exchange = D.FrameExchange(W, H, lr, dst=0) if world > 1 else None
    ctx = T.open_scene(scene, spp, seeds, precision=precision, rng_mode=rng_mode, devices=[lr], shard_index=rank, shard_count=world)
    if exchange is not None:
        ctx.set_frame(exchange.frame)

    def step():
        ctx.trace()                 # the kernel, CUDA-event timed inside the library on its own stream; with a frame attached
        return ctx.stats()["kernel_ms"]   # its epilogue stores the pixels into rank 0's frame: the gather is in this time

    for _ in range(warmup):
        step()
        env.flush.fill_(1)
    env.barrier()
    sampler = ClockSampler(lr) if clocks else None
    if sampler:
        sampler.start()
    wall0 = time.perf_counter()
    dev_ms, kern_ms, launches = 0.0, [], 0
    for _ in range(steps):
        k = step()
        dev_ms += k
        kern_ms.append(k)
        launches += ctx.stats()["kernel_launches"]
        if exchange is not None:
            exchange.barrier()                        # the frame on rank 0 is complete here (inside the host-clock figure)
        env.flush.fill_(1)                            # L2 flush between timed iterations (not in dev_ms)
    env.barrier()
    wall = time.perf_counter() - wall0
    clk = sampler.finish() if sampler else None
    dev_ms_max, wall_ms_max = env.allmax([dev_ms, wall * 1e3])
    launches = env.allsum(launches)
    out = {"scene": scene_name, "W": W, "H": H, "spp": spp, "precision": precision_name, "total_paths": total_paths,
           "value": total_paths * steps / (dev_ms_max / 1e3) / 1e6, "ms_per_step": dev_ms_max / steps,
           "wall_ms_per_step_host_clock": wall_ms_max / steps, "kernel_ms": sum(kern_ms) / len(kern_ms), "launches": launches,
           "clocks": clk, "rows": len(ctx.rows), "h2d_open": int(ctx.stats()["h2d_bytes"]), "steps": steps, "warmup": warmup,
           "workload": workload_name(scene_name, W, H, spp, ap, fl, label)}
    gathered = None
    if exchange is not None and rank == 0:
        gathered = exchange.frame.read()              # for the single-process check below

    # ---- end to end: the public call with host buffers, `steps` times -----------------------------------------
    if with_e2e:
        pinned_seeds = torch.from_numpy(seeds).pin_memory()
        seeds_np = pinned_seeds.numpy()
        pinned_out = torch.empty(H * W * 4 if rank == 0 else 4, dtype=torch.float64).pin_memory()
        import ctypes as C

        def e2e_step():
            if world == 1:
                job = T._Job(scene.objects, scene.triangles if scene.n_triangles else None, scene.groups if scene.n_groups else None,
                             scene.camera, scene.textures[0], scene.textures[1], scene.textures[2], seeds_np, spp, precision,
                             rng_mode, [lr], 0, 1, 0)
                err = C.create_string_buffer(512)
                if T.lib().ptc_render(C.byref(job.struct), pinned_out.data_ptr(), err, 512) != 0:
                    raise RuntimeError(err.value.decode())
                return
            c2 = T.open_scene(scene, spp, seeds_np, precision=precision, rng_mode=rng_mode, devices=[lr], shard_index=rank, shard_count=world)
            c2.set_frame(exchange.frame)
            c2.trace()
            exchange.barrier()
            if rank == 0:
                exchange.frame.read(pinned_out.data_ptr())
            c2.close()

        for _ in range(2):                 # untimed: the first one-shot calls of a process still grow the memory pool
            e2e_step()
        env.barrier()
        e0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        env.barrier()
        (e2e_s,) = env.allmax([time.perf_counter() - e0])
        out["e2e"] = {"value": total_paths * steps / e2e_s / 1e6, "unit": "Mpaths/s",
                      "h2d_bytes_per_step": env.allsum(out["h2d_open"]), "d2h_bytes_per_step": H * W * 32,
                      "ms_per_step": e2e_s / steps * 1e3, "steps": steps,
                      "api": ("ptc_render (C ABI, one call), pinned host buffers" if world == 1 else
                              "per rank ptc_open + ptc_set_frame + ptc_trace (pixels stored into rank 0's frame by the kernel, NVLink) "
                              "+ barrier + rank 0 ptc_frame_read + ptc_close (C ABI), pinned host buffers")}

    # ---- N > 1: the one-process-N-GPUs path (ptc_job.devices), checked against the gathered frame ------------------
    if world > 1:
        env.barrier()
        if rank == 0:
            t0 = time.perf_counter()
            one = T.render_scene(scene, spp, seeds, precision=precision, rng_mode=rng_mode, devices=list(range(world)))
            ms = (time.perf_counter() - t0) * 1e3
            out["single_process_multi_gpu"] = {"bit_identical_to_gathered": bool(np.array_equal(one, gathered)),
                                               "max_abs_diff": float(np.abs(one - gathered).max()), "ms": ms,
                                               "devices": world, "api": "ptc_render with ptc_job.devices = [0..N-1], one process"}
        env.barrier()
    ctx.close()
    if exchange is not None:
        exchange.close()
    return out


def roofline_of(env, m, flops_per_path, flops_source):
    fp64 = m["precision"] == "fp64"
    nominal, measured, source = env.pipe_peak(fp64)
    algo_bytes = m["rows"] * m["W"] * 40          # 8 B seed in + 32 B RGBA out per pixel
    r = {"bound": "fp64_pipe" if fp64 else "fp32_issue", "achieved": None, "peak": nominal, "unit": "TFLOP/s", "frac": None,
         "traffic": ncu_evidence("traffic", m["scene"], m["precision"]), "peak_source": source,
         "peak_fma_microbenchmark": measured, "kernel": "ptk::trace_kernel", "kernel_ms": m["kernel_ms"],
         "executed": ncu_evidence("executed", m["scene"], m["precision"]),
         "hbm_algorithmic_bytes_per_launch": algo_bytes,
         "hbm_gbs_achieved": algo_bytes / (m["kernel_ms"] / 1e3) / 1e9, "hbm_peak_gbs": env.peaks.get("hbm_gbs")}
    if flops_per_path is not None:
        achieved = flops_per_path * (m["total_paths"] / env.world) / (m["kernel_ms"] / 1e3) / 1e12
        r.update(achieved=achieved, frac=achieved / nominal, model_flops_per_path=flops_per_path, model_flops_source=flops_source)
    return r


def stored_flops(scene, W, H, ap):
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "model_flops.json")))
        key = f"{scene}_{W}x{H}_ap{ap:g}"
        return table.get(key), f"profiles/model_flops.json[{key}]"
    except Exception:
        return None, None


def main():
    a = parse_args()
    if a.impl == "reference":
        return run_reference(a)

    from pathtracer_ocl_b200 import scene as S, trace as T
    if T.lib().ptc_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: pathtracer_ocl_b200 has no CPU fallback")
    env = Env()
    if env.world > 1:
        a.gpus = env.world
    threads = os.cpu_count() or 1
    W, H, spp = a.width, a.height, a.samples
    scene = S.build_scene(a.scene, W, H, a.aperture, a.focal_length)
    seeds = S.make_seeds(0x5EED0002, W * H)

    m = measure(env, a.scene, W, H, spp, a.aperture, a.focal_length, a.precision, a.rng, a.steps, a.warmup, label=a.label,
                clocks=True, scene=scene, seeds=seeds)

    extras = env.world == 1 and not a.no_extras and a.label is not None and a.config == "2" and a.precision == "fp32"
    fp64 = None
    others = {}
    if extras:
        k, w = min(a.steps, 3), 3
        fp64 = measure(env, a.scene, W, H, spp, a.aperture, a.focal_length, "fp64", a.rng, k, w, label=a.label, scene=scene, seeds=seeds)
        for key, cap in (("3", None), ("4", None), ("5", 256), ("5e", 512), ("5c", 256)):
            sc_name, w_, h_, spp_, ap_, fl_, label_ = CONFIGS[key]
            s_ = spp_ if cap is None else min(spp_, cap)
            lab = label_ if cap is None else f"{label_}; bounded: {s_} of {spp_} spp per step, a rate metric"
            others[f"cfg{key}_{sc_name}"] = measure(env, sc_name, w_, h_, s_, ap_, fl_, "fp32", a.rng, k, w, label=lab, with_e2e=(cap is None))

    if env.rank != 0:
        if env.world > 1:
            env.dist.destroy_process_group()
        return 0

    # ---- CPU baseline + rooflines (rank 0; the CPU leg at N = 1 only) --------------------------------------------
    cpu, flops_per_path, flops_source = None, None, None
    if env.world == 1 and not a.no_cpu_baseline:
        cs = cpu_sample(a, scene, seeds, a.cpu_seconds, threads)
        flops_per_path, flops_source = cs["flops_per_path"], "oracle event counters on the cpu_baseline sample of this run"
        cpu = {"value": cs["mpaths"], "unit": "Mpaths/s", "cores": threads, "kind": cs["kind"], "note": cs["note"],
               "sample": f"{W}x{H} frame at {cs['spp']} of {spp} spp, {cs['seconds']:.1f} s (cost is linear in spp)",
               "extrapolated_full_config_wall_s": m["total_paths"] / (cs["mpaths"] * 1e6),
               "segments_per_path": cs["segments_per_path"]}
    if flops_per_path is None:
        flops_per_path, flops_source = stored_flops(a.scene, W, H, a.aperture)
    line = {
        "metric": "Mpaths/s", "value": m["value"], "unit": "Mpaths/s", "n_gpus": env.world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32" if a.precision == "fp32" else "f64", "data": "synthetic",
        "config": {"workload": m["workload"], "rng": a.rng, "sharding": f"interleaved 4-scanline tiles over {env.world} GPU(s)"
                   + ("; every rank's kernel stores into rank 0's frame (fused gather)" if env.world > 1 else ""),
                   "l2": "256 MB device memset between timed steps", "seeds": "splitmix64(0x5EED0002), one per pixel",
                   "kernel_version": T.lib().ptc_version().decode()},
        "wall_time_s_per_frame": m["ms_per_step"] / 1e3, "wall_ms_per_step_host_clock": m["wall_ms_per_step_host_clock"],
        "e2e": m["e2e"], "gpu_launches": m["launches"], "clocks": m["clocks"],
        "roofline": roofline_of(env, m, flops_per_path, flops_source),
    }
    if "single_process_multi_gpu" in m:
        line["single_process_multi_gpu"] = m["single_process_multi_gpu"]
    if cpu is not None:
        line["cpu_baseline"] = cpu

    def sub_record(mm, fpp, src):
        rec = {"workload": mm["workload"], "value": mm["value"], "unit": "Mpaths/s", "ms_per_step": mm["ms_per_step"],
               "steps": mm["steps"], "warmup": mm["warmup"], "dtype": "f32" if mm["precision"] == "fp32" else "f64",
               "gpu_launches": mm["launches"], "roofline": roofline_of(env, mm, fpp, src)}
        if "e2e" in mm:
            rec["e2e"] = mm["e2e"]
        return rec

    if fp64 is not None:
        line["fp64"] = sub_record(fp64, flops_per_path, flops_source)
    if others:
        line["other_configs"] = {}
        for key, mm in others.items():
            try:
                fpp = model_flops_small(mm["scene"], 0.0, 0.0, threads)
                src = "oracle event counters on a 192x144@4spp frame of the same scene (event mix per path is resolution-independent)"
            except Exception:
                fpp, src = stored_flops(mm["scene"], mm["W"], mm["H"], 0.0)
            line["other_configs"][key] = sub_record(mm, fpp, src)
    print(json.dumps(line), flush=True)
    if env.world > 1:
        env.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
