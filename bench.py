#!/usr/bin/env python
"""bench.py -- Mpaths/s of the path-tracing hot path on BASELINE.json's headline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision fp32|fp64]
                    [--rng parity|fast] [--scene NAME --width .. --height .. --samples ..]

A "step" is one full render of the workload (every pixel, every sample) by the CUDA path.  The
default workload is BASELINE.json configs[1]: the README reference scene, 1280x960 at 2048 spp,
aperture 0.15, focal length 1.6.  N > 1 (launched with torchrun, one process per GPU) shards the
same frame into interleaved 4-scanline tiles -- total work fixed, so "scaling": "strong" -- and
gathers the rows to rank 0 over NCCL inside the timed region.

`value`   : paths of the whole job / device time (CUDA events on the launching stream, max over
            ranks); scene, seeds and framebuffers already resident in HBM.
`e2e`     : the same metric through the one-shot C-ABI call ptc_render (the drop-in for the
            reference's ocl.Trace) with HOST buffers: scene flattening, allocation, H2D of scene +
            seeds, kernel, gather, D2H of the frame all inside the timed region.
`roofline`: SM FP32-issue roofline of the trace kernel: model flops per launch from
            the oracle's event counters x fixed weights, over the kernel's CUDA-event time.
`cpu_baseline` / `--impl reference`: the CPU restatement of the reference kernel (oracle, fp64) on
            the host cores.  Neither Go nor an OpenCL runtime exists in this image, so the
            reference itself cannot run; this is the "port" baseline, on a bounded sample.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--rng", default="parity", choices=["parity", "fast"])
    ap.add_argument("--scene", default="reference")
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--height", type=int, default=960)
    ap.add_argument("--samples", type=int, default=2048)
    ap.add_argument("--aperture", type=float, default=0.15)
    ap.add_argument("--focal-length", type=float, default=1.6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def workload_name(a):
    return (f"{a.scene} scene {a.width}x{a.height}@{a.samples}spp aperture={a.aperture:g} focal={a.focal_length:g} "
            f"(BASELINE.json configs[1])" if (a.scene, a.width, a.height, a.samples) == ("reference", 1280, 960, 2048)
            else f"{a.scene} scene {a.width}x{a.height}@{a.samples}spp aperture={a.aperture:g} focal={a.focal_length:g}")


# ---- CPU arm (oracle/ is allowed here as the timed baseline only) ----------------------------------------
# Two CPU implementations of the path exist: oracle/_ref -- the reference's OWN kernel source (tracer.cl) compiled for
# the host through oracle/cl_shim.hpp, kind "reference" -- and the oracle, a restatement of it (kind "port") that also
# counts events for the cost model.  The tests hold them bit-identical (tests/test_oracle_vs_reference.py).
def cpu_kernel(scene, seeds, threads):
    """(trace function, kind, note) of the CPU implementation to time: the compiled reference kernel when it is there --
    and when the scene stays inside the kernel's fixed 64-entry intersection arrays (tracer.cl:96-102), which the
    reference overflows silently and a CPU build would turn into memory corruption."""
    from oracle import oracle as O
    fits = O.trace(scene, seeds, 1, precision=1, nthreads=threads)[1]["max_intersections"] <= 60
    if fits and O.ref_lib() is not None:
        return (lambda scene, seeds, spp, threads: O.ref_trace(scene, seeds, spp, nthreads=threads), "reference",
                "the reference's own kernel source (internal/ocl/tracer.cl) compiled for the host CPU through oracle/cl_shim.hpp, "
                "work-items spread over all host threads; no OpenCL runtime exists in this image")
    return (lambda scene, seeds, spp, threads: O.trace(scene, seeds, spp, precision=1, nthreads=threads), "port",
            "CPU restatement of tracer.cl (oracle); the compiled reference kernel (oracle/_ref) is not present")


def cpu_sample(a, scene, seeds, target_s, threads):
    """Time the CPU implementation on the workload's frame at a reduced sample count (~target_s of CPU work); the event
    counters of the cost model come from the oracle on the same sample."""
    from oracle import oracle as O
    run, kind, note = cpu_kernel(scene, seeds, threads)
    px = a.width * a.height
    t0 = time.perf_counter()
    run(scene, seeds, 1, threads)
    t1 = time.perf_counter() - t0
    spp = int(max(2, min(a.samples, target_s / max(t1, 1e-3))))
    t0 = time.perf_counter()
    run(scene, seeds, spp, threads)
    dt = time.perf_counter() - t0
    _, cnt = O.trace(scene, seeds, max(2, spp // 4) if kind == "reference" else spp, precision=1, nthreads=threads)
    flops = O.model_flops(cnt, dof=a.aperture != 0.0)
    return dict(spp=spp, seconds=dt, mpaths=px * spp / dt / 1e6, flops_per_path=flops / cnt["paths"],
                segments_per_path=cnt["segments"] / cnt["paths"], kind=kind, note=note)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from pathtracer_ocl_b200 import scene as S
    threads = os.cpu_count() or 1
    scene = S.build_scene(a.scene, a.width, a.height, a.aperture, a.focal_length)
    seeds = S.make_seeds(0x5EED0002, a.width * a.height)
    run, kind, note = cpu_kernel(scene, seeds, threads)
    px = a.width * a.height
    budget = min(10.0, 150.0 / max(1, a.steps + a.warmup))
    t0 = time.perf_counter()
    run(scene, seeds, 1, threads)
    t1 = time.perf_counter() - t0
    spp = int(max(1, min(a.samples, budget / max(t1, 1e-3))))
    for _ in range(a.warmup):
        run(scene, seeds, spp, threads)
    times = []
    for _ in range(a.steps):
        t0 = time.perf_counter()
        run(scene, seeds, spp, threads)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    value = px * spp * a.steps / total / 1e6
    sample = f"{a.width}x{a.height} frame at {spp} of {a.samples} spp per step (cost is linear in spp)"
    line = {
        "impl": "reference", "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": total / a.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample},
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
            self._stop_evt.wait(0.1)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def main():
    a = parse_args()
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist
    from pathtracer_ocl_b200 import distributed as D, scene as S, trace as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        a.gpus = world
    if T.lib().ptc_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: pathtracer_ocl_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    W, H, spp = a.width, a.height, a.samples
    precision = T.FP64 if a.precision == "fp64" else T.FP32
    rng_mode = T.RNG_FAST if a.rng == "fast" else T.RNG_PARITY
    scene = S.build_scene(a.scene, W, H, a.aperture, a.focal_length)
    seeds = S.make_seeds(0x5EED0002, W * H)
    total_paths = W * H * spp

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    # ---- device-resident arm ("value") ----------------------------------------------------------------
    ctx = T.open_scene(scene, spp, seeds, precision=precision, rng_mode=rng_mode, devices=[local_rank],
                       shard_index=rank, shard_count=world)
    fb = D.framebuffer_tensor(ctx)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def step():
        ctx.trace()                                   # kernel(s) on the library's stream, CUDA-event timed inside
        k_ms = ctx.stats()["kernel_ms"]
        gather_ms = 0.0
        frame = None
        if world > 1:
            g0.record()
            frame = D.gather_frame(fb, H, W, 0, dst=0)
            g1.record()
            g1.synchronize()
            gather_ms = g0.elapsed_time(g1)
        return k_ms, gather_ms, frame

    for _ in range(a.warmup):
        step()
        flush.fill_(1)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    wall0 = time.perf_counter()
    dev_ms, kern_ms, launches = 0.0, [], 0
    for _ in range(a.steps):
        k, g, _ = step()
        dev_ms += k + g
        kern_ms.append(k)
        launches += ctx.stats()["kernel_launches"]
        flush.fill_(1)                                # L2 flush between timed iterations (not in dev_ms)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.finish()
    t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=dev)
    lt = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)      # kernels launched by all ranks in the timed region
    dev_ms_max, wall_ms_max = float(t[0]), float(t[1])
    launches = int(lt[0])
    value = total_paths * a.steps / (dev_ms_max / 1e3) / 1e6
    upload_stats = ctx.stats()

    # ---- end-to-end arm ("e2e"): one-shot public call with host buffers --------------------------------
    pinned_seeds = torch.from_numpy(seeds).pin_memory()
    my_rows = len(ctx.rows)
    pinned_out = torch.empty(H * W * 4 if rank == 0 else max(my_rows, 1) * W * 4, dtype=torch.float64).pin_memory()
    seeds_np = pinned_seeds.numpy()

    def e2e_step():
        if world == 1:
            job = T._Job(scene.objects, scene.triangles if scene.n_triangles else None, scene.groups if scene.n_groups else None,
                         scene.camera, scene.textures[0], scene.textures[1], scene.textures[2], seeds_np, spp, precision,
                         rng_mode, [local_rank], 0, 1, 0)
            import ctypes as C
            err = C.create_string_buffer(512)
            if T.lib().ptc_render(C.byref(job.struct), pinned_out.data_ptr(), err, 512) != 0:
                raise RuntimeError(err.value.decode())
            return
        c2 = T.open_scene(scene, spp, seeds_np, precision=precision, rng_mode=rng_mode, devices=[local_rank],
                          shard_index=rank, shard_count=world)
        c2.trace()
        frame = D.gather_frame(D.framebuffer_tensor(c2), H, W, 0, dst=0)
        if rank == 0:
            pinned_out.view(H, W, 4).copy_(frame, non_blocking=True)
        torch.cuda.synchronize()
        c2.close()

    e2e_step()
    barrier()
    e0 = time.perf_counter()
    e2e_steps = max(1, min(a.steps, 3))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - e0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_paths * e2e_steps / float(t[0]) / 1e6
    h2d = int(upload_stats["h2d_bytes"]) * world
    d2h = H * W * 32

    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- CPU baseline + roofline (rank 0, N=1 only for the CPU leg) ------------------------------------
    cpu = None
    flops_per_path = None
    if world == 1 and not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cs = cpu_sample(a, scene, seeds, a.cpu_seconds, threads)
        flops_per_path = cs["flops_per_path"]
        cpu = {"value": cs["mpaths"], "unit": "Mpaths/s", "cores": threads, "kind": cs["kind"], "note": cs["note"],
               "sample": f"{W}x{H} frame at {cs['spp']} of {spp} spp, {cs['seconds']:.1f} s (cost is linear in spp)",
               "extrapolated_full_config_wall_s": total_paths / (cs["mpaths"] * 1e6),
               "segments_per_path": cs["segments_per_path"]}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_mhz_peak = float(peaks.get("sm_max_mhz", 1965.0))
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    fp32_peak = sm_count * 128 * 2 * sm_mhz_peak * 1e6 / 1e12
    kernel_ms_avg = sum(kern_ms) / len(kern_ms)
    try:
        fma_measured = T.debug_fma_peak(local_rank)           # pure-FFMA kernel on this GPU, outside every timed region
    except Exception:
        fma_measured = None
    roofline = {"bound": "fp32_issue", "achieved": None, "peak": fp32_peak, "unit": "TFLOP/s", "frac": None, "traffic": None,
                "peak_source": f"{sm_count} SMs x 128 lanes x 2 x {sm_mhz_peak:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz)",
                "peak_fma_microbenchmark": fma_measured,
                "kernel": "ptk::trace_kernel", "kernel_ms": kernel_ms_avg,
                "hbm_algorithmic_bytes_per_launch": len(ctx.rows) * W * 40,
                "hbm_gbs_achieved": len(ctx.rows) * W * 40 / (kernel_ms_avg / 1e3) / 1e9,
                "hbm_peak_gbs": peaks.get("hbm_gbs")}
    flops_source = "oracle event counters on the cpu_baseline sample of this run"
    if flops_per_path is None:
        # the CPU leg is skipped (N > 1 or --no-cpu-baseline): use the figure recorded for this workload in
        # profiles/model_flops.json by an earlier N=1 run
        try:
            table = json.load(open(os.path.join(ROOT, "profiles", "model_flops.json")))
            key = f"{a.scene}_{W}x{H}_ap{a.aperture:g}"
            flops_per_path = table.get(key)
            flops_source = f"profiles/model_flops.json[{key}]"
        except Exception:
            flops_per_path = None
    if flops_per_path is not None:
        achieved = flops_per_path * (total_paths / world) / (kernel_ms_avg / 1e3) / 1e12
        roofline.update(achieved=achieved, frac=achieved / fp32_peak, model_flops_per_path=flops_per_path,
                        model_flops_source=flops_source)
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get(f"{a.scene}_{a.precision}")
        except Exception:
            pass

    line = {
        "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dev_ms_max / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32" if precision == T.FP32 else "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "rng": a.rng, "sharding": f"interleaved 4-scanline tiles over {world} GPU(s)",
                   "l2": "256 MB device memset between timed steps", "seeds": "splitmix64(0x5EED0002), one per pixel"},
        "wall_time_s_per_frame": dev_ms_max / a.steps / 1e3, "wall_ms_per_step_host_clock": wall_ms_max / a.steps,
        "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": float(t[0]) / e2e_steps * 1e3, "api": "ptc_render (C ABI), pinned host buffers"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
