#!/usr/bin/env python
"""bench.py — samples/sec of the path-tracing hot path on BASELINE.json's quoted configuration.

A "step" is one full pass of the hot path over the workload: every pixel of the 1200x800 random-
spheres scene (make-random-scene n=11 moving=true, scene seed 1, ~486 spheres), 10 samples per
pixel, depth 50 (BASELINE.json configs[1]).  `value` = samples/s with the scene resident in HBM
(kernels + reduce + resolve timed with CUDA events on the launching stream); `e2e` = the same
metric through the C-ABI calls a front end makes (rt_set_scene + rt_set_camera + render + image
back on the host), host<->device copies inside the timed region.

N > 1 (torchrun, one rank per GPU): every rank renders its own 10-spp sample slice of the same
frame (weak scaling: global spp = 10 N), the float sums are combined with one NCCL reduce over
NVLink and rank 0 resolves; time = max over ranks.

`--impl reference` times the CPU restatement of the reference (oracle/, double precision C++ with
OpenMP on all host threads) — the Clojure original cannot run here (no JVM in the image).
"""
from __future__ import annotations

import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FLOP_PER_TEST = 17.0            # oc (3) + oc.d (5) + oc.oc - r^2 (6) + b'^2 - a c' (3)
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # 74.45: SMs x lanes x 2 flop x max SM clock

WORKLOADS = {
    # name: (nx, ny, spp, depth, scene builder name, scene seed)
    "c2": (1200, 800, 10, 50, "random", 1),
    "c1": (200, 100, 100, 50, "random", 1),
    "c3": (3840, 2160, 1024, 50, "random", 1),          # use with --scaling strong at N GPUs (sample slices)
    "c3-slice": (3840, 2160, 16, 50, "random", 1),
    "c4": (1200, 800, 64, 50, "stress", 4),
    # config 5: brute-force scale sweep, 1920x1080, 64 spp
    "c5-100": (1920, 1080, 64, 50, "sweep:100", 5),
    "c5-1k": (1920, 1080, 64, 50, "sweep:1000", 5),
    "c5-10k": (1920, 1080, 64, 50, "sweep:10000", 5),
    "c5-100k": (1920, 1080, 64, 50, "sweep:100000", 5),
}


def build_scene(name, nx, ny, seed):
    import raytrace_clj_b200 as rt

    rng = random.Random(seed)
    if name == "random":
        sc = rt.scene.make_random_scene(nx, ny, 11, True, rng)
    elif name == "stress":
        sc = rt.scene.make_material_stress_scene(nx, ny, 11, rng)
    elif name.startswith("sweep:"):
        sc = rt.scene.make_scale_sweep_scene(nx, ny, int(name.split(":")[1]), rng)
    else:
        raise ValueError(name)
    flat = rt.native.marshal_world(sc["world"])
    cam_type, cam = rt.native.marshal_camera(sc["camera"])
    return flat, cam_type, cam


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def scene_bytes(flat):
    return flat.nbytes() + 24 * 4


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


def run_reference(args, world, rank):
    """The reference's own CPU implementation of the path — here its C++ restatement (oracle/), all host
    threads; each step = a bounded sample of the workload (2 of the 10 spp of every pixel)."""
    if rank != 0:
        return 0
    import oracle

    nx, ny, spp, depth, scene_name, scene_seed = WORKLOADS[args.workload]
    flat, cam_type, cam = build_scene(scene_name, nx, ny, scene_seed)
    S = oracle.Scene(flat)
    step_spp = max(1, min(spp, 2))
    cores = host_cores()            # torchrun exports OMP_NUM_THREADS=1: ask for every core explicitly
    for w in range(args.warmup):
        S.render_accumulate(cam_type, cam, nx, ny, 0, 1, depth, seed=w, n_threads=cores)
    t0 = time.perf_counter()
    samples = tests = 0
    for k in range(args.steps):
        _, c = S.render_accumulate(cam_type, cam, nx, ny, k * step_spp, step_spp, depth, seed=1, n_threads=cores)
        samples += c["samples"]
        tests += c["sphere_tests"]
    dt = time.perf_counter() - t0
    v = samples / dt
    sample = f"{nx}x{ny}, {step_spp} of {spp} spp per step, depth {depth}, brute force over {flat.n_spheres} spheres"
    line = {
        "impl": "reference", "metric": "samples_per_sec", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: random-spheres scene {nx}x{ny}, {spp} spp, depth {depth}, "
                               f"{flat.n_spheres} spheres (make-random-scene n=11 moving=true, scene seed {scene_seed})"},
        "tests_per_sec": tests / dt,
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "C++ double-precision restatement of raytrace-clj (oracle/); the Clojure original "
                                 "cannot run here (no JVM)"},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--variant", type=int, default=int(os.environ.get("RT_VARIANT", "1")), help="1 wavefront (default), 0 megakernel")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU renders the workload's spp (global spp = spp x N); strong: the spp are divided")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="default: min(steps, 10)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        return run_reference(args, world, rank)

    import torch
    import torch.distributed as dist

    import raytrace_clj_b200 as rt
    from raytrace_clj_b200 import build as rtbuild

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback")
    rtbuild.build_library()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    nx, ny, spp, depth, scene_name, scene_seed = WORKLOADS[args.workload]
    if args.scaling == "strong":
        if spp % world:
            raise SystemExit(f"--scaling strong: {spp} spp do not divide over {world} GPUs")
        spp //= world                                # per-GPU sample slice
    flat, cam_type, cam = build_scene(scene_name, nx, ny, scene_seed)
    r = rt.native.Renderer([local_rank])
    r.set_scene(flat)
    r.set_camera(cam_type, cam)
    info = r.device_info()

    d_sum = torch.zeros(ny, nx, 3, device=dev, dtype=torch.float32)
    d_rgb = torch.zeros(ny, nx, 3, device=dev, dtype=torch.uint8)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)   # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    total_spp = spp * world

    def step(seed):
        d_sum.zero_()
        r.render_accumulate_device(nx, ny, rank * spp, spp, d_sum.data_ptr(), max_depth=depth, seed=seed,
                                   variant=args.variant, stream=stream.cuda_stream, sync=False)
        if world > 1:
            dist.reduce(d_sum, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            r.resolve_device(nx, ny, total_spp, d_sum.data_ptr(), d_rgb.data_ptr(), stream=stream.cuda_stream, sync=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    fp32_peak = r.measure_fp32_peak() if rank == 0 else (0.0, 0.0)

    for w in range(args.warmup):
        step(1000 + w)
    barrier()
    r.reset_counters()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        flush.fill_(float(k))                       # L2 flush between timed iterations (untimed)
        ev[k][0].record(stream)
        d_sum.zero_()
        kev[k][0].record(stream)
        r.render_accumulate_device(nx, ny, rank * spp, spp, d_sum.data_ptr(), max_depth=depth, seed=1 + k,
                                   variant=args.variant, stream=stream.cuda_stream, sync=False)
        kev[k][1].record(stream)
        if world > 1:
            dist.reduce(d_sum, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            r.resolve_device(nx, ny, total_spp, d_sum.data_ptr(), d_rgb.data_ptr(), stream=stream.cuda_stream, sync=False)
        ev[k][1].record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    my_ms = sum(a.elapsed_time(b) for a, b in ev)
    my_kernel_ms = sum(a.elapsed_time(b) for a, b in kev)
    ctr = r.counters()
    launches = ctr["kernel_launches"]
    t = torch.tensor([my_ms, my_kernel_ms], device=dev, dtype=torch.float64)
    cnt = torch.tensor([ctr["samples"], ctr["rays"], ctr["sphere_tests"], ctr["candidates"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    total_ms, kernel_ms = float(t[0]), float(t[1])
    samples, rays, tests, cands = (float(x) for x in cnt)
    value = samples / (total_ms * 1e-3)

    # ---- end to end through the C ABI with host buffers ------------------------------------
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    out_img = np.empty((ny, nx, 3), np.uint8)
    host_rgb = torch.empty(ny, nx, 3, dtype=torch.uint8, pin_memory=True)

    def e2e_step(seed):
        r.set_scene(flat)                           # H2D: the marshalled SoA scene
        r.set_camera(cam_type, cam)
        if world == 1:
            r.render(nx, ny, spp, depth, seed=seed, variant=args.variant, linear=False, rgb8=True, out_rgb8=out_img)
        else:
            d_sum.zero_()
            r.render_accumulate_device(nx, ny, rank * spp, spp, d_sum.data_ptr(), max_depth=depth, seed=seed,
                                       variant=args.variant, stream=stream.cuda_stream, sync=False)
            dist.reduce(d_sum, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                r.resolve_device(nx, ny, total_spp, d_sum.data_ptr(), d_rgb.data_ptr(), stream=stream.cuda_stream, sync=False)
                host_rgb.copy_(d_rgb, non_blocking=True)
            torch.cuda.synchronize(dev)

    e2e_step(500)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(600 + k)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = (nx * ny * spp * world * e2e_steps) / float(te[0])

    # ---- stage shares: one extra (untimed) step with per-stage CUDA events (rank 0, wavefront only) ------
    stage = None
    if rank == 0 and args.variant == 1:
        r.reset_counters()
        r.set_profile(True)
        d_sum.zero_()
        r.render_accumulate_device(nx, ny, rank * spp, spp, d_sum.data_ptr(), max_depth=depth, seed=1,
                                   variant=args.variant, stream=stream.cuda_stream, sync=True)
        r.set_profile(False)
        pc = r.counters()
        stage = {k: pc[k + "_ns"] * 1e-6 for k in ("cull", "refine", "tiebreak", "shade")}
        stage["tests"] = pc["sphere_tests"]
    barrier()

    scene_desc = {"random": "make-random-scene n=11 moving=true", "stress": "make-random-scene n=11, 10/45/45 % Lambert/metal/glass"}.get(
        scene_name, "5 hero objects + static r=0.2 spheres on a grid, scene.clj:369-375 placement rule")
    if rank == 0:
        achieved_tflops = FLOP_PER_TEST * (tests / world) / (kernel_ms * 1e-3) / 1e12   # per GPU, whole render step
        peak_measured = max(fp32_peak)
        line = {
            "metric": "samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"{args.workload}: random-spheres scene {nx}x{ny}, {spp} spp per GPU, depth {depth}, "
                            f"{flat.n_spheres} spheres ({scene_desc}, scene seed {scene_seed})",
                "variant": "megakernel" if args.variant == 0 else "wavefront",
                "parallelism": f"sample-slice x{world}" if world > 1 else "single GPU",
                "l2": "flushed between timed iterations (256 MiB fill); the scene itself is staged in shared memory",
                "precision": "FP32 cull over all spheres, FP64 refine of survivors, FP32 shading",
            },
            "tests_per_sec": tests / (total_ms * 1e-3),
            "rays_per_sample": rays / samples,
            "cull_survivors_per_ray": cands / rays,
            "kernel_ms_per_step": kernel_ms / args.steps,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {
                "bound": "fp32", "achieved": achieved_tflops, "peak": peak_measured, "unit": "TFLOP/s",
                "frac": achieved_tflops / peak_measured if peak_measured else None,
                "peak_source": "measured on this box: FFMA-chain microbenchmark (rt_measure_fp32_peak), max of scalar "
                               "FFMA and packed FFMA2",
                "peak_ffma_tflops": fp32_peak[0], "peak_ffma2_tflops": fp32_peak[1],
                "peak_nominal": NOMINAL_FP32_TFLOPS, "frac_of_nominal": achieved_tflops / NOMINAL_FP32_TFLOPS,
                "flop_per_test": FLOP_PER_TEST,
                # dram__bytes_read.sum + dram__bytes_write.sum of the largest wf_cull launch of a C2 frame (a lane's 4.8 M
                # camera rays), from the committed `ncu --set full` capture (profiles/r01_ncu_wf_cull_common_summary.txt)
                "traffic": 343.5e6 if args.workload == "c2" and args.variant == 1 else None,
                "traffic_note": "bytes per wf_cull launch (ncu): 72 B per ray (ray record in, pairs and closest-hit words out); "
                                "the kernel is FP32-issue bound, not HBM bound (0.6 TB/s)",
                "note": "non-tensor FP32 pipe; HBM is not the bound (scene in shared memory). `achieved` divides the "
                        "17-flop tests by the time of the WHOLE render step (cull + refine + tie-break + shade kernels)",
                "dominant_kernel": None if not stage else {
                    "name": "wf_cull", "ms_per_step": stage["cull"],
                    "achieved": FLOP_PER_TEST * stage["tests"] / (stage["cull"] * 1e-3) / 1e12,
                    "frac": FLOP_PER_TEST * stage["tests"] / (stage["cull"] * 1e-3) / 1e12 / peak_measured,
                    "share_of_step": stage["cull"] / max(1e-9, sum(stage[k] for k in ("cull", "refine", "tiebreak", "shade"))),
                    "other_stages_ms": {k: stage[k] for k in ("refine", "tiebreak", "shade")},
                },
            },
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": scene_bytes(flat),
                    "d2h_bytes_per_step": nx * ny * 3, "steps": e2e_steps},
            "device": info,
        }
        if world == 1 and not args.no_cpu_baseline:
            import oracle

            S = oracle.Scene(flat)
            cores = host_cores()
            # bounded sample: about 15 s of CPU work at ~0.25 G tests/s per core (2.6 rays per sample)
            budget = 15.0 * 0.25e9 * cores / (nx * ny * 2.6 * flat.n_spheres)
            cpu_spp = int(max(1, min(spp, budget)))
            S.render_accumulate(cam_type, cam, nx, ny, 0, 1, depth, seed=9, n_threads=cores)          # warm-up
            t0 = time.perf_counter()
            _, c = S.render_accumulate(cam_type, cam, nx, ny, 0, cpu_spp, depth, seed=1, n_threads=cores)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {
                "value": c["samples"] / dt, "unit": "samples/s", "cores": cores, "kind": "port",
                "tests_per_sec": c["sphere_tests"] / dt, "seconds": dt,
                "sample": f"{nx}x{ny}, {cpu_spp} of {spp} spp, depth {depth}, brute force, double precision",
                "note": "C++ restatement of raytrace-clj (oracle/), OpenMP; the Clojure original cannot run here (no JVM)",
            }
        print(json.dumps(line))
    r.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
