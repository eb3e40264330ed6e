#!/usr/bin/env python
"""bench.py — headline benchmark of the render path (BASELINE.json: Mrays/s and 4K frame time at 1/2/4/8
B200, FP32-pipe roofline %).

A "step" = one full frame of the workload (default C3: 3840x2160, 16 spp, depth 5, 1024 spheres + plane,
the configuration the metric's "4K frame time at 1/2/4/8 B200" is quoted on), tile-partitioned over the N
GPUs of one box, frame assembled on rank 0.  A ray = one nearest-hit query (ray_color call with depth > 0).

  python bench.py --gpus 1 --steps K --warmup W          # this repo's CUDA path
  python bench.py --impl reference ...                    # the reference algorithm on the host cores (oracle port)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  `value` = whole-job Mrays/s with the scene resident in HBM and the frame
left in HBM; `e2e` = the same through the C ABI with host buffers (scene upload + BVH build + render + frame
download every step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "ray-tracer-s8_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "Mrays/s"
UNIT = "Mrays/s"


# --------------------------------------------------------------------------------------------------------
# FLOP model (add/sub/mul/div/sqrt = 1, FMA = 2, compares/min/max/int = 0), applied to
# the kernel's own counters.  Ray::new is counted inside "primary" (91) and "shade" (79/91), not per ray.
# --------------------------------------------------------------------------------------------------------
FLOPS = dict(primary=91, slab=12, sphere_test=24, sphere_hit=23, tri_base=20, tri_s1=10, tri_s2=16, tri_s3=6,
             tri_hit=15, shade_sphere=79, shade_tri=91, emissive=3, sky=22, pixel=9)


def algorithmic_flops(st: dict, pixels: int) -> float:
    return (st["primary"] * FLOPS["primary"] + st["slab_tests"] * FLOPS["slab"]
            + st["sphere_tests"] * FLOPS["sphere_test"] + st["sphere_hits"] * FLOPS["sphere_hit"]
            + st["tri_tests"] * FLOPS["tri_base"] + st["tri_stage"][0] * FLOPS["tri_s1"]
            + st["tri_stage"][1] * FLOPS["tri_s2"] + st["tri_stage"][2] * FLOPS["tri_s3"]
            + st["tri_hits"] * FLOPS["tri_hit"] + st["shades_sphere"] * FLOPS["shade_sphere"]
            + st["shades_tri"] * FLOPS["shade_tri"] + st["emissive"] * FLOPS["emissive"] + st["sky"] * FLOPS["sky"]
            + pixels * FLOPS["pixel"])


# --------------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md)
# --------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw)}


# --------------------------------------------------------------------------------------------------------
# CPU legs (the oracle: a port of the reference's algorithm — the Rust reference cannot be built here)
# --------------------------------------------------------------------------------------------------------
REF_DIVISIONS = 20  # the controller's split (ray-tracer-controller/src/main.rs:33-36): one band = one slave request


def sample_divisions(k: int):
    """k evenly spaced bands of the 20-division split: a bounded, representative sample of the frame."""
    return [int((i + 0.5) * REF_DIVISIONS / k) for i in range(k)]


def oracle_sample(O, sp, tr, cfg, divs, threads=0):
    """Render the given divisions with the oracle, each like one reference slave request (world ingest + BVH
    build + band render, ray-tracer-slave/src/main.rs:37-83); returns (rays, seconds)."""
    rays, secs = 0, 0.0
    for d in divs:
        p = O.make_params(cfg["width"], cfg["height"], REF_DIVISIONS, d, cfg["spp"], cfg["max_bounces"], 0)
        t0 = time.perf_counter()
        _, st = O.render_rows(sp, tr, p, threads=threads, want_stats=True)
        secs += time.perf_counter() - t0
        rays += st["rays"]
    return rays, secs


def pick_sample(O, sp, tr, cfg, target_s: float):
    """Size the sample for ~target_s seconds of CPU work from a one-band probe (the middle band)."""
    assert cfg["height"] % REF_DIVISIONS == 0
    _, s = oracle_sample(O, sp, tr, cfg, [REF_DIVISIONS // 2])
    k = 1
    for cand in (2, 4, 5, 10, 20):
        if cand * s <= target_s:
            k = cand
    return sample_divisions(k)


def describe_sample(cfg, divs):
    rows = cfg["height"] // REF_DIVISIONS
    return (f"{len(divs)} of the controller's {REF_DIVISIONS} divisions (bands {divs}, {rows} rows x {cfg['width']} px each, "
            f"{cfg['spp']} spp), each rendered like one slave request incl. BVH build")


def run_reference(args, cfg, sp, tr):
    """--impl reference: the reference's CPU algorithm (oracle port, all host threads) on the same config."""
    from oracle import oracle as O

    O.build()
    cores = O.hardware_threads()
    divs = pick_sample(O, sp, tr, cfg, target_s=4.0)
    for _ in range(args.warmup):
        oracle_sample(O, sp, tr, cfg, divs[:1])
    rays, secs = 0, 0.0
    for _ in range(args.steps):
        r, s = oracle_sample(O, sp, tr, cfg, divs)
        rays += r
        secs += s
    val = rays / secs / 1e6
    sample = describe_sample(cfg, divs) + " per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cfg),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "C++ oracle = operation-for-operation port of the Rust slave (reference BVH build + "
                                 "unordered traversal); omits the reference's per-ray heap allocations, so it flatters the CPU"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


def workload_config(args, cfg):
    return {"workload": f"{args.workload}: {cfg['width']}x{cfg['height']}, {cfg['spp']} spp, max_bounces {cfg['max_bounces']} "
                        f"(<= {cfg['max_bounces'] + 1} queries/sample), {cfg['n_spheres']} spheres"
                        + (" + plane (2 triangles)" if cfg["plane"] else "") + ", scene_seed 0, seed 0",
            "width": cfg["width"], "height": cfg["height"], "spp": cfg["spp"], "max_bounces": cfg["max_bounces"],
            "n_spheres": cfg["n_spheres"], "n_triangles": 2 if cfg["plane"] else 0,
            "partition": f"8x4-pixel tiles, rotating interleave over {args.gpus} GPU(s)"}


# --------------------------------------------------------------------------------------------------------
def _claim_stdout():
    """Keep fd 1 for the ONE JSON line: libraries (NCCL prints its version banner there) get stderr instead."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--mode", default=os.environ.get("RT_B200_GATHER", "p2p"), choices=["p2p", "nccl"])
    ap.add_argument("--intersector", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    from rt_b200 import scenes

    cfg = scenes.CONFIGS[args.workload]
    sp, tr = scenes.config_scene(args.workload)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            real_stdout.write(json.dumps(run_reference(args, cfg, sp, tr)) + "\n")
            real_stdout.flush()
        return 0

    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd, stdout=real_stdout.fileno())

    import torch

    import rt_b200 as rt
    from rt_b200 import multi

    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def allreduce(x, op):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=op)
        return float(t.item())

    torch.cuda.set_device(local_rank)
    ctx = rt.Context(local_rank)
    info = ctx.device_info()
    scene = ctx.scene(sp, tr)
    W, H = cfg["width"], cfg["height"]
    pixels = W * H
    params = rt.make_params(W, H, spp=cfg["spp"], max_bounces=cfg["max_bounces"], intersector=args.intersector)

    sched = multi.FrameScheduler(ctx, rank, world, mode=args.mode)
    gather_mode = args.mode if world > 1 else "none"
    try:
        sched.setup(W, H)
    except rt.RtError as e:
        if world > 1 and args.mode == "p2p":
            raise RuntimeError(f"peer-mapped frame unavailable ({e}); rerun with --mode nccl") from e
        raise

    # ---- counters for the roofline: one untimed instrumented frame (deterministic → same work as timed steps)
    pc = rt.make_params(W, H, spp=cfg["spp"], max_bounces=cfg["max_bounces"], intersector=args.intersector,
                        collect_counters=True)
    cst = sched.render(scene, pc, want_stats=True)
    fp32_peak_tflops, _ = ctx.measure_fp32_peak()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")  # > 126 MB L2

    def step():
        flush.zero_()
        torch.cuda.current_stream().synchronize()
        return sched.render(scene, params, want_stats=True)

    for _ in range(args.warmup):
        step()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    kernel_ms, rays_step = [], 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = step()
        kernel_ms.append(st["kernel_ms"])
        rays_step = st["rays"]
    barrier()
    elapsed = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    ROp = dist.ReduceOp if dist is not None else None
    elapsed_max = allreduce(elapsed, ROp.MAX if ROp else None)
    rays_total = allreduce(float(rays_step), ROp.SUM if ROp else None)
    value = rays_total * args.steps / elapsed_max / 1e6
    ms_per_step = elapsed_max / args.steps * 1e3

    # ---- e2e: through the C ABI with host buffers; scene upload (+BVH build) and frame download every step
    host_frame = ctx.pinned_empty((H, W, 3)) if rank == 0 else None
    scene_bytes = scene.device_bytes

    def e2e_step():
        sc = ctx.scene(sp, tr)                       # H2D: primitive SoA + BVH + materials
        if world == 1:
            ctx.render_frame(sc, params, out=host_frame)   # kernel + D2H of the frame
        else:
            sched.render(sc, params)
            if rank == 0:
                sched.download(host_frame)
        sc.close()

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_elapsed = allreduce(time.perf_counter() - t0, ROp.MAX if ROp else None)
    e2e_value = rays_total * args.steps / e2e_elapsed / 1e6

    # ---- roofline of the render kernel on rank 0 (FP32 CUDA-core pipe; HBM traffic is negligible here)
    k_ms = float(np.mean(kernel_ms))
    flops = algorithmic_flops(cst, pixels // world)
    achieved = flops / (k_ms * 1e-3) / 1e12
    alg_bytes = scene_bytes + (pixels // world) * 3
    roofline = {
        "bound": "fp32", "kernel": "render_kernel_lanes<%s>" % ("BVH" if cst["intersector_used"] == 2 else "BRUTE"),
        "achieved": achieved, "peak": fp32_peak_tflops, "unit": "TFLOP/s", "frac": achieved / fp32_peak_tflops,
        "peak_source": "measured live: FFMA-chain micro-benchmark (rt_measure_fp32_peak); MEASURED_PEAKS.json has no CUDA-core figure",
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture of this
        # kernel on this workload at 1 GPU (profiles/r1_k2_lanes_final4_ncu_summary.txt); not re-measured live
        "traffic": (0.965376e6 + 0.911360e6) if (args.workload == "C3" and world == 1) else None,
        "traffic_unit": "bytes per launch (ncu)",
        "algorithmic_flops_per_launch": flops, "kernel_ms_avg": k_ms, "kernel_share_of_step": k_ms / ms_per_step,
        "flops_per_ray": flops / max(1, cst["rays"]),
        "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9},
        # pipe / issue view of the same kernel from the committed ncu captures (what the >= 60 % FP32-pipe target is about):
        # K2 on C3 is bound by issue under divergence, K1 (brute force, C2) by the FMA pipe after the f32x2 packing
        "ncu": ({"source": "profiles/r1_k2_lanes_final4_ncu_summary.txt", "issue_slots_pct": 78.4, "fma_pipe_cycles_pct": 34.9,
                 "alu_pipe_pct": 60.0, "threads_per_warp_inst": 10.49, "warp_inst": 34.4e9} if cst["intersector_used"] == 2 else
                {"source": "profiles/r1_k1_brute_final_ncu_summary.txt", "issue_slots_pct": 64.3, "fma_pipe_cycles_pct": 60.7,
                 "threads_per_warp_inst": 24.0}) if args.workload in ("C3", "C2") and world == 1 else None,
        "simt_efficiency_query_level": cst["active_lane_iters"] / max(1, cst["total_lane_iters"]),
        "counters": {k: cst[k] for k in ("rays", "primary", "slab_tests", "sphere_tests", "sphere_exact", "sphere_hits",
                                         "tri_tests", "tri_stage", "tri_hits", "shades_sphere", "shades_tri",
                                         "emissive", "sky")},
        "launch": {k: cst[k] for k in ("grid_ctas", "cta_threads", "ctas_per_sm", "scene_in_smem", "dyn_smem_bytes")},
    }

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O

        O.build()
        divs = pick_sample(O, sp, tr, cfg, target_s=15.0)
        r, s = oracle_sample(O, sp, tr, cfg, divs)
        cpu_baseline = {"value": r / s / 1e6, "unit": UNIT, "cores": O.hardware_threads(), "kind": "port",
                        "sample": describe_sample(cfg, divs) + f": {r} rays in {s:.2f} s",
                        "note": "C++ oracle (-O2 -ffp-contract=off), std::thread row-parallel like rayon; a port of the "
                                "Rust slave, which cannot be built in this image"}

    if dist is not None:
        barrier()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "frame_ms": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, cfg), l2="flushed between steps: 256 MiB device memset (inside the bracket)",
                           gather=gather_mode, intersector=int(cst["intersector_used"]), device=info["name"]),
            "rays_per_step": rays_total, "mrays_per_s_per_gpu": value / world,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_elapsed / args.steps * 1e3,
                    "h2d_bytes_per_step": int(scene_bytes * world + 72 * world),
                    "d2h_bytes_per_step": int(pixels * 3 + 128 * world),
                    "what": "rt_scene_create (host BVH build + upload) + render + frame download to pinned host memory, per step"},
            "gpu_launches": args.steps * world,
            "clocks": clocks,
            "roofline": roofline,
        }
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    sched.close()
    scene.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
