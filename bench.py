#!/usr/bin/env python
"""bench.py -- TV-L1 frame-pairs/s at 600x800 (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over one synthetic 64-frame 600x800 uint8 clip (BASELINE.json configs[1]):
63 frame pairs, TV-L1 defaults (lambda 0.15, 5 scales, 5 warps, early exit active), all pairs in flight.
  value   : frame-pairs/s, clip already resident in HBM (device-pointer C ABI: teeflow_calc_clip).
  e2e     : the same through the host-buffer entry point (teeflow_calc_clip_host): H2D of the uint8 clip from
            pinned memory and D2H of the fp16 flow (the reference's stored result) inside the timed region.
  roofline: algorithmic bytes of the solver kernel (from the iteration counters the engine
            returns) / its device time, against the measured HBM copy bandwidth.
  cpu_baseline / --impl reference: the CPU port of OpenCV's DualTVL1 (oracle/, the only CPU implementation of
            the path that can run here -- cv2.optflow is not installable) on the host cores.
N > 1 (torchrun): every rank solves its own clip (frame pairs are independent; weak scaling), no data-path
collective; time is the max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

H, W, N_FRAMES = 600, 800, 64
METRIC = "tvl1_frame_pairs_per_s_600x800"
UNIT = "frame-pairs/s"


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(counters: np.ndarray, level_sizes, n_frames: int, f32_out: bool, f16_out: bool):
    """per level s with P_s pixels: 64 B per inner iteration, 16 B per median pass, 32 B per
    warp; up-sampling 8*P_s + 8*P_(s-1); finalize 8*P_0 read + 8*P_0 (f32) / 4*P_0 (f16) write; pyramid per
    frame (1 + 4)*P_0 + sum_(s>=1) 4*(P_(s-1) + P_s).  counters[pair, level, (K, M, W)]."""
    P = np.array([h * w for h, w in level_sizes], dtype=np.float64)
    c = counters.astype(np.float64)
    solver = float((c[:, :, 0] * 64 * P + c[:, :, 1] * 16 * P + c[:, :, 2] * 32 * P).sum())
    n_pairs = counters.shape[0]
    up = float(sum(8 * P[s] + 8 * P[s - 1] for s in range(1, len(P)))) * n_pairs
    fin = float(8 * P[0] + (8 * P[0] if f32_out else 0) + (4 * P[0] if f16_out else 0)) * n_pairs
    pyr = float((1 + 4) * P[0] + sum(4 * (P[s - 1] + P[s]) for s in range(1, len(P)))) * n_frames
    return dict(solver=solver + up + fin, pyramid=pyr, total=solver + up + fin + pyr)


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.idx = gpu_index
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_rate(frames: np.ndarray, n_pairs: int, repeats: int = 1):
    """Times the CPU port of OpenCV's DualTVL1 (faithful serial-float32 error sum) on `n_pairs` pairs."""
    from oracle import tvl1_oracle as O
    O.build()
    model, kind_name = O.create_reference_model()
    cores = O.set_threads(0)          # all host cores (torchrun exports OMP_NUM_THREADS=1)
    best = None
    for _ in range(repeats):
        t = time.perf_counter()
        for i in range(n_pairs):
            model.calc(frames[i], frames[i + 1], None)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    kind = "reference" if kind_name == "cv2.optflow" else "port"
    return n_pairs / best, cores, kind, kind_name


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from tee_optical_flow_b200.synth import make_clip
    sample_pairs = 2
    frames = make_clip(seed=0, n_frames=sample_pairs + 1, H=H, W=W)
    for _ in range(args.warmup):
        cpu_reference_rate(frames, 1)
    t0 = time.perf_counter()
    rates = []
    for _ in range(args.steps):
        r, cores, kind, kind_name = cpu_reference_rate(frames, sample_pairs)
        rates.append(r)
    total = time.perf_counter() - t0
    value = sample_pairs * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "synthetic 64-frame 600x800 uint8 clip, TV-L1 defaults (lambda 0.15, 5 scales, 5 warps)",
                   "sample": f"first {sample_pairs} frame pairs of the clip per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample_pairs} frame pairs x {args.steps} steps of the same 600x800 clip "
                                   f"({kind_name}; OpenMP on all host cores, serial float32 error sum like OpenCV)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def phase_probe(frames_dev, device):
    """Per-phase throughput of the strip ops at full resolution (see the call site).  Bytes per pixel are the
    algorithmic figures of SURVEY.md 8d: inner 64, median 16, warp 32, level-init (coarsest: writes u, p) 24."""
    import numpy as np
    from tee_optical_flow_b200.engine import TVL1Engine
    peak, _ = measured_peak()
    old = os.environ.get("TEEFLOW_GROUPS")
    os.environ["TEEFLOW_GROUPS"] = "1"
    try:
        eng = TVL1Engine(device=device, nscales=1, warps=1, max_slots=frames_dev.shape[0] - 1)
    finally:
        if old is None:
            os.environ.pop("TEEFLOW_GROUPS", None)
        else:
            os.environ["TEEFLOW_GROUPS"] = old
    n, Hh, Ww = frames_dev.shape
    px = (n - 1) * Hh * Ww
    eng.time_launches(12)
    ms = None
    for _ in range(3):
        eng._calc_clip_device(frames_dev, 1.0, True, False, True)
        t = eng.launch_times_ms()
        ms = t if ms is None else np.minimum(ms, t[:len(ms)])
    eng.close()
    inner_ms = float(ms[3])     # launch 3 is always a single iteration (the first of the loop)
    out = {"what": "63 pairs x 600x800 px per launch, one pyramid level, lockstep launches (best of 3 runs); launches "
                   "4.. are two-iteration passes when the temporal blocking is on",
           "px_per_launch": px, "launch_ms": [round(float(x), 4) for x in ms]}
    for name, t, b in (("level_init", float(ms[0]), 24), ("warp", float(ms[1]), 32), ("median", float(ms[2]), 16),
                       ("inner", inner_ms, 64)):
        gbs = px * b / (t * 1e-3) / 1e9
        out[name] = {"ms": t, "bytes_per_px": b, "algorithmic_GBps": gbs, "frac_of_peak": gbs / peak}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="teeflow", choices=["teeflow", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--slots", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "teeflow" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from tee_optical_flow_b200.engine import TVL1Engine
    from tee_optical_flow_b200.synth import make_clip

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: teeflow has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # every rank solves its own copy of the SAME synthetic clip (seed 0): TV-L1 cost is data dependent (early exit:
    # 70-91 ms per clip over seeds 0..7, tools/seed_times.py), so identical clips keep the per-GPU work fixed as N
    # grows -- a clean weak-scaling measurement.  Frame pairs are independent units, no data-path collective.
    frames_np = make_clip(seed=0, n_frames=N_FRAMES, H=H, W=W)
    frames_dev = torch.from_numpy(frames_np).to(dev)
    n_pairs = N_FRAMES - 1
    eng = TVL1Engine(device=local_rank, max_slots=args.slots)
    out16 = torch.empty((N_FRAMES, H, W, 2), dtype=torch.float16, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        flush.fill_(1)                                              # L2 flush between steps
        eng._calc_clip_device(frames_dev, 1.0, True, False, True, out_f16=out16)

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    solver_ms = pyr_ms = 0.0
    launches = kernel_launches = 0
    t_wall = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step_device()
        _, info = eng.last_counters()
        solver_ms += info["solver_ms"]; pyr_ms += info["pyramid_ms"]
        launches += info["solver_launches"]; kernel_launches += info["kernel_launches"]
    ev1.record()
    barrier()
    wall_s = time.perf_counter() - t_wall
    dev_s = ev0.elapsed_time(ev1) / 1e3
    counters, info = eng.last_counters()
    levels = eng.level_sizes(H, W)

    # ---- end to end: host (pinned) buffers in, fp16 flow out, through teeflow_calc_clip_host
    pin_in = torch.from_numpy(frames_np).pin_memory()
    pin_out = torch.empty((N_FRAMES, H, W, 2), dtype=torch.float16).pin_memory()
    in_np, out_np = pin_in.numpy(), pin_out.numpy()
    lib, hnd = eng._lib, eng._h

    def step_host():
        flush.fill_(1)
        rc = lib.teeflow_calc_clip_host(hnd, in_np.ctypes.data, 0, N_FRAMES, H, W, None, out_np.ctypes.data, 1.0, 1)
        if rc != 0:
            raise RuntimeError(lib.teeflow_last_error(hnd))

    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None

    # ---- BASELINE config 3 (informational, outside the headline metric): masked radial / longitudinal
    # decomposition + exact per-frame percentiles / angle mode of the stored fp16 flow, then (N > 1) the NCCL
    # all-gather of the per-frame waveform rows -- the only exchange of the path
    analysis = None
    try:
        from tee_optical_flow_b200.sharding import WAVEFORM_COLUMNS, gather_rows, pair_range
        from tee_optical_flow_b200.synth import make_masks
        nfr = N_FRAMES - 2
        rv = torch.from_numpy(make_masks(0, N_FRAMES, H, W)["rv"]).to(dev)
        cent = np.tile(np.array([[0.77 * H, 0.5 * W]]), (nfr, 1))
        eng.analyze_clip(out16, rv, cent, nfr)
        barrier()
        a0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            res = eng.analyze_clip(out16, rv, cent, nfr)
            rows = np.stack([res[k] for k in WAVEFORM_COLUMNS[:6]], axis=1).astype(np.float64)
            if world > 1:   # every rank contributes its clip's rows: gather of world * nfr rows
                lo, hi = pair_range(world * nfr, rank, world)
                gather_rows(rows[:hi - lo], world * nfr, rank, world, device=dev)
        barrier()
        analysis = {"ms_per_clip": 1e3 * (time.perf_counter() - a0) / reps, "frames": nfr,
                    "what": "teeflow_analyze_clip (mag p99, angle mode, radial/longitudinal p1/p99) + waveform gather"}
        # saliency input stage of config 3 ("saliency on"): StaticSaliencyFineGrained per frame on the GPU
        rgb = frames_dev.unsqueeze(-1).expand(-1, -1, -1, 3).contiguous()
        eng.compute_saliency(rgb)
        barrier()
        s0 = time.perf_counter()
        for _ in range(reps):
            eng.compute_saliency(rgb)
        barrier()
        analysis["saliency_ms_per_clip"] = 1e3 * (time.perf_counter() - s0) / reps
        analysis["saliency_what"] = "teeflow_saliency_fine_grained, 64 RGB frames 600x800 -> float32 maps"
    except Exception as e:  # informational only
        analysis = {"error": str(e)[:200]}

    # ---- per-phase roofline (rank 0, informational): the step kernel runs four different strip ops; with one slot
    # group and ONE pyramid level all 63 pairs step in lockstep at first, so launch 0 is a pure level-init, 1 a pure
    # warp, 2 a pure median and 3.. pure inner iterations over 63 x 600 x 800 px -- timed with CUDA events per launch
    phases = None
    if rank == 0:
        try:
            phases = phase_probe(frames_dev, local_rank)
        except Exception as e:  # informational only
            phases = {"error": str(e)[:200]}

    t_max = torch.tensor([dev_s, wall_s, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    dev_s, wall_s, e2e_s = [float(x) for x in t_max.tolist()]
    step_s = max(dev_s, 1e-9)

    if rank == 0:
        peak, peak_src = measured_peak()
        ab = algorithmic_bytes(counters, levels, N_FRAMES, f32_out=False, f16_out=True)
        per_launch_bytes = ab["solver"] / max(info["solver_launches"], 1)
        avg_launch_ms = solver_ms / max(launches, 1)
        achieved = per_launch_bytes / (avg_launch_ms * 1e-3) / 1e9 if avg_launch_ms > 0 else 0.0
        traffic = None
        tp = ROOT / "profiles" / "solver_traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        value = world * n_pairs * args.steps / step_s
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * step_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": "synthetic 64-frame 600x800 uint8 clip per GPU, TV-L1 defaults (lambda 0.15, 5 scales, "
                            "5 warps, inner 30 x outer 10 with early exit, median 5), 63 frame pairs in flight, "
                            "fp16 flow (N,H,W,2) out",
                "pairs_per_step_per_gpu": n_pairs, "slots": info["n_slots"],
                "l2": "256 MiB buffer written between steps (L2 flush, inside the timed region); solver state "
                      "3.4 GB >> 126 MB L2",
                "inner_iterations_per_pair_mean": float(counters[:, :, 0].sum() / n_pairs),
                "algorithmic_GB_per_pair": ab["total"] / n_pairs / 1e9,
            },
            "wall_ms_per_step": 1e3 * wall_s / args.steps,
            "kernel_time_fraction": (solver_ms + pyr_ms) / 1e3 / max(wall_s, 1e-9),
            "gpu_launches": int(kernel_launches),
            "e2e": {"value": world * n_pairs * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(frames_np.nbytes), "d2h_bytes_per_step": int(out_np.nbytes),
                    "api": "teeflow_calc_clip_host (pinned host buffers)"},
            "roofline": {"bound": "hbm", "kernel": "tvl1_step_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": per_launch_bytes, "avg_launch_ms": avg_launch_ms,
                         "launches_per_step": launches / args.steps,
                         "solver_share_of_step": solver_ms / 1e3 / step_s},
            "clocks": clocks,
            "analysis_config3": analysis,
            "phase_roofline": phases,
        }
        if world == 1 and not args.no_cpu_baseline:
            sample_pairs = 6
            rate, cores, kind, kind_name = cpu_reference_rate(frames_np, sample_pairs)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"first {sample_pairs} frame pairs of the same 600x800 clip "
                                              f"({kind_name}, OpenMP on all host cores)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
