#!/usr/bin/env python
"""bench.py -- whole-job PMD compression throughput on N B200s (one process per GPU).

  python bench.py --gpus N --steps K --warmup W              our arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...    the reference's CPU path (NumPy/SciPy
                                                             restatement, the JAX original cannot run here)

A "step" is one full localmd_decomposition of the synthetic movie named by BASELINE.json configs[1]:
512x512x20000 float32, 20x20 blocks, frames_to_init 5000, rank pruning on (BASELINE.md's projection
table is quoted at k = 1650 = 0.33 * 5000).  `value` is frames/s with the movie already resident in
HBM; `e2e` is the same call fed a HOST (pinned) numpy movie, host->device copies inside the timed region.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c2": dict(T=20000, d1=512, d2=512, block=20, frames_to_init=5000, n_cells=400, blob=(3.0, 5.0), bg_rank=2),
    "c3": dict(T=20000, d1=512, d2=512, block=32, frames_to_init=5000, n_cells=400, blob=(3.0, 5.0), bg_rank=2),
    # BASELINE.json configs[3]: 1024x1024x30000 widefield movie over 8 GPUs = 3750 frames (15.7 GB) per GPU
    "c4": dict(T=3750, d1=1024, d2=1024, block=40, frames_to_init=5000, n_cells=150, blob=(12.0, 25.0), bg_rank=8),
    "small": dict(T=4096, d1=128, d2=128, block=20, frames_to_init=1000, n_cells=40, blob=(3.0, 5.0), bg_rank=2),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strong", action="store_true", help="N > 1: ONE movie of the workload's size split over the GPUs (default: weak)")
    ap.add_argument("--stage-times", action="store_true", help="print per-stage GPU ms to stderr")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle-reason samples DURING the timed region.  In-process NVML (two light queries every 200 ms from a
    daemon thread) when nvidia_ml_py is importable; otherwise one `nvidia-smi -lms` child.  The child's full query every
    100 ms showed up as multi-millisecond outliers in single latency-bound stages of single steps (the final eigensolver's
    ~300 dependent launches), so it is only the fallback."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.proc, self.nvml, self._stop = index, [], None, None, threading.Event()
        self._go = threading.Event()     # set by begin(): sampling starts with the timed region, initialisation happens before it
        self.sm, self.reasons = [], set()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "250"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        names = [("hw_slowdown", n.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", n.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", n.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", n.nvmlClocksEventReasonSwPowerCap)]
        self._go.wait()
        while not self._stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                for name, bit in names:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.2)

    def begin(self):
        self._go.set()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self._go.set()
            self.thread.join(timeout=2)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# The bounded sample of the workload that the CPU legs run (the oracle needs ~9 min for one full C2 pass on 16 cores):
# same block size, same max_components / background rank / rank pruning, same init-window fraction t / T = 1/4 as C2,
# on a 128x128 crop x 4096 frames.  Throughput is reported in FULL-FIELD-OF-VIEW frame equivalents:
# sample pixel-frames / (d1 d2 of the workload) / seconds -- measured, not extrapolated per stage.
SAMPLE = dict(T=4096, d1=128, d2=128, t=1024, n_cells=25, sims=250)


def sample_inputs(w, seed=0):
    """(movie, draws, kwargs) of the bounded sample: every random quantity is host supplied so that the oracle and the
    CUDA path see the same inputs (BASELINE.json north_star: 'same host-supplied random sketch matrices')."""
    import oracle.pmd_oracle as O

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from synth import make_movie

    T, d1, d2, t, blk, r, K = SAMPLE["T"], SAMPLE["d1"], SAMPLE["d2"], SAMPLE["t"], w["block"], 50, 15
    movie = make_movie(T, d1, d2, n_cells=SAMPLE["n_cells"], seed=5, blob_sigma=w["blob"])
    rng = np.random.default_rng(seed)
    nb = len(O.tile_starts(d1, blk)) * len(O.tile_starts(d2, blk))
    n_bg = min(1000, T)
    sims = SAMPLE["sims"]
    prune_seed = int(rng.integers(0, 2**31))
    draws = O.Draws(
        bg_frames=rng.choice(T, n_bg, replace=False).tolist(),
        bg_sketch=rng.standard_normal((n_bg, K + 10), dtype=np.float32),
        init_frames=list(range(t, 2 * t)),
        sim_noise=[rng.standard_normal((blk, blk, t), dtype=np.float32) for _ in range(sims)],
        sim_sketch=[rng.standard_normal((t, 11), dtype=np.float32) for _ in range(sims)],
        block_sketches=[[rng.standard_normal((t // 10, r + 10), dtype=np.float32)] for _ in range(nb)],
        prune_sketch=lambda shape: np.random.default_rng(prune_seed).standard_normal(shape, dtype=np.float32),
    )
    kw = dict(block_sizes=[blk, blk], frame_range=t, rank_prune=True, max_components=r, background_rank=K)
    return movie, draws, kw


def sample_description(w, seconds):
    return ("restated reference (NumPy/SciPy float32 oracle, not JAX) on a bounded sample of the workload: %dx%d crop x %d frames, "
            "%dx%d blocks, %d init frames, max_components 50, background rank 15, rank_prune 0.33, %d threshold simulations; "
            "%.1f s measured; value = sample pixel-frames / (%d x %d) / seconds (full-FOV frame equivalents, measured, not "
            "extrapolated)" % (SAMPLE["d1"], SAMPLE["d2"], SAMPLE["T"], w["block"], w["block"], SAMPLE["t"], SAMPLE["sims"],
                               seconds, w["d1"], w["d2"]))


def cpu_reference_pass(w, inputs=None, keep_result=False):
    """One pass of the reference's algorithm (oracle/pmd_oracle.py) over the bounded sample on all host cores.
    Returns dict(seconds, value, stages, result)."""
    import oracle.pmd_oracle as O

    movie, draws, kw = inputs if inputs is not None else sample_inputs(w)
    tm = {}
    t0 = time.perf_counter()
    res = O.localmd_decomposition_oracle(movie, kw["block_sizes"], kw["frame_range"], draws, rank_prune=True,
                                         max_components=kw["max_components"], background_rank=kw["background_rank"], timings=tm)
    dt = time.perf_counter() - t0
    equiv_frames = SAMPLE["T"] * (SAMPLE["d1"] * SAMPLE["d2"]) / float(w["d1"] * w["d2"])
    return dict(seconds=dt, value=equiv_frames / dt, stages=tm, result=res if keep_result else None)


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; JAX is not installable here) on the box's host
    cores.  Every step is one full pass over the bounded sample; warm-up and step counts are honoured as given."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    inputs = sample_inputs(w)
    secs = []
    for i in range(args.warmup + args.steps):
        p = cpu_reference_pass(w, inputs)
        if i >= args.warmup:
            secs.append(p["seconds"])
    sec = float(np.mean(secs))
    equiv_frames = SAMPLE["T"] * (SAMPLE["d1"] * SAMPLE["d2"]) / float(w["d1"] * w["d2"])
    v = equiv_frames / sec
    cfg = workload_config(args.workload, args.gpus)
    cfg["workload"] = "bounded sample of: " + cfg["workload"]
    cfg["sample"] = dict(SAMPLE, block=w["block"], frames_equivalent_per_step=equiv_frames)
    cfg["timing"] = "host wall clock (perf_counter) around every oracle pass, all host cores"
    cfg["parallelism"] = "rank 0 only, %d host threads" % cores
    line = {
        "impl": "reference", "metric": "frames/sec compressed", "value": v, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "extrapolated": False,
        "config": cfg,
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample_description(w, sec)},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def parity_on_sample(w, dev, inputs, ref):
    """The CUDA path on the bounded sample with the SAME host-supplied draws as the oracle pass `ref`; the oracle is also
    evaluated in float64 (the reference algorithm in exact-ish arithmetic) to separate float32 rounding of the CPU
    restatement from differences of the CUDA path.  Returns the `parity` object of the JSON line."""
    import localmd_b200
    import oracle.pmd_oracle as O
    from oracle import parity as P

    movie, draws, kw = inputs
    det = {}
    arr = localmd_b200.localmd_decomposition(movie, draws=draws, details=det, device=dev, **kw)
    rep32 = P.parity_report(arr, det, ref)
    with O.precision(np.float64):
        ref64 = O.localmd_decomposition_oracle(movie, kw["block_sizes"], kw["frame_range"], draws, rank_prune=True,
                                               max_components=kw["max_components"], background_rank=kw["background_rank"])
    rep64 = P.parity_report(arr, det, ref64)
    fake = type("A", (), dict(u=ref.u, r=ref.r, s=ref.s, v=ref.vt))()
    det_o = dict(ranks=ref.ranks, sstat=det["sstat"], tstat=det["tstat"], thresholds=ref.thresholds)
    floor = P.parity_report(fake, det_o, ref64) if np.array_equal(ref.ranks, ref64.ranks) else None
    return {
        "sample": "%dx%dx%d crop, %dx%d blocks, t=%d, max_components %d, rank_prune 0.33, same host-supplied draws"
                  % (SAMPLE["d1"], SAMPLE["d2"], SAMPLE["T"], w["block"], w["block"], SAMPLE["t"], kw["max_components"]),
        "tolerances": {"ranks/CSR": "bit-exact outside eps_stat of a threshold", "s_rel": 1e-4, "angle_rad": 1e-3, "yhat_rel_fro": 1e-4},
        "vs_oracle_f64": rep64, "within_north_star_f64": P.within_north_star(rep64),
        "vs_oracle_f32": rep32,
        "oracle_f32_vs_f64": None if floor is None else {k: floor[k] for k in ("s_max_rel_err_lead", "angle_UR_max_rad", "angle_Vt_max_rad",
                                                                               "yhat_rel_fro_err")},
        "mean_rank": float(np.mean(det["ranks"])), "max_rank": int(np.max(det["ranks"])),
    }


def multi_gpu_parity(dev, group, rank, world):
    """N > 1: the frame-sharded decomposition of one small movie equals its single-GPU decomposition (same draws)."""
    import localmd_b200

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from synth import make_movie

    T, d1, d2 = 1024 * max(world, 4), 64, 72
    movie = make_movie(T, d1, d2, n_cells=8, seed=11)
    kw = dict(block_sizes=[16, 16], frame_range=1500, max_components=10, background_rank=3, seed=5, rank_prune=True)
    det = {}
    arr = localmd_b200.localmd_decomposition(movie, device=dev, group=group, details=det, **kw)
    out = None
    if rank == 0:
        det1 = {}
        one = localmd_b200.localmd_decomposition(movie, device=dev, details=det1, **kw)
        lead = one.s > 0.05 * one.s[0]
        frames = [0, 1023, 1024, T // 2, T - 1]
        rec, rec1 = arr[frames], one[frames]
        out = {"movie": "%dx%dx%d, 16x16 blocks, sharded x%d vs single GPU" % (d1, d2, T, world),
               "ranks_equal": bool(np.array_equal(det["ranks"], det1["ranks"])), "k": [int(len(arr.s)), int(len(one.s))],
               "s_max_rel_err_lead": float(np.max(np.abs(arr.s[: len(one.s)][lead] / one.s[lead] - 1))) if len(arr.s) >= len(one.s) else None,
               "recon_rel_err": float(np.linalg.norm(rec - rec1) / np.linalg.norm(rec1 - one.mean_img[None])),
               "mean_max_rel_err": float(np.max(np.abs(arr.mean_img / one.mean_img - 1)))}
        out["ok"] = bool(out["ranks_equal"] and out["s_max_rel_err_lead"] is not None and out["s_max_rel_err_lead"] <= 1e-4
                         and out["recon_rel_err"] <= 1e-4)
    import torch.distributed as dist

    dist.barrier()
    return out


def workload_config(name, n_gpus, strong=False):
    w = WORKLOADS[name]
    if strong and n_gpus > 1:
        w = dict(w, T=w["T"] // n_gpus)
    which = {"c2": "BASELINE.json configs[1]", "c3": "BASELINE.json configs[2]", "c4": "BASELINE.json configs[3], per-GPU shard"}
    gb = 4.0 * w["d1"] * w["d2"] * w["T"] / 1e9
    return {"workload": "synthetic %dx%dx%d float32 %s movie, block %dx%d, frames_to_init %d, rank_prune 0.33 (%s)"
                        % (w["d1"], w["d2"], w["T"], "widefield" if name == "c4" else "two-photon", w["block"], w["block"],
                           w["frames_to_init"], which.get(name, "reduced test size")),
            "frames_per_gpu": w["T"], "total_frames": w["T"] * n_gpus,
            "timing": "inputs (%.0f GB/GPU) larger than L2; CUDA events, max over ranks" % gb,
            "parallelism": "one movie of %d frames, frame-sharded x%d (blocks partitioned, NCCL reductions/gathers)"
                           % (w["T"] * n_gpus, n_gpus) if n_gpus > 1 else "single GPU"}


def run_ours(args):
    # stdout must carry exactly ONE JSON line, but libraries print there too (NCCL writes its version banner to file
    # descriptor 1 at communicator creation): descriptor 1 is pointed at stderr for the whole run and the JSON line is
    # written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch

    import localmd_b200
    from localmd_b200 import ops, sharding
    from localmd_b200.dataset import DeviceMovie
    from localmd_b200.synthetic import make_movie

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    w = WORKLOADS[args.workload]
    T, d1, d2, blk, t_init = w["T"], w["d1"], w["d2"], w["block"], w["frames_to_init"]
    # weak scaling: ONE movie of world * T frames, frame-sharded over the ranks (1024-aligned contiguous ranges);
    # the stats and projection passes run on the local shard, the block stage is partitioned by blocks, the rank-sized
    # reductions / gathers go over NCCL (DESIGN.md section "multi-GPU")
    t_total = T if args.strong else world * T
    lo, hi = sharding.shard_bounds(t_total, world)[rank]
    shard = make_movie(t_total, d1, d2, n_cells=w["n_cells"], blob_sigma=w["blob"], bg_rank=w["bg_rank"], seed=1234, device=dev,
                       frame_lo=lo, frame_hi=hi)
    movie = DeviceMovie.from_shard(shard, t_total, lo)
    kw = dict(block_sizes=[blk, blk], frame_range=t_init, rank_prune=True, seed=0, group=group)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    stage = {"__detail__": True}
    for i in range(args.warmup):
        localmd_b200.localmd_decomposition(movie, timings=stage if i == args.warmup - 1 else None, **kw)
    if args.stage_times and rank == 0:
        sys.stderr.write("stage ms: %s\n" % json.dumps({k: round(v, 2) for k, v in stage.items() if isinstance(v, float)}))
        sys.stderr.write("n_cols/ranks info: %s\n" % json.dumps(stage.get("__info__", {})))
    # the interpreter's cyclic collector off the timed region's critical path: everything allocated so far is frozen (never
    # rescanned), so the collections that still trigger on the jobs' short-lived tensors stay short
    import gc

    gc.collect()
    gc.freeze()
    # the sampler starts BEFORE the barrier: NVML initialisation takes tens of milliseconds on rank 0, and a rank that enters
    # the timed region late makes every other rank wait for it at the first collective (max over ranks = +10 ms per step at N = 2)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    n0 = ops.LAUNCHES["count"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per_step = []
    ev0.record()
    sampler.begin()
    for _ in range(args.steps):
        tms = {"__detail__": True}
        localmd_b200.localmd_decomposition(movie, timings=tms, **kw)
        per_step.append(tms)
        if args.stage_times and rank == 0:
            sys.stderr.write("step stages: %s\n" % json.dumps({k: round(v, 2) for k, v in tms.items() if isinstance(v, float) and "." not in k}))
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    pmdarray = None
    if rank == 0 and world == 1 and not args.no_e2e:
        # BASELINE.json configs[4]: PMDArray reconstruction (CSR U . R diag(s) Vt, un-normalised) on this decomposition:
        # full frames and a cropped slice, host ndarray out (the device->host copy of the frames is inside the timing)
        arr_r = localmd_b200.localmd_decomposition(movie, **kw)
        n_fr = 256
        pmdarray = {}
        for name_r, key_r in (("full_frames", (slice(1000, 1000 + n_fr),)),
                              ("crop_128x128", (slice(1000, 1000 + n_fr), slice(100, 228), slice(300, 428)))):
            arr_r[key_r]  # warm-up (device state, library handles)
            torch.cuda.synchronize(dev)
            t_r = time.perf_counter()
            out_r = arr_r[key_r]
            dt_r = time.perf_counter() - t_r
            pmdarray[name_r] = {"frames_per_s": n_fr / dt_r, "ms": dt_r * 1e3, "out_MB": out_r.nbytes / 1e6}
        del arr_r, out_r
    launches = ops.LAUNCHES["count"] - n0
    if world > 1:
        tms_t = torch.tensor([ms], device=dev)
        dist.all_reduce(tms_t, op=dist.ReduceOp.MAX)
        ms = float(tms_t.item())
    ms_step = ms / args.steps
    value = t_total / (ms_step / 1e3)

    # roofline of the dominant streaming kernel (K7 = pmd_project_stream): algorithmic bytes = every movie element
    # of the local shard read once + the Z rows written once, over the kernel's own CUDA-event time
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    info = per_step[-1].get("__info__", {})
    n_cols = int(info.get("n_cols", 0))
    k7_ms = float(np.mean([p["projection.stream"] for p in per_step]))
    pass_ms = float(np.mean([p["projection"] for p in per_step]))
    prep_ms = float(np.mean([p.get("projection.prep", 0.0) for p in per_step]))
    n_loc = hi - lo
    alg_bytes = 4.0 * d1 * d2 * n_loc + 4.0 * n_cols * n_loc
    achieved = alg_bytes / (k7_ms / 1e3) / 1e9
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "k7_traffic.json")))
        if tr.get("workload") == args.workload:
            traffic = tr.get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "pmd_project_stream_ts / project_ts_kernel (K7: U^T standardised movie, local + background columns; "
                                          "2-D TMA raw tiles, movie operand in tensor memory, TF32 + bf16-pair tcgen05 MMAs)",
                "achieved": achieved, "peak": peak, "peak_source": "measured" if peaks else "fallback", "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "ms": k7_ms, "algorithmic_bytes": alg_bytes,
                "projection_pass_ms": pass_ms, "n_cols": n_cols,
                "timed": "CUDA events immediately before / after the launch of the kernel on its stream (the job's own launch)",
                "prep_ms": prep_ms,   # before the launch, same stage: mixing-operand split, table uploads, coefficient images, zero fills
                }

    # the other full-movie pass (K1 = pmd_stats_pass_tc, mean + Welch noise estimate): timed alone on the resident shard
    roofline_k1 = None
    try:
        chunk = shard.view(shard.shape[0], -1)
        for _ in range(2):
            ops.stats_pass(chunk, t_total)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(5):
            ops.stats_pass(chunk, t_total)
        k1.record()
        torch.cuda.synchronize(dev)
        k1_ms = k0.elapsed_time(k1) / 5
        k1_bytes = 4.0 * d1 * d2 * n_loc
        roofline_k1 = {"bound": "hbm", "kernel": "pmd_stats_pass_tc / stats_tc_kernel (K1: per-pixel mean + Welch noise estimate; 2-D TMA, "
                                              "DFT as TF32 + bf16-pair tcgen05 MMAs)", "achieved": k1_bytes / (k1_ms / 1e3) / 1e9,
                       "peak": peak, "unit": "GB/s", "frac": k1_bytes / (k1_ms / 1e3) / 1e9 / peak, "ms": k1_ms,
                       "algorithmic_bytes": k1_bytes, "traffic": None}
        try:
            trk = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
            if trk.get("workload") == args.workload:
                roofline_k1["traffic"] = trk.get("dram_bytes_per_launch")
        except Exception:
            pass
    except Exception as exc:
        roofline_k1 = {"error": repr(exc)}

    e2e = None
    if not args.no_e2e:
        host_t = torch.empty(shard.shape, dtype=shard.dtype, pin_memory=True)
        host_t.copy_(shard)
        del movie, shard  # the freed HBM stays in torch's pool (a warm process would not re-cudaMalloc 21 GB per movie)
        e2e_runs = []
        # untimed warm-up passes (page-locked staging buffers, library handles), then timed passes: every pass copies
        # the whole movie host -> device and reads every factor of the result + one reconstructed frame back
        n_warm_e2e = 2  # the second warm pass re-uses the page-locked result buffers the first one released
        for it_e2e in range(n_warm_e2e + max(1, min(2, args.steps))):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            det = {}
            # a fresh movie object per pass: nothing of the previous pass's upload is reused
            src = host_t.numpy() if world == 1 else DeviceMovie.from_host_shard(host_t, t_total, lo, dev)
            arr = localmd_b200.localmd_decomposition(src, timings=det, **kw)
            t_dec = time.perf_counter() - t0
            if args.stage_times and rank == 0:
                sys.stderr.write("e2e decomposition wall %.1f ms\n" % (t_dec * 1e3))
            if rank == 0:
                tq = time.perf_counter()
                result = (arr.u, arr.r, arr.s, arr.v, arr.mean_img, arr.var_img)  # device -> host read of the compressed movie
                tr = time.perf_counter()
                frame = arr[t_total // 2, :, :]  # ... and of one reconstructed frame
                if args.stage_times:
                    sys.stderr.write("e2e result d2h %.1f ms, one frame %.1f ms\n" % ((tr - tq) * 1e3, (time.perf_counter() - tr) * 1e3))
            e1.record()
            barrier()
            wall = time.perf_counter() - t0
            if args.stage_times and rank == 0:
                sys.stderr.write("e2e stage ms: %s  wall %.1f ms\n" % (
                    json.dumps({k: round(v, 2) for k, v in det.items() if isinstance(v, float) and "." not in k}), wall * 1e3))
            if it_e2e >= n_warm_e2e:
                e2e_runs.append(max(e0.elapsed_time(e1), wall * 1e3))
            if rank == 0:
                d2h = int(arr.u.data.nbytes + arr.u.indices.nbytes + arr.u.indptr.nbytes + arr.r.nbytes + arr.s.nbytes
                          + arr.v.nbytes + 2 * 4 * d1 * d2 + frame.nbytes)
                del result, frame
            del arr, src  # page-locked result buffers go back to torch's host cache for the next pass
        e2e_ms = float(np.mean(e2e_runs))
        if world > 1:
            tt = torch.tensor([e2e_ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_ms = float(tt.item())
        if rank == 0:
            e2e = {"value": t_total / (e2e_ms / 1e3), "unit": "frames/s",
                   "h2d_bytes_per_step": int(det.get("__info__", {}).get("h2d_bytes", host_t.numel() * host_t.element_size())) * world,
                   "d2h_bytes_per_step": d2h, "ms": e2e_ms, "passes": len(e2e_runs)}

    mgpu = multi_gpu_parity(dev, group, rank, world) if world > 1 else None
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": "frames/sec compressed", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if args.strong and world > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args.workload, world, args.strong), "clocks": clocks,
        "gpu_launches": launches, "roofline": roofline, "roofline_k1": roofline_k1, "e2e": e2e, "pmdarray_reconstruction": pmdarray,
        "stage_ms": {k: round(float(np.mean([p[k] for p in per_step])), 3) for k in per_step[0]
                     if isinstance(per_step[0][k], float) and "." not in k},
    }
    if mgpu is not None:
        line["multi_gpu_parity"] = mgpu
    if not args.no_cpu_baseline and world == 1:
        # CPU leg (rank 0, N = 1 only): ONE measured oracle pass over the bounded sample, then the CUDA path on the same
        # sample and draws -> parity block
        inputs = sample_inputs(w)
        p = cpu_reference_pass(w, inputs, keep_result=True)
        line["cpu_baseline"] = {"value": p["value"], "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": sample_description(w, p["seconds"]), "seconds": p["seconds"],
                                "stage_s": {k: round(v, 3) for k, v in p["stages"].items()}}
        try:
            line["parity"] = parity_on_sample(w, dev, inputs, p["result"])
        except Exception as exc:  # the bench line must still be printed; a failed parity run is reported as such
            line["parity"] = {"error": repr(exc)}
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
