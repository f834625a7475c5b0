"""One C2 decomposition under torch.profiler (CUPTI): per-kernel timeline in launch order, GPU busy / idle time and the
largest idle gaps with their neighbouring kernels (host-bound stretches).  Usage: python scripts/profile_job.py [workload] [out]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import localmd_b200  # noqa: E402
from localmd_b200.dataset import DeviceMovie  # noqa: E402
from localmd_b200.synthetic import make_movie  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "job_timeline.txt")
w = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
shard = make_movie(w["T"], w["d1"], w["d2"], n_cells=w["n_cells"], blob_sigma=w["blob"], bg_rank=w["bg_rank"], seed=1234, device=dev,
                   frame_lo=0, frame_hi=w["T"])
movie = DeviceMovie.from_shard(shard, w["T"], 0)
kw = dict(block_sizes=[w["block"], w["block"]], frame_range=w["frames_to_init"], rank_prune=True, seed=0)
TIMED = os.environ.get("PROFILE_WITH_TIMINGS")   # the bench passes a timings dict (CUDA-event marks + a final synchronise)
for _ in range(3):
    localmd_b200.localmd_decomposition(movie, timings={"__detail__": True} if TIMED else None, **kw)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    localmd_b200.localmd_decomposition(movie, timings={"__detail__": True} if TIMED else None, **kw)
    torch.cuda.synchronize()
trace = os.path.join(os.path.dirname(out), "job_trace.json")
prof.export_chrome_trace(trace)
ev = json.load(open(trace))["traceEvents"]
os.remove(trace)
ker = sorted(((e["ts"], e["dur"], e["name"], e.get("args", {}).get("stream", 0)) for e in ev
              if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e), key=lambda x: x[0])
syncs = sorted((e["ts"], e["dur"], e["name"]) for e in ev if e.get("cat") == "cuda_runtime" and "ynchronize" in e["name"])
t0 = ker[0][0]
end = max(k[0] + k[1] for k in ker)
# union of busy intervals over all streams
busy, cur_s, cur_e, gaps = 0.0, None, None, []
prev_name = None
for ts, dur, nm, st in ker:
    if cur_e is None:
        cur_s, cur_e = ts, ts + dur
    elif ts > cur_e:
        busy += cur_e - cur_s
        gaps.append((ts - cur_e, cur_e - t0, prev_name, nm))
        cur_s, cur_e = ts, ts + dur
    else:
        cur_e = max(cur_e, ts + dur)
    prev_name = nm
busy += cur_e - cur_s
with open(out, "w") as f:
    f.write("# torch.profiler timeline of one %s decomposition (not a bench number: profiler attached)\n" % name)
    f.write("span %.2f ms, GPU busy (any stream) %.2f ms, idle %.2f ms, %d kernels/copies, %d host synchronisations\n"
            % ((end - t0) / 1e3, busy / 1e3, (end - t0 - busy) / 1e3, len(ker), len(syncs)))
    f.write("\nlargest idle gaps (us, at ms, after -> before):\n")
    for g, at, a, b in sorted(gaps, reverse=True)[:40]:
        f.write("%8.1f  @%7.2f  %s -> %s\n" % (g, at / 1e3, a[:70], b[:70]))
    f.write("\nhost synchronisations (at ms, us):\n")
    for ts, dur, nm in syncs:
        f.write("  @%7.2f %8.1f %s\n" % ((ts - t0) / 1e3, dur, nm))
    f.write("\ntimeline (start ms, us, stream, name):\n")
    for ts, dur, nm, st in ker:
        f.write("%8.3f %9.1f %3s  %s\n" % ((ts - t0) / 1e3, dur, st, nm[:130]))
print(open(out).read()[:6000])
