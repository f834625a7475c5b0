#!/bin/bash
# Multi-GPU bench lines of BASELINE.json's configurations: usage scripts/run_multi_gpu.sh N  (run under gpurun --gpus N)
N=$1
run() {  # name, extra bench args
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus $N --steps 3 --warmup 3 --stage-times --no-e2e "$@" > gpurun_out/r02_${name}_n$N.json 2> gpurun_out/r02_${name}_n$N.err
  echo "== $name N=$N rc=$?"; grep "stage ms" gpurun_out/r02_${name}_n$N.err | head -1
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_${name}_n$N.json"))
    print(d["config"]["workload"], "| ms/step %.1f | %.0f frames/s | K7 frac %.2f | parity %s" % (
        d["ms_per_step"], d["value"], d["roofline"]["frac"], d.get("multi_gpu_parity", {}).get("ok")))
    print({k: v for k, v in d["stage_ms"].items()})
except Exception as e:
    print("no json:", e)
PY
}
if [ "$N" = "8" ]; then run c2weak; run c4 --workload c4; fi
run c3strong --workload c3 --strong
