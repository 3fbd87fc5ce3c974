#!/bin/bash
# Multi-GPU session (gpurun --gpus N): DP consistency test, then bench lines under torchrun.
#   bash tools/gpu_multi.sh <N> "<bench args 1>" "<bench args 2>" ...
set -u
N=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.active --format=csv -lms 1000 > gpurun_out/clocks_multi.csv &
SMI=$!
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q 2>&1 | tail -3
i=0
for args in "$@"; do
  i=$((i+1))
  out=gpurun_out/bench_n${N}_$i.json
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+i)) bench.py --gpus $N $args > $out 2> gpurun_out/bench_n${N}_$i.err
  echo "[$i] rc=$? args: $args"
  python - <<PY
import json
try:
    d=json.loads(open("$out").read().strip().splitlines()[-1])
    print("   value", round(d["value"],1), "img/s  ms/step", round(d["ms_per_step"],3), "e2e", d["e2e"] and round(d["e2e"]["value"],1), "scaling", d["scaling"], "frac", round(d["config"]["step_frac_of_peak"],3), d["config"]["workload"][:90], d["clocks"])
except Exception as e:
    print("   no line:", e); print(open("gpurun_out/bench_n${N}_$i.err").read()[-1500:])
PY
done
kill $SMI
