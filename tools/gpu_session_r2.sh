#!/bin/bash
# One gpurun call of round 2; parts selected by words:  tests parity bench configs launches ncu
#   gpurun --timeout 1500 -- 'bash tools/gpu_session_r2.sh tests parity bench'
set -u
mkdir -p gpurun_out
what="${*:-tests parity bench}"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
if [[ $what == *tests* ]]; then
  timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
fi
if [[ $what == *parity* ]]; then
  timeout 900 python tools/gpu_parity_report.py gpurun_out/parity_report.json > gpurun_out/parity_report.log 2>&1
  echo "parity exit $?"; tail -60 gpurun_out/parity_report.log | cut -c1-260
fi
if [[ $what == *bench* ]]; then
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
  echo "bench config 2 exit $?"; cut -c1-1200 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_ref.json 2> gpurun_out/bench_c2_ref.err
  echo "bench reference exit $?"; cut -c1-1500 gpurun_out/bench_c2_ref.json; tail -3 gpurun_out/bench_c2_ref.err
fi
if [[ $what == *configs* ]]; then
  for c in 3 4 5; do
    timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
    echo "bench config $c exit $?"; cut -c1-1400 gpurun_out/bench_c$c.json; tail -3 gpurun_out/bench_c$c.err
  done
  timeout 600 python bench.py --config 5 --batch 512 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/bench_c5_b512.json 2> gpurun_out/bench_c5_b512.err
  echo "bench config 5 B=512 exit $?"; cut -c1-800 gpurun_out/bench_c5_b512.json; tail -3 gpurun_out/bench_c5_b512.err
fi
if [[ $what == *launches* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_vpt.csv \
      python bench.py --mode vpt --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-roofline --no-e2e --no-settle > gpurun_out/launches_vpt.log 2>&1
  echo "launch list vpt exit $?"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_coop.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-roofline --no-e2e --no-settle > gpurun_out/launches_coop.log 2>&1
  echo "launch list coop exit $?"
fi
if [[ $what == *ncu* ]]; then
  B="python bench.py --mode vpt --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-roofline --no-e2e"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16 -s 110 -c 12 -f -o gpurun_out/prof_gemm $B > gpurun_out/ncu_gemm.log 2>&1
  echo "ncu gemm exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmha_fwd -s 14 -c 2 -f -o gpurun_out/prof_fmha_fwd $B > gpurun_out/ncu_fmha_fwd.log 2>&1
  echo "ncu fmha_fwd exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmha_bwd -s 3 -c 2 -f -o gpurun_out/prof_fmha_bwd $B > gpurun_out/ncu_fmha_bwd.log 2>&1
  echo "ncu fmha_bwd exit $?"
fi
kill $SMI
ls -la gpurun_out | head -40
