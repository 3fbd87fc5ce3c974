#!/bin/bash
# One gpurun call: GPU tests, the three bench workloads, an ncu launch list and ncu --set full captures of the
# tensor-core kernels.  Everything lands under gpurun_out/ (scratch); summaries are copied to profiles/ by hand.
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh [tests] [bench] [launches] [ncu]'
set -u
mkdir -p gpurun_out
what="${*:-tests bench launches ncu}"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
if [[ $what == *tests* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
fi
if [[ $what == *bench* ]]; then
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_coop.json 2> gpurun_out/bench_coop.err
  echo "bench coop exit $?"; cut -c1-900 gpurun_out/bench_coop.json
  timeout 600 python bench.py --mode vpt --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_vpt.json 2> gpurun_out/bench_vpt.err
  echo "bench vpt exit $?"; cut -c1-900 gpurun_out/bench_vpt.json
  timeout 600 python bench.py --mode upt --classes 1000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_upt.json 2> gpurun_out/bench_upt.err
  echo "bench upt exit $?"; cut -c1-900 gpurun_out/bench_upt.json
  timeout 600 python bench.py --mode cocoop --batch 32 --classes 100 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cocoop.json 2> gpurun_out/bench_cocoop.err
  echo "bench cocoop exit $?"; cut -c1-600 gpurun_out/bench_cocoop.json
  timeout 300 python tools/gpu_preprocess_bench.py > gpurun_out/bench_input_pipeline.json 2> gpurun_out/bench_input_pipeline.err
  echo "bench input pipeline exit $?"; cut -c1-600 gpurun_out/bench_input_pipeline.json
fi
if [[ $what == *launches* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_vpt.csv \
      python bench.py --mode vpt --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-e2e > gpurun_out/launches_vpt.log 2>&1
  echo "launch list vpt exit $?"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_coop.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-e2e > gpurun_out/launches_coop.log 2>&1
  echo "launch list coop exit $?"
fi
if [[ $what == *ncu* ]]; then
  B="python bench.py --mode vpt --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-e2e"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16 -s 110 -c 12 -f -o gpurun_out/prof_gemm $B > gpurun_out/ncu_gemm.log 2>&1
  echo "ncu gemm exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmha_fwd -s 3 -c 2 -f -o gpurun_out/prof_fmha_fwd $B > gpurun_out/ncu_fmha_fwd.log 2>&1
  echo "ncu fmha_fwd exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmha_bwd -s 3 -c 2 -f -o gpurun_out/prof_fmha_bwd $B > gpurun_out/ncu_fmha_bwd.log 2>&1
  echo "ncu fmha_bwd exit $?"
fi
kill $SMI
ls -la gpurun_out | head -40
