#!/bin/bash
# Final single-GPU session of round 2: GPU tests, smoke(), parity report, bench lines of the BASELINE configurations and the
# reference arm.  Everything lands under gpurun_out/ (copied to profiles/ by hand).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python tools/gpu_parity_report.py gpurun_out/parity_report.json > gpurun_out/parity_report.log 2>&1; echo "parity exit $?"; tail -1 gpurun_out/parity_report.log
timeout 600 python bench.py --steps 30 --warmup 8 > gpurun_out/bench_c2_n1.json 2> gpurun_out/bench_c2_n1.err; echo "config 2 exit $?"; cut -c1-260 gpurun_out/bench_c2_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_reference.json 2> gpurun_out/bench_c2_reference.err; echo "reference exit $?"; cut -c1-200 gpurun_out/bench_c2_reference.json
for c in 3 4 5; do
  timeout 600 python bench.py --config $c --steps 15 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c${c}_n1.json 2> gpurun_out/bench_c${c}_n1.err; echo "config $c exit $?"; cut -c1-260 gpurun_out/bench_c${c}_n1.json
done
timeout 600 python bench.py --config 5 --batch 512 --steps 6 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/bench_c5_b512_n1.json 2> gpurun_out/bench_c5_b512_n1.err; echo "config 5 B=512 exit $?"; cut -c1-260 gpurun_out/bench_c5_b512_n1.json
timeout 600 python bench.py --config 1 --steps 10 --warmup 3 --no-eager-baseline > gpurun_out/bench_c1_n1.json 2> gpurun_out/bench_c1_n1.err; echo "config 1 exit $?"; cut -c1-260 gpurun_out/bench_c1_n1.json
kill $SMI
