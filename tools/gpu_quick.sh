#!/bin/bash
# Quick GPU check: GPU tests, kernel timings of the LayerNorm carry, parity report, config 2/3 bench lines.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python tools/gpu_bench_kernels.py lnfuse 2>&1 | tail -6
timeout 900 python tools/gpu_parity_report.py gpurun_out/parity_report.json > gpurun_out/parity_report.log 2>&1; echo "parity exit $?"; tail -50 gpurun_out/parity_report.log | cut -c1-200
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cut -c1-300 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cut -c1-300 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
