#!/bin/bash
# ncu session of round 2 (keeps gpurun_out/ small: reports are reduced to CSV on the box).
#   launch lists of one config-2 / config-3 bench step; --set full of one launch per GEMM shape of an image-tower layer,
#   of the attention forward / backward kernels.
set -u
mkdir -p gpurun_out
B="--steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline --no-roofline --no-e2e --no-settle"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_c3.csv \
    python bench.py --config 3 $B > gpurun_out/launches_c3.log 2>&1; echo "launch list config 3 exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py $B > gpurun_out/launches_c2.log 2>&1; echo "launch list config 2 exit $?"
timeout 600 ncu --set full --clock-control none -k regex:gemm_f16 -c 9 -f -o gpurun_out/prof_gemm_shapes \
    python tools/gpu_ncu_shapes.py > gpurun_out/ncu_gemm_shapes.log 2>&1; echo "ncu gemm shapes exit $?"
ncu -i gpurun_out/prof_gemm_shapes.ncu-rep --page raw --csv > gpurun_out/prof_gemm_shapes_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmha_fwd -s 13 -c 1 -f -o gpurun_out/prof_fmha_fwd \
    python bench.py --config 3 $B > gpurun_out/ncu_fmha_fwd.log 2>&1; echo "ncu fmha_fwd exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmha_bwd -s 3 -c 1 -f -o gpurun_out/prof_fmha_bwd \
    python bench.py --config 3 $B > gpurun_out/ncu_fmha_bwd.log 2>&1; echo "ncu fmha_bwd exit $?"
for k in fmha_fwd fmha_bwd; do ncu -i gpurun_out/prof_$k.ncu-rep --page raw --csv > gpurun_out/prof_${k}_raw.csv 2>/dev/null; done
du -sh gpurun_out; ls -la gpurun_out | grep -E "prof_|launches_|gemm_order"
