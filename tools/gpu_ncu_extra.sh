#!/bin/bash
# Extra ncu --set full captures: image-tower attention forward, LayerNorm kernels, input-pipeline kernels.
set -u
mkdir -p gpurun_out
B="python bench.py --mode vpt --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmha_fwd -s 14 -c 2 -f -o gpurun_out/prof_fmha_fwd_img $B > gpurun_out/ncu_fmha_fwd_img.log 2>&1
echo "ncu fmha_fwd image exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ln_ -s 40 -c 6 -f -o gpurun_out/prof_ln $B > gpurun_out/ncu_ln.log 2>&1
echo "ncu ln exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ln_bwd_fast -s 4 -c 2 -f -o gpurun_out/prof_ln_bwd $B > gpurun_out/ncu_ln_bwd.log 2>&1
echo "ncu ln_bwd exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"coeff_kernel|hpass_kernel|vpass_kernel" -s 9 -c 3 -f -o gpurun_out/prof_preprocess python tools/gpu_preprocess_bench.py --steps 2 --warmup 2 --cpu-images 2 > gpurun_out/ncu_preprocess.log 2>&1
echo "ncu preprocess exit $?"
