#!/bin/bash
# Run on the GPU box (via gpurun): launch lists + full captures of the dominant kernels.
# Usage: bash profiles/run_ncu.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-refcuda --cameras 1"
TCMD="python bench.py --workload train_step --steps 2 --warmup 3 --no-e2e --cameras 1 --refine-every 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.stdout 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}train_launches.csv $TCMD > gpurun_out/${TAG}train_launches.stdout 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render_bwd -s 4 -c 1 -f -o gpurun_out/${TAG}_render_bwd $CMD > gpurun_out/${TAG}_ncu_bwd.stdout 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render_fwd -s 6 -c 1 -f -o gpurun_out/${TAG}_render_fwd $CMD > gpurun_out/${TAG}_ncu_fwd.stdout 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pack_ranges -s 6 -c 1 -f -o gpurun_out/${TAG}_pack $CMD > gpurun_out/${TAG}_ncu_pack.stdout 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_preprocess_bwd -s 4 -c 1 -f -o gpurun_out/${TAG}_preprocess_bwd $CMD > gpurun_out/${TAG}_ncu_pbwd.stdout 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_preprocess$" -s 6 -c 1 -f -o gpurun_out/${TAG}_preprocess $CMD > gpurun_out/${TAG}_ncu_pre.stdout 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_adam -s 4 -c 1 -f -o gpurun_out/${TAG}_adam $TCMD > gpurun_out/${TAG}_ncu_adam.stdout 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ssim_fwd -s 4 -c 1 -f -o gpurun_out/${TAG}_ssim_fwd $TCMD > gpurun_out/${TAG}_ncu_ssimf.stdout 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ssim_bwd -s 4 -c 1 -f -o gpurun_out/${TAG}_ssim_bwd $TCMD > gpurun_out/${TAG}_ncu_ssimb.stdout 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ref_render_bwd -s 1 -c 1 -f -o gpurun_out/${TAG}_ref_render_bwd python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --cameras 1 > gpurun_out/${TAG}_ncu_refbwd.stdout 2>&1
ls -la gpurun_out/ | grep ${TAG}
