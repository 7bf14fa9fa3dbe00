#!/bin/bash
# Run on the GPU box (via gpurun): launch list + one full capture of the render kernels.
# Usage: bash profiles/run_ncu.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --cameras 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.stdout 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render_bwd -s 4 -c 1 -f -o gpurun_out/${TAG}_render_bwd $CMD > gpurun_out/${TAG}_ncu_bwd.stdout 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render_fwd -s 6 -c 1 -f -o gpurun_out/${TAG}_render_fwd $CMD > gpurun_out/${TAG}_ncu_fwd.stdout 2>&1
ls -la gpurun_out/
