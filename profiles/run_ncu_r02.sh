#!/bin/bash
# Round-2 evidence run (one GPU): launch list of the c3 bench step + one full ncu capture per own kernel.
# Usage: bash profiles/run_ncu_r02.sh <tag>     -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_<kernel>.ncu-rep
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-refcuda --cameras 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.stdout 2>&1
cap() {  # name regex skip
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/${TAG}_$1 $CMD > gpurun_out/${TAG}_$1.stdout 2>&1
}
cap render_bwd k_render_bwd 4
cap render_fwd k_render_fwd 6
cap bin_scatter k_bin_scatter 6
cap bin_count k_bin_count 6
cap bin_prefix k_bin_prefix 6
cap preprocess "k_preprocess$" 6
cap preprocess_bwd k_preprocess_bwd 4
ls -la gpurun_out/ | grep ${TAG}
