#!/bin/bash
# One full ncu capture of one kernel of the c3 bench step: bash profiles/ncu_one.sh <tag> <kernel-regex> [skip]
TAG=$1; K=$2; SKIP=${3:-4}
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-refcuda --cameras 1"
ncu --set full --clock-control none --import-source on -k "regex:$K" -s $SKIP -c 1 -f -o gpurun_out/${TAG} $CMD > gpurun_out/${TAG}.stdout 2>&1
