#!/bin/bash
# ncu launch list of the c3 bench step (per-launch durations, cold-cache + serialised: compare SHARES): bash profiles/launch_list.sh <tag>
TAG=$1
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-refcuda --cameras 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.stdout 2>&1
