#!/bin/bash
# Final single-GPU measurement set of round 2 (after the skip-work changes): bash profiles/final_r04.sh <tag>
TAG=${1:-r04e}
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/${TAG}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
python bench.py --steps 100 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
python bench.py --impl reference --steps 2 --warmup 1 --no-measured-configs > gpurun_out/${TAG}_reference_c3.json 2> gpurun_out/${TAG}_reference_c3.err
python bench.py --config c2 --steps 100 --no-cpu-baseline > gpurun_out/${TAG}_bench_c2.json 2> /dev/null
python bench.py --config c5 --steps 20 --no-cpu-baseline --no-refcuda --no-e2e > gpurun_out/${TAG}_bench_c5.json 2> /dev/null
python bench.py --config fixture --steps 100 --no-cpu-baseline > gpurun_out/${TAG}_bench_fixture.json 2> /dev/null
python bench.py --config fixture1m --steps 50 --no-cpu-baseline > gpurun_out/${TAG}_bench_fixture1m.json 2> /dev/null
python bench.py --forward-only --steps 100 > gpurun_out/${TAG}_bench_c3_forward_only.json 2> /dev/null
python bench.py --workload train_step --steps 100 --no-e2e > gpurun_out/${TAG}_train_c3.json 2> /dev/null
python bench.py --workload train_step --config c5 --steps 20 --no-e2e > gpurun_out/${TAG}_train_c5.json 2> /dev/null
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-refcuda --cameras 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.stdout 2>&1
cap() {  # name regex skip
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/${TAG}_$1 $CMD > gpurun_out/${TAG}_$1.stdout 2>&1
}
cap preprocess_bwd k_preprocess_bwd 4
cap bin_prefix k_bin_prefix 6
cap bin_scatter k_bin_scatter 6
cap bin_count k_bin_count 6
