#!/bin/bash
# Round-2 final single-GPU measurement set: bash profiles/final_1gpu.sh <tag>
TAG=$1
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/${TAG}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
python bench.py --steps 100 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
python bench.py --config c2 --steps 100 --no-cpu-baseline > gpurun_out/${TAG}_bench_c2.json 2> /dev/null
python bench.py --config c5 --steps 20 --no-cpu-baseline --no-refcuda --no-e2e > gpurun_out/${TAG}_bench_c5.json 2> /dev/null
python bench.py --config fixture --steps 100 --no-cpu-baseline > gpurun_out/${TAG}_bench_fixture.json 2> /dev/null
python bench.py --config fixture1m --steps 50 --no-cpu-baseline > gpurun_out/${TAG}_bench_fixture1m.json 2> /dev/null
python bench.py --forward-only --steps 100 > gpurun_out/${TAG}_bench_c3_forward_only.json 2> /dev/null
python bench.py --workload train_step --steps 100 --no-e2e > gpurun_out/${TAG}_train_c3.json 2> /dev/null
python bench.py --workload train_step --config c5 --steps 20 --no-e2e > gpurun_out/${TAG}_train_c5.json 2> /dev/null
python bench.py --workload touch_inputs --steps 20 > gpurun_out/${TAG}_bench_touch_inputs.json 2> /dev/null
bash profiles/run_ncu_r02.sh ${TAG} > /dev/null 2>&1
