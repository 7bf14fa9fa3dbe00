#!/bin/bash
# r04f (1 GPU): k_render_bwd templated on the contributor bytes (no run-time test on the single-GPU path)
TAG=r04f
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/${TAG}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
python bench.py --steps 100 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-refcuda --no-e2e > gpurun_out/${TAG}_bench_c3_b.json 2> /dev/null
python bench.py --workload train_step --steps 100 --no-e2e > gpurun_out/${TAG}_train_c3.json 2> /dev/null
