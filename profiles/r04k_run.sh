#!/bin/bash
python bench.py --steps 100 --no-cpu-baseline --no-refcuda > gpurun_out/r04k_bench_c3.json 2> gpurun_out/r04k_bench_c3.err
