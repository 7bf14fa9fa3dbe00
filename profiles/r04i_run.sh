#!/bin/bash
# r04i (1 GPU): racecheck of the final library with the product forward (mbarrier-tracked record arrival) and with the
# wait_group diagnostic variant; parity tests of the diagnostic variant
timeout 900 compute-sanitizer --tool racecheck --print-limit 6 python profiles/sanitize_case.py > gpurun_out/r04i_sanitizer_racecheck.log 2>&1
TGS_FWD_RECORD_SYNC=waitgroup timeout 900 compute-sanitizer --tool racecheck --print-limit 6 python profiles/sanitize_case.py > gpurun_out/r04i_sanitizer_racecheck_waitgroup.log 2>&1
TGS_FWD_RECORD_SYNC=waitgroup python -m pytest tests/test_gpu_parity.py tests/test_fullsize_parity.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r04i_waitgroup_tests.log
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r04i_gpu_tests.log
python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-refcuda --no-e2e > gpurun_out/r04i_bench_c3.json 2> /dev/null
