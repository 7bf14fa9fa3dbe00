#!/bin/bash
# r04b (1 GPU): zero-row shortcut of k_preprocess_bwd, A/B + the fraction of zero rows.  bash profiles/r04b_run.sh
TAG=r04b
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/${TAG}_gpu_tests.log
python profiles/zero_rows.py c3 > gpurun_out/${TAG}_zero_rows.json 2> gpurun_out/${TAG}_zero_rows.err
for Z in 0 1; do
  TGS_ZERO_SKIP=$Z python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-refcuda --no-e2e > gpurun_out/${TAG}_c3_zskip$Z.json 2> gpurun_out/${TAG}_c3_zskip$Z.err
  TGS_ZERO_SKIP=$Z python bench.py --config fixture1m --steps 50 --no-cpu-baseline --no-refcuda --no-e2e > gpurun_out/${TAG}_fixture1m_zskip$Z.json 2> /dev/null
done
TGS_ZERO_SKIP=1 python bench.py --workload train_step --steps 100 --no-e2e > gpurun_out/${TAG}_train_c3.json 2> /dev/null
