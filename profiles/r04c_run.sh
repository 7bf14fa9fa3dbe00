#!/bin/bash
# r04c (2-GPU box): contributor bytes behind the screen-gradient rows.  bash profiles/r04c_run.sh
TAG=r04c
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/${TAG}_gpu_tests.log
python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-refcuda > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err
for Z in 1 2; do
  TGS_ZERO_SKIP=$Z python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline --no-refcuda > gpurun_out/${TAG}_c3_2gpu_skip$Z.json 2> gpurun_out/${TAG}_c3_2gpu_skip$Z.err
done
python bench.py --config fixture1m --steps 50 --no-cpu-baseline --no-refcuda --no-e2e > gpurun_out/${TAG}_fixture1m.json 2> /dev/null
