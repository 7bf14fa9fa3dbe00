#!/bin/bash
# r04a (2-GPU box): live-chunk binning + peer rows requested up front.  bash profiles/r04a_run.sh
TAG=r04a
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/${TAG}_gpu_tests.log
python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-refcuda > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err
for G in 0 1 2; do
  TGS_PEER_GATHER=$G python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline --no-refcuda > gpurun_out/${TAG}_c3_2gpu_gather$G.json 2> gpurun_out/${TAG}_c3_2gpu_gather$G.err
done
