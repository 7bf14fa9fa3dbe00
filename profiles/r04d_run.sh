#!/bin/bash
# r04d (8-GPU box): final multi-GPU numbers of the round.  bash profiles/r04d_run.sh
TAG=r04d
run() { # name, gpus, extra env
  env $3 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $2 --steps 200 --warmup 20 --no-cpu-baseline --no-refcuda $4 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err
}
run c3_8gpu 8 TGS_X=0 ""
run c3_8gpu_noskip 8 TGS_ZERO_SKIP=0 "--no-e2e"
run c3_4gpu 4 TGS_X=0 "--no-e2e"
