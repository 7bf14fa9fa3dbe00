#!/bin/bash
# r04g (2-GPU box): NVLink counters of the gather kernel with and without the contributor bytes
M=nvlrx__bytes.sum,nvltx__bytes.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for F in 1 0; do
  FLAGS=$F ncu --metrics $M --clock-control none -k regex:k_preprocess_bwd --csv --log-file gpurun_out/r04g_peer_gather_flags$F.csv \
     python profiles/peer_gather_ncu.py > gpurun_out/r04g_peer_gather_flags$F.log 2>&1
done
