#!/bin/bash
# bs=256: full captures of the message kernels (one launch each of the second step)
set -u
tag=${1:-r2msg}
mkdir -p gpurun_out
export BS=${BS:-256} STEPS=2 PAMNET_NODE_MLP=${PAMNET_NODE_MLP:-tf32}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'msg_|trip_|edge_fwd' -s 36 -c 10 -o gpurun_out/${tag}_full \
   python tools/one_step.py > gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out/${tag}_*
