#!/bin/bash
# bs=256 (configs[2]): launch list + full captures of the node chain (throughput variant) and the GEMM kernel
set -u
tag=${1:-r2c3}
mkdir -p gpurun_out
export BS=256 STEPS=2 PAMNET_NODE_MLP=${PAMNET_NODE_MLP:-tf32}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
   python tools/one_step.py > gpurun_out/${tag}_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_mma -s 56 -c 6 -o gpurun_out/${tag}_chain_full \
   python tools/one_step.py > gpurun_out/${tag}_ncu_chain.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 62 -c 6 -o gpurun_out/${tag}_gemm_full \
   python tools/one_step.py > gpurun_out/${tag}_ncu_gemm.log 2>&1
ls -la gpurun_out/${tag}_*
