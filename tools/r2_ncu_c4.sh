#!/bin/bash
# configs[3] (RNA batch): one full capture per kernel family of a training step
set -u
mkdir -p gpurun_out
STEPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gemm_rows|gemm_cols|sbf_embed|chain_kernel|knn_warp|node_grad' -c 40 -o gpurun_out/r2c4_full python tools/one_step_c4.py > gpurun_out/r2c4_full.log 2>&1
tail -2 gpurun_out/r2c4_full.log
