#!/bin/bash
# Round-2 profiles: serialized launch list of the bench command + full captures of the two dominant kernels.
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile > gpurun_out/r02_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 100 -c 6 -o gpurun_out/r02_gemm_tc2_full \
    python tools/one_step.py > gpurun_out/r02_ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_mma -s 60 -c 4 -o gpurun_out/r02_chain_mma_full \
    python tools/one_step.py > gpurun_out/r02_ncu_chain.log 2>&1
ls -la gpurun_out/r02_*
