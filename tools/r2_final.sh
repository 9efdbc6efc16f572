#!/bin/bash
# final verification of the round: GPU suite, smoke, the three single-GPU configurations, profiles of the final build
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/final_c2.json 2> gpurun_out/final_c2.err
timeout 300 python bench.py --steps 100 --warmup 5 --vary 16 --no-cpu-baseline --no-profile > gpurun_out/final_c2_vary.json 2> gpurun_out/final_c2_vary.err
timeout 300 python bench.py --config c3 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/final_c3.json 2> gpurun_out/final_c3.err
timeout 300 python bench.py --config c4 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/final_c4.json 2> gpurun_out/final_c4.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
for f in c2 c2_vary c3 c4; do echo "== $f"; python tools/bench_brief.py < gpurun_out/final_$f.json; done
tail -c 600 gpurun_out/final_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile > gpurun_out/r02_ncu_bench.log 2>&1
STEPS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c4.csv \
    python tools/one_step_c4.py > gpurun_out/r02_ncu_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 100 -c 6 -o gpurun_out/r02_gemm_tc2_full \
    python tools/one_step.py > gpurun_out/r02_ncu_gemm.log 2>&1
ls -la gpurun_out/r02_* | head
