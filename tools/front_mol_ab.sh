#!/bin/bash
# First GPU call for the two paths written without a GPU -- the per-molecule front end (csrc/front_mol.cuh, PAMNET_FRONT=mol)
# and the device-side collation (csrc/collate.cuh): parity against the generic
# graph kernels, then an A/B of the headline bench.  Run under gpurun from the repo root:
#   gpurun --timeout 900 -- 'bash tools/front_mol_ab.sh'
# Outputs land in gpurun_out/front_mol_*.{log,json}.  Make the switch the default only if the parity log is green.
set -u
mkdir -p gpurun_out
PAMNET_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_front_mol.py tests/test_collate_host.py tests/test_gpu_optim.py tests/test_gpu_ops.py -q -m gpu > gpurun_out/front_mol_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/front_mol_parity.log
tail -3 gpurun_out/front_mol_parity.log
for rep in 1 2; do
  timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/front_mol_bench_generic_$rep.json 2> gpurun_out/front_mol_bench.err
  PAMNET_FRONT=mol timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/front_mol_bench_mol_$rep.json 2>> gpurun_out/front_mol_bench.err
done
for rep in 1 2; do
  timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --prefetch > gpurun_out/front_mol_bench_prefetch_$rep.json 2>> gpurun_out/front_mol_bench.err
  PAMNET_FRONT=mol timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --prefetch > gpurun_out/front_mol_bench_prefetchmol_$rep.json 2>> gpurun_out/front_mol_bench.err
  PAMNET_FRONT=mol timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --prefetch --fused-loss > gpurun_out/front_mol_bench_prefetchmolloss_$rep.json 2>> gpurun_out/front_mol_bench.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/front_mol_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.3f" % d["ms_per_step"], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "launches", d["gpu_launches"],
              "graph ms %.3f" % d.get("kernels", {}).get("graph", {}).get("ms_per_step", float("nan")))
    except Exception as exc:
        print(f, "unreadable:", exc)
PY
