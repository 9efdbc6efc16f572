#!/usr/bin/env python
"""bench.py -- molecules/s of one PAMNet training step (forward + L1 loss + backward, no optimizer) on
synthetic QM9-shaped batches, BASELINE.json's metric and config (dim=128, n_layer=6, batch 32 per GPU).

    python bench.py [--gpus N --steps K --warmup W]            our CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference [...]                     the reference's CPU algorithm (oracle port)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what every key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "molecules/sec PAMNet fwd+bwd (QM9 dim=128 L=6 bs=32)"
UNIT = "molecules/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--n-layer", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--fused-loss", action="store_true",
                    help="opt-in: L1 loss and its gradient in one launch (pamnet_b200.ops.l1_loss) instead of F.l1_loss")
    ap.add_argument("--prefetch", action="store_true",
                    help="opt-in: build the NEXT step's graph plan on a side stream right after backward (model.prefetch); "
                         "every step still contains one H2D copy (e2e) and one front end")
    return ap.parse_args()


def model_cfg(args):
    return types.SimpleNamespace(dataset="QM9", dim=args.dim, n_layer=args.n_layer, cutoff_l=5.0, cutoff_g=5.0,
                                 flow="source_to_target")


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm restated (oracle/), all host threads
# ------------------------------------------------------------------------------------------------
def cpu_step_fn(cfg, batch, seed=0):
    import torch
    from oracle import pamnet_oracle as O
    sd = O.init_state_dict(cfg, seed=seed)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    consts = O.sbf_constants()

    def step():
        for v in leaves.values():
            v.grad = None
        out = O.forward(leaves, cfg, batch, consts=consts)
        loss = (out - batch.y).abs().mean()
        loss.backward()
        return float(loss.detach())
    return step


def time_cpu(cfg, n_graphs, budget_s, steps=None, warmup=1):
    """Bounded sample: `n_graphs` molecules per step; returns (molecules/s, steps timed, cores)."""
    import torch
    from pamnet_b200.data import synthetic_qm9_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = synthetic_qm9_batch(n_graphs, seed=0)
    step = cpu_step_fn(cfg, batch)
    for _ in range(warmup):
        step()
    times = []
    t_end = time.perf_counter() + budget_s
    while (steps is None and time.perf_counter() < t_end and len(times) < 30) or (steps is not None and len(times) < steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    tot = sum(times)
    return n_graphs * len(times) / tot, len(times), cores, tot / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = model_cfg(args)
    # size the per-step sample so the whole run stays within a few minutes
    rate, _, cores, t_step = time_cpu(cfg, min(8, args.batch_size), budget_s=0, steps=1, warmup=1)
    per_mol = t_step / min(8, args.batch_size)
    budget = 150.0
    n = int(budget / max(per_mol * (args.steps + args.warmup), 1e-9))
    n = max(1, min(args.batch_size, n))
    rate, steps, cores, t_step = time_cpu(cfg, n, budget_s=0, steps=args.steps, warmup=args.warmup)
    sample = f"{n} of {args.batch_size} molecules per step, {steps} steps, oracle port (torch CPU, {cores} threads)"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, None),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def workload_config(args, sizes):
    c = {"workload": f"PAMNet QM9 target=7 dim={args.dim} n_layer={args.n_layer} batch_size={args.batch_size} per GPU, "
                     "fwd + L1 loss + bwd, synthetic ~20-atom molecules (BASELINE.json configs[1])",
         "parallelism": f"dp{args.gpus} molecule-sharded, one flat-gradient all-reduce" if args.gpus > 1 else "single GPU",
         "l2": "flushed (256 MiB write) before every timed step"}
    if sizes:
        c["sizes"] = sizes
    if getattr(args, "fused_loss", False):
        c["loss"] = "pamnet_b200.ops.l1_loss (value + gradient in one launch)"
    if getattr(args, "prefetch", False):
        c["prefetch"] = "next step's graph plan built on a side stream after backward (model.prefetch)"
    if os.environ.get("PAMNET_FRONT"):           # opt-in front end in effect
        c["front_end"] = os.environ["PAMNET_FRONT"]
    return c


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def algorithmic_bytes_step(sz, D, L, s=4):
    """B_step = L*(3*B_f + 2*B_i) + front end."""
    N, Eg, El, T = sz["N"], sz["E_g"], sz["E_l"], sz["T2"] + sz["T1"]
    p_pair = 2 * (11 * D * D + 11 * D) + (3 * D * D + D + D * D) + 2 * (3 * D * D + D) + 2 * (D * D + D) + 2 * D * D + 4 * D + 2
    b_f = s * (4 * N * D + Eg * D + El * D + T * D + 4 * N + p_pair)
    b_i = 4 * (2 * Eg + 2 * El + 2 * T)
    front = s * (3 * N + Eg + El + T) + 4 * (2 * (Eg + N))
    return L * (3 * b_f + 2 * b_i) + front


def run_ours(args):
    import torch
    import torch.distributed as dist
    import pamnet_b200
    from pamnet_b200 import Config, PAMNet, _lib
    from pamnet_b200.data import synthetic_qm9_batch
    from pamnet_b200.parallel import allreduce_gradients

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = model_cfg(args)
    torch.manual_seed(0)
    model = PAMNet(Config(**vars(cfg))).to(dev)
    host_batch = synthetic_qm9_batch(args.batch_size, seed=rank).pin_memory()
    dev_batch = host_batch.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    loss_buf = torch.zeros(1, device=dev)

    params = list(model.parameters())

    import torch.nn.functional as F
    l1 = pamnet_b200.ops.l1_loss if args.fused_loss else F.l1_loss

    def step(batch, sync_grads=True, next_batch=None):
        for p in params:             # == optimizer.zero_grad(set_to_none=True)
            p.grad = None
        out = model(batch)
        loss = l1(out, batch.y)             # main_qm9.py:108
        loss.backward()
        if next_batch is not None:          # --prefetch: the next step's front end overlaps this step's backward
            next_batch()
        if world > 1 and sync_grads:
            allreduce_gradients(model)
        return loss

    # device-resident loop: the batch has been in HBM since before the warm-up, the side stream need not wait for anything
    nxt_dev = (lambda: model.prefetch(dev_batch, wait_current=False)) if args.prefetch else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ("value") ----------------------------------------------------------------
    # nvidia-smi needs ~0.1 s to deliver its first sample and a timed region is ~0.1-0.2 s: the sampler starts before
    # the warm-up (already under load) and runs until the end of the end-to-end region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step(dev_batch, next_batch=nxt_dev)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = _lib.launch_count()
    barrier()
    for s0, s1 in ev:
        flush.fill_(1)
        s0.record()
        step(dev_batch, next_batch=nxt_dev)
        s1.record()
    barrier()
    launches = (_lib.launch_count() - launches0) // args.steps
    dev_ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    # (the clock sampler keeps running through the end-to-end timed region below: both are under load)

    # ---- end to end through the public API with host buffers ("e2e") ------------------------------------
    pending = []

    def h2d_and_plan():         # H2D copy from pinned memory + graph plan of the next batch, both on the side stream
        pending.append(model.prefetch(host_batch))

    def e2e_step():
        if args.prefetch:       # this step's batch was copied and planned during the previous step; copy + plan the next
            if not pending:
                h2d_and_plan()
            loss = step(pending.pop(0), next_batch=h2d_and_plan)
        else:
            b = host_batch.to(dev, non_blocking=True)      # H2D from pinned memory, inside the timed region
            loss = step(b)
        return loss.item()                                  # D2H read of the step's result

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    h2d = sum(v.numel() * v.element_size() for v in host_batch.__dict__.values() if isinstance(v, torch.Tensor))
    d2h = 4 + 2 * 64    # loss scalar + the two count read-backs of the graph build

    # max over ranks
    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = t.tolist()

    if rank != 0:
        if world > 1:
            dist.barrier()          # leave together with rank 0 (it still runs the collective-free profile pass)
            dist.destroy_process_group()
        return

    sz = model.last_plan.sizes
    sizes = {"G": args.batch_size, "N": int(sz.n_nodes), "E_l": int(sz.n_edges_l), "E_g": int(sz.n_edges_g),
             "T2": int(sz.n_t2), "T1": int(sz.n_t1)}
    value = world * args.batch_size / (dev_ms * 1e-3)
    e2e_value = world * args.batch_size / (e2e_ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, sizes), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
    }

    # ---- per-kernel-class event timing (separate pass; events perturb the step, so not the timed one) ----
    # rank 0 only and WITHOUT the gradient all-reduce: the other ranks have left, a collective here would never return
    if not args.no_profile:
        for _ in range(2):
            step(dev_batch, sync_grads=False)
        torch.cuda.synchronize()
        nprof = 5
        _lib.profile_begin()
        for _ in range(nprof):
            step(dev_batch, sync_grads=False)
        prof = _lib.profile_end()
        tot_ms = sum(v[0] for v in prof.values())
        kernels = {k: {"ms_per_step": v[0] / nprof, "launches_per_step": v[1] / nprof, "share": v[0] / tot_ms,
                       "alg_gb_per_s": (v[2] / 1e9) / (v[0] * 1e-3) if v[0] > 0 and v[2] > 0 else None}
                   for k, v in prof.items() if v[1]}
        dom = max(kernels, key=lambda k: kernels[k]["share"])
        d = prof[dom]
        achieved = (d[2] / 1e9) / (d[0] * 1e-3)
        traffic = None
        try:        # dram__bytes_read + dram__bytes_write per launch from the committed ncu --set full capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(dom, {}).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        hbm_view = {"achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "peak_source": peak_src}
        if d[3] > 0:
            # the dominant class is the 3xTF32 tcgen05 GEMM: each fp32-accurate multiply-add is three tf32 tensor-core
            # multiply-adds, so the executed tensor work is 3 x the fp32-equivalent flops; tf32 runs at half the bf16
            # rate, hence peak = measured dense bf16 / 2
            bf16 = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1650.0)))
            tf32_peak = bf16 / 2.0
            tens = 3.0 * d[3] / 1e12 / (d[0] * 1e-3)
            line["roofline"] = {"kernel": dom, "bound": "tensor", "achieved": tens, "peak": tf32_peak, "unit": "TFLOP/s",
                                "frac": tens / tf32_peak, "traffic": traffic,
                                "peak_source": "MEASURED_PEAKS.json dense bf16 (sustained) / 2 = tf32 rate" if peaks else
                                               "fallback 1650 TFLOP/s bf16 / 2",
                                "share_of_step": kernels[dom]["share"], "fp32_equivalent_tflops": tens / 3.0,
                                "hbm_view": hbm_view,
                                "note": "3xTF32 GEMM class: achieved = 3 x fp32-equivalent flops of its launches / their "
                                        "CUDA-event time (concurrent streams share the SMs); hbm_view = algorithmic operand "
                                        "bytes / the same time"}
        else:
            line["roofline"] = {"kernel": dom, "bound": "hbm", **hbm_view, "traffic": traffic,
                                "share_of_step": kernels[dom]["share"],
                                "note": "algorithmic bytes of the kernel's operands / CUDA-event duration per launch, "
                                        "averaged over its launches in a step"}
        b_step = algorithmic_bytes_step(sizes, args.dim, args.n_layer)
        line["step_roofline"] = {"algorithmic_bytes": b_step, "achieved": b_step / 1e9 / (dev_ms * 1e-3), "peak": hbm_peak,
                                 "unit": "GB/s", "frac": b_step / 1e9 / (dev_ms * 1e-3) / hbm_peak,
                                 "definition": "SURVEY.md 8(d) B_step / device ms_per_step"}
        line["kernels"] = kernels

    if not args.no_cpu_baseline and world == 1:      # reported at N = 1 only (the host cores are shared by all ranks)
        rate, steps, cores, t_step = time_cpu(cfg, args.batch_size, budget_s=15.0, warmup=1)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{steps} full steps of the same {args.batch_size}-molecule batch "
                                          f"({1e3 * t_step:.0f} ms/step), oracle port on torch CPU"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
