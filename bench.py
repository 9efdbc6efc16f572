#!/usr/bin/env python
"""bench.py -- one PAMNet training step (forward + L1 loss + backward, no optimizer) on synthetic / fixture batches,
BASELINE.json's metric.

    python bench.py [--gpus N --steps K --warmup W]            configs[1]: QM9 dim=128 L=6 bs=32 per GPU (the headline)
    python bench.py --config c3                                 configs[2]: bs=256, reduced-precision tensor-core node MLPs
    python bench.py --config c4                                 configs[3]: RNA-Puzzles dim=16 L=1, the first 8 natives
    python bench.py --impl reference [...]                      the reference's CPU algorithm (oracle port) on the host cores

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

UNIT = "molecules/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4"],
                    help="BASELINE.json configs[1] (default), configs[2] (bs=256, reduced-precision node MLPs), configs[3] (RNA)")
    ap.add_argument("--batch-size", type=int, default=None)
    ap.add_argument("--dim", type=int, default=None)
    ap.add_argument("--n-layer", type=int, default=None)
    ap.add_argument("--node-mlp", default=None, choices=["f32", "tf32"],
                    help="node-MLP precision: f32 (3xTF32, fp32-accurate; default except --config c3) or tf32 (single pass)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--torch-loss", action="store_true", help="F.l1_loss instead of the fused loss + gradient launch")
    ap.add_argument("--no-prefetch", action="store_true",
                    help="build every step's graph plan inline instead of on a side stream behind the previous backward")
    ap.add_argument("--loss-readback", default="early", choices=["early", "late"],
                    help="e2e arm: early = the step's loss is read back through a copy stream right after the loss kernel "
                         "(default); late = loss.item() after backward, as the reference loop does")
    ap.add_argument("--vary", type=int, default=1,
                    help="visit K distinct synthetic batches round-robin (different atom / edge / triplet counts every step, as "
                         "an epoch does) instead of re-running one batch")
    ap.add_argument("--allreduce", default="native", choices=["native", "buckets", "plain"],
                    help="N > 1: native = library-owned NCCL all-reduce issued bucket by bucket inside backward (default); "
                         "buckets = the same buckets from Python (OverlappedGradSync); plain = one all-reduce after backward")
    a = ap.parse_args()
    d = {"c2": (32, 128, 6), "c3": (256, 128, 6), "c4": (8, 16, 1)}[a.config]
    a.batch_size = a.batch_size or d[0]
    a.dim = a.dim or d[1]
    a.n_layer = a.n_layer or d[2]
    if a.node_mlp is None:
        a.node_mlp = "tf32" if a.config == "c3" else "f32"
    return a


def metric_name(args):
    if args.config == "c4":
        return "graphs/sec PAMNet fwd+bwd (RNA-Puzzles dim=16 L=1 bs=8, first 8 native structures)"
    return f"molecules/sec PAMNet fwd+bwd (QM9 dim={args.dim} L={args.n_layer} bs={args.batch_size})"


def model_cfg(args):
    if args.config == "c4":
        return types.SimpleNamespace(dataset="rna_native", dim=args.dim, n_layer=args.n_layer, cutoff_l=2.6, cutoff_g=20.0,
                                     flow="target_to_source")
    return types.SimpleNamespace(dataset="QM9", dim=args.dim, n_layer=args.n_layer, cutoff_l=5.0, cutoff_g=5.0,
                                 flow="source_to_target")


def make_batch(args, seed):
    """Host batch of the configuration (+ a state dict to load, or None for seeded default init)."""
    import torch
    from pamnet_b200.data import Batch, synthetic_qm9_batch
    if args.config == "c4":
        gold = torch.load(os.path.join(ROOT, "tests", "golden", "rna_c4.pt"), map_location="cpu", weights_only=False)
        n = min(args.batch_size, 8)
        sizes = gold["sizes"][:n]
        tot = sum(sizes)
        b = Batch(x=gold["x"][:tot].clone(), batch=torch.repeat_interleave(torch.arange(n), torch.tensor(sizes)),
                  y=gold["y"][:n].clone())
        return b, gold["state_dict"]
    return synthetic_qm9_batch(args.batch_size, seed=seed), None


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm restated (oracle/), all host threads
# ------------------------------------------------------------------------------------------------
def cpu_step_fn(cfg, batch, sd=None, seed=0):
    import torch
    from oracle import pamnet_oracle as O
    sd = sd if sd is not None else O.init_state_dict(cfg, seed=seed)
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


def sub_batch(args, batch, n):
    """The first n graphs of `batch` (bounded CPU sample)."""
    import torch
    from pamnet_b200.data import Batch
    if n >= int(batch.batch.max()) + 1:
        return batch
    keep = batch.batch < n
    f = {"x": batch.x[keep], "batch": batch.batch[keep], "y": batch.y[:n]}
    if getattr(batch, "pos", None) is not None:
        f["pos"] = batch.pos[keep]
        nk = int(keep.sum())
        e = batch.edge_index
        f["edge_index"] = e[:, (e[0] < nk) & (e[1] < nk)]
    return Batch(**f)


def time_cpu(args, cfg, n_graphs, steps, warmup=1, budget_s=None):
    """Bounded sample: the first `n_graphs` graphs of the workload per step; returns (units/s, steps timed, cores, s/step)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch, sd = make_batch(args, seed=0)
    batch = sub_batch(args, batch, n_graphs)
    step = cpu_step_fn(cfg, batch, sd)
    for _ in range(warmup):
        step()
    times = []
    t_end = time.perf_counter() + (budget_s or 1e9)
    while len(times) < steps and (time.perf_counter() < t_end or not times):
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
    unit = "graphs/s" if args.config == "c4" else UNIT
    # size the per-step sample so the whole run stays within a few minutes
    probe = 1 if args.config == "c4" else min(8, args.batch_size)
    _, _, cores, t_step = time_cpu(args, cfg, probe, steps=1, warmup=1)
    per = t_step / probe
    budget = 150.0
    n = int(budget / max(per * (args.steps + args.warmup), 1e-9))
    n = max(1, min(args.batch_size, n))
    rate, steps, cores, t_step = time_cpu(args, cfg, n, steps=args.steps, warmup=args.warmup)
    sample = f"{n} of {args.batch_size} graphs per step, {steps} steps, oracle port (torch CPU, {cores} threads)"
    line = {
        "impl": "reference", "metric": metric_name(args), "value": rate, "unit": unit, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic" if args.config != "c4" else "reference fixture (rna_native)",
        "config": workload_config(args, None),
        "cpu_baseline": {"value": rate, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def workload_config(args, sizes):
    if args.config == "c4":
        wl = ("PAMNet rna_native dim=16 n_layer=1 batch_size=8: the first 8 graphs of the reference's RNA-Puzzles native fixture "
              "(841-3771 atoms), shipped checkpoint, fwd + L1 loss + bwd (BASELINE.json configs[3])")
    else:
        idx = {"c2": 1, "c3": 2}[args.config]
        wl = (f"PAMNet QM9 target=7 dim={args.dim} n_layer={args.n_layer} batch_size={args.batch_size} per GPU, fwd + L1 loss + bwd, "
              f"synthetic ~20-atom molecules (BASELINE.json configs[{idx}])")
    c = {"workload": wl,
         "parallelism": (f"dp{args.gpus} molecule-sharded, " + {"plain": "one flat-gradient all-reduce after backward",
                         "buckets": "bucketed gradient all-reduce issued from Python, overlapped with backward",
                         "native": "bucketed ncclAllReduce issued by the library inside backward (two-layer buckets, head last)"}[
                         getattr(args, "allreduce", "native")]) if args.gpus > 1 else "single GPU",
         "l2": "flushed (256 MiB write) before every timed step of the device-resident arm; e2e arm: no flush, a step's working set (~0.5 GB of workspace at bs=32) is larger than the 126 MB L2",
         "node_mlp": "single-pass TF32 tensor-core node MLPs (reduced precision, PAMNET_NODE_MLP=tf32)" if args.node_mlp == "tf32"
                     else "3xTF32 tensor-core node MLPs (fp32-accurate)",
         "loss": "F.l1_loss" if getattr(args, "torch_loss", False) else "pamnet_b200.ops.l1_loss (value + gradient in one launch)",
         "e2e_loss_readback": "every step, from pinned memory, copied on a copy stream right after the step's loss kernel" if getattr(args, "loss_readback", "early") == "early"
                              else "loss.item() after backward, every step",
         "front_end": "inline" if getattr(args, "no_prefetch", False)
                      else "next step's H2D copy + graph plan on a side stream from a prefetch worker thread (model.prefetch_async), overlapping this step's backward; one plan and one H2D copy per step"}
    if getattr(args, "vary", 1) > 1:
        c["batches"] = f"{args.vary} distinct batches visited round-robin (sizes below: the last one)"
    if sizes:
        c["sizes"] = sizes
    for k in ("PAMNET_FRONT", "PAMNET_GEMM", "PAMNET_CHAIN", "PAMNET_STREAMS", "PAMNET_TC2_PROD"):
        if os.environ.get(k):
            c[k.lower()] = os.environ[k]
    return c


class ClockSampler:
    """SM clock and throttle reasons during the timed regions (B200_PROFILING.md clocks line), from NVML (the source
    `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` reads), queried from the TIMING thread between two steps
    -- after a step's closing event, before the next step's opening one -- while the GPU still works through what is
    queued.  Why not a background `nvidia-smi -lms` / NVML thread: a query normally takes ~10 us, but every few seconds
    the driver refreshes its cached values and the call takes 10-40 ms while holding a lock the launching threads need;
    with a 1.5 ms step that showed up as a 10-80 ms stall of one step in about every second 300-step run
    (`tools/nvml_cost.py`, profiles/README.md).  Between the event brackets the same refresh delays the next step's
    launch but is not attributed to a step.  BENCH_SAMPLER=smi keeps the recipe's background child for comparison;
    BENCH_SAMPLER=none disables sampling."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.mode = os.environ.get("BENCH_SAMPLER", "nvml")
        self.nv = self.h = None
        self.mx = None

    def start(self):
        if self.mode == "none":
            return
        if self.mode == "nvml":
            try:
                import pynvml as nv
                import torch
                nv.nvmlInit()
                # CUDA_VISIBLE_DEVICES may renumber: resolve through the PCI bus id of the CUDA device
                props = torch.cuda.get_device_properties(self.index)
                h = None
                if hasattr(props, "pci_bus_id"):
                    for i in range(nv.nvmlDeviceGetCount()):
                        hi = nv.nvmlDeviceGetHandleByIndex(i)
                        if nv.nvmlDeviceGetPciInfo(hi).bus == props.pci_bus_id:
                            h = hi
                            break
                if h is None:
                    h = nv.nvmlDeviceGetHandleByIndex(self.index)
                self.mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)      # (slow call: once, before anything is timed)
                self.nv, self.h = nv, h
                return
            except Exception:
                self.mode = "smi"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def sample_now(self):
        """One NVML sample from the calling thread (no-op in the other modes)."""
        if self.nv is None:
            return
        try:
            sm = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
            mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            self.rows.append([str(sm), str(self.mx)] + ["Active" if mask & b else "Not Active" for _, b in self.BITS])
        except Exception:
            pass

    def wait_first(self, timeout=3.0):
        """nvidia-smi mode: block until the first sample arrived (its start-up is over) -- before the timed regions."""
        t_end = time.perf_counter() + timeout
        while self.proc is not None and not self.rows and time.perf_counter() < t_end:
            time.sleep(0.01)

    def stop(self):
        if self.mode == "none" or (self.proc is None and self.nv is None):
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler unavailable"], "samples": 0}
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = [n for n, _ in self.BITS]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm),
                "source": "NVML, queried between steps of the timed regions (GPU busy with the queued step)" if self.nv is not None
                          else "nvidia-smi -lms 200 child"}


def algorithmic_bytes_step(sz, D, L, s=4):
    """B_step = L*(3*B_f + 2*B_i) + front end."""
    N, Eg, El, T = sz["N"], sz["E_g"], sz["E_l"], sz["T2"] + sz["T1"]
    p_pair = 2 * (11 * D * D + 11 * D) + (3 * D * D + D + D * D) + 2 * (3 * D * D + D) + 2 * (D * D + D) + 2 * D * D + 4 * D + 2
    b_f = s * (4 * N * D + Eg * D + El * D + T * D + 4 * N + p_pair)
    b_i = 4 * (2 * Eg + 2 * El + 2 * T)
    front = s * (3 * N + Eg + El + T) + 4 * (2 * (Eg + N))
    return L * (3 * b_f + 2 * b_i) + front


def run_ours(args):
    if args.node_mlp == "tf32":
        os.environ["PAMNET_NODE_MLP"] = "tf32"       # read once by the library
    import torch
    import torch.distributed as dist
    import pamnet_b200
    from pamnet_b200 import Config, PAMNet, _lib
    from pamnet_b200.parallel import NativeGradSync, OverlappedGradSync, allreduce_gradients

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
    unit = "graphs/s" if args.config == "c4" else UNIT
    torch.manual_seed(0)
    host_batch, sd = make_batch(args, seed=rank)
    # --vary K: K distinct batches (different atom / edge / triplet counts) visited round-robin, as an epoch does
    host_pool = [host_batch] + [make_batch(args, seed=1000 * (i + 1) + rank)[0] for i in range(max(args.vary, 1) - 1)]
    model = PAMNet(Config(**vars(cfg)))
    if sd is not None:
        model.load_state_dict(sd)
    model = model.to(dev)
    host_pool = [b.pin_memory() for b in host_pool]
    host_batch = host_pool[0]
    dev_pool = [b.to(dev) for b in host_pool]
    dev_batch = dev_pool[0]
    turn = {"dev": 0, "e2e": 0}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    prefetch = not args.no_prefetch

    params = list(model.parameters())

    import torch.nn.functional as F
    l1 = F.l1_loss if args.torch_loss else pamnet_b200.ops.l1_loss
    sync = OverlappedGradSync(model) if (world > 1 and args.allreduce == "buckets") else None
    native = NativeGradSync(model) if (world > 1 and args.allreduce == "native") else None

    # --loss-readback early (default): the step's loss is complete when its own kernel is, long before backward ends -- it
    # is copied to pinned host memory on a copy stream behind an event recorded right after the loss launch, and the host
    # reads it there.  `late` = loss.item() after backward (the reference loop, main_qm9.py:109): the host then cannot
    # enqueue the next forward before the whole step has drained.
    early = args.loss_readback == "early"
    copy_stream = torch.cuda.Stream(device=dev) if early else None
    host_loss = torch.zeros(1, dtype=torch.float32).pin_memory() if early else None
    loss_ready = {"ev": None}

    def step(batch, sync_grads=True, next_batch=None, read_loss=False):
        if sync is not None:
            sync.wait()                     # the previous step's collectives read the gradient buffer backward is about to zero
        model.zero_grad()                   # optimizer.zero_grad() of main_qm9.py:106 (the module's O(1) form: grads stay attached, backward overwrites)
        out = model(batch)
        if next_batch is not None:          # the next step's front end: started on the prefetch worker while this thread
            next_batch()                    # enqueues the loss and the backward pass
        loss = l1(out, batch.y)             # main_qm9.py:108
        if read_loss and early:             # D2H of this step's loss: ordered after the loss kernel only
            ev = torch.cuda.Event()
            ev.record()
            copy_stream.wait_event(ev)
            with torch.cuda.stream(copy_stream):
                host_loss.copy_(loss.detach().reshape(1), non_blocking=True)
                done = torch.cuda.Event()
                done.record(copy_stream)
            loss.record_stream(copy_stream)
            loss_ready["ev"] = done
        loss.backward()
        if world > 1 and sync_grads:        # enqueue the collectives first: they overlap what backward still has queued
            if sync is not None:
                sync()
            elif native is None:            # (native: the all-reduce happened inside backward)
                allreduce_gradients(model)
        return loss

    # device-resident loop: the batch has been in HBM since before the warm-up, the side stream need not wait for anything
    fut_dev = []

    def plan_next_dev():
        fut_dev.append(model.prefetch_async(dev_pool[(turn["dev"] + 1) % len(dev_pool)], wait_current=False))
    nxt_dev = plan_next_dev if prefetch else None

    def dev_step():
        if fut_dev:
            fut_dev.pop(0).result()         # the plan requested during the previous step
        loss = step(dev_pool[turn["dev"] % len(dev_pool)], next_batch=nxt_dev)
        turn["dev"] += 1
        return loss

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
    if os.environ.get("BENCH_SWITCH"):
        sys.setswitchinterval(float(os.environ["BENCH_SWITCH"]))
    # at least 10 untimed steps (and one visit of every batch of --vary): allocator pools of both threads, tensor-map and
    # plan caches, clocks; the JSON line reports the count actually run
    warmup = max(args.warmup, 10, len(dev_pool))
    for _ in range(warmup):
        dev_step()
    if rank == 0:
        sampler.wait_first()            # NVML start-up is over before anything is timed
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    import gc
    gc.collect()
    gc.disable()                # a collector pause inside a 1.5 ms step is host jitter, not kernel time (re-enabled after the timed regions)
    launches0 = _lib.launch_count()
    ar_ms = []
    barrier()
    every = max(10, args.steps // 4)       # clock samples: ~4 per region, each right after a step's closing event
    for i, (s0, s1) in enumerate(ev):
        flush.fill_(1)
        s0.record()
        dev_step()
        if sync is not None:
            sync.wait()                     # the step ends when its gradients are reduced
        s1.record()
        if rank == 0 and i % every == every // 2:
            sampler.sample_now()
        if sync is not None and len(ar_ms) < 8:
            ar_ms.append(sync.allreduce_ms())       # (synchronises: only for a few steps)
    barrier()
    launches = (_lib.launch_count() - launches0) // args.steps
    per_step = [a.elapsed_time(b) for a, b in ev]
    dev_ms = sum(per_step) / args.steps
    # (the clock sampler keeps running through the end-to-end timed region below: both are under load)

    # ---- end to end through the public API with host buffers ("e2e") ------------------------------------
    pending = []

    if fut_dev:
        fut_dev.pop(0).result()

    def h2d_and_plan():         # H2D copy from pinned memory + graph plan of the next batch: prefetch worker, side stream
        pending.append(model.prefetch_async(host_pool[turn["e2e"] % len(host_pool)]))
        turn["e2e"] += 1

    def e2e_step():
        if prefetch:            # this step's batch was copied and planned during the previous step; copy + plan the next
            if not pending:
                h2d_and_plan()
            loss = step(pending.pop(0).result(), next_batch=h2d_and_plan, read_loss=True)
        else:
            b = host_pool[turn["e2e"] % len(host_pool)].to(dev, non_blocking=True)      # H2D from pinned memory, inside the timed region
            turn["e2e"] += 1
            loss = step(b, read_loss=True)
        if sync is not None:
            sync.wait()
        if early:                                           # D2H read of the step's result (host waits for the copy only)
            loss_ready["ev"].synchronize()
            return float(host_loss[0])
        return loss.item()

    for _ in range(3):
        e2e_step()
    if rank == 0:
        sampler.sample_now()        # under load (the warm-up steps are still running), outside the wall-clock region below
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)) / args.steps
    gc.enable()
    clocks = sampler.stop() if rank == 0 else None
    h2d = sum(v.numel() * v.element_size() for v in host_batch.__dict__.values() if isinstance(v, torch.Tensor))
    d2h = 4 + 64    # loss scalar + the count read-back of the graph build

    # max over ranks (and the spread: per-rank skew)
    skew = None
    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
        tmin = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        skew = {"ms_per_step_min_rank": float(tmin[0]), "ms_per_step_max_rank": float(t[0])}
        dev_ms, e2e_ms = t.tolist()

    if rank != 0:
        if world > 1:
            dist.barrier()          # leave together with rank 0 (it still runs the collective-free profile pass)
            dist.destroy_process_group()
        return

    sz = model.last_plan.sizes
    sizes = {"G": int(sz.n_graphs), "N": int(sz.n_nodes), "E_l": int(sz.n_edges_l), "E_g": int(sz.n_edges_g),
             "T2": int(sz.n_t2), "T1": int(sz.n_t1)}
    n_units = int(sz.n_graphs)
    value = world * n_units / (dev_ms * 1e-3)
    e2e_value = world * n_units / (e2e_ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    line = {
        "metric": metric_name(args), "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.node_mlp == "f32" else "f32 (node MLPs: tf32 single pass)",
        "data": "synthetic" if args.config != "c4" else "reference fixture (rna_native, first 8 graphs)",
        "config": workload_config(args, sizes), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": unit, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "ms_per_step_stats": {"median": sorted(per_step)[len(per_step) // 2], "min": min(per_step), "max": max(per_step),
                              "slow_steps": [[i, round(v, 3)] for i, v in enumerate(per_step)
                                             if v > 1.5 * sorted(per_step)[len(per_step) // 2]][:8]},
    }
    if world > 1:
        line["allreduce_ms"] = {"exposed_after_backward": sum(ar_ms) / len(ar_ms) if ar_ms else None,
                                "how": "CUDA events: compute stream idle -> last bucket reduced (OverlappedGradSync)" if sync is not None
                                       else ("issued inside backward by the library (not timed separately)" if native is not None
                                             else "single blocking all-reduce (not timed separately)"), **(skew or {})}

    # ---- per-kernel-class event timing (separate passes; events perturb the step, so not the timed ones) ----
    # rank 0 only and WITHOUT the gradient all-reduce: the other ranks have left, a collective here would never return
    if native is not None:
        native.enable(False)                # rank 0 is alone from here on
    if not args.no_profile:
        def profile_pass(nprof=5):
            for _ in range(2):
                step(dev_batch, sync_grads=False)
            torch.cuda.synchronize()
            _lib.profile_begin()
            for _ in range(nprof):
                step(dev_batch, sync_grads=False)
            return _lib.profile_end(), nprof
        prof, nprof = profile_pass()
        tot_ms = sum(v[0] for v in prof.values())
        kernels = {k: {"ms_per_step": v[0] / nprof, "launches_per_step": v[1] / nprof, "share": v[0] / tot_ms,
                       "alg_gb_per_s": (v[2] / 1e9) / (v[0] * 1e-3) if v[0] > 0 and v[2] > 0 else None}
                   for k, v in prof.items() if v[1]}
        dom = max(kernels, key=lambda k: kernels[k]["share"])
        # The concurrent pass charges a launch for SMs it shares with other streams (its class times can sum to more than
        # the step).  A second pass on ONE stream (PAMNET_STREAMS=1 semantics through the model's switch) gives the isolated
        # per-class times; the roofline uses those.
        iso = None
        try:
            os.environ["PAMNET_STREAMS"] = "1"
            prof1, n1 = profile_pass()
            iso = {k: {"ms_per_step": v[0] / n1, "launches_per_step": v[1] / n1} for k, v in prof1.items() if v[1]}
        finally:
            os.environ.pop("PAMNET_STREAMS", None)
        d = prof1[dom] if iso and dom in prof1 else prof[dom]
        dn = n1 if iso and dom in prof1 else nprof
        achieved = (d[2] / 1e9) / (d[0] * 1e-3)
        traffic = None
        try:        # dram__bytes_read + dram__bytes_write per launch from this round's committed ncu --set full capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get(dom, {}).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        hbm_view = {"achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "peak_source": peak_src}
        if d[3] > 0:
            # the dominant class is the 3xTF32 tcgen05 GEMM: each fp32-accurate multiply-add is three (gemm_tc2: four) tf32
            # tensor-core multiply-adds; the roofline counts three.  tf32 runs at half the bf16 rate: peak = measured bf16 / 2
            bf16 = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1650.0)))
            tf32_peak = bf16 / 2.0
            tens = 3.0 * d[3] / 1e12 / (d[0] * 1e-3)
            line["roofline"] = {"kernel": dom, "bound": "tensor", "achieved": tens, "peak": tf32_peak, "unit": "TFLOP/s",
                                "frac": tens / tf32_peak, "traffic": traffic,
                                "peak_source": "MEASURED_PEAKS.json dense bf16 (sustained) / 2 = tf32 rate" if peaks else
                                               "fallback 1650 TFLOP/s bf16 / 2",
                                "share_of_step": kernels[dom]["share"], "fp32_equivalent_tflops": tens / 3.0,
                                "class_ms_per_step_isolated": d[0] / dn, "class_ms_per_step_concurrent": kernels[dom]["ms_per_step"],
                                "hbm_view": hbm_view,
                                "note": "3xTF32 GEMM class: achieved = 3 x fp32-equivalent flops of its launches / their CUDA-event "
                                        "time in a single-stream pass (isolated); hbm_view = algorithmic operand bytes / the same time"}
        else:
            line["roofline"] = {"kernel": dom, "bound": "hbm", **hbm_view, "traffic": traffic,
                                "share_of_step": kernels[dom]["share"],
                                "class_ms_per_step_isolated": d[0] / dn, "class_ms_per_step_concurrent": kernels[dom]["ms_per_step"],
                                "note": "algorithmic bytes of the kernel's operands / CUDA-event duration per launch in a "
                                        "single-stream pass, averaged over its launches in a step"}
        b_step = algorithmic_bytes_step(sizes, args.dim, args.n_layer)
        line["step_roofline"] = {"algorithmic_bytes": b_step, "achieved": b_step / 1e9 / (dev_ms * 1e-3), "peak": hbm_peak,
                                 "unit": "GB/s", "frac": b_step / 1e9 / (dev_ms * 1e-3) / hbm_peak,
                                 "definition": "SURVEY.md 8(d) B_step / device ms_per_step"}
        line["kernels"] = kernels
        if iso:
            line["kernels_isolated"] = iso

    if not args.no_cpu_baseline and world == 1:      # reported at N = 1 only (the host cores are shared by all ranks)
        n_cpu = args.batch_size if args.config == "c2" else (32 if args.config == "c3" else 2)
        rate, steps, cores, t_step = time_cpu(args, cfg, n_cpu, steps=30, budget_s=15.0, warmup=1)
        line["cpu_baseline"] = {"value": rate, "unit": unit, "cores": cores, "kind": "port",
                                "sample": f"{steps} steps of the first {n_cpu} graphs of the same batch "
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
