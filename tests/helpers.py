"""Shared test helpers: golden loading, oracle runs, tolerance rule (SURVEY.md 8(c) ladder)."""
import os
import types

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), map_location="cpu", weights_only=False)


def cfg_of(gold):
    return types.SimpleNamespace(**gold["config"])


def batch_of(gold):
    from pamnet_b200.data import Batch
    return Batch(**gold["batch"])


def rel_err(a, b):
    """max |a-b| / max |b| (tensor-level relative error used throughout the parity tests)."""
    a, b = a.double().cpu(), b.double().cpu()
    denom = b.abs().max().clamp(min=1e-30)
    return float((a - b).abs().max() / denom)


def oracle_step(sd, cfg, batch, simple=False, loss="l1", dtype=torch.float32):
    """forward + loss + backward of the oracle; returns (out, loss, grads-by-key)."""
    from oracle import pamnet_oracle as O
    leaves = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    out = O.forward(leaves, cfg, batch, simple=simple)
    y = batch.y.to(dtype)
    l = (out - y).abs().mean() if loss == "l1" else ((out - y) ** 2).mean()
    l.backward()
    return out.detach(), l.detach(), {k: v.grad for k, v in leaves.items()}


def ladder_ok(new, ref32, ref64, tol=1e-5):
    """SURVEY.md 8(c) precision ladder: fp64 is truth; accept err_new <= max(tol, 2*err_ref),
    both measured against fp64 relative to max|fp64|.  Returns (ok, err_new, err_ref)."""
    ref64 = ref64.double().cpu()
    scale = float(ref64.abs().max().clamp(min=1e-30))
    err_new = float((new.double().cpu() - ref64).abs().max()) / scale
    err_ref = float((ref32.double().cpu() - ref64).abs().max()) / scale
    return err_new <= max(tol, 2 * err_ref), err_new, err_ref


def elementwise_ok(new, ref32, ref64, rtol=1e-4, atol_frac=5e-6):
    """Element-wise companion of ladder_ok (which is a max-norm and would hide a wrong small-magnitude row):
    |new - fp64| <= max(rtol * |fp64|, 2 * |ref32 - fp64|) + atol, atol = atol_frac * max|fp64| ... per element, where the
    fp32 reference's own element-wise error widens the bound the way SURVEY.md 8(c) allows.  Returns (ok, worst ratio)."""
    new, ref32, ref64 = new.double().cpu(), ref32.double().cpu(), ref64.double().cpu()
    atol = atol_frac * float(ref64.abs().max().clamp(min=1e-30))
    bound = torch.maximum(rtol * ref64.abs(), 2.0 * (ref32 - ref64).abs()) + atol
    ratio = float(((new - ref64).abs() / bound).max())
    return ratio <= 1.0, ratio
