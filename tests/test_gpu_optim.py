"""GPU: the fused clip + Adam + EMA step (csrc/optim.cu, pamnet_b200.FusedAdamEMA) against the reference's own
per-step sequence -- clip_grad_norm_ -> torch.optim.Adam.step() -> EMA.__call__ (main_qm9.py:111-112,117;
utils/ema.py:13-20, restated below) -- driven by the SAME gradients, so the comparison is pure arithmetic."""
import pytest
import torch
from torch.nn.utils import clip_grad_norm_

pytestmark = pytest.mark.gpu


class RefEMA:
    """utils/ema.py:3-20 restated (test infrastructure)."""

    def __init__(self, model, decay):
        self.decay = decay
        self.shadow = {n: p.data.clone() for n, p in model.named_parameters() if p.requires_grad}

    def __call__(self, model, num_updates=99999):
        decay = min(self.decay, (1.0 + num_updates) / (10.0 + num_updates))
        for n, p in model.named_parameters():
            if p.requires_grad:
                self.shadow[n] = ((1.0 - decay) * p.data + decay * self.shadow[n]).clone()


def _pair(dataset="QM9", dim=32, n_layer=2):
    from pamnet_b200 import Config, PAMNet
    torch.manual_seed(0)
    a = PAMNet(Config(dataset, dim, n_layer, 5.0, 5.0)).cuda()
    b = PAMNet(Config(dataset, dim, n_layer, 5.0, 5.0)).cuda()
    b.load_state_dict(a.state_dict())
    return a, b


@pytest.mark.parametrize("max_norm,wd", [(1000.0, 0.0), (0.05, 0.0), (1000.0, 1e-2), (0.05, 1e-2)])
def test_fused_step_matches_reference_sequence(max_norm, wd):
    from pamnet_b200 import FusedAdamEMA, synthetic_qm9_batch
    ref, fus = _pair()
    batch = synthetic_qm9_batch(6, seed=3).to("cuda")
    opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=wd, amsgrad=False)
    ema_ref = RefEMA(ref, 0.999)
    opt_fus = FusedAdamEMA(fus, lr=1e-3, weight_decay=wd, max_norm=max_norm, ema_decay=0.999)
    for it in range(4):
        opt_fus.zero_grad(set_to_none=True)
        out = fus(batch)
        torch.nn.functional.l1_loss(out, batch.y).backward()
        # identical gradients for the reference sequence
        opt_ref.zero_grad(set_to_none=True)
        for (n, pr), (_, pf) in zip(ref.named_parameters(), fus.named_parameters()):
            pr.grad = None if pf.grad is None else pf.grad.detach().clone()
        norm_ref = clip_grad_norm_(ref.parameters(), max_norm=max_norm, norm_type=2)
        opt_ref.step()
        ema_ref(ref, num_updates=it)
        opt_fus.step(num_updates=it)
        torch.cuda.synchronize()
        assert abs(float(opt_fus.total_norm()) - float(norm_ref)) <= 2e-6 * float(norm_ref)
        # keep the two forward passes on identical weights: compare, then resynchronise exactly
        for (n, pr), (_, pf) in zip(ref.named_parameters(), fus.named_parameters()):
            scale = float(pr.abs().max()) + 1e-12
            assert float((pr - pf).abs().max()) <= 2e-6 * scale + 1e-9, (it, n)
            sh = opt_fus.shadow[fus._offsets[[k for k, _ in fus._param_list].index(n)]:][:pr.numel()].view(pr.shape)
            assert float((ema_ref.shadow[n] - sh).abs().max()) <= 2e-6 * scale + 1e-9, (it, n, "shadow")
    # the tensor without a gradient on QM9 (init_linear, models.py:35) is never touched, even with weight decay
    a0 = dict(ref.named_parameters())["init_linear.weight"]
    a1 = dict(fus.named_parameters())["init_linear.weight"]
    assert torch.equal(a0, a1)


def test_lr_scheduler_drives_fused_optimizer_and_ema_swap():
    from pamnet_b200 import FusedAdamEMA, synthetic_qm9_batch
    _, model = _pair()
    batch = synthetic_qm9_batch(4, seed=1).to("cuda")
    opt = FusedAdamEMA(model, lr=1e-3, max_norm=1000.0, ema_decay=0.9)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.5)          # main_qm9.py:92
    before = model._flat.detach().clone()
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        torch.nn.functional.l1_loss(model(batch), batch.y).backward()
        opt.step()
        sched.step()
    assert abs(opt.param_groups[0]["lr"] - 2.5e-4) < 1e-12
    after = model._flat.detach().clone()
    assert not torch.equal(before, after)
    opt.ema_assign()                                                          # utils/ema.py:22-27
    assert torch.equal(model._flat.detach(), opt.shadow)
    out_ema = model(batch)
    opt.ema_resume()                                                          # utils/ema.py:29-33
    assert torch.equal(model._flat.detach(), after)
    assert torch.isfinite(out_ema).all()


def test_optimizer_state_dict_round_trip_resumes():
    """Checkpoint / resume: moments, EMA shadow and step count travel through state_dict(); a resumed run continues
    exactly like the uninterrupted one."""
    from pamnet_b200 import FusedAdamEMA, synthetic_qm9_batch
    model_a, model_b = _pair()
    batch = synthetic_qm9_batch(4, seed=1).to("cuda")

    def steps(model, opt, n):
        for _ in range(n):
            opt.zero_grad(set_to_none=True)
            torch.nn.functional.l1_loss(model(batch), batch.y).backward()
            opt.step()

    opt_a = FusedAdamEMA(model_a, lr=1e-3, max_norm=1000.0, ema_decay=0.9)
    steps(model_a, opt_a, 2)
    ckpt_opt, ckpt_model = opt_a.state_dict(), {k: v.detach().clone() for k, v in model_a.state_dict().items()}
    steps(model_a, opt_a, 2)
    model_b.load_state_dict(ckpt_model)
    opt_b = FusedAdamEMA(model_b, lr=1e-3, max_norm=1000.0, ema_decay=0.9)
    opt_b.load_state_dict(ckpt_opt)
    assert opt_b.num_steps == 2
    steps(model_b, opt_b, 2)
    rel = lambda x, y: float((x - y).abs().max() / y.abs().max())
    assert rel(model_b._flat.detach(), model_a._flat.detach()) < 1e-6          # split-K atomics: equal up to summation order
    assert rel(opt_b.shadow, opt_a.shadow) < 1e-6 and rel(opt_b.exp_avg_sq, opt_a.exp_avg_sq) < 1e-5
    with pytest.raises(KeyError):
        opt_b.load_state_dict(torch.optim.Adam(model_b.parameters()).state_dict())


def test_step_without_backward_is_a_no_op_and_frozen_block_is_skipped():
    """torch.optim.Adam skips parameters whose grad is None: a step() without a preceding backward changes nothing, and a
    block frozen AFTER the optimizer was built (requires_grad False) keeps its values while the rest updates."""
    from pamnet_b200 import FusedAdamEMA, synthetic_qm9_batch
    model, _ = _pair()
    batch = synthetic_qm9_batch(4, seed=1).to("cuda")
    opt = FusedAdamEMA(model, lr=1e-2, max_norm=1000.0, ema_decay=0.9)
    before = model._flat.detach().clone()
    opt.zero_grad(set_to_none=True)
    opt.step()
    assert torch.equal(model._flat.detach(), before) and opt.num_steps == 0
    for p in model.global_layer[0].parameters():
        p.requires_grad_(False)
    frozen = {k: p.detach().clone() for k, p in model.global_layer[0].named_parameters()}
    torch.nn.functional.l1_loss(model(batch), batch.y).backward()
    opt.step()
    for k, p in model.global_layer[0].named_parameters():
        assert torch.equal(p.detach(), frozen[k]), k
    assert not torch.equal(model._flat.detach(), before)          # everything else moved
