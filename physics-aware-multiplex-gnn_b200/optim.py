"""Fused optimizer step for the PAMNet modules of this package.

The reference's training step ends with three Python loops over 390 tensors (main_qm9.py:111-112,117):

    clip_grad_norm_(model.parameters(), max_norm=1000, norm_type=2)
    optimizer.step()                     # optim.Adam(lr, weight_decay, amsgrad=False), main_qm9.py:91
    ema(model)                           # utils/ema.py:13-20

``FusedAdamEMA.step()`` does the same arithmetic in two CUDA launches on the module's flat parameter / gradient
buffers (csrc/optim.cu) with no host synchronisation.  It is a ``torch.optim.Optimizer`` with one parameter group,
so the reference's LR schedulers (ExponentialLR + GradualWarmupScheduler, main_qm9.py:92-93) drive it unchanged.
There is no CPU fallback.
"""
import torch

from . import _lib


class FusedAdamEMA(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=None, ema_decay=None):
        if not hasattr(model, "_flat"):
            raise TypeError("FusedAdamEMA needs a pamnet_b200 PAMNet / PAMNet_s module (flat parameter storage)")
        super().__init__(list(model.parameters()), dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.model = model
        self.max_norm = None if max_norm is None else float(max_norm)
        self.ema_decay = None if ema_decay is None else float(ema_decay)
        self.num_steps = 0
        flat = model._flat
        if not flat.is_cuda:
            raise _lib.PamnetError("FusedAdamEMA runs on CUDA parameters only: move the model to a GPU first")
        self.exp_avg = torch.zeros_like(flat, requires_grad=False)
        self.exp_avg_sq = torch.zeros_like(flat, requires_grad=False)
        # utils/ema.py:9-11: the shadow starts as a copy of the parameters at construction time
        self.shadow = flat.detach().clone() if self.ema_decay is not None else None
        self._original = None
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=flat.device)
        self._skip, self._n_skip = None, 0
        self._skip_ranges()

    def _skip_ranges(self):
        """[begin, end) element ranges the kernel must leave untouched: tensors without a gradient RIGHT NOW (grad is None,
        requires_grad False, or unused on this dataset) -- torch.optim.Adam skips exactly those.  Re-evaluated every step
        (a later requires_grad change or a step without a preceding backward is honoured); adjacent tensors are merged."""
        m = self.model
        ranges = []
        for (_, p), off, used in zip(m._param_list, m._offsets, m._param_used):
            if used and p.requires_grad and p.grad is not None:
                continue
            end = off + p.numel()
            if ranges and off - ranges[-1][1] < 64:       # padding between tensors (128 B alignment) carries no parameter
                ranges[-1][1] = end
            else:
                ranges.append([off, end])
        if len(ranges) > 4:
            raise ValueError("more than 4 separate gradient-free parameter ranges: the fused step supports the reference "
                             "configurations (unused init_linear / embeddings) and frozen contiguous blocks only")
        flat = [x for r in ranges for x in r]
        self._skip = (_lib.c_i64 * max(len(flat), 1))(*flat)
        self._n_skip = len(ranges)
        return ranges

    def _flat_grad(self):
        """The flat gradient buffer, valid when every p.grad is the view backward attached (the normal case after
        zero_grad(set_to_none=True) + backward); gradients assigned by other means are packed into it first."""
        m = self.model
        views = m._grad_views()
        for (_, p), v, used in zip(m._param_list, views, m._param_used):
            if not (used and p.requires_grad) or p.grad is None:
                continue
            if p.grad is not v:
                v.copy_(p.grad)
        return m._gflat

    @torch.no_grad()
    def step(self, closure=None, num_updates=99999):
        if closure is not None:
            raise NotImplementedError("closures are not used by the reference trainers")
        m = self.model
        if not m._aliased(full=False):          # e.g. someone replaced param.data wholesale
            m._flatten()
        flat = m._flat
        if self.exp_avg.device != flat.device:
            raise RuntimeError("the model moved to another device after the optimizer was built")
        ranges = self._skip_ranges()
        if len(ranges) == 1 and ranges[0][0] == 0 and ranges[0][1] >= m._offsets[-1] + m._param_list[-1][1].numel():
            return None                      # no parameter has a gradient (step without backward): torch.optim.Adam does nothing
        g = self._flat_grad()
        group = self.param_groups[0]
        beta1, beta2 = group["betas"]
        self.num_steps += 1
        decay = 0.0
        if self.shadow is not None:             # utils/ema.py:14
            decay = min(self.ema_decay, (1.0 + num_updates) / (10.0 + num_updates))
        lib = _lib.load()
        with torch.cuda.device(flat.device):            # the library launches on the current device
            _lib.check(lib.pamnet_optimizer_step(
                flat.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                _lib.ptr(self.shadow), flat.numel(), self._skip, self._n_skip, self.num_steps, float(group["lr"]),
                float(beta1), float(beta2), float(group["eps"]), float(group["weight_decay"]),
                float(self.max_norm) if self.max_norm is not None else 0.0, float(decay), 1,
                self._sumsq.data_ptr(), torch.cuda.current_stream(flat.device).cuda_stream), "optimizer_step")

    # ---- checkpoint / resume: the moments and the EMA shadow live in flat buffers, not in Optimizer.state ----------
    def state_dict(self):
        sd = super().state_dict()
        sd["pamnet_flat"] = {"exp_avg": self.exp_avg.detach().clone(), "exp_avg_sq": self.exp_avg_sq.detach().clone(),
                             "shadow": None if self.shadow is None else self.shadow.detach().clone(),
                             "num_steps": int(self.num_steps)}
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        sd = dict(state_dict)
        flat = sd.pop("pamnet_flat", None)
        # validate BEFORE touching any state: a rejected checkpoint must leave param_groups as they were
        if flat is None:
            raise KeyError("not a FusedAdamEMA state_dict (no 'pamnet_flat' entry)")
        if flat["exp_avg"].numel() != self.exp_avg.numel():
            raise ValueError("optimizer state belongs to a different model configuration")
        if (flat["shadow"] is None) != (self.shadow is None):
            raise ValueError("EMA shadow present in one of (checkpoint, optimizer) only")
        super().load_state_dict(sd)
        self.exp_avg.copy_(flat["exp_avg"])
        self.exp_avg_sq.copy_(flat["exp_avg_sq"])
        if (flat["shadow"] is None) != (self.shadow is None):
            raise ValueError("EMA shadow present in one of (checkpoint, optimizer) only")
        if self.shadow is not None:
            self.shadow.copy_(flat["shadow"])
        self.num_steps = int(flat["num_steps"])

    def total_norm(self):
        """Gradient norm of the last step as clip_grad_norm_ returns it (device scalar, no synchronisation)."""
        return self._sumsq.sqrt().float()[0]

    # ---- utils/ema.py:22-33 on the flat buffers (one copy each instead of a loop over the tensors) ----------------
    @torch.no_grad()
    def ema_assign(self):
        if self.shadow is None:
            raise RuntimeError("built without ema_decay")
        flat = self.model._flat
        self._original = flat.detach().clone()
        flat.data.copy_(self.shadow)

    @torch.no_grad()
    def ema_resume(self):
        if self._original is None:
            raise RuntimeError("ema_resume() without ema_assign()")
        self.model._flat.data.copy_(self._original)
        self._original = None
