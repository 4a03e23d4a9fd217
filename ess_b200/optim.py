"""Drop-in `RAdam` (reference: utils/radam.py:7-80) as ONE multi-tensor kernel launch per parameter group.

Same constructor and per-parameter state (`step`, `exp_avg`, `exp_avg_sq`) and the same update rule:
second moment first, then first moment, variance-rectified step once N_sma >= 5, plain momentum step
before that, optional decoupled weight decay.  The reference walks `group['params']` in Python with ~10 ATen
launches per tensor (34 tensors in the supervised step, 34 + 45 in the UDA step); here all tensors of a group that
share the step count are updated by a single `essb_radam_multi_step` launch (up to 48 tensors per launch); the
step-dependent scalars are computed on the host.  `torch.optim.lr_scheduler.ExponentialLR` (the reference's
scheduler, training/base_trainer.py:64-66,388-389) works unchanged: `lr` is read from the group at every step."""
import functools
import math

import torch
from torch.optim.optimizer import Optimizer

import ctypes as C

from . import ops
from ._lib import RADAM_MAX, RadamMulti, call


@functools.lru_cache(maxsize=64)
def rectification(step, beta1, beta2):
    """(rectified?, step_size) of RAdam at `step` (radam.py:53-66).  N_sma is the length of the approximated
    simple moving average; below 5 the adaptive term is switched off ("more conservative", :60)."""
    b2t = beta2 ** step
    sma_inf = 2.0 / (1.0 - beta2) - 1.0
    sma = sma_inf - 2.0 * step * b2t / (1.0 - b2t)
    bias1 = 1.0 - beta1 ** step
    if sma >= 5:
        r = math.sqrt((1.0 - b2t) * (sma - 4.0) / (sma_inf - 4.0) * (sma - 2.0) / sma * sma_inf / (sma_inf - 2.0))
        return True, r / bias1
    return False, 1.0 / bias1


class RAdam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            lr, eps, wd = group['lr'], group['eps'], group['weight_decay']
            by_step = {}
            for p in group['params']:
                g = p.grad
                if g is None:
                    continue
                if g.is_sparse:
                    raise RuntimeError('RAdam does not support sparse gradients')
                ops.require_cuda_any(p)
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError('ess_b200.RAdam: parameters must be contiguous fp32')
                if g.dtype != torch.float32 or g.device != p.device:
                    raise RuntimeError('ess_b200.RAdam: gradients must be fp32 on the parameter\'s device')
                if not g.is_contiguous():
                    g = g.contiguous()
                st = self.state[p]
                if not st:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p)
                    st['exp_avg_sq'] = torch.zeros_like(p)
                st['step'] += 1
                by_step.setdefault((st['step'], p.device), []).append((p, g, st))
            for (step, device), items in by_step.items():
                rectified, step_size = rectification(step, float(beta1), float(beta2))
                with torch.cuda.device(device):
                    stream = ops._stream(device)
                    for i in range(0, len(items), RADAM_MAX):
                        chunk = items[i:i + RADAM_MAX]
                        d = RadamMulti()
                        d.count = len(chunk)
                        for j, (p, g, st) in enumerate(chunk):
                            d.p[j], d.g[j] = p.data_ptr(), g.data_ptr()
                            d.m[j], d.v[j], d.n[j] = st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(), p.numel()
                        d.beta1, d.beta2, d.step_lr, d.eps, d.wd_lr = beta1, beta2, step_size * lr, eps, wd * lr
                        d.rectified = int(rectified)
                        call('essb_radam_multi_step', C.byref(d), stream)
                # the kernel wrote the parameters through raw pointers: tell autograd / the packed-weight caches
                # (keyed on Parameter._version) that the tensors changed in place
                torch._C._increment_version([p for p, _, _ in items])
        return loss
