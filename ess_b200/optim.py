"""Drop-in `RAdam` (reference: utils/radam.py:7-80) with the per-tensor update fused into one kernel.

Same constructor and per-parameter state (`step`, `exp_avg`, `exp_avg_sq`) and the same update rule:
second moment first, then first moment, variance-rectified step once N_sma >= 5, plain momentum step
before that, optional decoupled weight decay.  The ~10 ATen launches per tensor of the reference become a
single elementwise kernel (`essb_radam_step`); the step-dependent scalars are computed on the host."""
import functools
import math

import torch
from torch.optim.optimizer import Optimizer

from . import ops
from ._lib import call


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
        stream = ops._stream()
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            lr, eps, wd = group['lr'], group['eps'], group['weight_decay']
            for p in group['params']:
                g = p.grad
                if g is None:
                    continue
                if g.is_sparse:
                    raise RuntimeError('RAdam does not support sparse gradients')
                ops.require_cuda(p)
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError('ess_b200.RAdam: parameters must be contiguous fp32')
                if not g.is_contiguous():
                    g = g.contiguous()
                st = self.state[p]
                if not st:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p)
                    st['exp_avg_sq'] = torch.zeros_like(p)
                st['step'] += 1
                rectified, step_size = rectification(st['step'], float(beta1), float(beta2))
                call('essb_radam_step', ops._p(p), ops._p(g), ops._p(st['exp_avg']), ops._p(st['exp_avg_sq']), p.numel(),
                     beta1, beta2, step_size * lr, eps, wd * lr, int(rectified), stream)
                # the kernel wrote p through its raw pointer: tell autograd / the packed-weight caches (keyed on
                # Parameter._version) that the tensor changed in place
                torch._C._increment_version([p])
        return loss
