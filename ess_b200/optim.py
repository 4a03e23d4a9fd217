"""Drop-in `RAdam` (reference: utils/radam.py:7-80) with the per-tensor update fused into one kernel.

Same constructor, same state (`step`, `exp_avg`, `exp_avg_sq`), same 10-slot step buffer and the
same N_sma >= 5 rectification rule; the ~10 ATen launches per tensor of the reference become one."""
import math

import torch
from torch.optim.optimizer import Optimizer

from . import ops
from ._lib import call


class RAdam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.buffer = [[None, None, None] for _ in range(10)]
        super().__init__(params, defaults)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError('RAdam does not support sparse gradients')
                ops.require_cuda(p)
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError('ess_b200.RAdam: parameters must be contiguous fp32')
                grad = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                state = self.state[p]
                if len(state) == 0:
                    state['step'] = 0
                    state['exp_avg'] = torch.zeros_like(p)
                    state['exp_avg_sq'] = torch.zeros_like(p)
                state['step'] += 1
                buffered = self.buffer[int(state['step'] % 10)]
                if state['step'] == buffered[0]:
                    N_sma, step_size = buffered[1], buffered[2]
                else:
                    buffered[0] = state['step']
                    beta2_t = beta2 ** state['step']
                    N_sma_max = 2 / (1 - beta2) - 1
                    N_sma = N_sma_max - 2 * state['step'] * beta2_t / (1 - beta2_t)
                    buffered[1] = N_sma
                    if N_sma >= 5:
                        step_size = math.sqrt((1 - beta2_t) * (N_sma - 4) / (N_sma_max - 4) * (N_sma - 2) /
                                              N_sma * N_sma_max / (N_sma_max - 2)) / (1 - beta1 ** state['step'])
                    else:
                        step_size = 1.0 / (1 - beta1 ** state['step'])
                    buffered[2] = step_size
                call('essb_radam_step', ops._p(p), ops._p(grad), ops._p(state['exp_avg']), ops._p(state['exp_avg_sq']),
                     p.numel(), beta1, beta2, step_size * group['lr'], group['eps'],
                     group['weight_decay'] * group['lr'], int(N_sma >= 5), ops._stream())
        return loss
