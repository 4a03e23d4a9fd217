"""Data-parallel training of the ESS hot path: one process per GPU, minibatch sharded by sample,
weights replicated, ONE gradient exchange per iteration over NCCL (NVLink 5 / NVSwitch).

The reference is single-process; three reductions couple the batch there, so a plain "mean of
per-shard losses" would not reproduce a single-process run at the global batch:
  1. event normalisation statistics are taken over the whole batch tensor per window
     (e2vid/utils/inference_utils.py:98-107)      -> all-reduce(SUM) of the [T,3] (sum, sumsq, nnz);
  2. Dice sums / CE valid-pixel count run over the whole batch (utils/loss_functions.py:85-88,15)
                                                   -> all-reduce(SUM) of the [2+3K] partial sums;
  3. weight gradients                              -> all-reduce(SUM) of ONE flat fp32 bucket, no 1/N
                                                      (each rank already differentiates the global loss).
Payloads are tiny (26.8 MB of gradients): the design goal is exact global-batch semantics and a
single launch, not bandwidth.  Works with backend "nccl" on GPUs and "gloo" on CPU (tests).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device('cuda', local_rank))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, world, local_rank


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def allreduce_sum_(t):
    """In-place SUM all-reduce (no-op for a single process). Returns t."""
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def shard_batch(n_global, rank, world):
    """Contiguous sample range [lo, hi) of this rank (global batch must divide evenly)."""
    if n_global % world:
        raise ValueError('global batch %d is not divisible by world size %d' % (n_global, world))
    per = n_global // world
    return rank * per, (rank + 1) * per


def attach(reconstructor=None, task_loss=None):
    """Install the global-batch reduction hooks on an ImageReconstructor / TaskLoss."""
    if reconstructor is not None:
        reconstructor.stats_reduce_fn = allreduce_sum_
    if task_loss is not None:
        task_loss.reduce_fn = allreduce_sum_


class GradBucket:
    """One flat fp32 buffer holding every trainable parameter's gradient.

    `p.grad` of each parameter is a view into the buffer, so autograd accumulates straight into it
    and the exchange is a single all-reduce launch with no packing copy."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()
        off = 0
        for p in self.params:      # re-attach in case an optimizer replaced .grad (zero_grad(set_to_none))
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * off:
                p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def allreduce_(self):
        return allreduce_sum_(self.flat)
