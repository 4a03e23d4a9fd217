"""Drop-in `TaskLoss` (reference: utils/loss_functions.py:6-24; DiceLoss :96-135, BinaryDiceLoss :63-90,
torch.nn.CrossEntropyLoss(ignore_index) :15) as one fused forward kernel + one backward kernel.

`forward(predict [N,K,H,W], target [N,H,W] int64) -> 0-d tensor` with autograd.  `gamma`, `alpha`,
`weight`, `reduction` are accepted and ignored exactly as in the reference.  The per-class partial
sums are exposed so that a data-parallel caller can all-reduce them between the two kernels and every
rank back-propagates the GLOBAL-batch loss; `reduce_fn` is that hook.
"""
import torch

from . import ops


class _TaskLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, predict, target, K, ignore_index, use_dice, use_ce, reduce_fn):
        ops.require_cuda(predict, target)
        if predict.dim() != 4 or predict.shape[1] != K:
            raise AssertionError('predict & target shape do not match')      # loss_functions.py:120
        # the kernels read `target` as int64 [N*H*W] matching the logits pixel for pixel: validate instead of
        # misreading (torch's CrossEntropyLoss raises on both conditions, loss_functions.py:15)
        if target.is_floating_point() or target.dtype == torch.bool:
            raise RuntimeError('TaskLoss: expected integer class labels, got %s' % target.dtype)
        if tuple(target.shape) != (predict.shape[0], predict.shape[2], predict.shape[3]):
            raise ValueError('TaskLoss: expected target of shape %s for logits %s, got %s' %
                             ((predict.shape[0], predict.shape[2], predict.shape[3]), tuple(predict.shape), tuple(target.shape)))
        logits = predict.detach().float().permute(0, 2, 3, 1)
        if not logits.is_contiguous():
            logits = logits.contiguous()
        target = target.long().contiguous()
        if TaskLoss.check_labels:
            bad = (target != ignore_index) & ((target < 0) | (target >= K))
            if bool(bad.any()):
                raise IndexError('TaskLoss: labels outside [0, %d) that are not ignore_index=%d' % (K, ignore_index))
        sums = ops.task_loss_sums(logits, target, K, ignore_index)
        if reduce_fn is not None:
            sums = reduce_fn(sums)
        loss = ops.task_loss_finish(sums, K, ignore_index, use_dice, use_ce)
        ctx.save_for_backward(logits, target, sums)
        ctx.cfg = (K, ignore_index, use_dice, use_ce)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        logits, target, sums = ctx.saved_tensors
        K, ignore_index, use_dice, use_ce = ctx.cfg
        gs = g.detach().float().contiguous().view(1)
        dl = ops.task_loss_bwd(logits, target, K, ignore_index, sums, use_dice, use_ce, gs)
        return ops.as_nchw(dl), None, None, None, None, None, None


class TaskLoss(torch.nn.Module):
    # debug switch: verify on the device that every label is in [0, K) or == ignore_index (one reduction + a host
    # sync per call, so off on the hot path; torch's CrossEntropyLoss device-asserts on such labels)
    check_labels = False

    def __init__(self, losses=['cross_entropy'], gamma=2.0, num_classes=13, alpha=None, weight=None,
                 ignore_index=None, reduction='mean'):
        super().__init__()
        self.losses = losses
        self.weight = weight
        self.gamma = gamma
        self.alpha = alpha
        self.ignore_index = ignore_index
        self.num_classes = num_classes
        self.reduce_fn = None      # set by ess_b200.dp for global-batch semantics

    def forward(self, predict, target):
        use_dice, use_ce = 'dice' in self.losses, 'cross_entropy' in self.losses
        if not (use_dice or use_ce):
            return 0                                                            # loss_functions.py:18-24
        ign = self.ignore_index if self.ignore_index is not None else -100   # CrossEntropyLoss default
        ops.require_cuda_any(predict, target)
        if target.device != predict.device:
            raise RuntimeError('TaskLoss: predict on %s but target on %s' % (predict.device, target.device))
        with ops.on_device_of(predict):
            return _TaskLossFn.apply(predict, target, self.num_classes, int(ign), use_dice, use_ce, self.reduce_fn)
