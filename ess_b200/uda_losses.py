"""UDA consistency losses as fused kernels: `symJSDivLoss` (reference: utils/loss_functions.py:27-37) and the
`torch.nn.L1Loss()` the reference uses for its cycle losses (training/ess_trainer.py:29-30, 211-255, 303-330)."""
import torch

from . import ops


def _pm(t):
    """logical NCHW (or any [N, C, ...]) -> contiguous pixel-major fp32; zero-copy for our channels_last views."""
    if t.dim() == 4:
        v = t.detach().float().permute(0, 2, 3, 1)
        return v if v.is_contiguous() else v.contiguous()
    return t.detach().float().contiguous()


def _back(g, like):
    return g.permute(0, 3, 1, 2) if like.dim() == 4 else g


class _L1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ops.require_cuda(a, b)
        if a.shape != b.shape:
            raise RuntimeError('L1Loss: shape mismatch %s vs %s' % (tuple(a.shape), tuple(b.shape)))
        pa, pb = _pm(a), _pm(b)
        sums = ops.l1_sum(pa, pb)
        ctx.save_for_backward(pa, pb)
        ctx.dim4 = a.dim() == 4
        return (sums / pa.numel()).float().view(())

    @staticmethod
    def backward(ctx, g):
        pa, pb = ctx.saved_tensors
        da = ops.l1_bwd(pa, pb, g.detach().float().contiguous().view(1))
        ga = da.permute(0, 3, 1, 2) if ctx.dim4 else da
        return (ga if ctx.needs_input_grad[0] else None), (-ga if ctx.needs_input_grad[1] else None)


class L1Loss(torch.nn.Module):
    """torch.nn.L1Loss(reduction='mean')."""

    def forward(self, input, target):
        with ops.on_device_of(input):
            return _L1Fn.apply(input, target)


class _JSFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, predict, target):
        ops.require_cuda(predict, target)
        if predict.shape != target.shape or predict.dim() != 4:
            raise RuntimeError('symJSDivLoss expects two [N, K, H, W] tensors of equal shape')
        K = predict.shape[1]
        pp, pt = _pm(predict), _pm(target)
        sums, _ = ops.jsdiv(pp, pt, K, want_sum=True)
        ctx.save_for_backward(pp, pt)
        ctx.K = K
        return (sums / (pp.numel())).float().view(())

    @staticmethod
    def backward(ctx, g):
        if ctx.needs_input_grad[1]:
            raise NotImplementedError('symJSDivLoss: gradient w.r.t. the target is not built (the reference always '
                                      'passes a no-grad target)')
        pp, pt = ctx.saved_tensors
        _, dp = ops.jsdiv(pp, pt, ctx.K, want_sum=False, gscale=g.detach().float().contiguous().view(1), want_grad=True)
        return dp.permute(0, 3, 1, 2), None


class symJSDivLoss(torch.nn.Module):
    def forward(self, predict, target):
        with ops.on_device_of(predict):
            return _JSFn.apply(predict, target)
