"""Teacher-forced parity of the SemSegE2VID decoder (`-m gpu`): every node of the static program, forward AND
backward, in the shipped `bf16x3` mode (and `fp32`), held to the 1e-3 contract against the fp64 oracle.

End-to-end weight gradients are chaotic (one ReLU sign flip behind an InstanceNorm moves everything upstream of
it), which is why an end-to-end max-norm comparison cannot hold 1e-3 for any fp32-level
implementation -- the reference's own fp32 run differs from its fp64 run by more.  SURVEY.md s7.3(c) prescribes
the remedy used here: feed each layer IDENTICAL inputs.  `module._probe` (a test hook of
`ess_b200.semseg._DecoderFn`) hands every tensor the executor produces -- each conv output / materialised
tensor in the forward pass, each gradient w.r.t. a node's output in the backward pass -- to this test, which
(1) compares it with the fp64 teacher (helpers.program_forward, pinned on CPU to the oracle by
tests/test_host_logic.py::test_decoder_program_equals_oracle) and (2) replaces it by the teacher's value, so the
next node starts from exact inputs and no deviation can propagate.  The executor's real code path runs: the same
wiring, channel offsets, concat segments, residual accumulation, tcgen05 fwd / dgrad / wgrad, IN backward to
operand planes.  All 34 parameter gradients and the three input gradients are asserted at 1e-3.

The only elements excluded are gradient entries AT a ReLU kink (|IN(y)| < 1e-4 in the teacher): there the
subgradient is decided by the last bit of the normalisation and is arbitrary in any implementation.
"""
import pytest
import torch

from helpers import O, make_labels, make_latents, make_semseg, program_forward, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3
KINK = 1e-4


def _nhwc32(t):
    return t.detach().float().permute(0, 2, 3, 1).contiguous().cuda()


class _Teacher:
    """fwd/bwd probe: records the relative error of everything the executor produces and substitutes the teacher."""

    def __init__(self, T64, G64, kink):
        self.T, self.G, self.kink = T64, G64, kink
        self.err_f, self.err_b = {}, {}
        self.n_planes_only = 0

    def fwd(self, tid, y):
        ref = self.T[tid].detach()
        self.err_f[tid] = rel_err(y[..., :ref.shape[1]].permute(0, 3, 1, 2), ref)
        return _nhwc32(ref)

    def bwd(self, tid, gy, planes):
        from ess_b200 import ops
        ref = self.G[tid]
        C = ref.shape[1]
        keep = (~self.kink[tid]) if tid in self.kink else None

        def err(t_nhwc):
            a = t_nhwc[..., :C].permute(0, 3, 1, 2).double().cpu()
            d = (a - ref).abs()
            if keep is not None:
                d = d * keep
            return float(d.max() / ref.abs().max())

        errs = []
        if gy is not None:
            errs.append(err(gy))
        if planes is not None:
            errs.append(err(planes[0].float() + planes[1].float()))
            if gy is None:
                self.n_planes_only += 1
        self.err_b[tid] = max(errs)
        g32 = _nhwc32(ref)
        new_planes = None
        if planes is not None:
            ld = planes[0].shape[-1]
            N, H, W, _ = g32.shape
            new_planes = ops.split_bf16(ops.Seg(g32), N, H, W, c_pad=ld)
            assert float(new_planes[0][..., C:].float().abs().max() if ld > C else 0.0) == 0.0
        return (g32 if gy is not None else None), new_planes


def _run(mode, K, B, H, W, skip_connect=True, freeze=False):
    import ess_b200
    kw = dict(skip_connect=True, skip_type='concat') if skip_connect else dict(skip_connect=False, skip_type='sum')
    dec = make_semseg(K, **kw)
    lat = make_latents(B, H, W)
    labels = make_labels(B, H, W, K)
    # ---- teacher: fp64, every node output with its gradient
    sd = {k: v.detach().double().requires_grad_(True) for k, v in dec.state_dict().items()}
    lat64 = {k: v.double().requires_grad_(k != 1) for k, v in lat.items()}
    outs, T64 = program_forward(dec, sd, lat64, retain=True)
    loss64 = O.task_loss(outs[1], labels, K)
    loss64.backward()
    G64 = {tid: t.grad for tid, t in T64.items() if t is not None and t.grad is not None}
    kink = {}
    from ess_b200.semseg import _Conv
    for nd in dec._nodes:
        if isinstance(nd, _Conv) and nd.stats:     # conv outputs that feed an InstanceNorm (+ReLU)
            kink[nd.out] = torch.nn.functional.instance_norm(T64[nd.out].detach(), eps=1e-5).abs() < KINK
    # ---- executor under teacher forcing
    dec = dec.cuda()
    dec.mode = mode
    if freeze:
        for p in dec.parameters():
            p.requires_grad = False
    probe = _Teacher(T64, G64, kink)
    dec._probe = probe
    lat_g = {k: v.clone().cuda().requires_grad_(k != 1) for k, v in lat.items()}
    pred = dec(lat_g)
    crit = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255)
    loss = crit(pred[1], labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    return dec, probe, sd, lat64, lat_g, loss, loss64


@pytest.mark.parametrize('mode', ['bf16x3', 'fp32'])
@pytest.mark.parametrize('K,B,H,W', [(11, 2, 64, 96), (6, 1, 40, 56)])
def test_decoder_teacher_forced_all_gradients(mode, K, B, H, W):
    dec, probe, sd, lat64, lat_g, loss, loss64 = _run(mode, K, B, H, W)
    worst_f = max(probe.err_f.values())
    worst_b = max(probe.err_b.values())
    print('%s: %d forward tensors, worst %.2e; %d node gradients (%d as bf16 planes only), worst %.2e' %
          (mode, len(probe.err_f), worst_f, len(probe.err_b), probe.n_planes_only, worst_b))
    assert len(probe.err_f) == len(dec._nodes)
    assert worst_f < TOL, sorted(probe.err_f.items(), key=lambda kv: -kv[1])[:3]
    assert abs(float(loss) - float(loss64)) < TOL * abs(float(loss64))
    assert worst_b < TOL, sorted(probe.err_b.items(), key=lambda kv: -kv[1])[:3]
    if mode == 'bf16x3':
        assert probe.n_planes_only > 0           # the IN-backward -> operand-planes path was exercised
    report = {}
    for n, p in dec.named_parameters():
        ref = sd[n].grad
        assert p.grad is not None, n
        if n.endswith('bias') and not n.startswith('decoder_scale_5'):
            assert float(p.grad.abs().max()) < 5e-6 and float(ref.abs().max()) < 1e-9, n   # cancelled by the IN
            continue
        report[n] = rel_err(p.grad, ref)
    print('parameter gradients: worst %.2e (%s)' % (max(report.values()), max(report, key=report.get)))
    assert len(report) == 18 and max(report.values()) < TOL, sorted(report.items(), key=lambda kv: -kv[1])[:4]
    for key in (8, 4, 2):
        assert rel_err(lat_g[key].grad, lat64[key].grad) < TOL, key


def test_decoder_teacher_forced_no_skip_variant():
    dec, probe, sd, lat64, lat_g, loss, loss64 = _run('bf16x3', 5, 2, 32, 48, skip_connect=False)
    assert max(probe.err_f.values()) < TOL and max(probe.err_b.values()) < TOL
    for n, p in dec.named_parameters():
        if n.endswith('weight') or n.startswith('decoder_scale_5'):
            assert rel_err(p.grad, sd[n].grad) < TOL, n
    assert rel_err(lat_g[8].grad, lat64[8].grad) < TOL


def test_decoder_teacher_forced_frozen_parameters():
    """UDA usage (training/ess_trainer.py:133-137): frozen parameters, input gradients only."""
    dec, probe, sd, lat64, lat_g, loss, loss64 = _run('bf16x3', 6, 1, 32, 48, freeze=True)
    assert all(p.grad is None for p in dec.parameters())
    assert max(probe.err_b.values()) < TOL
    for key in (8, 4, 2):
        assert rel_err(lat_g[key].grad, lat64[key].grad) < TOL, key
