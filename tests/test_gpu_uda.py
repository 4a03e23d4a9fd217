"""UDA companion path (`-m gpu`): train-mode BatchNorm kernels, the consistency
losses, the StyleEncoderE2VID drop-in and a full ESSModel.train_step-shaped iteration vs the oracle
(which is itself pinned to the reference trainer in tests/test_oracle.py)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import E2VID_CFG, O, make_e2vid, make_events, make_labels, make_semseg, rel_err, sd_cpu

pytestmark = pytest.mark.gpu
TOL = 1e-3


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().cuda()


def nchw(x):
    return x.permute(0, 3, 1, 2).cpu()


@pytest.mark.parametrize('C,relu,with_res', [(64, True, True), (128, True, False), (256, False, False)])
def test_batchnorm_train_kernels(C, relu, with_res):
    from ess_b200 import ops
    g = torch.Generator().manual_seed(0)
    N, H, W = 3, 9, 11
    x = (torch.randn(N, C, H, W, generator=g) * 1.5 + 0.4).requires_grad_(True)
    gamma = (torch.rand(C, generator=g) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, generator=g) * 0.2).requires_grad_(True)
    res = torch.randn(N, C, H, W, generator=g).requires_grad_(True) if with_res else None
    y = F.batch_norm(x, None, None, gamma, beta, training=True, eps=1e-5)
    if with_res:
        y = y + res
    if relu:
        y = torch.relu(y)
    dout = torch.randn(y.shape, generator=g)
    wrt = [x, gamma, beta] + ([res] if with_res else [])
    grads = torch.autograd.grad(y, wrt, dout)
    xd = nhwc(x.detach())
    rows = N * H * W
    mean, rstd = ops.in_stats(xd.view(1, 1, rows, C))
    mean, rstd = mean.view(-1), rstd.view(-1)
    assert rel_err(mean.cpu(), x.detach().mean((0, 2, 3))) < 1e-5
    a = (gamma.detach().cuda() * rstd).contiguous()
    b = (beta.detach().cuda() - mean * a).contiguous()
    out = ops.affine_act(xd, a, b, res=nhwc(res.detach()) if with_res else None, relu=relu)
    assert rel_err(nchw(out), y.detach()) < 2e-5
    dx, dgamma, dbeta, gres = ops.bn_backward(nhwc(dout), out if relu else None, xd, mean, rstd,
                                              gamma.detach().cuda().contiguous())
    assert rel_err(nchw(dx), grads[0]) < 1e-4
    assert rel_err(dgamma.cpu(), grads[1]) < 1e-4 and rel_err(dbeta.cpu(), grads[2]) < 1e-4
    if with_res:
        assert rel_err(nchw(gres), grads[3]) < 1e-6


@pytest.mark.parametrize('K', [5, 11, 19])
def test_uda_losses(K):
    import ess_b200
    g = torch.Generator().manual_seed(1)
    a = (torch.randn(2, K, 9, 13, generator=g) * 3).requires_grad_(True)
    b = torch.randn(2, K, 9, 13, generator=g) * 3
    ref = O.sym_js_div_loss(a, b)
    gref, = torch.autograd.grad(ref * 1.7, [a])
    ad = a.detach().cuda().requires_grad_(True)
    loss = ess_b200.symJSDivLoss()(ad, b.cuda())
    (loss * 1.7).backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    assert rel_err(ad.grad, gref) < 1e-4
    x = torch.randn(2, 64, 7, 9, generator=g).requires_grad_(True)
    y = torch.randn(2, 64, 7, 9, generator=g)
    ref = F.l1_loss(x, y)
    gx, = torch.autograd.grad(ref * 0.3, [x])
    xd = x.detach().cuda().requires_grad_(True)
    l1 = ess_b200.L1Loss()(xd, y.cuda())
    (l1 * 0.3).backward()
    assert abs(float(l1) - float(ref)) < 1e-6 and rel_err(xd.grad, gx) < 1e-6


def _make_style_encoder(seed=3):
    import ess_b200
    torch.manual_seed(seed)
    m = ess_b200.StyleEncoderE2VID(1, skip_connect=True)
    g = torch.Generator().manual_seed(seed + 1)
    for mod in m.modules():                      # non-trivial affine parameters / running statistics
        if isinstance(mod, torch.nn.BatchNorm2d):
            with torch.no_grad():
                mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) + 0.5)
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
    return m


@pytest.mark.parametrize('mode', ['fp32', 'bf16x3'])
def test_style_encoder_forward_backward_vs_oracle(mode):
    B, H, W = 2, 64, 96
    m = _make_style_encoder()
    m.mode = mode
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, 1, H, W, generator=g)
    targets = {2: torch.randn(B, 64, H // 2, W // 2, generator=g), 4: torch.randn(B, 128, H // 4, W // 4, generator=g),
               8: torch.randn(B, 256, H // 8, W // 8, generator=g)}
    params = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in sd.items()}
    stats = {}
    out_r = O.style_encoder_forward(params, x, True, True, stats)
    loss_r = sum(F.l1_loss(out_r[k], targets[k]) for k in (2, 4, 8))
    names = [k for k, v in params.items() if v.requires_grad]
    g_r = dict(zip(names, torch.autograd.grad(loss_r, [params[k] for k in names])))

    import ess_b200
    m = m.cuda().train()
    out = m(x.cuda())
    assert set(out.keys()) == {1, 2, 4, 8}
    for k in (2, 4, 8):
        assert out[k].shape == out_r[k].shape
        assert rel_err(out[k], out_r[k]) < TOL, k
    l1 = ess_b200.L1Loss()
    loss = sum(l1(out[k], targets[k].cuda()) for k in (2, 4, 8))
    loss.backward()
    assert abs(float(loss.detach()) - float(loss_r)) < TOL * float(loss_r)
    # running statistics were updated exactly once with momentum 0.1
    sd_new = m.state_dict()
    for pfx, (rm, rv) in stats.items():
        assert rel_err(sd_new[pfx + '.running_mean'], rm) < 1e-4 and rel_err(sd_new[pfx + '.running_var'], rv) < 1e-4
    worst = 0.0
    for n, p in m.named_parameters():
        ref = g_r[n]
        l2 = float((p.grad.cpu().double() - ref.double()).norm() / ref.double().norm().clamp(min=1e-30))
        worst = max(worst, l2)
        assert l2 < 2e-2, (n, l2)            # chaotic ReLU/L1-sign flips bound the end-to-end agreement
    print('worst relative L2 gradient error', worst)
    # eval mode (running statistics), forward only
    m.eval()
    with torch.no_grad():
        oe = m(x.cuda())
    oe_r = O.style_encoder_forward({k: v.cpu() for k, v in m.state_dict().items()}, x, True, False)
    assert rel_err(oe[8], oe_r[8]) < TOL


def test_uda_train_step_vs_oracle():
    """One ESSModel.train_step-shaped iteration (training/ess_trainer.py:103-148, DSEC branch) assembled from
    the drop-in modules, against oracle.uda_step on identical weights and inputs."""
    import ess_b200
    B, T, C, H, W, K = 2, 2, 5, 32, 64, 6
    e2vid = make_e2vid(mode='bf16x3')
    enc = _make_style_encoder()
    dec = make_semseg(K)
    e_sd, enc_sd, dec_sd = sd_cpu(e2vid), {k: v.detach().clone() for k, v in enc.state_dict().items()}, sd_cpu(dec)
    g = torch.Generator().manual_seed(9)
    img_a = torch.rand(B, 1, H, W, generator=g)
    labels_a = make_labels(B, H, W, K)
    data_b = make_events(B, T, C, H, W)
    ol, g_enc, g_dec = O.uda_step(e_sd, E2VID_CFG, enc_sd, dec_sd, img_a, labels_a, data_b, T, C, K)

    e2vid, enc, dec = e2vid.cuda(), enc.cuda().train(), dec.cuda()
    rec = ess_b200.ImageReconstructor(e2vid, H, W, C, 'cuda')
    task = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255)
    l1, js = ess_b200.L1Loss(), ess_b200.symJSDivLoss()
    # img_train_step (:150-180) + trainTaskStep (:182-194, latents detached for DSEC)
    lat_fake = enc(img_a.cuda())
    pred_a = dec({k: v.detach() for k, v in lat_fake.items()})
    t_img = task(pred_a[1], labels_a.cuda())
    for p in enc.parameters():
        p.requires_grad = False
    t_img.backward()
    for p in enc.parameters():
        p.requires_grad = True
    # event_train_step (:257-301)
    img_fake, _, lat_real = rec.unroll(data_b.cuda(), T, C)
    lat_real = {k: v.detach() for k, v in lat_real.items()}
    lat_fake = enc(img_fake.detach())
    e_loss = l1(lat_fake[2], lat_real[2]) + l1(lat_fake[4], lat_real[4]) + l1(lat_fake[8], lat_real[8])
    pred_second = dec(lat_fake)
    with torch.no_grad():
        pred_first_ng = dec(lat_real)
    e_loss = e_loss + js(pred_second[1], pred_first_ng[1]) + l1(pred_second[2], pred_first_ng[2]) + \
        l1(pred_second[4], pred_first_ng[4])
    pred_first = dec(lat_real)
    with torch.no_grad():
        pred_second_ng = dec({k: v.detach() for k, v in lat_fake.items()})
    t_loss = js(pred_first[1], pred_second_ng[1]) + l1(pred_first[2], pred_second_ng[2]) + \
        l1(pred_first[4], pred_second_ng[4])
    for p in dec.parameters():
        p.requires_grad = False
    e_loss.backward()
    for p in dec.parameters():
        p.requires_grad = True
    t_loss.backward()
    for name, ours, ref in (('task_img', t_img, ol['task_img']), ('e_loss', e_loss, ol['e_loss']),
                            ('t_loss', t_loss, ol['t_loss'])):
        assert abs(float(ours.detach()) - ref) < TOL * max(abs(ref), 1e-3), (name, float(ours.detach()), ref)

    def cos(a, b):
        a, b = a.double().flatten().cpu(), b.double().flatten()
        return float((a @ b) / (a.norm() * b.norm()).clamp(min=1e-30))

    for n, p in enc.named_parameters():
        assert p.grad is not None, n
        assert cos(p.grad, g_enc[n]) > 0.99, (n, cos(p.grad, g_enc[n]))
    for n, p in dec.named_parameters():
        if n.endswith('weight'):
            assert cos(p.grad, g_dec[n]) > 0.99, (n, cos(p.grad, g_dec[n]))
