"""Module-level parity (`-m gpu`): the drop-in nn.Modules vs the golden fixture (generated from the
reference) and vs the CPU oracle with identical state_dicts.  Calls go through the reference-facing
module surface -> C ABI."""
import pytest
import torch

from helpers import (E2VID_CFG, O, make_e2vid, make_events, make_labels, make_latents, make_semseg, rel_err, sd_cpu)

pytestmark = pytest.mark.gpu
TOL = 1e-3


def test_golden_end_to_end_fp32(golden):
    """tiny config of tests/golden: reconstructor T-loop -> decoder -> loss -> grads -> mIoU."""
    import ess_b200
    d = golden['dims']
    m = ess_b200.E2VIDRecurrent(golden['cfg'], mode='fp32')
    m.load_state_dict(golden['e2vid_sd'])
    m = m.cuda().eval()
    rec = ess_b200.ImageReconstructor(m, d['H'], d['W'], d['C'], 'cuda')
    data = golden['data'].cuda()
    for i in range(d['T']):
        img, states, latent = rec.update_reconstruction(data[:, i * d['C']:(i + 1) * d['C']])
    assert rel_err(img, golden['img']) < TOL
    for k in (1, 2, 4, 8):
        assert latent[k].shape == golden['latent'][k].shape
        assert rel_err(latent[k], golden['latent'][k]) < TOL
    for (h, c), (gh, gc) in zip(states, golden['states']):
        assert rel_err(h, gh) < TOL and rel_err(c, gc) < TOL
    # fused unroll == per-window loop
    img2, states2, latent2 = rec.unroll(data, d['T'], d['C'])
    assert rel_err(img2, golden['img']) < TOL and rel_err(latent2[8], golden['latent'][8]) < TOL

    dec = ess_b200.SemSegE2VID(golden['cfg']['base_num_channels'] * 8, d['K'], skip_connect=True, skip_type='concat')
    dec.load_state_dict(golden['semseg_sd'])
    dec = dec.cuda()
    lat = {k: v.cuda() for k, v in golden['latent'].items()}
    pred = dec(lat)
    assert set(pred.keys()) == {8, 4, 2, 1}
    for k in (1, 2, 4):
        assert rel_err(pred[k], golden['pred'][k]) < TOL
    crit = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], gamma=2.0, num_classes=d['K'], ignore_index=255,
                             reduction='mean')
    loss = crit(pred[1], golden['labels'].cuda())
    assert abs(float(loss) - float(golden['loss'])) < TOL * abs(float(golden['loss']))
    loss.backward()
    for n, p in dec.named_parameters():
        ref = golden['grads'][n]
        assert p.grad is not None, n
        # abs term: conv biases in front of an InstanceNorm have zero true gradient (rounding noise only)
        assert (p.grad.cpu() - ref).abs().max() <= TOL * ref.abs().max() + 5e-6, n
    met = ess_b200.MetricsSemseg(d['K'], 255, ['c%d' % i for i in range(d['K'])])
    met.update_batch(pred[1].argmax(1), golden['labels'].cuda())
    s = met.get_metrics_summary()
    assert torch.equal(s['cm'], golden['confusion'])                    # integer work: bit-exact
    assert abs(float(s['mean_iou']) - float(golden['mean_iou'])) < 1e-9


@pytest.mark.parametrize('mode,tol', [('fp32', 1e-3), ('bf16x3', 1e-3), ('f16f8', 1e-3), ('bf16', 6e-2)])
def test_e2vid_lightweight_vs_oracle(mode, tol):
    """E2VID-lightweight architecture (10.7 M params), 3 windows, all four precision modes."""
    import ess_b200
    B, T, C, H, W = 2, 3, 5, 64, 96
    m = make_e2vid(mode=mode)
    sd = sd_cpu(m)
    data = make_events(B, T, C, H, W)
    img_r, st_r, lat_r = O.encoder_unroll(sd, E2VID_CFG, data, T, C)
    m = m.cuda()
    rec = ess_b200.ImageReconstructor(m, H, W, C, 'cuda')
    for i in range(T):
        img, st, lat = rec.update_reconstruction(data[:, i * C:(i + 1) * C].cuda())
    errs = {k: rel_err(lat[k], lat_r[k]) for k in (1, 2, 4, 8)}
    errs['img'] = rel_err(img, img_r)
    errs['c2'] = rel_err(st[2][1], st_r[2][1])
    print(mode, errs)
    assert max(errs.values()) < tol, errs


@pytest.mark.parametrize('mode', ['bf16x3', 'f16f8'])
def test_full_length_unroll_logits_within_contract(mode):
    """The contract end to end at the benchmark's recurrence depth: T = 20 windows through the recurrent encoder in the
    tensor-core modes, then the decoder -- latents, image and LOGITS within 1e-3 of the fp32 oracle (the f16f8 operand
    scheme was selected by CPU emulation, tools/precision_emul.py; this is its measurement on the device)."""
    import ess_b200
    B, T, C, H, W, K = 1, 20, 5, 64, 96, 11
    m = make_e2vid(mode=mode)
    sd = sd_cpu(m)
    data = make_events(B, T, C, H, W)
    img_r, st_r, lat_r = O.encoder_unroll(sd, E2VID_CFG, data, T, C)
    dec = make_semseg(K)
    with torch.no_grad():
        logits_r = O.semseg_forward(sd_cpu(dec), lat_r)[1]
    m, dec = m.cuda(), dec.cuda()
    dec.mode = mode
    rec = ess_b200.ImageReconstructor(m, H, W, C, 'cuda')
    img, st, lat = rec.unroll(data.cuda(), T, C)
    with torch.no_grad():
        logits = dec({k: v for k, v in lat.items()})[1]
    errs = {k: rel_err(lat[k], lat_r[k]) for k in (1, 2, 4, 8)}
    errs.update(img=rel_err(img, img_r), c0=rel_err(st[0][1], st_r[0][1]), logits=rel_err(logits, logits_r))
    print(mode, 'T=20:', {k: '%.1e' % v for k, v in errs.items()})
    assert max(errs.values()) < TOL, errs


@pytest.mark.parametrize('variant', ['convgru', 'upsample_conv', 'no_norm', 'reflect_pad', 'concat', 'in_norm'])
def test_e2vid_variants_fp32(variant):
    import ess_b200
    cfg = dict(E2VID_CFG, base_num_channels=8, num_bins=3)
    H, W = 32, 40
    if variant == 'convgru':
        cfg['recurrent_block_type'] = 'convgru'
    elif variant == 'upsample_conv':
        cfg['use_upsample_conv'] = True
    elif variant == 'no_norm':
        cfg.pop('norm')
    elif variant == 'concat':
        cfg['skip_type'] = 'concat'
    elif variant == 'in_norm':
        cfg['norm'] = 'IN'
    else:
        H, W = 30, 43
    m = make_e2vid(cfg, mode='fp32')
    sd = sd_cpu(m)
    data = make_events(2, 2, 3, H, W)
    img_r, st_r, lat_r = O.encoder_unroll(sd, cfg, data, 2, 3)
    m = m.cuda()
    rec = ess_b200.ImageReconstructor(m, H, W, 3, 'cuda')
    img, st, lat = rec.unroll(data.cuda(), 2, 3)
    assert img.shape == img_r.shape
    assert rel_err(img, img_r) < TOL and rel_err(lat[8], lat_r[8]) < TOL and rel_err(lat[1], lat_r[1]) < TOL


@pytest.mark.parametrize('mode', ['bf16x3', 'f16f8'])
@pytest.mark.parametrize('variant', ['upsample_conv', 'concat', 'upsample_concat', 'convgru_concat'])
def test_e2vid_variants_tensor_core(variant, mode):
    """UpsampleConvLayer (submodules.py:65-93: bilinear x2 + conv5x5) and skip_type='concat' (unet.py:175-179) image
    decoders on the tcgen05 path (the concatenated skip tensor is a second K segment of the launch), at tensor-core
    widths, vs the oracle: image and latents within 1e-3.  No fallback warning may fire."""
    import warnings
    import ess_b200
    cfg = dict(E2VID_CFG)
    if 'upsample' in variant:
        cfg['use_upsample_conv'] = True
    if 'concat' in variant:
        cfg['skip_type'] = 'concat'
    if 'convgru' in variant:
        cfg['recurrent_block_type'] = 'convgru'
    H, W = 48, 64
    m = make_e2vid(cfg, mode=mode)
    sd = sd_cpu(m)
    data = make_events(2, 2, 5, H, W)
    img_r, st_r, lat_r = O.encoder_unroll(sd, cfg, data, 2, 5)
    m = m.cuda()
    rec = ess_b200.ImageReconstructor(m, H, W, 5, 'cuda')
    with warnings.catch_warnings():
        warnings.simplefilter('error', RuntimeWarning)
        img, st, lat = rec.unroll(data.cuda(), 2, 5)
    assert img.shape == img_r.shape
    errs = dict(img=rel_err(img, img_r), l8=rel_err(lat[8], lat_r[8]), l1=rel_err(lat[1], lat_r[1]))
    print(variant, mode, {k: '%.1e' % v for k, v in errs.items()})
    assert max(errs.values()) < TOL, errs


@pytest.mark.parametrize('mode', ['bf16x3', 'f16f8'])
def test_e2vid_instance_norm_in_tensor_core_mode(mode):
    """norm='IN' (submodules.py:21-22,149-151): the conv layers' InstanceNorm2d(track_running_stats=True) is folded like an
    eval BatchNorm (running statistics, no affine) and the encoder stays on the tcgen05 kernels; the two resblocks use
    per-sample statistics, so the image decoder takes the fp32 kernels (and says so)."""
    import ess_b200
    cfg = dict(E2VID_CFG, norm='IN')
    H, W = 48, 64
    m = make_e2vid(cfg, mode=mode)
    sd = sd_cpu(m)
    data = make_events(2, 2, 5, H, W)
    img_r, st_r, lat_r = O.encoder_unroll(sd, cfg, data, 2, 5)
    m = m.cuda()
    rec = ess_b200.ImageReconstructor(m, H, W, 5, 'cuda')
    from ess_b200 import e2vid as E
    E._WARNED.clear()
    with pytest.warns(RuntimeWarning, match='image decoder'):
        img, st, lat = rec.unroll(data.cuda(), 2, 5)
    errs = dict(img=rel_err(img, img_r), l8=rel_err(lat[8], lat_r[8]), l2=rel_err(lat[2], lat_r[2]), c2=rel_err(st[2][1], st_r[2][1]))
    print(mode, {k: '%.1e' % v for k, v in errs.items()})
    assert max(errs.values()) < TOL, errs


def test_fallback_to_cuda_core_path_warns():
    """A configuration outside the tensor-core path (base_num_channels = 8) still computes on the fp32 kernels, but says so."""
    import ess_b200
    cfg = dict(E2VID_CFG, base_num_channels=8, num_bins=3)
    m = make_e2vid(cfg, mode='bf16x3').cuda()
    rec = ess_b200.ImageReconstructor(m, 32, 40, 3, 'cuda')
    from ess_b200 import e2vid as E
    E._WARNED.clear()
    with pytest.warns(RuntimeWarning, match='CUDA-core'):
        rec.unroll(make_events(1, 1, 3, 32, 40).cuda(), 1, 3)


def _oracle_semseg_grads(dec, lat, labels, K, want_inputs=False, dtype=torch.float32, **kw):
    params = {k: v.detach().cpu().to(dtype).clone().requires_grad_(True) for k, v in dec.state_dict().items()}
    lat_c = {k: v.detach().cpu().to(dtype).clone().requires_grad_(want_inputs) for k, v in lat.items()}
    pred = O.semseg_forward(params, lat_c, **kw)
    loss = O.task_loss(pred[1], labels.cpu(), K)
    wrt = list(params.values()) + ([lat_c[8], lat_c[4], lat_c[2]] if want_inputs and kw.get('skip_connect', True)
                                   else ([lat_c[8]] if want_inputs else []))
    grads = torch.autograd.grad(loss, wrt)
    return pred, loss, dict(zip(list(params.keys()) + ['in8', 'in4', 'in2'][:len(wrt) - len(params)], grads))


@pytest.mark.parametrize('mode', ['fp32', 'bf16x3'])
@pytest.mark.parametrize('K,H,W', [(11, 64, 96), (6, 40, 56)])
def test_semseg_forward_backward_vs_oracle(K, H, W, mode):
    """Decoder forward + backward through the module surface vs the oracle, one fixed draw, both modes.

    Forward values and the loss must meet the 1e-3 contract outright.  End-to-end weight gradients are
    chaotic: a ReLU input within the implementation's rounding error of zero lands on the other side, which
    zeroes one gradient element and shifts everything upstream of it.  A forward deviation of
    1e-5 (bf16x3) flips ~10 of the 1.5 M ReLU inputs of this test and moves every upstream layer's gradient by
    ~8e-3 in relative L2 -- measured on CPU by injecting 1e-5 noise into the fp32 oracle's conv outputs, i.e. a
    property of the function, not of the kernels.  The 1e-3 gradient contract is therefore enforced where it is
    well-posed -- per layer on identical inputs, all 34 gradients, in tests/test_gpu_teacher.py -- and here the
    end-to-end gradients get flip-robust norm bounds per parameter: cosine > 0.999 and relative L2 < 3e-2 against
    the fp64 oracle (a wiring defect -- dropped segment, wrong channel offset, missed upsample backward -- moves
    these by O(1)).  The survey's criterion err <= max(1e-3, 2*err(ref32, fp64)) is reported per layer."""
    import ess_b200
    B = 2
    labels = make_labels(B, H, W, K).cuda()
    crit = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255)
    dec = make_semseg(K).cuda()
    dec.mode = mode
    lat = make_latents(B, H, W, seed=5, device='cuda')
    pred_r, loss_r, g_r = _oracle_semseg_grads(dec, lat, labels, K, skip_connect=True, skip_type='concat')
    _, _, g_64 = _oracle_semseg_grads(dec, lat, labels, K, dtype=torch.float64, skip_connect=True, skip_type='concat')
    pred = dec(lat)
    loss = crit(pred[1], labels)
    loss.backward()
    for k in (1, 2, 4):
        assert rel_err(pred[k], pred_r[k]) < TOL
    assert abs(float(loss.detach()) - float(loss_r)) < TOL * abs(float(loss_r))
    table, within = [], 0
    for n, p in dec.named_parameters():
        r64 = g_64[n].flatten()
        a = p.grad.cpu().double().flatten()
        if n.endswith('bias') and not n.startswith('decoder_scale_5'):
            assert float((a - r64).abs().max()) < 5e-6, n            # zero true gradient: rounding noise only
            continue
        l2 = float((a - r64).norm() / r64.norm())
        l2_ref = float((g_r[n].double().flatten() - r64).norm() / r64.norm())
        cos = float((a @ r64) / (a.norm() * r64.norm()))
        table.append((n, l2, l2_ref, cos))
        within += l2 <= max(1e-3, 2 * l2_ref)
    print('%s: rel-L2 vs fp64 worst %.2e (reference fp32: %.2e); min cosine %.6f; %d/%d layers within max(1e-3, 2*ref)' %
          (mode, max(t[1] for t in table), max(t[2] for t in table), min(t[3] for t in table), within, len(table)))
    assert len(table) == 18
    for n, l2, l2_ref, cos in table:
        assert cos > 0.999 and l2 < 3e-2, (n, l2, l2_ref, cos)


def test_semseg_input_grads_with_frozen_params():
    """UDA usage (training/ess_trainer.py:133-137): parameters frozen, gradients flow to the inputs."""
    import ess_b200
    K, B, H, W = 6, 1, 32, 48
    dec = make_semseg(K).cuda()
    lat = make_latents(B, H, W, device='cuda')
    labels = make_labels(B, H, W, K).cuda()
    _, _, g_r = _oracle_semseg_grads(dec, lat, labels, K, want_inputs=True, skip_connect=True, skip_type='concat')
    for p in dec.parameters():
        p.requires_grad = False
    lat_g = {k: v.clone().requires_grad_(k != 1) for k, v in lat.items()}
    pred = dec(lat_g)
    crit = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255)
    crit(pred[1], labels).backward()
    assert all(p.grad is None for p in dec.parameters())
    for key, name in ((8, 'in8'), (4, 'in4'), (2, 'in2')):
        a, b = lat_g[key].grad.double().flatten().cpu(), g_r[name].double().flatten()
        cos = float((a @ b) / (a.norm() * b.norm()))
        l2 = float((a - b).norm() / b.norm())
        # bounded, not 1e-3: a single ReLU flip moves the input gradients of this tiny case by ~1e-2
        assert cos > 0.999 and l2 < 5e-2, (key, cos, l2)
    # and under no_grad nothing is recorded
    with torch.no_grad():
        out = dec(lat)
    assert not out[1].requires_grad


def test_semseg_no_skip_variant():
    import ess_b200
    K, B, H, W = 5, 2, 32, 32
    dec = make_semseg(K, skip_connect=False, skip_type='sum').cuda()
    lat = make_latents(B, H, W, device='cuda')
    labels = make_labels(B, H, W, K).cuda()
    pred_r, loss_r, g_r = _oracle_semseg_grads(dec, lat, labels, K, skip_connect=False, skip_type='sum')
    pred = dec(lat)
    assert set(pred.keys()) == set(pred_r.keys())
    for k in pred_r:
        assert rel_err(pred[k], pred_r[k]) < TOL
    crit = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255)
    crit(pred[1], labels).backward()
    n = 'decoder_scale_5.0.weight'
    assert rel_err(dict(dec.named_parameters())[n].grad, g_r[n]) < 5e-3


@pytest.mark.parametrize('mode', ['fp32', 'bf16x3'])
def test_semseg_input_index_map_variant(mode):
    """SemSegE2VID(skip_connect=False, input_index_map=True) (style_networks.py:35-62,89-106): two coordinate channels
    are concatenated to in[8], the three INSResBlocks run 258 channels wide (executed zero-padded to 260), then
    258 -> 128 -> 64 -> 32 -> K.  Forward at 1e-3, every parameter gradient and the input gradient by the flip-robust
    norm criterion against the fp64 oracle; state_dict shapes are the reference's."""
    import ess_b200
    K, B, H, W = 5, 2, 32, 48
    dec = make_semseg(K, skip_connect=False, skip_type='sum', input_index_map=True)
    assert dec.decoder_scale_1[0].model[0].weight.shape == (258, 258, 3, 3)
    assert dec.decoder_scale_2[1].model[0].weight.shape == (128, 258, 3, 3)
    dec = dec.cuda()
    dec.mode = mode
    lat = make_latents(B, H, W, device='cuda')
    labels = make_labels(B, H, W, K).cuda()
    kw = dict(skip_connect=False, skip_type='sum', input_index_map=True)
    pred_r, loss_r, g_r = _oracle_semseg_grads(dec, lat, labels, K, want_inputs=True, dtype=torch.float64, **kw)
    lat_g = {k: v.clone().requires_grad_(k == 8) for k, v in lat.items()}
    pred = dec(lat_g)
    assert set(pred.keys()) == set(pred_r.keys())
    for k in pred_r:
        assert pred[k].shape == pred_r[k].shape and rel_err(pred[k], pred_r[k]) < TOL, k
    crit = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255)
    loss = crit(pred[1], labels)
    assert abs(float(loss.detach()) - float(loss_r)) < TOL * abs(float(loss_r))
    loss.backward()
    for n, p in list(dec.named_parameters()) + [('in8', lat_g[8])]:
        a, r = p.grad.detach().cpu().double().flatten(), g_r[n].flatten()
        assert p.grad.shape == p.shape
        if n.endswith('bias') and not n.startswith('decoder_scale_5'):
            assert float((a - r).abs().max()) < 5e-6, n
            continue
        cos, l2 = float(a @ r / (a.norm() * r.norm())), float((a - r).norm() / r.norm())
        assert cos > 0.999 and l2 < 3e-2, (n, cos, l2)
    # second call re-uses the cached coordinates; a different batch size rebuilds them
    assert rel_err(dec({k: v for k, v in lat.items()})[1], pred_r[1]) < TOL
    lat1 = {k: v[:1].contiguous() for k, v in lat.items()}
    assert dec(lat1)[1].shape[0] == 1


def test_full_size_properties_dsec():
    """BASELINE.json full size (440x640, C=5): properties that need no CPU oracle -- the tcgen05 path
    agrees with the exact-fp32 path, states carry, an all-zero window leaves zero-input statistics."""
    import ess_b200
    B, T, C, H, W = 1, 2, 5, 440, 640
    data = make_events(B, T, C, H, W).cuda()
    outs = {}
    for mode in ('fp32', 'bf16x3', 'f16f8'):
        m = make_e2vid(mode=mode).cuda()
        rec = ess_b200.ImageReconstructor(m, H, W, C, 'cuda')
        img, st, lat = rec.unroll(data, T, C)
        outs[mode] = (img, st, lat)
        assert lat[8].shape == (B, 256, 55, 80) and lat[1].shape == (B, 32, 440, 640) and img.shape == (B, 1, H, W)
        assert bool(torch.isfinite(lat[8]).all()) and float(img.min()) >= 0 and float(img.max()) <= 1
    for mode in ('bf16x3', 'f16f8'):
        for k in (1, 2, 4, 8):
            assert rel_err(outs[mode][2][k], outs['fp32'][2][k]) < TOL, (mode, k)
        assert rel_err(outs[mode][0], outs['fp32'][0]) < TOL, mode


@pytest.mark.parametrize('mode', ['fp32', 'bf16x3', 'f16f8'])
def test_e2vid_module_forward_signature(mode):
    """E2VIDRecurrent.forward(event_tensor NCHW, prev_states) itself (no reconstructor): two chained calls,
    states passed back in, outputs are logical-NCHW tensors of the reference's shapes."""
    B, C, H, W = 1, 5, 32, 64
    m = make_e2vid(mode=mode)
    sd = sd_cpu(m)
    ev = make_events(B, 2, C, H, W)
    r1 = O.e2vid_recurrent_forward(sd, E2VID_CFG, ev[:, :C], None)
    r2 = O.e2vid_recurrent_forward(sd, E2VID_CFG, ev[:, C:], r1[1])
    m = m.cuda()
    with torch.no_grad():                 # as the reference trainers call it (frozen encoder): the fused inference path
        o1 = m(ev[:, :C].cuda(), None)
        o2 = m(ev[:, C:].cuda(), o1[1])
    assert not o2[2][8].requires_grad
    assert o2[0].shape == (B, 1, H, W) and o2[2][8].shape == (B, 256, H // 8, W // 8)
    assert rel_err(o2[0], r2[0]) < TOL
    for k in (1, 2, 4, 8):
        assert rel_err(o2[2][k], r2[2][k]) < TOL
    for (h, c), (hr, cr) in zip(o2[1], r2[1]):
        assert rel_err(h, hr) < TOL and rel_err(c, cr) < TOL
    # states handed back as plain contiguous NCHW copies (not our channels_last views) also work
    st = [(h.contiguous(), c.contiguous()) for (h, c) in o1[1]]
    with torch.no_grad():
        o2b = m(ev[:, C:].cuda(), st)
    assert rel_err(o2b[2][8], r2[2][8]) < TOL


@pytest.mark.parametrize('mode,block', [('f16f8', 'convlstm'), ('fp32', 'convlstm'), ('f16f8', 'convgru')])
def test_e2vid_backprop_through_time(mode, block):
    """The differentiable path of E2VIDRecurrent (gradient mode + trainable encoder; SURVEY.md s8f 'later'): three
    chained windows, a loss on the last window's latents and states, gradients w.r.t. EVERY encoder parameter (head,
    three stride-2 convs + eval BatchNorm affine, the gate convolutions of three ConvLSTM / ConvGRU cells) and w.r.t. the events of all
    windows vs autograd through the oracle.  Forward values must agree with the fused inference path as well."""
    import warnings
    B, T, C, H, W = 2, 3, 5, 32, 48
    CFG = dict(E2VID_CFG, recurrent_block_type=block)
    lstm = block == 'convlstm'
    m = make_e2vid(CFG, mode=mode)
    sd = {k: (v.detach().clone().double().requires_grad_(v.is_floating_point() and 'running' not in k and 'num_batches' not in k)
              if v.is_floating_point() else v.detach().clone()) for k, v in m.state_dict().items()}
    ev = make_events(B, T, C, H, W)
    ev64 = ev.double().requires_grad_(True)
    g = torch.Generator().manual_seed(3)
    wts = {k: torch.randn(1, generator=g).item() for k in (1, 2, 4, 8, 'h0', 'c2')}

    def loss_of(lat, st):
        h0, s2 = (st[0][0], st[2][1]) if lstm else (st[0], st[2])
        return sum(wts[k] * lat[k].pow(2).mean() for k in (1, 2, 4, 8)) + wts['h0'] * h0.sum() * 1e-3 + wts['c2'] * s2.pow(2).mean()

    st = None
    for t in range(T):
        _, st, lat = O.e2vid_recurrent_forward(sd, CFG, ev64[:, t * C:(t + 1) * C], st, with_image=False)
    loss_r = loss_of(lat, st)
    names = [k for k, v in sd.items() if v.is_floating_point() and v.requires_grad and not k.startswith('unetrecurrent.resblocks')
             and not k.startswith('unetrecurrent.decoders') and not k.startswith('unetrecurrent.pred')]
    grads_r = torch.autograd.grad(loss_r, [sd[k] for k in names] + [ev64])
    m = m.cuda()
    evc = ev.cuda().requires_grad_(True)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        st = None
        for t in range(T):
            img, st, lat = m(evc[:, t * C:(t + 1) * C], st, with_image=(t == T - 1))
    assert lat[8].requires_grad and img is not None and not img.requires_grad
    loss = loss_of(lat, st)
    assert abs(float(loss) - float(loss_r)) < 1e-3 * abs(float(loss_r))
    loss.backward()
    params = dict(m.named_parameters())
    worst = 0.0
    for k, gr in zip(names, grads_r[:-1]):
        ga = params[k].grad
        assert ga is not None, k
        e = float((ga.detach().cpu().double() - gr).norm() / gr.norm().clamp_min(1e-30))
        worst = max(worst, e)
        assert e < 2e-3, (k, e)
    e_in = float((evc.grad.cpu().double() - grads_r[-1]).norm() / grads_r[-1].norm())
    print(mode, block, 'BPTT: loss %.6f (oracle %.6f), worst parameter-gradient rel-L2 %.1e, event-gradient %.1e'
          % (float(loss), float(loss_r), worst, e_in))
    assert e_in < 2e-3
    # and the fused inference path computes the same forward
    with torch.no_grad():
        st2 = None
        for t in range(T):
            _, st2, lat2 = m(ev[:, t * C:(t + 1) * C].cuda(), st2, with_image=False)
    assert rel_err(lat2[8], lat[8]) < TOL and rel_err(st2[2][1] if lstm else st2[2], st[2][1] if lstm else st[2]) < TOL


def test_training_trajectory_and_miou_parity():
    """Learnable synthetic task: labels = argmax of a fixed random 'teacher' decoder on the
    same latents.  Train the decoder for 25 RAdam steps with our modules (GPU, bf16x3) and with the oracle
    (CPU fp32, oracle.radam_step), same init, same data: loss curves and final mIoU must agree."""
    import ess_b200
    from ess_b200.optim import RAdam
    K, B, H, W, steps = 5, 2, 32, 48, 25
    lat = make_latents(B, H, W, seed=11)
    teacher = make_semseg(K, seed=123)
    with torch.no_grad():
        labels = O.semseg_forward(sd_cpu(teacher), lat)[1].argmax(1)
    labels[:, :2] = 255
    dec = make_semseg(K, seed=6)
    params = {k: v.detach().clone().requires_grad_(True) for k, v in dec.state_dict().items()}
    states = {k: {} for k in params}
    ref_losses = []
    for _ in range(steps):
        pred = O.semseg_forward(params, lat)
        loss = O.task_loss(pred[1], labels, K)
        grads = torch.autograd.grad(loss, list(params.values()))
        ref_losses.append(float(loss))
        with torch.no_grad():
            for (k, p), g in zip(params.items(), grads):
                O.radam_step(p, g, states[k], 5e-4, (0., 0.999))
    with torch.no_grad():
        conf_ref = O.confusion_matrix(O.semseg_forward(params, lat)[1].argmax(1), labels, K, 255)
    miou_ref = float(O.confusion_to_iou(conf_ref)[0])

    dec = dec.cuda()
    crit = ess_b200.TaskLoss(losses=['dice', 'cross_entropy'], num_classes=K, ignore_index=255)
    opt = RAdam(dec.parameters(), lr=5e-4, weight_decay=0., betas=(0., 0.999))
    lat_d = {k: v.cuda() for k, v in lat.items()}
    lab_d = labels.cuda()
    losses = []
    for _ in range(steps):
        opt.zero_grad()
        loss = crit(dec(lat_d)[1], lab_d)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    met = ess_b200.MetricsSemseg(K, 255, ['c%d' % i for i in range(K)])
    with torch.no_grad():
        met.update_batch(dec(lat_d)[1].argmax(1), lab_d)
    miou = float(met.get_metrics_summary()['mean_iou'])
    print('loss first/last ours %.5f/%.5f oracle %.5f/%.5f; mIoU ours %.3f oracle %.3f' %
          (losses[0], losses[-1], ref_losses[0], ref_losses[-1], miou, miou_ref))
    assert losses[-1] < losses[0] and ref_losses[-1] < ref_losses[0]          # both actually learn
    assert abs(losses[0] - ref_losses[0]) < 1e-3 * ref_losses[0]
    assert max(abs(a - b) for a, b in zip(losses, ref_losses)) < 1e-2 * ref_losses[0]   # trajectories track
    assert abs(miou - miou_ref) < 1.0                                             # mIoU in percent points


@pytest.mark.parametrize('mode', ['bf16x3', 'f16f8'])
def test_e2vid_ten_bins_config5(mode):
    """BASELINE config 5 uses C=10 voxel bins: the tensor-core head conv then runs with 16-channel pixels
    (two 64-wide K chunks per kernel row)."""
    import ess_b200
    cfg = dict(E2VID_CFG, num_bins=10)
    B, T, C, H, W = 1, 2, 10, 32, 64
    m = make_e2vid(cfg, mode=mode)
    sd = sd_cpu(m)
    data = make_events(B, T, C, H, W)
    img_r, st_r, lat_r = O.encoder_unroll(sd, cfg, data, T, C)
    m = m.cuda()
    rec = ess_b200.ImageReconstructor(m, H, W, C, 'cuda')
    img, st, lat = rec.unroll(data.cuda(), T, C)
    assert rel_err(img, img_r) < TOL
    for k in (1, 2, 4, 8):
        assert rel_err(lat[k], lat_r[k]) < TOL, k


@pytest.mark.parametrize('mode', ['bf16x3', 'f16f8'])
def test_ddd17_shape_config2(mode):
    """BASELINE config 2: raw DDD17 width 346 is reflect-padded to 352 (L3/R3) and the logits stay at the
    padded size; K=6.  Checked against the oracle at B=1, T=2 (CPU finishes in seconds)."""
    import ess_b200
    B, T, C, H, W, K = 1, 2, 5, 200, 346, 6
    m = make_e2vid(mode=mode)
    sd = sd_cpu(m)
    data = make_events(B, T, C, H, W)
    img_r, st_r, lat_r = O.encoder_unroll(sd, E2VID_CFG, data, T, C)
    m = m.cuda()
    rec = ess_b200.ImageReconstructor(m, H, W, C, 'cuda')
    img, st, lat = rec.unroll(data.cuda(), T, C)
    assert img.shape == (B, 1, 200, 352) and lat[8].shape == (B, 256, 25, 44)
    assert rel_err(img, img_r) < TOL
    for k in (1, 2, 4, 8):
        assert rel_err(lat[k], lat_r[k]) < TOL, k
    dec = make_semseg(K).cuda()
    pred = dec({k: v.detach() for k, v in lat.items()})
    pred_r = O.semseg_forward(sd_cpu(dec), lat_r)
    assert pred[1].shape == (B, K, 200, 352)
    assert rel_err(pred[1], pred_r[1]) < TOL


@pytest.mark.parametrize('mode', ['bf16x3', 'f16f8'])
@pytest.mark.parametrize('B,H,W', [(3, 200, 346), (2, 48, 64), (4, 56, 80)])
def test_row_stacked_levels_match_dense_and_oracle(mode, B, H, W, monkeypatch):
    """Row-stacked levels (E2VIDRecurrent._stack_plan: the batch as ONE tall image with shared zero rows, masked stores,
    tall stride-2 convs): (1) equal to the dense layout up to the fp32 accumulation order (2e-5), (2) within 1e-3 of the
    oracle over 3 windows incl. the image decoder, (3) the same through FOREIGN states (cloned tensors the module
    has to re-stack) and through the per-window module call, (4) returned states / latents have the reference shapes."""
    import ess_b200
    from ess_b200.e2vid import E2VIDRecurrent as E
    T, C = 3, 5
    Hp, Wp = (H + 7) // 8 * 8, (W + 7) // 8 * 8
    plan = E._stack_plan(B, Hp, Wp, 3)
    assert any(rows > (Hp >> (i + 1)) for i, (rows, _) in enumerate(plan)), plan     # stacking is active at this shape
    m = make_e2vid(mode=mode)
    sd = sd_cpu(m)
    data = make_events(B, T, C, H, W)
    img_r, st_r, lat_r = O.encoder_unroll(sd, E2VID_CFG, data, T, C)
    m = m.cuda()
    rec = ess_b200.ImageReconstructor(m, H, W, C, 'cuda')
    img, st, lat = rec.unroll(data.cuda(), T, C)
    assert lat[8].shape == lat_r[8].shape and st[2][0].shape == st_r[2][0].shape
    errs = dict(img=rel_err(img, img_r), **{'l%d' % k: rel_err(lat[k], lat_r[k]) for k in (1, 2, 4, 8)},
                c2=rel_err(st[2][1], st_r[2][1]), c0=rel_err(st[0][1], st_r[0][1]))
    print(mode, (B, H, W), plan, {k: '%.1e' % v for k, v in errs.items()})
    assert max(errs.values()) < TOL, errs
    # (1) dense layout, same kernels
    monkeypatch.setenv('ESS_B200_STACK', '0')
    img_d, st_d, lat_d = rec.unroll(data.cuda(), T, C)
    monkeypatch.delenv('ESS_B200_STACK')
    # same arithmetic per output pixel; only the fp32 accumulation ORDER may differ where the dispatcher picks another
    # kernel variant for the other tiling (CTA pairs vs classic: tap-major vs chunk-major K loop)
    same = dict(img=rel_err(img_d, img), **{'l%d' % k: rel_err(lat_d[k], lat[k]) for k in (1, 2, 4, 8)},
                **{'s%d' % i: max(rel_err(a[0], b[0]), rel_err(a[1], b[1])) for i, (a, b) in enumerate(zip(st_d, st))})
    assert max(same.values()) < 2e-5, same
    # (3) foreign states: two windows, the second one from clones of the first one's states
    ev = data.cuda()
    rec2 = ess_b200.ImageReconstructor(m, H, W, C, 'cuda')
    _, s1, _ = rec2.update_reconstruction(ev[:, :C])
    rec2.last_states_for_each_channel['grayscale'] = [(h.clone(), c.clone()) for h, c in s1]
    img2, s2, lat2 = rec2.update_reconstruction(ev[:, C:2 * C])
    rec3 = ess_b200.ImageReconstructor(m, H, W, C, 'cuda')
    rec3.update_reconstruction(ev[:, :C])
    img3, s3, lat3 = rec3.update_reconstruction(ev[:, C:2 * C])
    assert torch.equal(img2, img3) and torch.equal(lat2[8], lat3[8]) and torch.equal(s2[1][1], s3[1][1])


@pytest.mark.parametrize('mode', ['f16f8', 'fp32'])
def test_unroll_cuda_graph_matches_launch_by_launch(mode):
    """ImageReconstructor.unroll(graph=True): the T encoder steps replayed as ONE CUDA graph give bit-identical results
    to the launch-by-launch loop -- on first use (capture + replay), on new data at the same address (replay only), at
    further addresses (one graph each, then the shared static-input graph), and after the weights change (re-capture)."""
    import ess_b200
    B, T, C, H, W = 2, 3, 5, 48, 64
    m = make_e2vid(mode=mode).cuda()
    rec = ess_b200.ImageReconstructor(m, H, W, C, 'cuda')
    ref = ess_b200.ImageReconstructor(m, H, W, C, 'cuda')

    def check(x):
        img, st, lat = rec.unroll(x, T, C, graph=True)
        img_e, st_e, lat_e = ref.unroll(x.clone(), T, C, graph=False)
        assert torch.equal(img, img_e) and all(torch.equal(lat[k], lat_e[k]) for k in (1, 2, 4, 8))
        assert all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(st, st_e))
        assert rec.last_states_for_each_channel['grayscale'] is st

    x = make_events(B, T, C, H, W, seed=1).cuda()
    check(x)
    assert len(rec._graphs) == 1
    x.copy_(make_events(B, T, C, H, W, seed=2).cuda())        # same address, new contents: replay only
    check(x)
    assert len(rec._graphs) == 1
    keep = [x]
    for seed in range(3, 3 + rec.MAX_ADDRESS_GRAPHS + 2):      # more addresses than address-keyed graphs
        keep.append(make_events(B, T, C, H, W, seed=seed).cuda())
        check(keep[-1])
    assert len(rec._graphs) == rec.MAX_ADDRESS_GRAPHS + 1 and any(k[0] == 'static' for k in rec._graphs)
    with torch.no_grad():
        m.unetrecurrent.head.conv2d.bias.add_(0.01)            # new parameter version: stale graphs are dropped
    check(x)
    assert len(rec._graphs) == 1


@pytest.mark.parametrize('mode', ['bf16x3', 'f16f8'])
def test_convgru_in_tensor_core_mode(mode):
    """ConvGRU checkpoints (`recurrent_block_type='convgru'`, model.py:77-80) in bf16x3 mode: head and encoder convs on
    tcgen05, each GRU cell as two tcgen05 launches (GRU_UR epilogue: update gate + planes of prev_state*reset;
    GRU_OUT epilogue: out gate + blend), ess_b200.e2vid._gru_tc."""
    import ess_b200
    cfg = dict(E2VID_CFG, recurrent_block_type='convgru')
    B, T, C, H, W = 1, 2, 5, 32, 64
    m = make_e2vid(cfg, mode=mode)
    sd = sd_cpu(m)
    data = make_events(B, T, C, H, W)
    img_r, st_r, lat_r = O.encoder_unroll(sd, cfg, data, T, C)
    m = m.cuda()
    rec = ess_b200.ImageReconstructor(m, H, W, C, 'cuda')
    img, st, lat = rec.unroll(data.cuda(), T, C)
    assert torch.is_tensor(st[0]) and st[0].shape == st_r[0].shape      # GRU state is a tensor, not a tuple
    assert rel_err(img, img_r) < TOL
    for k in (1, 2, 4, 8):
        assert rel_err(lat[k], lat_r[k]) < TOL, k


@pytest.mark.parametrize('mode', ['fp32', 'bf16x3', 'f16f8'])
def test_reconstructor_flip_and_hot_pixels(mode, tmp_path):
    """ImageReconstructor with the EventPreprocessor options of e2vid/utils/inference_utils.py:73-93 (hot-pixel
    file, flip) on a width that needs reflect padding (30 -> 32), per-window and fused-unroll forms."""
    import types
    import ess_b200
    B, T, C, H, W = 2, 2, 5, 24, 30
    hot = [(3, 5), (29, 0), (10, 23)]
    f = tmp_path / 'hot.txt'
    f.write_text('\n'.join('%d,%d' % xy for xy in hot))
    m = make_e2vid(mode=mode)
    sd = sd_cpu(m)
    data = make_events(B, T, C, H, W)
    for x, y in hot:
        data[:, :, y, x] = 7.0
    ref_in = data.clone()
    states = None
    for i in range(T):
        win = ref_in[:, i * C:(i + 1) * C]
        img_r, states, lat_r = O.reconstructor_step(sd, E2VID_CFG, win, states, hot_pixels=hot, flip=True)
    opts = types.SimpleNamespace(flip=True, hot_pixels_file=str(f), no_normalize=False, no_recurrent=False, color=False)
    rec = ess_b200.ImageReconstructor(m.cuda(), H, W, C, 'cuda', opts)
    d = data.clone().cuda()
    for i in range(T):
        img, st, lat = rec.update_reconstruction(d[:, i * C:(i + 1) * C].contiguous())
    tol = TOL if mode != 'fp32' else 2e-5
    assert rel_err(img, img_r) < tol
    for k in (1, 2, 4, 8):
        assert rel_err(lat[k], lat_r[k]) < tol
    d2 = data.clone().cuda()
    img_u, _, lat_u = rec.unroll(d2, T, C)
    assert float(d2[0, 0, 5, 3]) == 0.0                       # zeroed in place, all T*C channels
    assert rel_err(lat_u[8], lat_r[8]) < tol and rel_err(img_u, img_r) < tol
