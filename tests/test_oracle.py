"""Pins the oracle: against the committed golden fixture (generated from the reference) and, where
/root/reference is present (build container only), against the live imported reference modules."""
import pytest
import torch

from helpers import O, rel_err
from oracle import ref_shim


def test_oracle_matches_golden(golden):
    d = golden['dims']
    img, states, latent = O.encoder_unroll(golden['e2vid_sd'], golden['cfg'], golden['data'], d['T'], d['C'])
    assert rel_err(img, golden['img']) < 1e-5
    for k in (1, 2, 4, 8):
        assert rel_err(latent[k], golden['latent'][k]) < 1e-5
    for (h, c), (gh, gc) in zip(states, golden['states']):
        assert rel_err(h, gh) < 1e-5 and rel_err(c, gc) < 1e-5
    lat = {k: v for k, v in golden['latent'].items()}
    pred = O.semseg_forward(golden['semseg_sd'], lat)
    for k in (1, 2, 4):
        assert rel_err(pred[k], golden['pred'][k]) < 1e-5
    loss = O.task_loss(golden['pred'][1], golden['labels'], d['K'])
    assert abs(float(loss) - float(golden['loss'])) < 1e-5
    conf = O.confusion_matrix(golden['pred'][1].argmax(1), golden['labels'], d['K'], 255)
    assert torch.equal(conf, golden['confusion'])
    miou, _, acc = O.confusion_to_iou(conf)
    assert abs(float(miou) - float(golden['mean_iou'])) < 1e-9 and abs(float(acc) - float(golden['acc'])) < 1e-9


def test_oracle_gradients_match_golden(golden):
    d = golden['dims']
    params = {k: v.clone().requires_grad_(True) for k, v in golden['semseg_sd'].items()}
    pred = O.semseg_forward(params, golden['latent'])
    loss = O.task_loss(pred[1], golden['labels'], d['K'])
    grads = torch.autograd.grad(loss, list(params.values()))
    for (n, g) in zip(params.keys(), grads):
        ref = golden['grads'][n]
        # biases that feed an InstanceNorm have an exactly-zero true gradient: pure rounding noise
        # (|g| ~ 1e-6), hence the absolute term
        assert (g - ref).abs().max() <= 1e-4 * ref.abs().max() + 5e-6, n


def test_edge_cases_event_normalize():
    z = torch.zeros(2, 3, 8, 8)
    assert torch.equal(O.event_normalize(z), z)           # no non-zeros: unchanged (inference_utils.py:100)
    x = torch.zeros(1, 1, 4, 4)
    x[0, 0, 1, 1], x[0, 0, 2, 2] = 2.0, 4.0
    y = O.event_normalize(x)
    assert y[0, 0, 0, 0] == 0 and abs(float(y[0, 0, 1, 1]) + 1) < 1e-6 and abs(float(y[0, 0, 2, 2]) - 1) < 1e-6
    assert O.crop_padding(200, 346, 3) == (3, 3, 0, 0)    # DDD17 raw width 346 -> 352
    assert O.crop_padding(440, 640, 3) == (0, 0, 0, 0)


needs_ref = pytest.mark.skipif(not ref_shim.available(), reason='/root/reference not present')


@needs_ref
def test_oracle_vs_live_reference_lightweight():
    ref_shim.install()
    from e2vid.image_reconstructor import ImageReconstructor
    m = ref_shim.make_reference_e2vid()
    B, T, C, H, W = 1, 2, 5, 32, 40
    torch.manual_seed(0)
    data = torch.randn(B, T * C, H, W) * (torch.rand(B, T * C, H, W) < 0.2)
    rec = ImageReconstructor(m, H, W, C, 'cpu', ref_shim.e2vid_options())
    for i in range(T):
        img, st, lat = rec.update_reconstruction(data[:, i * C:(i + 1) * C])
    img2, st2, lat2 = O.encoder_unroll(m.state_dict(), ref_shim.E2VID_LIGHTWEIGHT_CFG, data, T, C)
    assert rel_err(img2, img) < 1e-6
    for k in lat:
        assert rel_err(lat2[k], lat[k]) < 1e-6


@needs_ref
@pytest.mark.parametrize('variant', ['convgru', 'upsample_conv', 'no_norm', 'reflect_pad', 'in_norm', 'in_norm_upsample'])
def test_oracle_vs_live_reference_variants(variant):
    ref_shim.install()
    from e2vid.image_reconstructor import ImageReconstructor
    cfg = dict(ref_shim.E2VID_LIGHTWEIGHT_CFG, base_num_channels=8, num_bins=3)
    H, W = 32, 40
    if variant == 'convgru':
        cfg['recurrent_block_type'] = 'convgru'
    elif variant == 'upsample_conv':
        cfg['use_upsample_conv'] = True
    elif variant == 'no_norm':
        cfg.pop('norm')
    elif variant.startswith('in_norm'):       # InstanceNorm2d(track_running_stats=True) in the conv layers, plain IN in the resblocks
        cfg['norm'] = 'IN'
        cfg['use_upsample_conv'] = variant.endswith('upsample')
    else:
        H, W = 30, 43
    m = ref_shim.make_reference_e2vid(cfg)
    torch.manual_seed(1)
    data = torch.randn(2, 6, H, W) * (torch.rand(2, 6, H, W) < 0.3)
    rec = ImageReconstructor(m, H, W, 3, 'cpu', ref_shim.e2vid_options())
    for i in range(2):
        img, st, lat = rec.update_reconstruction(data[:, i * 3:(i + 1) * 3])
    img2, st2, lat2 = O.encoder_unroll(m.state_dict(), cfg, data, 2, 3)
    assert rel_err(img2, img) < 1e-6
    assert rel_err(lat2[8], lat[8]) < 1e-6


@needs_ref
def test_oracle_vs_live_reference_semseg_variants():
    ref_shim.install()
    from models.style_networks import SemSegE2VID
    from utils.loss_functions import TaskLoss, symJSDivLoss
    torch.manual_seed(3)
    lat = {1: torch.randn(2, 8, 32, 32), 2: torch.randn(2, 16, 16, 16), 4: torch.randn(2, 32, 8, 8),
           8: torch.randn(2, 64, 4, 4)}
    for kw in (dict(skip_connect=True, skip_type='concat'), dict(skip_connect=False),
               dict(skip_connect=False, input_index_map=True)):
        dec = SemSegE2VID(input_c=64, output_c=5, **kw)
        out = dec(lat)
        out2 = O.semseg_forward(dec.state_dict(), lat, **kw)
        assert set(out.keys()) == set(out2.keys())
        for k in out:
            assert rel_err(out2[k], out[k]) < 1e-6
    lab = torch.randint(0, 5, (2, 32, 32))
    lab[0, :3] = 255
    for losses in (['dice', 'cross_entropy'], ['dice'], ['cross_entropy']):
        tl = TaskLoss(losses=losses, num_classes=5, ignore_index=255)
        assert abs(float(tl(out[1], lab)) - float(O.task_loss(out2[1], lab, 5, 255, losses))) < 1e-6
    a, b = torch.randn(2, 5, 8, 8), torch.randn(2, 5, 8, 8)
    assert abs(float(symJSDivLoss()(a, b)) - float(O.sym_js_div_loss(a, b))) < 1e-7


@needs_ref
def test_reference_trainer_runs_with_shim():
    """The real ESSSupervisedModel.train_step executes on CPU through the shim (config 1 plumbing)."""
    import os
    cfg = dict(ref_shim.E2VID_LIGHTWEIGHT_CFG, num_bins=1, base_num_channels=32)
    m = ref_shim.make_reference_e2vid(cfg)
    path = ref_shim.save_synthetic_checkpoint(m, cfg)
    try:
        t = ref_shim.make_supervised_trainer(ref_shim.supervised_settings(path, (64, 64), 11, 2, 1))
        torch.manual_seed(0)
        data = torch.randn(2, 2, 64, 64) * (torch.rand(2, 2, 64, 64) < 0.2)
        labels = torch.randint(0, 11, (2, 64, 64))
        labels[:, :3] = 255
        # oracle.supervised_step (the checker of __graft_entry__.smoke) vs the trainer's own first iteration:
        # same loss and the same gradient on every decoder parameter (ess_supervised_trainer.py:92-152)
        e2vid_sd = {k: v.detach().clone() for k, v in t.front_end_sensor_b.state_dict().items()}
        dec_sd = {k: v.detach().clone() for k, v in t.task_backend.state_dict().items()}
        first = float(t.train_step([data, labels])[2])
        loss_o, _, grads_o = O.supervised_step(e2vid_sd, cfg, dec_sd, data, labels, 2, 1, 11)
        assert abs(first - float(loss_o)) < 1e-5 * max(1.0, abs(first))
        for n, p_ in t.task_backend.named_parameters():
            assert (p_.grad - grads_o[n]).abs().max() <= 1e-4 * grads_o[n].abs().max() + 5e-6, n
        losses = [first] + [float(t.train_step([data, labels])[2]) for _ in range(2)]
        assert all(torch.isfinite(torch.tensor(losses)))
    finally:
        os.remove(path)


@needs_ref
def test_oracle_radam_vs_live_reference():
    ref_shim.install()
    from utils.radam import RAdam
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(7, 5))
    w2 = w.detach().clone()
    opt = RAdam([w], lr=5e-4, weight_decay=0., betas=(0., 0.999))
    state = {}
    for i in range(9):       # crosses the N_sma >= 5 switch
        g = torch.randn(7, 5)
        w.grad = g.clone()
        opt.step()
        O.radam_step(w2, g, state, 5e-4, (0., 0.999))
        assert torch.allclose(w2, w.detach(), rtol=1e-6, atol=1e-8), i


@needs_ref
def test_oracle_uda_step_vs_live_reference_trainer():
    """oracle.uda_step restates ESSModel.train_step (DSEC branch): same losses and the same gradients on the
    image encoder and on the decoder as the UNMODIFIED reference trainer run through the shim."""
    import os
    cfg = dict(ref_shim.E2VID_LIGHTWEIGHT_CFG, num_bins=2)
    m = ref_shim.make_reference_e2vid(cfg)
    path = ref_shim.save_synthetic_checkpoint(m, cfg)
    try:
        H, W, K, T, C, B = 32, 48, 5, 2, 2, 2
        t = ref_shim.make_uda_trainer(ref_shim.uda_settings(path, (H, W), K, T, C))
        enc_sd = {k: v.detach().clone() for k, v in t.front_end_sensor_a.state_dict().items()}
        dec_sd = {k: v.detach().clone() for k, v in t.task_backend.state_dict().items()}
        e2vid_sd = {k: v.detach().clone() for k, v in t.front_end_sensor_b.state_dict().items()}
        g = torch.Generator().manual_seed(0)
        img_a = torch.rand(B, 1, H, W, generator=g)
        labels_a = torch.randint(0, K, (B, H, W), generator=g)
        labels_a[:, :2] = 255
        data_b = torch.randn(B, T * C, H, W, generator=g) * (torch.rand(B, T * C, H, W, generator=g) < 0.3)
        labels_b = torch.randint(0, K, (B, H, W), generator=g)
        losses, _, final = t.train_step([[img_a, labels_a], [data_b, labels_b]])
        ol, g_enc, g_dec = O.uda_step(e2vid_sd, cfg, enc_sd, dec_sd, img_a, labels_a, data_b, T, C, K)
        assert abs(ol['task_img'] - float(losses['semseg_sensor_a_loss'])) < 1e-5
        assert abs(ol['task_img'] + ol['e_loss'] + ol['t_loss'] - float(final)) < 1e-4
        for n, p in t.front_end_sensor_a.named_parameters():
            assert p.grad is not None, n
            assert (p.grad - g_enc[n]).abs().max() <= 1e-4 * g_enc[n].abs().max() + 1e-7, n
        for n, p in t.task_backend.named_parameters():
            assert (p.grad - g_dec[n]).abs().max() <= 1e-4 * g_dec[n].abs().max() + 5e-6, n
    finally:
        os.remove(path)


@needs_ref
def test_oracle_voxel_grid_dsec_vs_live_reference():
    ref_shim.install()
    from DSEC.dataset.representations import VoxelGrid
    g = torch.Generator().manual_seed(0)
    n, C, H, W = 5000, 5, 24, 32
    x = torch.rand(n, generator=g) * (W + 1) - 0.7          # includes out-of-range and negative-fraction coordinates
    y = torch.rand(n, generator=g) * (H + 1) - 0.7
    pol = (torch.rand(n, generator=g) > 0.5).float()
    t = torch.sort(torch.rand(n, generator=g))[0] * 1e3
    ref = VoxelGrid(C, H, W, normalize=False).convert(x, y, pol, t)
    assert torch.allclose(O.voxel_grid_dsec(x, y, pol, t, C, H, W), ref, rtol=0, atol=1e-5)


@needs_ref
def test_oracle_flip_and_hot_pixels_vs_live_reference(tmp_path):
    """EventPreprocessor options (inference_utils.py:73-93): hot-pixel removal (in place) and the H/W flip."""
    import numpy as np
    ref_shim.install()
    from e2vid.image_reconstructor import ImageReconstructor
    if not hasattr(np, 'int'):
        np.int = int          # the reference predates numpy 1.24 (inference_utils.py:76 uses np.int)
    cfg = dict(ref_shim.E2VID_LIGHTWEIGHT_CFG, base_num_channels=8, num_bins=3)
    m = ref_shim.make_reference_e2vid(cfg)
    B, C, H, W = 2, 3, 24, 30            # W = 30 -> reflect pad to 32: the flip must precede the padding
    hot = [(3, 5), (29, 0), (10, 23)]
    f = tmp_path / 'hot.txt'
    f.write_text('\n'.join('%d,%d' % xy for xy in hot))
    torch.manual_seed(1)
    data = torch.randn(B, C, H, W) * (torch.rand(B, C, H, W) < 0.3)
    for x, y in hot:
        data[:, :, y, x] = 7.0           # make sure the hot pixels matter
    rec = ImageReconstructor(m, H, W, C, 'cpu', ref_shim.e2vid_options(flip=True, hot_pixels_file=str(f)))
    d1, d2 = data.clone(), data.clone()
    img, st, lat = rec.update_reconstruction(d1)
    img2, st2, lat2 = O.reconstructor_step(m.state_dict(), cfg, d2, None, hot_pixels=hot, flip=True)
    assert torch.equal(d1, d2) and float(d2[0, 0, 5, 3]) == 0.0     # both zero the caller's tensor in place
    assert rel_err(img2, img) < 1e-6
    for k in lat:
        assert rel_err(lat2[k], lat[k]) < 1e-6


@needs_ref
@pytest.mark.parametrize('separate_pol', [True, False])
def test_oracle_voxel_grid_ddd17_vs_live_reference(separate_pol):
    """generate_voxel_grid (datasets/data_util.py:54-126, bilinear in time) run from the unmodified reference (with
    the `np.int` alias numpy >= 1.24 dropped) vs the oracle restatement, incl. out-of-range pixels and polarity 0."""
    import numpy as np
    ref_shim.install()
    if not hasattr(np, 'int'):
        np.int = int
    from datasets.data_util import generate_voxel_grid
    rng = np.random.RandomState(0)
    n, C, H, W = 4000, 5, 20, 28
    ev = np.zeros((n, 4), np.float64)
    ev[:, 0] = rng.randint(-2, W + 2, n)                      # x, some outside the image
    ev[:, 1] = rng.randint(-2, H + 2, n)                      # y
    ev[:, 2] = np.sort(rng.rand(n)) * 1e4 + 17.0              # t, increasing
    ev[:, 3] = rng.randint(0, 2, n)                           # polarity in {0, 1}
    ref = generate_voxel_grid(ev.copy(), (H, W), C, separate_pol=separate_pol)
    out = O.voxel_grid_ddd17(ev.copy(), (H, W), C, separate_pol=separate_pol)
    assert out.shape == ref.shape
    assert np.allclose(out.numpy(), ref, rtol=0, atol=1e-5)
