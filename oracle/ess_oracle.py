"""CPU oracle for the ESS hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain PyTorch-CPU (fp32 or fp64) restatement of the reference's algorithm for the path named in
BASELINE.json `north_star`: event normalisation -> reflect pad -> E2VID recurrent encoder (+ image
decoder) -> SemSegE2VID decoder -> CE+Dice task loss -> confusion-matrix mIoU.  Every function cites
the reference file:line (relative to /root/reference) it restates.  It works on raw `state_dict`s so
the same code runs in fp32 and fp64.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this module, and only as the checker / the timed CPU baseline.  Nothing under `ess_b200/`
imports it.

PINNING: the reference ships no tests, golden vectors or fixtures, so upstream
parity is unpinned.  This oracle is pinned instead against the *imported reference modules run in the
build container* (`tests/test_oracle.py`, skipped where /root/reference is absent) and
against committed golden tensors generated from the reference by `tests/golden/make_golden.py`.
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5   # torch.nn.BatchNorm2d default, e2vid/model/submodules.py:20
IN_EPS = 1e-5   # torch.nn.InstanceNorm2d default, models/style_networks.py:163


# ------------------------------------------------------------------------------------------------
# a1 / a2: event pre-processing (e2vid/utils/inference_utils.py:84-109, 311-338)
# ------------------------------------------------------------------------------------------------
def event_normalize(events, hot_pixels=(), flip=False):
    """EventPreprocessor.__call__ (default options: no hot pixels, no flip, normalise on).

    inference_utils.py:88-93: hot pixels (x, y) are zeroed IN PLACE in the caller's tensor, then the tensor is
    flipped over H and W.  inference_utils.py:96-107: mean/std over the non-zero entries of the WHOLE
    [B,C,H,W] tensor, zeros stay zero.  If there is no non-zero entry the tensor is returned unchanged."""
    for x, y in hot_pixels:
        events[:, :, y, x] = 0
    if flip:
        events = torch.flip(events, dims=[2, 3])
    nonzero = events != 0
    nnz = nonzero.sum()
    if nnz > 0:
        mean = events.sum() / nnz
        std = torch.sqrt((events ** 2).sum() / nnz - mean ** 2)
        events = nonzero.to(events.dtype) * (events - mean) / std
    return events


def crop_padding(height, width, num_encoders):
    """CropParameters.__init__ (inference_utils.py:311-330): (left, right, top, bottom)."""
    f = 2 ** num_encoders
    hc = int(f * math.ceil(height / f))
    wc = int(f * math.ceil(width / f))
    top = math.ceil(0.5 * (hc - height))
    bottom = math.floor(0.5 * (hc - height))
    left = math.ceil(0.5 * (wc - width))
    right = math.floor(0.5 * (wc - width))
    return left, right, top, bottom


def reflect_pad(events, num_encoders=3):
    """CropParameters.pad = ReflectionPad2d (inference_utils.py:330)."""
    h, w = events.shape[-2:]
    pads = crop_padding(h, w, num_encoders)
    if any(pads):
        events = F.pad(events, pads, mode='reflect')
    return events


# ------------------------------------------------------------------------------------------------
# a4-a10: E2VIDRecurrent (e2vid/model/{model,unet,submodules}.py)
# ------------------------------------------------------------------------------------------------
def _bn_eval(x, sd, prefix):
    """nn.BatchNorm2d in eval mode (running statistics), submodules.py:19-20,26-27."""
    return F.batch_norm(x, sd[prefix + '.running_mean'], sd[prefix + '.running_var'],
                        sd[prefix + '.weight'], sd[prefix + '.bias'], training=False, eps=BN_EPS)


def _in_eval(x, sd, prefix):
    """nn.InstanceNorm2d(C, track_running_stats=True) in eval mode (submodules.py:21-22,50-51,80-81): F.instance_norm with
    use_input_stats=False, i.e. the RUNNING statistics, no affine parameters."""
    rm, rv = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
    return (x - rm.view(1, -1, 1, 1)) / torch.sqrt(rv.view(1, -1, 1, 1) + BN_EPS)


def conv_layer(x, sd, prefix, stride, padding, norm, activation='relu'):
    """ConvLayer.forward (submodules.py:24-31): conv (+bias unless BN) -> norm -> activation."""
    out = F.conv2d(x, sd[prefix + '.conv2d.weight'], sd.get(prefix + '.conv2d.bias'),
                   stride=stride, padding=padding)
    if norm == 'BN':
        out = _bn_eval(out, sd, prefix + '.norm_layer')
    elif norm == 'IN':
        out = _in_eval(out, sd, prefix + '.norm_layer')
    elif norm is not None:
        raise NotImplementedError('norm=%r' % (norm,))
    if activation == 'relu':
        out = torch.relu(out)
    return out


def convlstm(x, prev_state, sd, prefix):
    """ConvLSTM.forward (submodules.py:190-230). Gate order: in, remember, out, cell (:216)."""
    w, b = sd[prefix + '.Gates.weight'], sd[prefix + '.Gates.bias']
    hidden = w.shape[0] // 4
    if prev_state is None:
        z = x.new_zeros((x.shape[0], hidden) + tuple(x.shape[2:]))
        prev_state = (z, z)
    prev_h, prev_c = prev_state
    gates = F.conv2d(torch.cat((x, prev_h), 1), w, b, padding=w.shape[-1] // 2)
    i, f, o, g = gates.chunk(4, 1)
    i, f, o, g = torch.sigmoid(i), torch.sigmoid(f), torch.sigmoid(o), torch.tanh(g)
    cell = f * prev_c + i * g
    hid = o * torch.tanh(cell)
    return hid, cell


def convgru(x, prev_state, sd, prefix):
    """ConvGRU.forward (submodules.py:255-273)."""
    wu, bu = sd[prefix + '.update_gate.weight'], sd[prefix + '.update_gate.bias']
    wr, br = sd[prefix + '.reset_gate.weight'], sd[prefix + '.reset_gate.bias']
    wo, bo = sd[prefix + '.out_gate.weight'], sd[prefix + '.out_gate.bias']
    hidden = wu.shape[0]
    pad = wu.shape[-1] // 2
    if prev_state is None:
        prev_state = x.new_zeros((x.shape[0], hidden) + tuple(x.shape[2:]))
    stacked = torch.cat([x, prev_state], 1)
    update = torch.sigmoid(F.conv2d(stacked, wu, bu, padding=pad))
    reset = torch.sigmoid(F.conv2d(stacked, wr, br, padding=pad))
    out = torch.tanh(F.conv2d(torch.cat([x, prev_state * reset], 1), wo, bo, padding=pad))
    return prev_state * (1 - update) + out * update


def residual_block(x, sd, prefix, norm):
    """ResidualBlock.forward (submodules.py:157-172)."""
    out = F.conv2d(x, sd[prefix + '.conv1.weight'], sd.get(prefix + '.conv1.bias'), padding=1)
    if norm == 'BN':
        out = _bn_eval(out, sd, prefix + '.bn1')
    elif norm == 'IN':                        # nn.InstanceNorm2d(C) WITHOUT running statistics (submodules.py:149-151):
        out = F.instance_norm(out, eps=BN_EPS)  # per-sample statistics in train and eval mode alike
    out = torch.relu(out)
    out = F.conv2d(out, sd[prefix + '.conv2.weight'], sd.get(prefix + '.conv2.bias'), padding=1)
    if norm == 'BN':
        out = _bn_eval(out, sd, prefix + '.bn2')
    elif norm == 'IN':
        out = F.instance_norm(out, eps=BN_EPS)
    return torch.relu(out + x)


def upsample_layer(x, sd, prefix, norm, use_upsample_conv):
    """TransposedConvLayer.forward (submodules.py:53-62) / UpsampleConvLayer.forward (:83-93)."""
    if use_upsample_conv:
        x = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)
        out = F.conv2d(x, sd[prefix + '.conv2d.weight'], sd.get(prefix + '.conv2d.bias'), padding=2)
    else:
        out = F.conv_transpose2d(x, sd[prefix + '.transposed_conv2d.weight'],
                                 sd.get(prefix + '.transposed_conv2d.bias'),
                                 stride=2, padding=2, output_padding=1)
    if norm == 'BN':
        out = _bn_eval(out, sd, prefix + '.norm_layer')
    elif norm == 'IN':
        out = _in_eval(out, sd, prefix + '.norm_layer')
    return torch.relu(out)


def e2vid_config_defaults(config):
    """BaseE2VID.__init__ / E2VIDRecurrent.__init__ defaults (e2vid/model/model.py:9-44,73-80)."""
    return dict(num_bins=int(config['num_bins']),
                skip_type=str(config.get('skip_type', 'sum')),
                num_encoders=int(config.get('num_encoders', 4)),
                base_num_channels=int(config.get('base_num_channels', 32)),
                num_residual_blocks=int(config.get('num_residual_blocks', 2)),
                norm=(str(config['norm']) if 'norm' in config else None),
                use_upsample_conv=bool(config.get('use_upsample_conv', True)),
                recurrent_block_type=str(config.get('recurrent_block_type', 'convlstm')))


def e2vid_recurrent_forward(sd, config, x, prev_states, with_image=True):
    """E2VIDRecurrent.forward -> UNetRecurrent.forward (model.py:93-100, unet.py:145-181).

    Returns (img, states, latent) with latent = {1: head, 2: blocks[0], 4: blocks[1], 8: blocks[2]}
    (unet.py:172; latent[8] is the last recurrent hidden state BEFORE the residual blocks)."""
    cfg = e2vid_config_defaults(config)
    norm, ne = cfg['norm'], cfg['num_encoders']
    skip = (lambda a, b: a + b) if cfg['skip_type'] == 'sum' else (lambda a, b: torch.cat([a, b], 1))
    p = 'unetrecurrent.'
    x = conv_layer(x, sd, p + 'head', 1, 2, None)                      # unet.py:131-132,153
    head = x
    if prev_states is None:
        prev_states = [None] * ne
    blocks, states = [], []
    for i in range(ne):                                               # unet.py:162-166
        x = conv_layer(x, sd, p + 'encoders.%d.conv' % i, 2, 2, norm)  # submodules.py:107,111
        rp = p + 'encoders.%d.recurrent_block' % i
        if cfg['recurrent_block_type'] == 'convlstm':
            state = convlstm(x, prev_states[i], sd, rp)
            x = state[0]
        else:
            state = convgru(x, prev_states[i], sd, rp)
            x = state
        blocks.append(x)
        states.append(state)
    latent = {1: head}
    for i in range(ne):
        latent[2 ** (i + 1)] = blocks[i]
    if ne != 3:
        latent = {1: head, 2: blocks[0], 4: blocks[1], 8: blocks[2]}   # unet.py:172 (fixed keys)
    if not with_image:
        return None, states, latent
    for j in range(cfg['num_residual_blocks']):                       # unet.py:169-170
        x = residual_block(x, sd, p + 'resblocks.%d' % j, norm)
    for i in range(ne):                                               # unet.py:175-176
        x = upsample_layer(skip(x, blocks[ne - i - 1]), sd, p + 'decoders.%d' % i, norm,
                           cfg['use_upsample_conv'])
    img = torch.sigmoid(conv_layer(skip(x, head), sd, p + 'pred', 1, 0, norm, activation=None))
    return img, states, latent


def reconstructor_step(sd, config, event_tensor, last_states, with_image=True, hot_pixels=(), flip=False):
    """ImageReconstructor.update_reconstruction (e2vid/image_reconstructor.py:82-163), default
    options: normalise -> reflect pad -> model -> carry states.  All under no_grad (:83)."""
    with torch.no_grad():
        ev = event_normalize(event_tensor, hot_pixels, flip)
        ev = reflect_pad(ev, e2vid_config_defaults(config)['num_encoders'])
        return e2vid_recurrent_forward(sd, config, ev, last_states, with_image)


def encoder_unroll(sd, config, data, num_windows, channels_per_window, with_image_last=True):
    """The trainer's T-loop (training/ess_supervised_trainer.py:126-130): state reset, then T windows
    of C channels each sliced from data[B, T*C, H, W]; returns the LAST (img, states, latent)."""
    states, img, latent = None, None, None
    for i in range(num_windows):
        ev = data[:, i * channels_per_window:(i + 1) * channels_per_window]
        need_img = with_image_last and i == num_windows - 1
        img, states, latent = reconstructor_step(sd, config, ev, states, with_image=need_img)
    return img, states, latent


# ------------------------------------------------------------------------------------------------
# a11-a14: SemSegE2VID (models/style_networks.py)
# ------------------------------------------------------------------------------------------------
def _relu_ins_conv(x, sd, prefix):
    """ReLUINSConv2d.forward (style_networks.py:158-169): conv(+bias) -> IN -> ReLU."""
    x = F.conv2d(x, sd[prefix + '.model.0.weight'], sd[prefix + '.model.0.bias'], padding=1)
    return torch.relu(F.instance_norm(x, eps=IN_EPS))


def _ins_res_block(x, sd, prefix):
    """INSResBlock.forward (style_networks.py:172-193): conv-IN-ReLU-conv-IN, += x, no last ReLU."""
    out = F.conv2d(x, sd[prefix + '.model.0.weight'], sd[prefix + '.model.0.bias'], padding=1)
    out = torch.relu(F.instance_norm(out, eps=IN_EPS))
    out = F.conv2d(out, sd[prefix + '.model.3.weight'], sd[prefix + '.model.3.bias'], padding=1)
    out = F.instance_norm(out, eps=IN_EPS)
    return out + x


def _up2(x):
    """f.interpolate(scale_factor=2, mode='nearest') (style_networks.py:77) ==
    InterpolationLayer expand/reshape (models/submodules.py:17-19)."""
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def semseg_forward(sd, input_dict, skip_connect=True, skip_type='concat', input_index_map=False):
    """SemSegE2VID.forward (style_networks.py:69-107). `sd` uses the module's state_dict keys."""
    sz_in = input_dict[1].shape[3]
    skip = (lambda a, b: a + b) if skip_type == 'sum' else (lambda a, b: torch.cat([a, b], 1))
    x = input_dict[8]
    out = {8: x}

    def put(t):
        assert sz_in % t.shape[3] == 0
        out[sz_in // t.shape[3]] = t

    if skip_connect:
        for i in range(5):
            x = _ins_res_block(x, sd, 'decoder_scale_1.%d' % i)
        x = _relu_ins_conv(x, sd, 'decoder_scale_1.5')
        x = skip(_up2(x), input_dict[4])
        x = _relu_ins_conv(x, sd, 'decoder_scale_2.0')
        x = _relu_ins_conv(x, sd, 'decoder_scale_2.1')
        put(x)
        x = skip(_up2(x), input_dict[2])
        x = _relu_ins_conv(x, sd, 'decoder_scale_3.0')
        x = _relu_ins_conv(x, sd, 'decoder_scale_3.1')
        put(x)
        x = _relu_ins_conv(_up2(x), sd, 'decoder_scale_4.0')
        x = F.conv2d(x, sd['decoder_scale_5.0.weight'], sd['decoder_scale_5.0.bias'])
        put(x)
    else:
        if input_index_map:                                            # style_networks.py:90-97
            xc = torch.arange(x.size(2), dtype=x.dtype)
            yc = torch.arange(x.size(3), dtype=x.dtype)
            coords = torch.stack(torch.meshgrid([xc, yc], indexing='ij'), 0)
            x = torch.cat([x, coords[None].repeat(x.size(0), 1, 1, 1)], 1)
        for i in range(3):
            x = _ins_res_block(x, sd, 'decoder_scale_1.%d' % i)
        x = _relu_ins_conv(_up2(x), sd, 'decoder_scale_2.1')
        put(x)
        x = _relu_ins_conv(_up2(x), sd, 'decoder_scale_3.1')
        put(x)
        x = _relu_ins_conv(_up2(x), sd, 'decoder_scale_4.1')
        x = F.conv2d(x, sd['decoder_scale_5.0.weight'], sd['decoder_scale_5.0.bias'])
        put(x)
    return out


# ------------------------------------------------------------------------------------------------
# a15: TaskLoss (utils/loss_functions.py)
# ------------------------------------------------------------------------------------------------
def dice_loss(predict, target, num_classes, ignore_index):
    """DiceLoss.forward + BinaryDiceLoss.forward (loss_functions.py:114-135, 80-90), smooth=1, p=2.
    Sums run over the whole batch and all pixels; class `ignore_index` (if < K) is skipped (:128)
    but the mean still divides by K (:135)."""
    mask = target != ignore_index
    tgt = target * mask
    onehot = torch.zeros((target.shape[0], num_classes) + tuple(target.shape[1:]), dtype=predict.dtype,
                         device=predict.device)
    onehot.scatter_(1, tgt.unsqueeze(1), 1)
    onehot = onehot * mask.unsqueeze(1)
    p = F.softmax(predict, dim=1) * mask.unsqueeze(1)
    total = 0
    for k in range(num_classes):
        if k != ignore_index:
            num = torch.sum(p[:, k] * onehot[:, k]) * 2 + 1
            den = torch.sum(p[:, k] ** 2 + onehot[:, k] ** 2) + 1
            total = total + (1 - num / den)
    return total / num_classes


def task_loss(predict, target, num_classes, ignore_index=255, losses=('dice', 'cross_entropy')):
    """TaskLoss.forward (loss_functions.py:17-24): Dice + CrossEntropyLoss(ignore_index) (:15)."""
    total = 0
    if 'dice' in losses:
        total = total + dice_loss(predict, target, num_classes, ignore_index)
    if 'cross_entropy' in losses:
        total = total + F.cross_entropy(predict, target, ignore_index=ignore_index)
    return total


def sym_js_div_loss(predict, target):
    """symJSDivLoss.forward (loss_functions.py:27-37). nn.KLDivLoss() default = element-wise mean."""
    ps = predict.softmax(dim=1).clamp(min=1e-10)
    ts = target.softmax(dim=1).clamp(min=1e-10)
    return 0.5 * F.kl_div(ps.log(), ts, reduction='mean') + 0.5 * F.kl_div(ts.log(), ps, reduction='mean')


# ------------------------------------------------------------------------------------------------
# a16: confusion matrix / mIoU (evaluation/metrics.py)
# ------------------------------------------------------------------------------------------------
def confusion_matrix(y_hat_lbl, y_lbl, num_classes, ignore_label):
    """semseg_compute_confusion (evaluation/metrics.py:4-24): conf[y, y_hat] via bincount."""
    mask = y_lbl != ignore_label
    x = y_hat_lbl[mask] + num_classes * y_lbl[mask]
    return torch.bincount(x.long(), minlength=num_classes ** 2).view(num_classes, num_classes).long()


def confusion_to_iou(conf):
    """semseg_accum_confusion_to_iou / _to_acc (evaluation/metrics.py:27-38)."""
    conf = conf.double()
    diag = conf.diag()
    iou = 100 * diag / (conf.sum(1) + conf.sum(0) - diag).clamp(min=1e-12)
    acc = 100 * diag.sum() / conf.sum().clamp(min=1e-12)
    return iou.mean(), iou, acc


# ------------------------------------------------------------------------------------------------
# whole supervised step (training/ess_supervised_trainer.py:92-152) used by the CPU baseline
# ------------------------------------------------------------------------------------------------
def supervised_step(e2vid_sd, e2vid_cfg, semseg_sd, data, labels, num_windows, channels, num_classes,
                    ignore_index=255, with_image_last=True):
    """Encoder unroll (no grad) -> SemSegE2VID fwd -> TaskLoss -> backward w.r.t. decoder params.
    Returns (loss, logits, grads dict).  `semseg_sd` tensors are treated as leaf parameters."""
    _, _, latent = encoder_unroll(e2vid_sd, e2vid_cfg, data, num_windows, channels, with_image_last)
    latent = {k: v.detach() for k, v in latent.items()}               # :145-146
    params = {k: v.detach().clone().requires_grad_(True) for k, v in semseg_sd.items()}
    pred = semseg_forward(params, latent)
    loss = task_loss(pred[1], labels, num_classes, ignore_index)
    grads = torch.autograd.grad(loss, list(params.values()))
    return loss.detach(), pred[1].detach(), dict(zip(params.keys(), grads))


# ------------------------------------------------------------------------------------------------
# optimizer: RAdam (utils/radam.py:15-80), restated functionally for one tensor
# ------------------------------------------------------------------------------------------------
def radam_step(p, grad, state, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """One RAdam update of tensor `p` in place; `state` = dict(step, exp_avg, exp_avg_sq) (radam.py:33-75)."""
    beta1, beta2 = betas
    if not state:
        state.update(step=0, exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p))
    state['exp_avg_sq'].mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    state['exp_avg'].mul_(beta1).add_(grad, alpha=1 - beta1)
    state['step'] += 1
    t = state['step']
    beta2_t = beta2 ** t
    n_max = 2 / (1 - beta2) - 1
    n_sma = n_max - 2 * t * beta2_t / (1 - beta2_t)
    if weight_decay != 0:
        p.add_(p, alpha=-weight_decay * lr)
    if n_sma >= 5:
        step_size = math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_max - 4) * (n_sma - 2) / n_sma * n_max /
                              (n_max - 2)) / (1 - beta1 ** t)
        p.addcdiv_(state['exp_avg'], state['exp_avg_sq'].sqrt().add_(eps), value=-step_size * lr)
    else:
        p.add_(state['exp_avg'], alpha=-1.0 / (1 - beta1 ** t) * lr)
    return p


# ------------------------------------------------------------------------------------------------
# UDA companion path: StyleEncoderE2VID + ESSModel.train_step (DSEC branch)
# ------------------------------------------------------------------------------------------------
def _bn(x, sd, prefix, training, stats_out=None):
    """nn.BatchNorm2d (torchvision ResNet), train mode = batch statistics (+ running-stat update
    returned through stats_out), eval mode = running statistics."""
    if training:
        if stats_out is not None:
            m = x.shape[0] * x.shape[2] * x.shape[3]
            mean = x.mean((0, 2, 3))
            var_u = x.var((0, 2, 3), unbiased=True) if m > 1 else x.var((0, 2, 3), unbiased=False)
            stats_out[prefix] = (0.9 * sd[prefix + '.running_mean'] + 0.1 * mean.detach(),
                                 0.9 * sd[prefix + '.running_var'] + 0.1 * var_u.detach())
        return F.batch_norm(x, None, None, sd[prefix + '.weight'], sd[prefix + '.bias'], training=True, eps=BN_EPS)
    return F.batch_norm(x, sd[prefix + '.running_mean'], sd[prefix + '.running_var'], sd[prefix + '.weight'],
                        sd[prefix + '.bias'], training=False, eps=BN_EPS)


def _basic_block(x, sd, prefix, stride, training, stats_out):
    """torchvision.models.resnet.BasicBlock.forward."""
    out = F.conv2d(x, sd[prefix + '.conv1.weight'], None, stride=stride, padding=1)
    out = torch.relu(_bn(out, sd, prefix + '.bn1', training, stats_out))
    out = F.conv2d(out, sd[prefix + '.conv2.weight'], None, stride=1, padding=1)
    out = _bn(out, sd, prefix + '.bn2', training, stats_out)
    identity = x
    if prefix + '.downsample.0.weight' in sd:
        identity = F.conv2d(x, sd[prefix + '.downsample.0.weight'], None, stride=stride)
        identity = _bn(identity, sd, prefix + '.downsample.1', training, stats_out)
    return torch.relu(out + identity)


def style_encoder_forward(sd, x, skip_connect=True, training=True, stats_out=None):
    """StyleEncoderE2VID.forward (models/style_networks.py:128-145): conv7x7 s2 -> bn1 -> relu -> layer1
    (encoder_scale_1), layer2 (encoder_scale_2), layer3 (encoder_scale_3); out keys by width ratio."""
    out = {1: x}
    sz_in = x.shape[3]

    def put(t):
        assert sz_in % t.shape[3] == 0
        out[sz_in // t.shape[3]] = t

    y = F.conv2d(x, sd['encoder_scale_1.0.weight'], None, stride=2, padding=3)
    y = torch.relu(_bn(y, sd, 'encoder_scale_1.1', training, stats_out))
    for b in range(2):
        y = _basic_block(y, sd, 'encoder_scale_1.3.%d' % b, 1, training, stats_out)
    if skip_connect:
        put(y)
    for b in range(2):
        y = _basic_block(y, sd, 'encoder_scale_2.%d' % b, 2 if b == 0 else 1, training, stats_out)
    if skip_connect:
        put(y)
    for b in range(2):
        y = _basic_block(y, sd, 'encoder_scale_3.%d' % b, 2 if b == 0 else 1, training, stats_out)
    put(y)
    return out


def uda_step(e2vid_sd, e2vid_cfg, enc_sd, dec_sd, img_a, labels_a, data_b, num_windows, channels, num_classes,
             ignore_index=255, w_task=1.0, w_kl=1.0, w_cycle=1.0, w_cycle_task=1.0):
    """ESSModel.train_step, DSEC branch, train_on_event_labels=False (training/ess_trainer.py:103-148 with
    img_train_step :150-180, event_train_step :257-301, trainCycleStep :211-255, TasktrainCycleStep :303-330).
    Returns (losses dict, grads of the image encoder dict, grads of the decoder dict)."""
    enc = {k: v.detach().clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in enc_sd.items()}
    dec = {k: v.detach().clone().requires_grad_(True) for k, v in dec_sd.items()}
    enc_params = [v for v in enc.values() if v.requires_grad]
    dec_params = list(dec.values())
    g_enc = [torch.zeros_like(p) for p in enc_params]
    g_dec = [torch.zeros_like(p) for p in dec_params]
    l1 = torch.nn.functional.l1_loss

    def accumulate(loss, params, grads):
        gs = torch.autograd.grad(loss, params, allow_unused=True, retain_graph=False)
        for a, g in zip(grads, gs):
            if g is not None:
                a += g

    # ---- images: encoder forward (train-mode BN), latents detached for DSEC (:188), decoder + task loss
    lat_fake = style_encoder_forward(enc, img_a, True, True)
    pred_a = semseg_forward(dec, {k: v.detach() for k, v in lat_fake.items()})
    t_img = task_loss(pred_a[1], labels_a, num_classes, ignore_index) * w_task
    accumulate(t_img, dec_params, g_dec)                                  # :120-125 (encoder params frozen)

    # ---- events: T windows through the frozen E2VID (no_grad, :277-280)
    img_fake, _, lat_real = encoder_unroll(e2vid_sd, e2vid_cfg, data_b, num_windows, channels, True)
    lat_real = {k: v.detach() for k, v in lat_real.items()}
    lat_fake = style_encoder_forward(enc, img_fake.detach(), True, True)  # :282
    # trainCycleStep (:211-255)
    e_loss = (l1(lat_fake[2], lat_real[2]) + l1(lat_fake[4], lat_real[4]) + l1(lat_fake[8], lat_real[8])) * w_cycle
    pred_second = semseg_forward(dec, lat_fake)
    with torch.no_grad():
        pred_first_ng = semseg_forward(dec, lat_real)
    e_loss = e_loss + sym_js_div_loss(pred_second[1], pred_first_ng[1])
    e_loss = e_loss + l1(pred_second[2], pred_first_ng[2]) * w_cycle_task
    e_loss = e_loss + l1(pred_second[4], pred_first_ng[4]) * w_cycle_task
    # TasktrainCycleStep (:303-330)
    pred_first = semseg_forward(dec, lat_real)
    with torch.no_grad():
        pred_second_ng = semseg_forward(dec, {k: v.detach() for k, v in lat_fake.items()})
    t_loss = sym_js_div_loss(pred_first[1], pred_second_ng[1]) * w_kl
    t_loss = t_loss + l1(pred_first[2], pred_second_ng[2]) * w_cycle_task
    t_loss = t_loss + l1(pred_first[4], pred_second_ng[4]) * w_cycle_task
    accumulate(e_loss, enc_params, g_enc)          # :133-137: back_end frozen for e_loss -> encoder grads only
    accumulate(t_loss, dec_params, g_dec)          # :138
    losses = dict(task_img=float(t_img), e_loss=float(e_loss), t_loss=float(t_loss))
    names_e = [k for k, v in enc.items() if v.requires_grad]
    return losses, dict(zip(names_e, g_enc)), dict(zip(dec.keys(), g_dec))


# ------------------------------------------------------------------------------------------------
# events -> voxel grid
# ------------------------------------------------------------------------------------------------
def voxel_grid_dsec(x, y, pol, time, channels, height, width):
    """VoxelGrid.convert with normalize=False (DSEC/dataset/representations.py:15-44)."""
    C, H, W = channels, height, width
    grid = torch.zeros((C, H, W), dtype=torch.float)
    t_norm = (C - 1) * (time - time[0]) / (time[-1] - time[0])
    x0, y0, t0 = x.int(), y.int(), t_norm.int()
    value = 2 * pol - 1
    for xlim in [x0, x0 + 1]:
        for ylim in [y0, y0 + 1]:
            for tlim in [t0, t0 + 1]:
                mask = (xlim < W) & (xlim >= 0) & (ylim < H) & (ylim >= 0) & (tlim >= 0) & (tlim < C)
                w = value * (1 - (xlim - x).abs()) * (1 - (ylim - y).abs()) * (1 - (tlim - t_norm).abs())
                index = H * W * tlim.long() + W * ylim.long() + xlim.long()
                grid.put_(index[mask], w[mask], accumulate=True)
    return grid


def voxel_grid_ddd17(events, shape, nr_temporal_bins, separate_pol=True):
    """generate_voxel_grid (datasets/data_util.py:54-126) restated with numpy; `events` = [N,4] float64 rows
    [x, y, t, polarity].  (The reference function uses `np.int`, removed in numpy >= 1.24; tests/test_oracle.py
    restores that alias and pins this restatement against the live reference function.)"""
    import numpy as np
    events = np.array(events, dtype=np.float64)
    height, width = shape
    n = nr_temporal_bins
    pos = np.zeros((n, height, width), np.float32).ravel()
    neg = np.zeros((n, height, width), np.float32).ravel()
    first, last = events[0, 2], events[-1, 2]
    dT = last - first
    if dT == 0:
        dT = 1.0
    xs, ys = events[:, 0].astype(np.int64), events[:, 1].astype(np.int64)
    ts = (n - 1) * (events[:, 2] - first) / dT
    pols = events[:, 3].copy()
    pols[pols == 0] = -1
    tis = ts.astype(np.int64)
    dts = ts - tis
    vl, vr = np.abs(pols) * (1.0 - dts), np.abs(pols) * dts
    valid = (xs < width) & (xs >= 0) & (ys < height) & (ys >= 0) & (ts >= 0) & (ts < n)
    for grid, sel in ((pos, pols == 1), (neg, pols != 1)):
        m = (tis < n) & sel & valid
        np.add.at(grid, xs[m] + ys[m] * width + tis[m] * width * height, vl[m])
        m = ((tis + 1) < n) & sel & valid
        np.add.at(grid, xs[m] + ys[m] * width + (tis[m] + 1) * width * height, vr[m])
    pos, neg = pos.reshape(n, height, width), neg.reshape(n, height, width)
    if separate_pol:
        return torch.from_numpy(np.concatenate([pos, neg], 0))
    return torch.from_numpy(pos - neg)
