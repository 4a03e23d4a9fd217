"""Shared helpers for the parity tests (seeded synthetic inputs of SURVEY.md s8d, error metrics)."""
import torch

from oracle import ess_oracle as O

E2VID_CFG = dict(num_bins=5, skip_type='sum', recurrent_block_type='convlstm', num_encoders=3, base_num_channels=32,
                 num_residual_blocks=2, norm='BN', use_upsample_conv=False)


def rel_err(a, b):
    """max-norm relative error |a-b|_inf / |b|_inf (the criterion of SURVEY.md s7.3)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def make_events(B, T, C, H, W, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, T * C, H, W, generator=g) * (torch.rand(B, T * C, H, W, generator=g) < 0.2)


def make_labels(B, H, W, K, seed=99):
    g = torch.Generator().manual_seed(seed)
    lab = torch.randint(0, K, (B, H, W), generator=g)
    lab[:, :5] = 255
    return lab


def randomize_bn_(module, seed=7):
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            with torch.no_grad():
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
        elif isinstance(m, torch.nn.InstanceNorm2d) and m.track_running_stats:      # norm='IN' conv layers
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)


def make_e2vid(cfg=None, seed=6, mode='fp32'):
    """Our drop-in module with seeded random weights and non-trivial BN statistics (CPU tensors)."""
    import ess_b200
    torch.manual_seed(seed)
    m = ess_b200.E2VIDRecurrent(dict(cfg or E2VID_CFG), mode=mode)
    randomize_bn_(m)
    return m.eval()


def make_semseg(K=11, seed=6, input_c=256, **kw):
    import ess_b200
    torch.manual_seed(seed)
    kw.setdefault('skip_connect', True)
    kw.setdefault('skip_type', 'concat')
    return ess_b200.SemSegE2VID(input_c, K, **kw)


def sd_cpu(module, dtype=torch.float32):
    return {k: v.detach().cpu().to(dtype) if v.is_floating_point() else v.detach().cpu()
            for k, v in module.state_dict().items()}


def make_latents(B, H, W, base=32, seed=5, device='cpu'):
    g = torch.Generator().manual_seed(seed)
    return {1: torch.randn(B, base, H, W, generator=g).to(device),
            2: torch.randn(B, 2 * base, H // 2, W // 2, generator=g).to(device),
            4: torch.randn(B, 4 * base, H // 4, W // 4, generator=g).to(device),
            8: torch.randn(B, 8 * base, H // 8, W // 8, generator=g).to(device)}




def program_forward(dec, sd, lat, retain=False):
    """The decoder's static node program (ess_b200.semseg._build_program) interpreted with plain torch ops in the
    dtype of `sd` / `lat` -- test infrastructure for the teacher-forced parity tests.  tests/test_host_logic.py pins
    it to O.semseg_forward (the reference restatement), so the program's wiring (sources, upsampling, concat order,
    residuals) is checked on CPU; the GPU tests then use its per-node tensors and gradients as the teacher.
    Returns (outs dict {4,2,1}, T) with T = {tensor id: tensor} for every node (retain_grad'ed when retain)."""
    import torch.nn.functional as F
    from ess_b200.semseg import _Conv
    T = {0: lat[8], 1: lat.get(4), 2: lat.get(2)}

    def IN(t):
        return F.instance_norm(t, eps=1e-5)

    def up(t, ups):
        return t.repeat_interleave(2, 2).repeat_interleave(2, 3) if ups else t

    for nd in dec._nodes:
        if isinstance(nd, _Conv):
            parts = []
            for (sid, xf, ups) in nd.srcs:
                a = T[sid]
                if xf == 'nr':
                    a = torch.relu(IN(a))
                parts.append(up(a, ups))
            y = F.conv2d(torch.cat(parts, 1) if len(parts) > 1 else parts[0], sd[nd.w], sd[nd.b], padding=nd.k // 2)
        else:
            y = IN(T[nd.src])
            if nd.relu:
                y = torch.relu(y)
            if nd.res is not None:
                y = y + T[nd.res]
        if retain and y.requires_grad:
            y.retain_grad()
        T[nd.out] = y
    sz = lat[1].shape[3]
    return {sz // T[o].shape[3]: T[o] for o in dec._outs}, T


def seeded_state(template_sd, seed):
    """Deterministic weights for a module from its state_dict TEMPLATE (keys / shapes / dtypes) and a seed, without
    storing them: a CPU torch.Generator stream consumed in key order.  Shared by tests/golden/make_golden_tc.py (which
    loads them into the REFERENCE modules) and the tests (which load them into ours), so a fixture at the real channel
    widths (10.7 M + 6.7 M parameters) stays ~1 MB.  Conv weights U(-1, 1)/sqrt(fan_in) (N(0, 0.02) inside the decoder's
    IN blocks, style_networks.py:152-155), biases N(0, 0.1), BatchNorm affine/statistics non-trivial."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in template_sd.items():
        if not v.is_floating_point():
            out[k] = v.clone()
        elif k.endswith('running_var'):
            out[k] = torch.rand(v.shape, generator=g) + 0.5
        elif k.endswith('running_mean'):
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        elif v.dim() == 4:
            if k.startswith('decoder_scale_') and '.model.' in k:
                out[k] = torch.randn(v.shape, generator=g) * 0.02
            else:
                fan_in = v.shape[1] * v.shape[2] * v.shape[3]
                out[k] = (torch.rand(v.shape, generator=g) * 2 - 1) / fan_in ** 0.5
        elif 'norm_layer.weight' in k or '.bn1.weight' in k or '.bn2.weight' in k:
            out[k] = torch.rand(v.shape, generator=g) + 0.5
        else:
            out[k] = torch.randn(v.shape, generator=g) * 0.1
    return out


def sample_indices(numel, n, seed):
    """n seeded positions of a flattened tensor (the gradient samples kept by the compact fixture)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, numel, (min(n, numel),), generator=g)


__all__ = ['O', 'rel_err', 'make_events', 'make_labels', 'make_e2vid', 'make_semseg', 'sd_cpu', 'make_latents',
           'E2VID_CFG', 'randomize_bn_', 'program_forward', 'seeded_state', 'sample_indices']
