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


__all__ = ['O', 'rel_err', 'make_events', 'make_labels', 'make_e2vid', 'make_semseg', 'sd_cpu', 'make_latents',
           'E2VID_CFG', 'randomize_bn_']
