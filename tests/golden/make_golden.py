"""Generates tests/golden/ess_tiny.pt by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only:  python tests/golden/make_golden.py
The fixture pins the oracle (oracle/ess_oracle.py) and the CUDA path to the reference's own outputs
at a tiny configuration (the reference ships no golden vectors of its own).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

CFG = dict(num_bins=2, skip_type='sum', recurrent_block_type='convlstm', num_encoders=3, base_num_channels=4,
           num_residual_blocks=2, norm='BN', use_upsample_conv=False)
B, T, H, W, K = 2, 2, 32, 48, 6


def main():
    ref_shim.install()
    from e2vid.image_reconstructor import ImageReconstructor
    from evaluation.metrics import MetricsSemseg
    from utils.loss_functions import TaskLoss
    torch.manual_seed(1234)
    C = CFG['num_bins']
    model = ref_shim.make_reference_e2vid(CFG, seed=6)
    g = torch.Generator().manual_seed(1234)
    data = torch.randn(B, T * C, H, W, generator=g) * (torch.rand(B, T * C, H, W, generator=g) < 0.2)
    labels = torch.randint(0, K, (B, H, W), generator=g)
    labels[:, :5] = 255
    rec = ImageReconstructor(model, H, W, C, 'cpu', ref_shim.e2vid_options())
    for i in range(T):
        img, states, latent = rec.update_reconstruction(data[:, i * C:(i + 1) * C])
    from models.style_networks import SemSegE2VID
    torch.manual_seed(6)
    dec = SemSegE2VID(input_c=CFG['base_num_channels'] * 8, output_c=K, skip_connect=True, skip_type='concat')
    latent_d = {k: v.detach() for k, v in latent.items()}
    pred = dec(latent_d)
    crit = TaskLoss(losses=['dice', 'cross_entropy'], gamma=2.0, num_classes=K, ignore_index=255, reduction='mean')
    loss = crit(pred[1], labels)
    loss.backward()
    metrics = MetricsSemseg(K, 255, ['c%d' % i for i in range(K)])
    metrics.update_batch(pred[1].argmax(1), labels)
    summ = metrics.get_metrics_summary()
    out = dict(
        cfg=CFG, dims=dict(B=B, T=T, C=C, H=H, W=W, K=K),
        e2vid_sd={k: v.clone() for k, v in model.state_dict().items()},
        semseg_sd={k: v.detach().clone() for k, v in dec.state_dict().items()},
        data=data, labels=labels,
        img=img, latent={k: v.clone() for k, v in latent.items()},
        states=[(h.clone(), c.clone()) for (h, c) in states],
        pred={k: v.detach().clone() for k, v in pred.items() if k != 8},
        loss=loss.detach().clone(),
        grads={n: p.grad.clone() for n, p in dec.named_parameters()},
        confusion=summ['cm'].clone(), mean_iou=summ['mean_iou'].clone(), acc=summ['acc'].clone(),
        torch_version=torch.__version__,
    )
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ess_tiny.pt')
    torch.save(out, path)
    print('wrote', path, os.path.getsize(path), 'bytes; loss', float(loss), 'mIoU', float(summ['mean_iou']))


if __name__ == '__main__':
    main()
