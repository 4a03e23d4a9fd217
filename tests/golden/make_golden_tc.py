"""Generates tests/golden/ess_tc_small.pt by running the UNMODIFIED reference (/root/reference) on CPU at the REAL
channel widths (E2VID-lightweight base 32, decoder input 256) -- the configuration whose layers run on the tcgen05
kernels -- at a small spatial size.

Run in the build container only:  python tests/golden/make_golden_tc.py
The 17.4 M weights are not stored: both this script and the tests draw them from helpers.seeded_state (a seeded CPU
generator stream over the state_dict template).  Stored: inputs, the reference's latents / image / logits / loss,
and for each of the 34 decoder gradients its L2 norm plus 4096 seeded samples (helpers.sample_indices).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim  # noqa: E402
from helpers import sample_indices, seeded_state  # noqa: E402

CFG = dict(ref_shim.E2VID_LIGHTWEIGHT_CFG)
B, T, H, W, K = 1, 2, 32, 48, 6
E2VID_SEED, SEMSEG_SEED, NSAMP = 101, 202, 4096


def main():
    ref_shim.install()
    from e2vid.image_reconstructor import ImageReconstructor
    from utils.loss_functions import TaskLoss
    C = CFG['num_bins']
    model = ref_shim.make_reference_e2vid(CFG, seed=6)
    model.load_state_dict(seeded_state(model.state_dict(), E2VID_SEED))
    model.eval()
    g = torch.Generator().manual_seed(4321)
    data = torch.randn(B, T * C, H, W, generator=g) * (torch.rand(B, T * C, H, W, generator=g) < 0.2)
    labels = torch.randint(0, K, (B, H, W), generator=g)
    labels[:, :3] = 255
    rec = ImageReconstructor(model, H, W, C, 'cpu', ref_shim.e2vid_options())
    for i in range(T):
        img, states, latent = rec.update_reconstruction(data[:, i * C:(i + 1) * C])
    dec = ref_shim.make_reference_semseg(K)
    dec.load_state_dict(seeded_state(dec.state_dict(), SEMSEG_SEED))
    pred = dec({k: v.detach() for k, v in latent.items()})
    crit = TaskLoss(losses=['dice', 'cross_entropy'], gamma=2.0, num_classes=K, ignore_index=255, reduction='mean')
    loss = crit(pred[1], labels)
    loss.backward()
    grads = {}
    for j, (n, p) in enumerate(dec.named_parameters()):
        idx = sample_indices(p.numel(), NSAMP, 1000 + j)
        grads[n] = dict(norm=float(p.grad.double().norm()), samples=p.grad.flatten()[idx].clone())
    out = dict(cfg=CFG, dims=dict(B=B, T=T, C=C, H=H, W=W, K=K), e2vid_seed=E2VID_SEED, semseg_seed=SEMSEG_SEED,
               nsamp=NSAMP, data=data, labels=labels, img=img.clone(),
               latent={k: v.clone() for k, v in latent.items()},
               c2=states[2][1].clone(),
               pred={k: v.detach().clone() for k, v in pred.items() if k != 8},
               loss=loss.detach().clone(), grads=grads,
               weight_checksums={k: float(v.double().sum()) for k, v in list(model.state_dict().items())[:4]},
               torch_version=torch.__version__)
    path = os.path.join(HERE, 'ess_tc_small.pt')
    torch.save(out, path)
    print('wrote', path, os.path.getsize(path), 'bytes; loss', float(loss))


if __name__ == '__main__':
    main()
