"""The drop-ins wired into the UNMODIFIED reference code (build container only: needs /root/reference).
No GPU here, so this checks everything up to the first kernel launch: `load_model` (eval(arch) + strict
load_state_dict) builds OUR encoder from a reference checkpoint, the reference trainers construct OUR decoder /
loss / optimizer / reconstructor, and a training step then fails LOUDLY on CPU instead of silently falling back."""
import os

import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason='/root/reference not present')


@pytest.fixture()
def patched():
    ref_shim.install()
    import ess_b200.reference_patch as rp
    rp.install()
    yield rp
    rp.uninstall()


def test_load_model_builds_our_encoder_from_a_reference_checkpoint(patched):
    import ess_b200
    ref_model = None
    patched.uninstall()
    ref_model = ref_shim.make_reference_e2vid()                    # reference class, reference init
    path = ref_shim.save_synthetic_checkpoint(ref_model, ref_shim.E2VID_LIGHTWEIGHT_CFG)
    patched.install()
    try:
        import contextlib
        import io
        from e2vid.utils.loading_utils import load_model
        with contextlib.redirect_stdout(io.StringIO()):
            model, decoder = load_model(path)
        assert isinstance(model, ess_b200.E2VIDRecurrent)
        assert model.num_bins == 5 and model.num_encoders == 3      # read by CropParameters (image_reconstructor.py:65)
        sd, ref_sd = model.state_dict(), ref_model.state_dict()
        assert list(sd.keys()) == list(ref_sd.keys())
        assert all(torch.equal(sd[k], ref_sd[k]) for k in sd)
    finally:
        os.remove(path)


def test_reference_trainers_construct_the_dropins_and_fail_loudly_without_cuda(patched):
    import ess_b200
    import ess_b200.optim
    cfg = dict(ref_shim.E2VID_LIGHTWEIGHT_CFG, num_bins=2)
    patched.uninstall()
    m = ref_shim.make_reference_e2vid(cfg)
    path = ref_shim.save_synthetic_checkpoint(m, cfg)
    patched.install()
    try:
        t = ref_shim.make_supervised_trainer(ref_shim.supervised_settings(path, (32, 48), 6, 2, 2))
        assert isinstance(t.front_end_sensor_b, ess_b200.E2VIDRecurrent)
        assert isinstance(t.task_backend, ess_b200.SemSegE2VID)
        assert isinstance(t.task_loss, ess_b200.TaskLoss)
        assert isinstance(t.reconstructor, ess_b200.ImageReconstructor)
        assert isinstance(t.optimizers_dict['optimizer_back'], ess_b200.optim.RAdam)
        assert not any(p.requires_grad for p in t.front_end_sensor_b.parameters())     # frozen by the trainer (:45-46)
        data = torch.randn(2, 4, 32, 48)
        labels = torch.randint(0, 6, (2, 32, 48))
        if not torch.cuda.is_available():
            with pytest.raises(RuntimeError, match='CUDA'):
                t.train_step([data, labels])
        u = ref_shim.make_uda_trainer(ref_shim.uda_settings(path, (32, 48), 6, 2, 2))
        assert isinstance(u.front_end_sensor_a, ess_b200.StyleEncoderE2VID)
        assert isinstance(u.cycle_pred_loss, ess_b200.symJSDivLoss)
        assert set(u.optimizers_dict) == {'optimizer_front_sensor_a', 'optimizer_back'}
    finally:
        os.remove(path)
