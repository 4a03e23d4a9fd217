"""Import shim for running the UNMODIFIED reference (/root/reference) as a parity oracle.

TEST INFRASTRUCTURE ONLY (see oracle/ess_oracle.py header).  /root/reference exists only in the build
container; everything here degrades to `available() == False` elsewhere (e.g. on the GPU box), and
callers skip.  Implements the six-point shim of SURVEY.md s8c:
  1. stub modules absent from the image (albumentations, tensorboardX, matplotlib[.pyplot]);
  2. resolve the `datasets` name clash with HuggingFace `datasets`;
  3. replace CudaTimer by a no-op (it needs a CUDA driver and injects 5 device syncs per window);
  4. synthetic checkpoint for the untouched `load_model`;
  5. `resnet18(pretrained=True)` -> `weights=None` (no network);
  6. trainers built with object.__new__ (bypasses BaseTrainer.__init__, which needs real datasets).
"""
import contextlib
import os
import sys
import tempfile
import types

REF_ROOT = os.environ.get('ESS_REFERENCE_ROOT', '/root/reference')
_installed = False


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'e2vid'))


class _NullTimer(contextlib.ContextDecorator):
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    """Make `import e2vid...`, `import models...`, `import utils...`, `import training...` resolve
    to the reference tree.  Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError('reference tree not present at %s' % REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Anything()

        def __getattr__(self, k):
            return _Anything()

    def _modattr(value_fn):
        def _ga(k):
            if k.startswith('__'):
                raise AttributeError(k)
            return value_fn()
        return _ga

    # (1) absent third-party modules
    for mod in ('albumentations', 'tensorboardX', 'hdf5plugin'):
        try:
            __import__(mod)
        except Exception:
            sys.modules.pop(mod, None)
            _stub(mod, __getattr__=_modattr(lambda: _Anything))
    try:
        import matplotlib  # noqa: F401
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        cm = types.SimpleNamespace(Blues=None)
        plt = _stub('matplotlib.pyplot', cm=cm, __getattr__=_modattr(_Anything))
        _stub('matplotlib', pyplot=plt, cm=cm, use=lambda *a, **k: None)
        _stub('matplotlib.cm', Blues=None)
    try:
        import h5py  # noqa: F401
    except Exception:
        _stub('h5py', File=_Anything, __getattr__=_modattr(lambda: _Anything))

    # (2) `datasets`: the reference directory has no __init__.py and loses to site-packages
    ds = types.ModuleType('datasets')
    ds.__path__ = [os.path.join(REF_ROOT, 'datasets')]
    sys.modules['datasets'] = ds
    # same for the generic names `utils`, `models`, `training`, `evaluation`, `config`
    for name in ('utils', 'models', 'training', 'evaluation', 'config', 'e2vid', 'DSEC'):
        cur = sys.modules.get(name)
        if cur is not None and not str(getattr(cur, '__file__', '') or '').startswith(REF_ROOT) \
                and REF_ROOT not in list(getattr(cur, '__path__', [])):
            del sys.modules[name]

    # (3) CudaTimer -> no-op
    import e2vid.utils.timers as timers
    timers.CudaTimer = _NullTimer
    import e2vid.utils.inference_utils as iu
    iu.CudaTimer = _NullTimer
    import e2vid.image_reconstructor as ir
    ir.CudaTimer = _NullTimer

    # (5) no network for torchvision weights
    import torchvision.models as tvm
    _orig = tvm.resnet18

    def _resnet18(pretrained=False, **kw):
        kw.pop('weights', None)
        return _orig(weights=None, **kw)

    tvm.resnet18 = _resnet18
    _installed = True


E2VID_LIGHTWEIGHT_CFG = dict(num_bins=5, skip_type='sum', recurrent_block_type='convlstm', num_encoders=3,
                             base_num_channels=32, num_residual_blocks=2, norm='BN', use_upsample_conv=False)


def randomize_bn_(module, generator=None):
    """Non-trivial eval-mode BN: running stats and affine drawn at random."""
    import torch
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=generator) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=generator) + 0.5)
            with torch.no_grad():
                m.weight.copy_(torch.rand(m.weight.shape, generator=generator) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=generator) * 0.1)
        elif isinstance(m, torch.nn.InstanceNorm2d) and m.track_running_stats:      # norm='IN' conv layers
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=generator) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=generator) + 0.5)


def make_reference_e2vid(cfg=None, seed=6):
    """Instantiate the reference E2VIDRecurrent under the reference's seed (train.py:17)."""
    install()
    import torch
    from e2vid.model.model import E2VIDRecurrent
    torch.manual_seed(seed)
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        m = E2VIDRecurrent(dict(cfg or E2VID_LIGHTWEIGHT_CFG))
    g = torch.Generator().manual_seed(seed + 1)
    randomize_bn_(m, g)
    return m.eval()


def make_reference_semseg(num_classes=11, seed=6, **kw):
    install()
    import torch
    from models.style_networks import SemSegE2VID
    torch.manual_seed(seed)
    kw.setdefault('skip_connect', True)
    kw.setdefault('skip_type', 'concat')
    return SemSegE2VID(input_c=256, output_c=num_classes, **kw)


def save_synthetic_checkpoint(model, cfg, path=None):
    """(4) `{'arch','model','state_dict'}` file consumed by the untouched loading_utils.load_model."""
    import torch
    if path is None:
        fd, path = tempfile.mkstemp(suffix='.pth.tar')
        os.close(fd)
    torch.save({'arch': 'E2VIDRecurrent', 'model': dict(cfg), 'state_dict': model.state_dict()}, path)
    return path


def e2vid_options(**over):
    """The argparse defaults of e2vid/options/inference_options.py:1-66 that the path reads."""
    d = dict(use_gpu=False, no_recurrent=False, color=False, auto_hdr=False, no_normalize=False,
             hot_pixels_file=None, flip=False, Imin=0.0, Imax=1.0, auto_hdr_median_filter_size=10,
             unsharp_mask_amount=0.3, unsharp_mask_sigma=1.0, bilateral_filter_sigma=0.0)
    d.update(over)
    return types.SimpleNamespace(**d)


def supervised_settings(ckpt_path, img_size, num_classes, T, C, device='cpu'):
    return types.SimpleNamespace(
        path_to_model=ckpt_path, img_size_b=list(img_size), nr_temporal_bins_b=C, gpu_device=device,
        e2vid_config=e2vid_options(), semseg_num_classes=num_classes, semseg_ignore_label=255,
        semseg_class_names=['c%d' % i for i in range(num_classes)], skip_connect_task=True,
        skip_connect_task_type='concat', lr_back=5e-4, task_loss=['dice', 'cross_entropy'],
        require_paired_data_train_b=False, require_paired_data_val_b=False, nr_events_data_b=T,
        input_channels_b=C, weight_task_loss=1, semseg_color_map=None, dataset_name_b='DSEC_events')


def make_supervised_trainer(settings, device='cpu'):
    """(6) ESSSupervisedModel without BaseTrainer.__init__."""
    install()
    import io
    import torch
    from training.ess_supervised_trainer import ESSSupervisedModel
    t = object.__new__(ESSSupervisedModel)
    t.settings = settings
    t.device = torch.device(device)
    t.is_training = True
    t.step_count = 0
    t.epoch_count = 0
    t.do_val_training_epoch = False
    with contextlib.redirect_stdout(io.StringIO()):
        t.init_fn()
    t.visualize_epoch = lambda: False
    return t


def uda_settings(ckpt_path, img_size, num_classes, T, C, device='cpu'):
    """Fields of config/settings.py read by ESSModel.init_fn / train_step (DSEC branch, settings_DSEC.yaml)."""
    s = supervised_settings(ckpt_path, img_size, num_classes, T, C, device)
    s.dataset_name_b = 'DSEC_events'
    s.input_channels_a = 1
    s.skip_connect_encoder = True
    s.lr_front = 5e-4
    s.weight_KL_loss = 1.0
    s.weight_cycle_loss = 1.0
    s.weight_cycle_task_loss = 1.0
    s.require_paired_data_train_a = False
    s.train_on_event_labels = False
    s.semseg_label_val_b = True
    return s


def make_uda_trainer(settings, device='cpu'):
    """ESSModel (training/ess_trainer.py) without BaseTrainer.__init__."""
    install()
    import io
    import torch
    from training.ess_trainer import ESSModel
    t = object.__new__(ESSModel)
    t.settings = settings
    t.device = torch.device(device)
    t.is_training = True
    t.step_count = 0
    t.epoch_count = 0
    t.do_val_training_epoch = False
    with contextlib.redirect_stdout(io.StringIO()):
        t.init_fn()
    t.visualize_epoch = lambda: False
    return t
