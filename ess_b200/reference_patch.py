"""One-call wiring of the drop-ins into an importable reference tree (uzh-rpg/ess on `sys.path`).

    import ess_b200.reference_patch as rp
    rp.install()            # before the trainer object is constructed

Rebinds the names the reference trainers resolve at call time; no reference file is
edited.  `uninstall()` restores the originals."""
import importlib

_saved = []


def _swap(module, name, new):
    _saved.append((module, name, getattr(module, name, None)))
    setattr(module, name, new)


def install(reconstructor=True, uda=True):
    import ess_b200
    import ess_b200.optim
    if _saved:
        return
    lu = importlib.import_module('e2vid.utils.loading_utils')
    _swap(lu, 'E2VIDRecurrent', ess_b200.E2VIDRecurrent)            # eval(arch) at loading_utils.py:16
    trainers = [importlib.import_module('training.ess_supervised_trainer')]
    if uda:
        trainers.append(importlib.import_module('training.ess_trainer'))
    for mod in trainers:
        _swap(mod, 'SemSegE2VID', ess_b200.SemSegE2VID)
        _swap(mod, 'TaskLoss', ess_b200.TaskLoss)
        _swap(mod, 'MetricsSemseg', ess_b200.MetricsSemseg)
        if reconstructor:
            _swap(mod, 'ImageReconstructor', ess_b200.ImageReconstructor)
    if uda:
        # NB: like the reference, ess_b200.StyleEncoderE2VID starts from ImageNet resnet18 weights (hub cache or
        # download) and warns loudly when they cannot be obtained -- see ess_b200/style_encoder.py
        _swap(trainers[-1], 'StyleEncoderE2VID', ess_b200.StyleEncoderE2VID)
        _swap(trainers[-1], 'symJSDivLoss', ess_b200.symJSDivLoss)
    radam = importlib.import_module('utils.radam')
    _swap(radam, 'RAdam', ess_b200.optim.RAdam)


def uninstall():
    while _saved:
        module, name, old = _saved.pop()
        if old is None:
            delattr(module, name)
        else:
            setattr(module, name, old)
