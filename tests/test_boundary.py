"""The checkpoint contract of the drop-in modules, checked anywhere (CPU, no reference needed): ordered state_dict
keys, shapes and dtypes equal the reference modules' (fixture tests/golden/state_dict_keys.json, generated from the
unmodified reference by tests/golden/make_keys.py), so `load_model`'s strict `load_state_dict`
(e2vid/utils/loading_utils.py:19) and `utils/saver.py` checkpoints round-trip."""
import json
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'state_dict_keys.json')))


def listing(module):
    return [[k, list(v.shape), str(v.dtype)] for k, v in module.state_dict().items()]


@pytest.mark.parametrize('name', ['e2vid_lightweight_convlstm', 'e2vid_convgru', 'e2vid_upsample_conv_10bins',
                                  'e2vid_instance_norm'])
def test_e2vid_state_dict_contract(name):
    import ess_b200
    m = ess_b200.E2VIDRecurrent(dict(KEYS[name]['cfg']), mode='fp32')
    assert listing(m) == KEYS[name]['state_dict']
    assert m.num_bins == KEYS[name]['cfg']['num_bins'] and m.num_encoders == 3


@pytest.mark.parametrize('name', ['semseg_skip_concat_k11', 'semseg_no_skip_k6', 'semseg_no_skip_index_map_k5'])
def test_semseg_state_dict_contract(name):
    import ess_b200
    m = ess_b200.SemSegE2VID(**KEYS[name]['args'])
    assert listing(m) == KEYS[name]['state_dict']
    assert [n for n, _ in m.named_parameters()] == [k for k, _, _ in KEYS[name]['state_dict']]   # all entries are parameters


def test_style_encoder_state_dict_contract():
    import ess_b200
    m = ess_b200.StyleEncoderE2VID(**KEYS['style_encoder']['args'])
    assert listing(m) == KEYS['style_encoder']['state_dict']


def test_checkpoint_round_trip_is_strict():
    """A state_dict saved from one instance loads strictly into another and reproduces every tensor."""
    import ess_b200
    a = ess_b200.SemSegE2VID(256, 11, skip_connect=True, skip_type='concat')
    b = ess_b200.SemSegE2VID(256, 11, skip_connect=True, skip_type='concat')
    missing = b.load_state_dict(a.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))
    with pytest.raises(RuntimeError):
        b.load_state_dict({k: v for k, v in list(a.state_dict().items())[:-1]}, strict=True)
