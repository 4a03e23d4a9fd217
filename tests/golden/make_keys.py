"""Generates tests/golden/state_dict_keys.json from the UNMODIFIED reference (/root/reference), build container only:
    python tests/golden/make_keys.py
Ordered (key, shape) lists of the reference modules' state_dicts -- the checkpoint contract of the drop-ins
(e2vid/utils/loading_utils.py:19 strict load_state_dict; utils/saver.py:20,38,58 round-trips `back_end`)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402


def listing(module):
    return [[k, list(v.shape), str(v.dtype)] for k, v in module.state_dict().items()]


def main():
    ref_shim.install()
    from models.style_networks import SemSegE2VID, StyleEncoderE2VID
    out = {}
    for name, over in (('e2vid_lightweight_convlstm', {}), ('e2vid_convgru', dict(recurrent_block_type='convgru')),
                       ('e2vid_upsample_conv_10bins', dict(use_upsample_conv=True, num_bins=10)),
                       ('e2vid_instance_norm', dict(norm='IN'))):
        cfg = dict(ref_shim.E2VID_LIGHTWEIGHT_CFG, **over)
        out[name] = dict(cfg=cfg, state_dict=listing(ref_shim.make_reference_e2vid(cfg)))
    torch.manual_seed(0)
    out['semseg_skip_concat_k11'] = dict(args=dict(input_c=256, output_c=11, skip_connect=True, skip_type='concat'),
                                         state_dict=listing(SemSegE2VID(256, 11, skip_connect=True, skip_type='concat')))
    out['semseg_no_skip_k6'] = dict(args=dict(input_c=256, output_c=6, skip_connect=False),
                                    state_dict=listing(SemSegE2VID(256, 6, skip_connect=False)))
    out['semseg_no_skip_index_map_k5'] = dict(args=dict(input_c=256, output_c=5, skip_connect=False, input_index_map=True),
                                              state_dict=listing(SemSegE2VID(256, 5, skip_connect=False, input_index_map=True)))
    out['style_encoder'] = dict(args=dict(input_dim=1, skip_connect=True),
                                state_dict=listing(StyleEncoderE2VID(1, skip_connect=True)))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'state_dict_keys.json')
    json.dump(out, open(path, 'w'))
    print('wrote', path, {k: len(v['state_dict']) for k, v in out.items()})


if __name__ == '__main__':
    main()
