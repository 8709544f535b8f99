"""Live run of the reference's own Python (through the test-only shim) beside the oracle.
Only possible in the build container, where /root/reference exists; skipped elsewhere."""
import os
import warnings

import pytest
import torch

from mv2d_b200 import synth
from oracle import mv2d_oracle as O

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present')
EXP = REF + '/configs/mv2d/exp/'


def _head(name):
    from oracle import ref_shim
    return ref_shim.build_reference_head(EXP + name)


def test_state_dict_names_and_shapes_match_reference(state_dicts):
    head, _ = _head('mv2d_r50_frcnn_single_frame_roi_1408x512_ep72.py')
    ref_sd = head.state_dict()
    sd = state_dicts(6)
    assert set(ref_sd) == set(sd)
    assert sum(v.numel() for v in ref_sd.values()) == 14019973
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(sd[k].shape), k


@pytest.mark.parametrize('mode,cfg_name,V', [
    ('S', 'mv2d_r50_frcnn_single_frame_roi_1408x512_ep24.py', 6),
    ('T', 'mv2d_r50_frcnn_two_frames_1408x512_ep24.py', 12)])
def test_oracle_equals_live_reference(mode, cfg_name, V, state_dicts):
    head, _ = _head(cfg_name)
    sd = state_dicts(6)
    head.load_state_dict(sd)
    feat, boxes, metas = synth.make_sample(21, V, 5, cam_jitter_deg=3.0)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter('ignore')
        pe = head.position_encoding([feat], metas)[0]
        out = head._bbox_forward([torch.cat([feat, pe], 1)], [b.clone() for b in boxes], metas)
        fn = O.mv2d_s_forward if mode == 'S' else O.mv2d_t_forward
        cls, box = fn(sd, feat, boxes, metas, O.make_cfg(mode))
    assert (cls - torch.stack(out['cls_scores'])).abs().max() < 1e-4
    assert (box - torch.stack(out['bbox_preds'])).abs().max() < 2e-4
