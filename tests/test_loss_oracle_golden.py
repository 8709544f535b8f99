"""Row f3 (training targets / losses): the oracle restatement against golden vectors written by the reference's own
HungarianAssigner3D / loss_single / dn_loss_single (oracle/make_loss_golden.py)."""
import json

import numpy as np
import pytest
import torch

from conftest import golden_path
from mv2d_b200 import synth
from oracle import mv2d_oracle as O

LOSS_CASES = ['loss_s_dn', 'loss_t_dn', 'loss_s_cfg2', 'loss_s_small', 'loss_s_one']


def load_loss_case(name):
    g = dict(np.load(golden_path(name)))
    src = bytes(g['src']).decode()
    gt_spec = json.loads(bytes(g['gt_spec']).decode())
    s = dict(np.load(golden_path(src)))
    gt_boxes, gt_labels, _ = synth.make_dn_inputs(gt_spec)
    return g, s, gt_boxes, gt_labels


@pytest.mark.parametrize('name', LOSS_CASES)
def test_loss_restatement_matches_reference(name):
    g, s, gt_boxes, gt_labels = load_loss_case(name)
    cls, box = torch.from_numpy(s['cls_scores']), torch.from_numpy(s['bbox_preds'])
    for l in range(cls.shape[0]):
        lc, lb, asg = O.loss_single(cls[l], box[l], gt_boxes, gt_labels)
        assert np.array_equal(asg.numpy(), g['assigned'][l]), f'layer {l}: Hungarian assignment differs'
        assert abs(float(lc) - g['loss_cls'][l]) <= 1e-5 * abs(g['loss_cls'][l]) + 1e-7
        assert abs(float(lb) - g['loss_bbox'][l]) <= 1e-5 * abs(g['loss_bbox'][l]) + 1e-7
    if 'dn_loss_cls' in g:
        pad = int(s['dn_pad'])
        known = gt_boxes.repeat(pad // gt_boxes.shape[0], 1)
        labels = torch.from_numpy(s['dn_labels'])
        for l in range(cls.shape[0]):
            a, b = O.dn_loss_single(torch.from_numpy(s['dn_cls'][l]), torch.from_numpy(s['dn_box'][l]), known, labels, pad,
                                    float(g['dn_split']), neg_bbox_loss=bool(g['dn_neg_bbox_loss']))
            assert abs(float(a) - g['dn_loss_cls'][l]) <= 1e-5 * abs(g['dn_loss_cls'][l]) + 1e-7
            assert abs(float(b) - g['dn_loss_bbox'][l]) <= 1e-5 * abs(g['dn_loss_bbox'][l]) + 1e-7


def test_loss_without_ground_truth():
    """No GT boxes: every query is background, the box loss is zero (hungarian_assigner_3d.py:109-115)."""
    cls, box = torch.randn(7, 10), torch.randn(7, 10)
    lc, lb, asg = O.loss_single(cls, box, torch.zeros(0, 9), torch.zeros(0, dtype=torch.long))
    assert (asg == -1).all() and float(lb) == 0.0 and float(lc) > 0
