"""Next row f2: the oracle's restatement of the detections hand-off (process_2d_detections, box_iou, complement_2d_gt)
against golden vectors written by the reference's own three methods (oracle/make_f2_golden.py runs their unmodified
source)."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import ROOT

CASES = sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'f2_*.npz')))


def split_per_class(det):
    return [det[det[:, 5] == c][:, :5] for c in range(10)]


@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[3:-4] for p in CASES])
def test_oracle_handoff_matches_reference(path):
    from oracle import mv2d_oracle as O
    g = np.load(path)
    det = O.process_2d_detections([split_per_class(g['det_in'])], float(g['min_size']))[0]
    assert np.array_equal(det.numpy(), g['det_filtered'])
    gts = torch.cat([torch.from_numpy(g['gt_boxes']), torch.ones(len(g['gt_labels']), 1),
                     torch.from_numpy(g['gt_labels']).float()[:, None]], 1)
    out = O.complement_2d_gt(det, gts, float(g['thr']), float(g['min_size']))
    assert np.array_equal(out.numpy(), g['out'])


def test_there_are_golden_cases():
    assert len(CASES) >= 5
