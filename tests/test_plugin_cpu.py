"""CPU: the registry surface builds from our configs and -- unchanged -- from the reference's
experiment configs, with exactly the reference's state_dict names and shapes."""
import os

import pytest

from conftest import ROOT
from mv2d_b200.plugin.build import build_roi_head

OURS = [os.path.join(ROOT, 'configs', 'mv2d_b200', n) for n in ('mv2d_s_r50_1408x512.py', 'mv2d_t_r50_1408x512.py')]
REF_DIR = '/root/reference/configs/mv2d/exp'
REF = [os.path.join(REF_DIR, n) for n in sorted(os.listdir(REF_DIR))] if os.path.isdir(REF_DIR) else []


@pytest.mark.parametrize('path', OURS + REF)
def test_roi_head_builds_with_reference_state_dict(path, state_dicts):
    head = build_roi_head(path, device=None)
    sd, ref = head.state_dict(), state_dicts(6)
    assert set(sd) == set(ref), (set(sd) ^ set(ref))
    for k, v in ref.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    head.load_state_dict(ref, strict=True)
    assert type(head).__name__ == ('MV2DTHead' if 'two_frames' in path or '_t_' in path else 'MV2DSHead')
    assert head.box_corr_module.topk == (20 if head.MODE == 'T' else 1)


def test_config_delete_and_base_semantics(tmp_path):
    from mv2d_b200.config import Config
    (tmp_path / 'base.py').write_text("a = dict(x=1, y=dict(p=1, q=2))\nb = 3\n")
    (tmp_path / 'child.py').write_text("_base_ = ['./base.py']\na = dict(y=dict(_delete_=True, r=5), z=7)\n")
    cfg = Config.fromfile(str(tmp_path / 'child.py'))
    assert cfg.a.x == 1 and cfg.a.z == 7 and dict(cfg.a.y) == {'r': 5} and cfg.b == 3


def test_process_2d_detections_matches_reference_semantics():
    """detectors/mv2d.py:60-86: per-class arrays -> [n,6] with label column, min-size filter (8 px)."""
    import numpy as np
    import torch
    from mv2d_b200.plugin.modules import MV2D
    shell = MV2D.__new__(MV2D)
    torch.nn.Module.__init__(shell)
    shell.train_cfg = None
    shell.test_cfg = dict(detection_proposal=dict(min_bbox_size=8))
    rng = np.random.default_rng(0)
    res = []
    for _ in range(3):
        per_cls = []
        for c in range(10):
            n = int(rng.integers(0, 4))
            xy = rng.uniform(0, 500, size=(n, 2))
            wh = rng.uniform(2, 40, size=(n, 2))
            per_cls.append(np.concatenate([xy, xy + wh, rng.uniform(0, 1, size=(n, 1))], 1).astype(np.float32))
        res.append(per_cls)
    out = shell.process_2d_detections(res, 'cpu')
    for view, det in zip(res, out):
        ref = np.concatenate([np.concatenate([b, np.full((len(b), 1), c, np.float32)], 1) for c, b in enumerate(view)], 0)
        ref = ref[((ref[:, 2:4] - ref[:, 0:2]) >= 8).all(1)]
        assert det.shape[1] == 6 and np.array_equal(det.numpy(), ref)


@pytest.mark.parametrize('path', OURS + REF)
def test_neck_and_training_side_build_from_the_registry(path):
    """The configs' neck (one-level FPN) and train_cfg.rcnn (HungarianAssigner3D with its match costs, stage loss
    weights) build through the registries under the reference's type names; the neck carries mmdet's parameter names."""
    from mv2d_b200.config import Config
    from mv2d_b200.plugin import modules  # noqa: F401
    from mv2d_b200.registry import NECKS, build_from_cfg
    cfg = Config.fromfile(path)
    neck = build_from_cfg(dict(cfg['model']['neck']), NECKS)
    assert set(neck.state_dict()) == {'lateral_convs.0.conv.weight', 'lateral_convs.0.conv.bias',
                                      'fpn_convs.0.conv.weight', 'fpn_convs.0.conv.bias'}
    assert tuple(neck.state_dict()['fpn_convs.0.conv.weight'].shape) == (256, 256, 3, 3) and neck.start_level == 2
    head = build_roi_head(path, device=None, train=True)
    a = head.bbox_head.assigner
    assert type(a).__name__ == 'HungarianAssigner3D' and a.cls_cost.weight == 2.0 and a.reg_cost.weight == 0.25
    assert head.stage_loss_weights == [0.1] * 6
    kw = head.bbox_head._loss_kwargs()
    assert kw['code_weights'] == [1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.5, 1.5, 2.0, 2.0] and kw['cls_loss_weight'] == 2.0 and kw['bbox_loss_weight'] == 0.25
