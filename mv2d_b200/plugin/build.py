"""Build the hot-path head from a config file (ours under configs/mv2d_b200/, or an unchanged
reference experiment config configs/mv2d/exp/*.py)."""
import copy

from ..config import Config
from ..registry import HEADS, build_from_cfg
from . import modules  # noqa: F401  (registers the types)


def roi_head_cfg(cfg):
    """The ``model.roi_head`` dict with the shared block merged in (our configs keep the common part
    at top level; the reference configs carry everything under model.roi_head)."""
    rh = copy.deepcopy(dict(cfg['model']['roi_head']))
    if 'roi_head' in cfg:   # configs/mv2d_b200/* layout
        base = copy.deepcopy(dict(cfg['roi_head']))
        base.update(rh)
        rh = base
    return rh


def build_roi_head(cfg_path, device='cuda', train=False):
    """HEADS.build(cfg.model.roi_head) with test_cfg=cfg.model.test_cfg.rcnn and train_cfg=None (evaluation) or
    cfg.model.train_cfg.rcnn (train=True: assigner / stage loss weights for the loss row) --
    what MV2D.__init__ does (reference detectors/mv2d.py:34-38)."""
    cfg = Config.fromfile(cfg_path)
    rh = roi_head_cfg(cfg)
    test = dict(cfg['model'].get('test_cfg', {}) or {}).get('rcnn')
    tr = dict(cfg['model'].get('train_cfg', {}) or {}).get('rcnn') if train else None
    rh.update(train_cfg=tr, test_cfg=test)
    head = build_from_cfg(rh, HEADS).eval()
    return head.to(device) if device is not None else head
