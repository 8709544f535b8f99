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


def build_roi_head(cfg_path, device='cuda'):
    """HEADS.build(cfg.model.roi_head) with train_cfg=None, test_cfg=cfg.model.test_cfg.rcnn --
    what MV2D.__init__ does (reference detectors/mv2d.py:34-38)."""
    cfg = Config.fromfile(cfg_path)
    rh = roi_head_cfg(cfg)
    test = dict(cfg['model'].get('test_cfg', {}) or {}).get('rcnn')
    rh.update(train_cfg=None, test_cfg=test)
    head = build_from_cfg(rh, HEADS).eval()
    return head.to(device) if device is not None else head
