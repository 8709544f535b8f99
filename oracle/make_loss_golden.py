"""TEST INFRASTRUCTURE ONLY.  Golden vectors for the training targets / losses row: runs the
REFERENCE's own HungarianAssigner3D.assign, BBox3DL1Cost, normalize_bbox and CrossAttentionBoxHead.loss_single /
dn_loss_single (unmodified files under /root/reference, imported through oracle/ref_shim.py; the mmdet 2.25.1
FocalLoss / L1Loss / FocalLossCost / PseudoSampler they call are restated in the shim) on the decoder outputs stored
in tests/golden/<case>.npz.  Run in the build container:   python -m oracle.make_loss_golden
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mv2d_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

# (golden with the decoder outputs, GT spec); the *_dn cases also carry denoising-query outputs
CASES = {'loss_s_dn': ('s_dn', None), 'loss_t_dn': ('t_dn', None),
         'loss_s_cfg2': ('s_cfg2', dict(num_gt=30, seed=91)), 'loss_s_small': ('s_small', dict(num_gt=40, seed=92)),
         'loss_s_one': ('s_one', dict(num_gt=3, seed=93))}
CFG = '/root/reference/configs/mv2d/exp/mv2d_r50_frcnn_single_frame_roi_1408x512_ep72.py'


class _Boxes:
    """The two attributes CrossAttentionBoxHead.loss reads from LiDARInstance3DBoxes (cross_attention_head.py:450-452)."""

    def __init__(self, gt):
        self.gravity_center, self.tensor = gt[:, :3], torch.cat([gt[:, :3], gt[:, 3:]], 1)


def build_head():
    import copy
    Assigner, _ = ref_shim.install_loss_support()
    cfg = ref_shim.load_reference_config(CFG)
    bh = copy.deepcopy(cfg['model']['roi_head']['bbox_head'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        head = ref_shim.build_from_cfg(dict(bh, train_cfg=None), ref_shim.HEADS)
    tc = cfg['model']['train_cfg']['rcnn']
    tc = tc[0] if isinstance(tc, (list, tuple)) else tc
    a = dict(tc['assigner'])
    a.pop('type')
    head.assigner = Assigner(**a)
    head.sampler = ref_shim.PseudoSampler()
    lc, lb = dict(bh['loss_cls']), dict(bh['loss_bbox'])
    lc.pop('type'); lb.pop('type')
    head.loss_cls, head.loss_bbox = ref_shim.FocalLoss(**lc), ref_shim.L1Loss(**lb)
    return head, cfg


def main():
    head, cfg = build_head()
    for name, (src, gt_spec) in CASES.items():
        g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', f'{src}.npz')))
        spec = json.loads(bytes(g['spec']).decode())
        gt_spec = gt_spec or spec['dn']
        gt_boxes, gt_labels, _ = synth.make_dn_inputs(gt_spec)
        cls, box = torch.from_numpy(g['cls_scores']), torch.from_numpy(g['bbox_preds'])
        L, N = cls.shape[:2]
        out = dict(src=np.frombuffer(src.encode(), dtype=np.uint8),
                   gt_spec=np.frombuffer(json.dumps(gt_spec).encode(), dtype=np.uint8))
        lcs, lbs, asg = [], [], []
        with torch.no_grad():
            for l in range(L):
                d = head.loss([_Boxes(gt_boxes)], [gt_labels], {'cls_scores': [cls[l].clone()], 'bbox_preds': [box[l].clone()]})
                lcs.append(float(d['loss_cls'])); lbs.append(float(d['loss_bbox']))
                r = head.assigner.assign(box[l], cls[l], gt_boxes, gt_labels)
                asg.append((r.gt_inds - 1).numpy())
            out.update(loss_cls=np.array(lcs, np.float64), loss_bbox=np.array(lbs, np.float64), assigned=np.stack(asg))
            if 'dn_cls' in g:
                mode_split = 0.75 if spec['mode'] == 'S' else 0.6        # denoise_split: mv2d_s_head.py:26 / exp two_frames :47
                pad = int(g['dn_pad'])
                known_labels = torch.from_numpy(g['dn_labels'])
                known_boxes = gt_boxes.repeat(pad // gt_boxes.shape[0], 1)
                dc, db = [], []
                for l in range(L):
                    a, b = head.dn_loss_single(torch.from_numpy(g['dn_cls'][l]).clone(), torch.from_numpy(g['dn_box'][l]).clone(),
                                               known_boxes.clone(), known_labels.clone(), pad,
                                               cfg['model']['roi_head'].get('pc_range', None), mode_split,
                                               neg_bbox_loss=spec['mode'] == 'T')     # exp two_frames config :45
                    dc.append(float(a)); db.append(float(b))
                out.update(dn_loss_cls=np.array(dc, np.float64), dn_loss_bbox=np.array(db, np.float64), dn_split=np.float64(mode_split),
                           dn_neg_bbox_loss=np.int64(spec['mode'] == 'T'))
        path = os.path.join(ROOT, 'tests', 'golden', f'{name}.npz')
        np.savez_compressed(path, **out)
        print(name, 'L', L, 'N', N, 'G', gt_boxes.shape[0], 'loss_cls', lcs[-1], 'loss_bbox', lbs[-1],
              'pos', int((out['assigned'][-1] >= 0).sum()), os.path.getsize(path), 'B')


if __name__ == '__main__':
    main()
