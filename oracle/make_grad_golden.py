"""TEST INFRASTRUCTURE ONLY.  Golden vectors for the training step (BASELINE configs[3]): runs the
REFERENCE's own MV2DSHead forward (position encoding -> _bbox_forward) and CrossAttentionBoxHead.loss with torch
autograd switched on (unmodified files under /root/reference through oracle/ref_shim.py), sums the per-layer losses
with the config's stage_loss_weights as MV2DSHead.forward_train does (mv2d_s_head.py:278-305) and stores

  * the loss values,
  * the gradients of the inputs of the decoder slice (reference points, RoI feature tokens, RoI position tokens --
    the gathered [N,M,...] gradients scattered back to the N RoIs),
  * every parameter gradient of roi_head (bbox_head.* from mv2d_decoder_train_backward, query_generator.* and
    position_encoding.* from mv2d_front_train_backward) and d loss / d feat, big tensors subsampled.

The single-frame cases assemble the loss layer by layer with bbox_head.loss (so the inputs of the decoder slice can be
hooked); the result was checked to be identical to the reference's own head.forward_train (same total, bit-identical
d feat).  The two-frame case calls head.forward_train itself.

Run in the build container:   python -m oracle.make_grad_golden
"""
import copy
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
from oracle.make_golden import CFG  # noqa: E402

# name -> (hot-path case, decoder layers, GT spec)
CASES = {
    'grad_s_small': (dict(synth.CASES['s_small'], num_layers=2), dict(num_gt=6, seed=71)),
    'grad_s_mid': (dict(mode='S', seed=21, num_views=6, boxes_per_view=[12, 10, 11, 9, 12, 10], num_layers=3), dict(num_gt=20, seed=72)),
    'grad_s_one': (dict(synth.CASES['s_one'], num_layers=2), dict(num_gt=3, seed=73)),
    # padded images: masked cells in the sine branch and in the position embedding (img_shape < pad_shape)
    'grad_s_pad': (dict(synth.CASES['s_pad'], num_layers=2), dict(num_gt=5, seed=74)),
}
SUB = 97          # stride of the subsample kept for tensors with more than KEEP_FULL elements
KEEP_FULL = 4096


class _Boxes:
    """The two attributes CrossAttentionBoxHead.loss reads from LiDARInstance3DBoxes (cross_attention_head.py:450-452)."""

    def __init__(self, gt):
        self.gravity_center, self.tensor = gt[:, :3], torch.cat([gt[:, :3], gt[:, 3:]], 1)


def sub(t):
    t = t.detach().reshape(-1)
    return (t if t.numel() <= KEEP_FULL else t[::SUB]).numpy().copy()


def run_reference(spec, gt_spec, weight_seed=0):
    Assigner, _ = ref_shim.install_loss_support()
    cfg = ref_shim.load_reference_config(CFG['S'])
    roi_head = copy.deepcopy(cfg['model']['roi_head'])
    roi_head['bbox_head']['transformer']['decoder']['num_layers'] = spec['num_layers']
    roi_head.update(train_cfg=None, test_cfg=ref_shim.ConfigDict(cfg['model']['test_cfg']['rcnn']))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        head = ref_shim.build_from_cfg(roi_head, ref_shim.HEADS).eval()     # eval: dropout off, as the CUDA path
    tc = cfg['model']['train_cfg']['rcnn']
    tc = tc[0] if isinstance(tc, (list, tuple)) else tc
    a = dict(tc['assigner'])
    a.pop('type')
    bh = head.bbox_head
    bh.assigner, bh.sampler = Assigner(**a), ref_shim.PseudoSampler()
    lc, lb = dict(roi_head['bbox_head']['loss_cls']), dict(roi_head['bbox_head']['loss_bbox'])
    lc.pop('type'); lb.pop('type')
    bh.loss_cls, bh.loss_bbox = ref_shim.FocalLoss(**lc), ref_shim.L1Loss(**lb)
    stage_w = list(tc['stage_loss_weights'])[:spec['num_layers']]
    sd = synth.make_state_dict(weight_seed, num_layers=spec['num_layers'])
    head.load_state_dict(sd)
    feat, boxes, metas = synth.case_inputs(spec)
    gt_boxes, gt_labels, _ = synth.make_dn_inputs(gt_spec)
    feat = feat.clone().requires_grad_(True)

    cap = {}

    def grab(mod, args, kwargs):      # the slice inputs, as CrossAttentionBoxHead.forward receives them
        ref, x, masks, pos = args[:4]
        for t in (ref, x, pos):
            t.retain_grad()
        cap.update(ref=ref, x=x, pos=pos)
    hook = bh.register_forward_pre_hook(grab, with_kwargs=True)
    orig = head.box_corr_module.gen_box_roi_correlation

    def wrap(*a, **k):
        c, m = orig(*a, **k)
        cap['corr'], cap['corr_mask'] = c.clone(), m.clone()
        return c, m
    head.box_corr_module.gen_box_roi_correlation = wrap

    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        pe = head.position_encoding([feat], metas)[0]
        out = head._bbox_forward([torch.cat([feat, pe], dim=1)], [b.clone() for b in boxes], metas)
        L = spec['num_layers']
        total, lcs, lbs = 0.0, [], []
        for l in range(L):
            d = bh.loss([_Boxes(gt_boxes)], [gt_labels], {'cls_scores': [out['cls_scores'][l]], 'bbox_preds': [out['bbox_preds'][l]]})
            total = total + stage_w[l] * (d['loss_cls'] + d['loss_bbox'])
            lcs.append(float(d['loss_cls'])); lbs.append(float(d['loss_bbox']))
        total.backward()
    hook.remove()
    N, M = cap['corr'].shape
    corr, mask = cap['corr'], cap['corr_mask']
    # gathered [N,M,C,7,7] gradients back onto the N RoIs (a masked-out slot contributes exactly zero)
    idx = corr.reshape(-1)
    d_feat_tok = torch.zeros(N, 256, 7, 7).index_add_(0, idx, cap['x'].grad.reshape(N * M, 256, 7, 7))
    d_pos_tok = torch.zeros(N, 256, 7, 7).index_add_(0, idx, cap['pos'].grad.reshape(N * M, 256, 7, 7))
    g = dict(
        spec=np.frombuffer(json.dumps(spec).encode(), dtype=np.uint8),
        gt_spec=np.frombuffer(json.dumps(gt_spec).encode(), dtype=np.uint8),
        stage_loss_weights=np.array(stage_w, np.float64),
        loss_cls=np.array(lcs, np.float64), loss_bbox=np.array(lbs, np.float64), loss=np.float64(float(total)),
        cls_scores=torch.stack(out['cls_scores']).detach().numpy(), bbox_preds=torch.stack(out['bbox_preds']).detach().numpy(),
        corr=corr.numpy(), corr_mask=mask.numpy(),
        d_ref=cap['ref'].grad.reshape(N, 3).numpy().copy(),
        d_roi_feat_sub=sub(d_feat_tok), d_roi_pos_sub=sub(d_pos_tok),
        d_feat_sub=sub(feat.grad),
        sub_stride=np.int64(SUB), keep_full=np.int64(KEEP_FULL),
    )
    for name, prm in head.named_parameters():
        if prm.grad is not None:
            g['dparam.' + name] = sub(prm.grad)
    return g


# two-frame head WITH denoising queries (the configuration the reference trains MV2D-T with): the gradients come from the
# reference's own MV2DSHead.forward_train (inherited by MV2DTHead: mv2d_s_head.py:236-307), i.e. its loss dict summed
# as mmdet's _parse_losses does.  Pins the oracle for the rows whose backward comes next.
CASES_T = {'grad_t_dn': dict(synth.CASES['t_dn'], num_layers=2)}


def run_reference_forward_train_t(spec, weight_seed=0):
    Assigner, _ = ref_shim.install_loss_support()
    cfg = ref_shim.load_reference_config(CFG['T'])
    roi_head = copy.deepcopy(cfg['model']['roi_head'])
    roi_head['bbox_head']['transformer']['decoder']['num_layers'] = spec['num_layers']
    roi_head.update(train_cfg=None, test_cfg=ref_shim.ConfigDict(cfg['model']['test_cfg']['rcnn']))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        head = ref_shim.build_from_cfg(roi_head, ref_shim.HEADS).eval()     # children in eval: dropout off
    tc = cfg['model']['train_cfg']['rcnn']
    tc = tc[0] if isinstance(tc, (list, tuple)) else tc
    a = dict(tc['assigner'])
    a.pop('type')
    bh = head.bbox_head
    bh.assigner, bh.sampler = Assigner(**a), ref_shim.PseudoSampler()
    lc, lb = dict(roi_head['bbox_head']['loss_cls']), dict(roi_head['bbox_head']['loss_bbox'])
    lc.pop('type'); lb.pop('type')
    bh.loss_cls, bh.loss_bbox = ref_shim.FocalLoss(**lc), ref_shim.L1Loss(**lb)
    head.stage_loss_weights = list(tc['stage_loss_weights'])[:spec['num_layers']]
    head.load_state_dict(synth.make_state_dict(weight_seed, num_layers=spec['num_layers']))
    feat, boxes, metas = synth.case_inputs(spec)
    gt_boxes, gt_labels, rand = synth.make_dn_inputs(spec['dn'])
    feat = feat.clone().requires_grad_(True)
    head.training = True                      # only the head's flag: the denoising branch (mv2d_t_head.py:91-98)
    patched = (torch.Tensor.cuda, torch.rand_like)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.rand_like = lambda t, *a, **k: rand.clone()
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            losses = head.forward_train([feat], metas, [b.clone() for b in boxes], None, None, None, None,
                                        [_Boxes(gt_boxes)], [gt_labels], None)
    finally:
        torch.Tensor.cuda, torch.rand_like = patched
        head.training = False
    total = sum(v for k, v in losses.items() if 'loss' in k)      # mmdet BaseDetector._parse_losses
    total.backward()
    g = dict(
        spec=np.frombuffer(json.dumps(spec).encode(), dtype=np.uint8),
        stage_loss_weights=np.array(head.stage_loss_weights, np.float64),
        denoise_weight=np.float64(head.denoise_weight), denoise_split=np.float64(head.denoise_split),
        neg_bbox_loss=np.int64(bool(head.neg_bbox_loss)),
        loss=np.float64(float(total.detach())),
        loss_names=np.frombuffer(json.dumps(sorted(losses)).encode(), dtype=np.uint8),
        loss_values=np.array([float(losses[k].detach()) for k in sorted(losses)], np.float64),
        d_feat_sub=sub(feat.grad), sub_stride=np.int64(SUB), keep_full=np.int64(KEEP_FULL),
    )
    for name, prm in head.named_parameters():
        if prm.grad is not None:
            g['dparam.' + name] = sub(prm.grad)
    return g


def main():
    names = sys.argv[1:] or (list(CASES) + list(CASES_T))
    for name in [n for n in names if n in CASES_T]:
        g = run_reference_forward_train_t(CASES_T[name])
        path = os.path.join(ROOT, 'tests', 'golden', f'{name}.npz')
        np.savez_compressed(path, **g)
        print(name, 'loss', float(g['loss']), 'tensors', sum(k.startswith('dparam.') for k in g), os.path.getsize(path) // 1024, 'KiB')
    names = [n for n in names if n in CASES]
    for name in names:
        spec, gt_spec = CASES[name]
        g = run_reference(spec, gt_spec)
        path = os.path.join(ROOT, 'tests', 'golden', f'{name}.npz')
        np.savez_compressed(path, **g)
        print(name, 'N', g['corr'].shape[0], 'loss', float(g['loss']), 'tensors', sum(k.startswith('dparam.') for k in g),
              os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    main()
