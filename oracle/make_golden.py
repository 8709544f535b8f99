"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the REFERENCE's own
Python (unmodified files under /root/reference, imported through oracle/ref_shim.py) on the
seeded synthetic cases of mv2d_b200/synth.py.  Run in the build container:

    python -m oracle.make_golden            # all cases
    python -m oracle.make_golden s_small    # one case

The fixtures hold outputs only (inputs and weights are regenerated from their seeds), plus a
few stage-level intermediates captured with forward hooks so each CUDA stage can be checked
against the reference itself, not only against the restatement.
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

CFG = {'S': '/root/reference/configs/mv2d/exp/mv2d_r50_frcnn_single_frame_roi_1408x512_ep72.py',
       'T': '/root/reference/configs/mv2d/exp/mv2d_r50_frcnn_two_frames_1408x512_ep72.py'}
PE_SUBSAMPLE = 251  # pe is 17-35 MB; keep every 251st element (prime, hits all channels)


def run_reference(spec, weight_seed=0):
    import copy
    ref_shim.install()
    cfg = ref_shim.load_reference_config(CFG[spec['mode']])
    roi_head = copy.deepcopy(cfg['model']['roi_head'])
    roi_head['bbox_head']['transformer']['decoder']['num_layers'] = spec['num_layers']
    if 'dn' in spec:
        roi_head['use_denoise'] = True     # the single-frame exp configs ship with it off; the code path exists
    roi_head.update(train_cfg=None, test_cfg=ref_shim.ConfigDict(cfg['model']['test_cfg']['rcnn']))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        head = ref_shim.build_from_cfg(roi_head, ref_shim.HEADS).eval()
    sd = synth.make_state_dict(weight_seed, num_layers=spec['num_layers'])
    head.load_state_dict(sd)
    feat, boxes, metas = synth.case_inputs(spec)
    cap = {}
    head.query_generator.register_forward_hook(
        lambda m, i, o: cap.__setitem__('center_lidar', o[0].detach().clone()))
    head.bbox_head.transformer.register_forward_hook(
        lambda m, i, o: cap.__setitem__('outs_dec', o[0].detach().clone()))
    head.bbox_head.query_embedding.register_forward_hook(
        lambda m, i, o: cap.__setitem__('query_pos', o.detach().clone()))
    orig_s = head.box_corr_module.gen_box_roi_correlation
    orig_t = head.box_corr_module.gen_box_correlation

    def wrap_s(*a, **k):
        c, m = orig_s(*a, **k)
        cap['corr'], cap['corr_mask'] = c.clone(), m.clone()
        return c, m

    def wrap_t(*a, **k):
        r = orig_t(*a, **k)
        cap['key_mask'] = r.clone()
        return r

    head.box_corr_module.gen_box_roi_correlation = wrap_s
    head.box_corr_module.gen_box_correlation = wrap_t
    dn_patch = None
    if 'dn' in spec:
        # training-mode forward with denoising queries (row a20).  Only the HEAD's `training` flag is raised
        # (children stay in eval: dropout off); .cuda() is a no-op here and the noise is injected.
        gt_boxes, gt_labels, rand = synth.make_dn_inputs(spec['dn'])

        class _Boxes:   # the two attributes prepare_for_dn reads from LiDARInstance3DBoxes (mv2d_s_head.py:42)
            gravity_center = gt_boxes[:, :3]
            tensor = torch.cat([gt_boxes[:, :3], gt_boxes[:, 3:]], 1)
        metas[0] = dict(metas[0], gt_bboxes_3d=_Boxes(), gt_labels_3d=gt_labels)
        head.training = True
        dn_patch = (torch.Tensor.cuda, torch.rand_like)
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.rand_like = lambda t, *a, **k: rand.clone()
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter('ignore')
        pe = head.position_encoding([feat], metas)[0]
        x = [torch.cat([feat, pe], dim=1)]
        out = head._bbox_forward(x, [b.clone() for b in boxes], metas)
        if dn_patch is not None:
            torch.Tensor.cuda, torch.rand_like = dn_patch
            head.training = False
        L = spec['num_layers']
        N = out['rois'].shape[0]
        # MV2DHead.simple_test tail (mv2d_head.py:262-265): get_bboxes = NMSFreeCoder.decode +
        # z-shift; box_type_3d replaced by an identity wrapper (LiDARInstance3DBoxes is mmdet3d)
        metas_bt = [dict(m, box_type_3d=lambda t, d: t) for m in metas]
        dec = head.bbox_head.get_bboxes(
            {'cls_scores': [out['cls_scores'][-1].clone()],
             'bbox_preds': [out['bbox_preds'][-1].clone()]}, metas_bt)[0]
        decoded = dict(bboxes=dec[0], scores=dec[1], labels=dec[2])
    g = dict(
        spec=np.frombuffer(json.dumps(spec).encode(), dtype=np.uint8),
        cls_scores=torch.stack(out['cls_scores']).numpy(),
        bbox_preds=torch.stack(out['bbox_preds']).numpy(),
        rois=out['rois'].numpy(),
        intrinsics=out['intrinsics'].numpy(), extrinsics=out['extrinsics'].numpy(),
        center_lidar=cap['center_lidar'].numpy(),
        query_pos=cap['query_pos'].reshape(-1, 256)[-N:].numpy(),
        outs_dec=cap['outs_dec'].reshape(L, -1, 256)[:, -N:].numpy(),
        pe_sub=pe.flatten()[::PE_SUBSAMPLE].numpy().copy(),
        roi_feat_sub=out['bbox_feats'].flatten()[::PE_SUBSAMPLE].numpy().copy(),
        dec_boxes=decoded['bboxes'].numpy(), dec_scores=decoded['scores'].numpy(),
        dec_labels=decoded['labels'].numpy(),
    )
    if 'corr' in cap:
        g['corr'] = cap['corr'].numpy()
        g['corr_mask'] = cap['corr_mask'].numpy()
    if 'key_mask' in cap:
        g['key_mask_packed'] = np.packbits(cap['key_mask'].numpy().reshape(N, -1), axis=1)
    if out.get('dn_mask_dict'):
        md = out['dn_mask_dict']
        kc, kb = md['output_known_lbs_bboxes']
        g['dn_cls'] = kc[:, 0].numpy()          # [L, pad, 10]
        g['dn_box'] = kb[:, 0].numpy()
        g['dn_labels'] = md['known_lbs_bboxes'][0].numpy()
        g['dn_pad'] = np.int64(md['pad_size'])
    return g


def main():
    names = sys.argv[1:] or list(synth.CASES)
    os.makedirs(os.path.join(ROOT, 'tests', 'golden'), exist_ok=True)
    for name in names:
        g = run_reference(synth.CASES[name])
        path = os.path.join(ROOT, 'tests', 'golden', f'{name}.npz')
        # full-size cases: keep only the final outputs + small stage tensors
        if g['rois'].shape[0] > 100:
            for k in ('outs_dec', 'query_pos', 'intrinsics', 'extrinsics'):
                g.pop(k, None)
        np.savez_compressed(path, **g)
        print(name, 'N =', g['rois'].shape[0], os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    main()
