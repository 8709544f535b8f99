"""CPU: the differentiable restatement of the training slice (oracle.decoder_slice_loss, torch autograd) against the
gradients the REFERENCE's own Python produced (tests/golden/grad_*.npz, written by oracle/make_grad_golden.py):
loss values, gradients of the slice inputs and of every bbox_head parameter."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from mv2d_b200 import synth
from oracle import mv2d_oracle as O

CASES = ['grad_s_small', 'grad_s_mid', 'grad_s_one']


def load(name):
    g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', f'{name}.npz')))
    return g, json.loads(bytes(g['spec']).decode()), json.loads(bytes(g['gt_spec']).decode())


def sub(t, g):
    t = t.detach().reshape(-1)
    return (t if t.numel() <= int(g['keep_full']) else t[::int(g['sub_stride'])]).numpy()


def slice_inputs(spec):
    """The slice's inputs from the oracle's own front end (no grad): ref, roi_feat, roi_pe, corr, mask."""
    sd = synth.make_state_dict(0, num_layers=spec['num_layers'])
    feat, boxes, metas = synth.case_inputs(spec)
    cfg = O.make_cfg('S', num_layers=spec['num_layers'])
    with torch.no_grad():
        _, _, st = O.mv2d_s_forward(sd, feat, boxes, metas, cfg, return_stages=True)
    return sd, cfg, st


def oracle_grads(spec, gt_spec, stage_w):
    sd, cfg, st = slice_inputs(spec)
    gt_boxes, gt_labels, _ = synth.make_dn_inputs(gt_spec)
    sd = {k: (v.clone().requires_grad_(True) if k.startswith('bbox_head.') else v) for k, v in sd.items()}
    ref = st['ref'].clone().requires_grad_(True)
    roi_feat = st['roi_feat'].clone().requires_grad_(True)
    roi_pe = st['roi_pe'].clone().requires_grad_(True)
    total, cls, box, per = O.decoder_slice_loss(sd, ref, roi_feat, roi_pe, st['corr'], st['corr_mask'], gt_boxes, gt_labels, cfg,
                                                stage_loss_weights=stage_w)
    total.backward()
    return dict(sd=sd, ref=ref, roi_feat=roi_feat, roi_pe=roi_pe, total=total, cls=cls, box=box, per=per, st=st)


@pytest.mark.parametrize('name', CASES)
def test_gradient_oracle_matches_reference(name):
    g, spec, gt_spec = load(name)
    r = oracle_grads(spec, gt_spec, list(g['stage_loss_weights']))
    assert np.array_equal(r['st']['corr'].numpy(), g['corr'])
    np.testing.assert_allclose(float(r['total'].detach()), float(g['loss']), rtol=2e-5)
    np.testing.assert_allclose([float(p[0].detach()) for p in r['per']], g['loss_cls'], rtol=2e-5)
    np.testing.assert_allclose([float(p[1].detach()) for p in r['per']], g['loss_bbox'], rtol=2e-5)

    def close(a, b, what):
        # relative to the tensor's largest entry, with an absolute floor: layer 0's self-attention in_proj weight has
        # an exactly-zero true gradient (its values are all the bias), autograd leaves 1e-9 round-off there
        scale = max(float(np.abs(b).max()), 1e-5)
        err = float(np.abs(a - b).max()) / scale
        assert err < 2e-3, f'{what}: max error {err:.2e} of the largest gradient entry'
    close(r['ref'].grad.numpy(), g['d_ref'], 'd_ref')
    close(sub(r['roi_feat'].grad, g), g['d_roi_feat_sub'], 'd_roi_feat')
    close(sub(r['roi_pe'].grad, g), g['d_roi_pos_sub'], 'd_roi_pos')
    n = 0
    for k, v in r['sd'].items():
        if k.startswith('bbox_head.') and ('dparam.' + k) in g:
            grad = v.grad if v.grad is not None else torch.zeros_like(v)
            close(sub(grad, g), g['dparam.' + k], k)
            n += 1
    assert n >= 6 + 34 * spec['num_layers']


def full_oracle_grads(spec, gt_spec, stage_w, sd=None):
    """Autograd through the whole restated hot path: gradients of all weights and of the feature map."""
    sd = sd or synth.make_state_dict(0, num_layers=spec['num_layers'])
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and k != 'bbox_head.code_weights' else v) for k, v in sd.items()}
    feat, boxes, metas = synth.case_inputs(spec)
    feat = feat.clone().requires_grad_(True)
    gt_boxes, gt_labels, _ = synth.make_dn_inputs(gt_spec) if gt_spec['num_gt'] > 0 else (torch.zeros(0, 9), torch.zeros(0, dtype=torch.long), None)
    cfg = O.make_cfg('S', num_layers=spec['num_layers'])
    total, cls, box, per = O.hot_path_loss(sd, feat, boxes, metas, gt_boxes, gt_labels, cfg, stage_loss_weights=stage_w)
    total.backward()
    return dict(sd=sd, feat=feat, total=total, cls=cls, box=box, per=per, gt=(gt_boxes, gt_labels), inputs=(boxes, metas))


@pytest.mark.parametrize('name', ['grad_s_small', 'grad_s_mid', 'grad_s_one', 'grad_s_pad'])
def test_full_chain_gradient_oracle_matches_reference(name):
    """All 96 hot-path parameter gradients and d loss / d feat of the restated full path vs the reference's autograd."""
    g, spec, gt_spec = load(name)
    r = full_oracle_grads(spec, gt_spec, list(g['stage_loss_weights']))
    np.testing.assert_allclose(float(r['total'].detach()), float(g['loss']), rtol=2e-5)

    def close(a, b, what):
        scale = max(float(np.abs(b).max()), 1e-5)
        err = float(np.abs(a - b).max()) / scale
        assert err < 2e-3, f'{what}: max error {err:.2e} of the largest gradient entry'
    close(sub(r['feat'].grad, g), g['d_feat_sub'], 'd_feat')
    n = 0
    for k, v in r['sd'].items():
        if ('dparam.' + k) in g:
            grad = v.grad if v.grad is not None else torch.zeros_like(v)
            close(sub(grad, g), g['dparam.' + k], k)
            n += 1
    assert n == 6 + 34 * spec['num_layers'] + 22


def test_two_frame_denoising_gradient_oracle_matches_reference_forward_train():
    """The rows whose backward comes next (two-frame head, denoising queries): the differentiable restatement against
    the loss dict and the gradients of the reference's own MV2DTHead.forward_train + autograd (grad_t_dn.npz)."""
    g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'grad_t_dn.npz')))
    spec = json.loads(bytes(g['spec']).decode())
    names = json.loads(bytes(g['loss_names']).decode())
    sd = synth.make_state_dict(0, num_layers=spec['num_layers'])
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and k != 'bbox_head.code_weights' else v) for k, v in sd.items()}
    feat, boxes, metas = synth.case_inputs(spec)
    feat = feat.clone().requires_grad_(True)
    gt_boxes, gt_labels, rand = synth.make_dn_inputs(spec['dn'])
    cfg = O.make_cfg('T', num_layers=spec['num_layers'])
    total, losses = O.hot_path_loss_t(sd, feat, boxes, metas, gt_boxes, gt_labels, rand, cfg,
                                      stage_loss_weights=list(g['stage_loss_weights']), denoise_weight=float(g['denoise_weight']),
                                      neg_bbox_loss=bool(g['neg_bbox_loss']))
    assert sorted(losses) == names
    np.testing.assert_allclose([float(losses[k].detach()) for k in names], g['loss_values'], rtol=5e-5)
    total.backward()

    def close(a, b, what):
        scale = max(float(np.abs(b).max()), 1e-5)
        err = float(np.abs(a - b).max()) / scale
        assert err < 2e-3, f'{what}: max error {err:.2e} of the largest gradient entry'
    close(sub(feat.grad, g), g['d_feat_sub'], 'd_feat')
    n = 0
    for k, v in sd.items():
        if ('dparam.' + k) in g:
            close(sub(v.grad if v.grad is not None else torch.zeros_like(v), g), g['dparam.' + k], k)
            n += 1
    assert n == 6 + 34 * spec['num_layers'] + 22
