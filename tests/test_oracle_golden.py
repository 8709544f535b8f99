"""The oracle (oracle/mv2d_oracle.py) against the golden vectors written by the REFERENCE's own
Python (oracle/make_golden.py).  CPU only.  Tolerance: the north-star gate
|out - ref| <= 1e-3 + 1e-3*|ref| is the pass bar for the product; the oracle itself must sit
an order of magnitude inside it (same arithmetic, different op order)."""
import json

import numpy as np
import pytest
import torch

from conftest import golden_path
from mv2d_b200 import synth
from oracle import mv2d_oracle as O

CASES = list(synth.CASES)
PE_SUB = 251


def load_golden(name):
    g = dict(np.load(golden_path(name)))
    spec = json.loads(bytes(g.pop('spec')).decode())
    return spec, g


def close(a, b, atol, rtol=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert err.max() <= 0, f'max |d| = {np.abs(a - b).max():.3e}'


def corr_as_sets(corr, mask):
    return [set(int(c) for c, m in zip(cr, mr) if m) for cr, mr in zip(corr, mask)]


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_golden(name, state_dicts):
    spec, g = load_golden(name)
    assert spec == json.loads(json.dumps(synth.CASES[name])), 'fixture is stale: re-run oracle.make_golden'
    sd = state_dicts(spec['num_layers'])
    feat, boxes, metas = synth.case_inputs(spec)
    cfg = O.make_cfg(spec['mode'], num_layers=spec['num_layers'])
    fn = O.mv2d_s_forward if spec['mode'] == 'S' else O.mv2d_t_forward
    kw = {}
    if 'dn' in spec:   # row a20: training-mode forward with denoising queries
        gt_boxes, gt_labels, rand = synth.make_dn_inputs(spec['dn'])
        kw['dn'] = dict(gt_boxes=gt_boxes, gt_labels=gt_labels, rand=rand)
    with torch.no_grad():
        cls, box, st = fn(sd, feat, boxes, metas, cfg, return_stages=True, **kw)
    if 'dn' in spec:
        assert st['dn']['cls'].shape[1] == int(g['dn_pad']) == 10 * spec['dn']['num_gt']
        close(st['dn']['cls'], g['dn_cls'], 1e-4, 1e-4)
        close(st['dn']['box'], g['dn_box'], 1e-4, 1e-4)
        assert np.array_equal(st['dn']['labels'].numpy(), g['dn_labels'])
    # stage level
    close(st['rois'], g['rois'], 0)
    close(st['pe'].flatten()[::PE_SUB], g['pe_sub'], 1e-4)
    close(st['roi_feat'].flatten()[::PE_SUB], g['roi_feat_sub'], 1e-5)
    close(st['center_lidar'], g['center_lidar'], 2e-4)
    if 'intrinsics' in g:
        close(st['intrinsics'], g['intrinsics'], 1e-9, 1e-12)
        close(st['extrinsics'], g['extrinsics'], 0)
        close(st['query_pos'], g['query_pos'], 1e-4)
        close(st['outs_dec'], g['outs_dec'], 2e-4)
    if spec['mode'] == 'S':
        assert corr_as_sets(st['corr'].numpy(), st['corr_mask'].numpy()) == \
            corr_as_sets(g['corr'], g['corr_mask'])
    else:
        N = st['key_mask'].shape[0]
        km = np.packbits(st['key_mask'].numpy().reshape(N, -1), axis=1)
        assert np.array_equal(km, g['key_mask_packed'])
    # final outputs: an order of magnitude inside the north-star tolerance
    close(cls, g['cls_scores'], 1e-4, 1e-4)
    close(box, g['bbox_preds'], 1e-4, 1e-4)
    # next row f1: NMS-free decode
    b, s, l = O.nms_free_decode(torch.from_numpy(g['cls_scores'][-1]),
                                torch.from_numpy(g['bbox_preds'][-1]), cfg)
    close(b, g['dec_boxes'], 1e-6)
    close(s, g['dec_scores'], 1e-7)
    assert np.array_equal(l.numpy(), g['dec_labels'])


def test_roi_align_restatement_vs_torchvision():
    """Independent pin of the RoIAlign restatement (mmcv semantics == torchvision aligned=True,
    sampling_ratio=0)."""
    from torchvision.ops import roi_align as tv
    feat, boxes, _ = synth.make_sample(9, 6, 6)
    boxes[0] = torch.cat([boxes[0], torch.tensor([[-20.0, -10.0, 90.0, 70.0, 1, 0],
                                                  [1300.0, 400.0, 1500.0, 600.0, 1, 0],
                                                  [10.0, 10.0, 12.0, 13.0, 1, 0]])])
    rois = O.bbox2roi(boxes)
    a = O.roi_align(feat, rois, 7, 1 / 16)
    b = tv(feat, rois, (7, 7), 1 / 16, 0, True)
    close(a, b, 2e-5)


def test_s_gather_attention_equals_dense_masked(state_dicts):
    """Invariant the CUDA S-path relies on: attending to the gathered RoI tokens of the matched
    RoIs (reference formulation, duplicated keys + key_padding_mask) equals attention over the
    UNIQUE token set with a per-query mask."""
    sd = state_dicts(1)
    spec = dict(synth.CASES['s_small'], num_layers=1)
    feat, boxes, metas = synth.case_inputs(spec)
    cfg = O.make_cfg('S', num_layers=1)
    with torch.no_grad():
        cls, box, st = O.mv2d_s_forward(sd, feat, boxes, metas, cfg, return_stages=True)
        N = st['rois'].shape[0]
        mem = st['roi_feat'].permute(0, 2, 3, 1).reshape(N * 49, 1, 256)
        pos = st['roi_pe'].permute(0, 2, 3, 1).reshape(N * 49, 1, 256)
        vis = torch.zeros((N, N), dtype=torch.bool)
        for n, (cr, mr) in enumerate(zip(st['corr'], st['corr_mask'])):
            vis[n, cr[mr]] = True
        cross = ~vis[:, :, None].expand(N, N, 49).reshape(N, N * 49)
        outs = O.decoder(sd, st['query_pos'][:, None], mem, pos, cfg, cross_mask=cross)
        cls2, box2 = O.branches(sd, outs, st['ref'][:, None], cfg)
    close(cls2.flatten(1, 2), cls, 2e-5)
    close(box2.flatten(1, 2), box, 2e-4)
