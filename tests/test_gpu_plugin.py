"""GPU: the reference-facing module interface (registry-built heads) against the golden vectors
of the reference's own run -- the tests read like calls into mmdet3d_plugin."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, golden_path
from mv2d_b200 import synth

pytestmark = pytest.mark.gpu
CFG = {'S': os.path.join(ROOT, 'configs', 'mv2d_b200', 'mv2d_s_r50_1408x512.py'),
       'T': os.path.join(ROOT, 'configs', 'mv2d_b200', 'mv2d_t_r50_1408x512.py')}
_HEADS = {}


def head(mode, state_dicts):
    from mv2d_b200.plugin.build import build_roi_head
    if mode not in _HEADS:
        h = build_roi_head(CFG[mode], device='cuda')
        h.load_state_dict(state_dicts(6), strict=True)
        _HEADS[mode] = h
    return _HEADS[mode]


def golden(name):
    g = dict(np.load(golden_path(name)))
    return json.loads(bytes(g.pop('spec')).decode()), g


def close(a, b, atol=1e-3, rtol=1e-3):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert (np.abs(a - b) <= atol + rtol * np.abs(b)).all(), f'max |d| = {np.abs(a - b).max():.3e}'


@pytest.mark.parametrize('name', ['s_small', 's_cfg2', 't_small'])
def test_head_bbox_forward_and_simple_test(name, state_dicts):
    spec, g = golden(name)
    h = head(spec['mode'], state_dicts)
    feat, boxes, metas = synth.case_inputs(spec)
    res = h._bbox_forward([feat.cuda()], [b.cuda() for b in boxes], metas)
    close(torch.stack(res['cls_scores']), g['cls_scores'])
    close(torch.stack(res['bbox_preds']), g['bbox_preds'])
    close(res['rois'], g['rois'], 0, 0)
    (b, s, l), = h.simple_test([feat.cuda()], [b.cuda() for b in boxes], metas)
    # decode on our own last-layer outputs: compare with the reference decode of ITS outputs
    assert b.shape == g['dec_boxes'].shape
    close(s, g['dec_scores'], 1e-3, 1e-3)


def test_training_mode_forward_with_denoising_queries(state_dicts, monkeypatch):
    """head.train() + use_denoise: _bbox_forward prepends the denoising queries built from img_metas[0]'s GT
    (mv2d_t_head.py:91-118) and returns the reference's dn_mask_dict; checked against the oracle with the
    same noise."""
    from oracle import mv2d_oracle as O
    spec = dict(synth.CASES['t_dn'], num_layers=6)
    h = head('T', state_dicts)
    feat, boxes, metas = synth.case_inputs(spec)
    gt_boxes, gt_labels, rand = synth.make_dn_inputs(spec['dn'])

    class Boxes:      # the two attributes read from LiDARInstance3DBoxes
        gravity_center = gt_boxes[:, :3].cuda()
        tensor = gt_boxes.cuda()
    metas = [dict(metas[0], gt_bboxes_3d=Boxes(), gt_labels_3d=gt_labels.cuda())] + list(metas[1:])
    monkeypatch.setattr(torch, 'rand', lambda *a, **k: rand.clone().to(k.get('device', 'cpu')))
    h.train()
    try:
        res = h._bbox_forward([feat.cuda()], [b.cuda() for b in boxes], metas)
    finally:
        h.eval()
    monkeypatch.undo()
    with torch.no_grad():
        cls, box, st = O.mv2d_t_forward(state_dicts(6), feat, boxes, metas, O.make_cfg('T'), return_stages=True,
                                        dn=dict(gt_boxes=gt_boxes, gt_labels=gt_labels, rand=rand))
    md = res['dn_mask_dict']
    assert md['pad_size'] == 70 and md['known_lbs_bboxes'][1].shape == (70, 9)
    assert torch.equal(md['known_lbs_bboxes'][0].cpu(), st['dn']['labels'])
    assert torch.equal(md['map_known_indice'].cpu(), torch.arange(70))
    close(torch.stack(res['cls_scores']), cls)
    close(torch.stack(res['bbox_preds']), box)
    close(md['output_known_lbs_bboxes'][0][:, 0], st['dn']['cls'])
    close(md['output_known_lbs_bboxes'][1][:, 0], st['dn']['box'])
    # eval mode: no denoising queries
    assert h._bbox_forward([feat.cuda()], [b.cuda() for b in boxes], metas)['dn_mask_dict'] is None


def test_submodule_interfaces(state_dicts):
    """PE.forward and BoxCorrelation.gen_* through the reference's call signatures."""
    spec, g = golden('s_small')
    h = head('S', state_dicts)
    feat, boxes, metas = synth.case_inputs(spec)
    pe = h.position_encoding([feat.cuda()], metas)[0]
    assert pe.shape == feat.shape
    close(pe.contiguous().flatten()[::251], g['pe_sub'], 3e-3, 1e-3)
    rois = torch.from_numpy(g['rois']).cuda()
    corr, mask = h.box_corr_module.gen_box_roi_correlation(rois, [len(b) for b in boxes], metas)
    ours = [set(int(c) for c, m in zip(cr, mr) if m) for cr, mr in zip(corr.cpu().numpy(), mask.cpu().numpy())]
    ref = [set(int(c) for c, m in zip(cr, mr) if m) for cr, mr in zip(g['corr'], g['corr_mask'])]
    assert ours == ref and corr.shape == g['corr'].shape
    # T: dense bool mask
    spec, g = golden('t_small')
    ht = head('T', state_dicts)
    feat, boxes, metas = synth.case_inputs(spec)
    rois = torch.from_numpy(g['rois']).cuda()
    km = ht.box_corr_module.gen_box_correlation(rois, [len(b) for b in boxes], metas, feat, 16)
    ref_bits = np.unpackbits(g['key_mask_packed'], axis=1)[:, :km[0].numel()]
    assert np.array_equal(km.view(km.shape[0], -1).cpu().numpy().astype(np.uint8), ref_bits)


def test_missing_library_fails_loudly(monkeypatch):
    """No fallback: if the .so is not there the product path raises."""
    from mv2d_b200 import lib
    monkeypatch.setattr(lib, '_lib', None)
    monkeypatch.setattr(lib, 'LIB_PATH', '/nonexistent/libmv2d_b200.so')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        lib.load()


def test_bbox_head_loss_interface_matches_reference_golden(state_dicts):
    """Row f3 through the reference-facing interface: bbox_head.loss(gt_boxes, gt_labels, {'cls_scores': [..],
    'bbox_preds': [..]}) per layer, as mv2d_s_head.py:281-286 calls it, and bbox_head.dn_loss_single, with the
    cost / loss weights read from the config (train_cfg.assigner, loss_cls, loss_bbox, code_weights)."""
    from test_loss_oracle_golden import load_loss_case
    from mv2d_b200.plugin.build import build_roi_head
    h = build_roi_head(CFG['S'], device='cuda', train=True)
    h.load_state_dict(state_dicts(6), strict=True)
    assert type(h.bbox_head.assigner).__name__ == 'HungarianAssigner3D' and h.stage_loss_weights == [0.1] * 6
    g, s, gt_boxes, gt_labels = load_loss_case('loss_s_dn')
    cls, box = torch.from_numpy(s['cls_scores']).cuda(), torch.from_numpy(s['bbox_preds']).cuda()

    class Boxes:        # LiDARInstance3DBoxes look-alike, as the reference passes it
        gravity_center = gt_boxes[:, :3].cuda()
        tensor = torch.cat([gt_boxes[:, :3], gt_boxes[:, 3:]], 1).cuda()
    for l in range(cls.shape[0]):
        d = h.bbox_head.loss([Boxes()], [gt_labels.cuda()], {'cls_scores': [cls[l]], 'bbox_preds': [box[l]]})
        close(d['loss_cls'], g['loss_cls'][l], 1e-6, 2e-5)
        close(d['loss_bbox'], g['loss_bbox'][l], 1e-6, 2e-5)
        pad = int(s['dn_pad'])
        known = gt_boxes.repeat(pad // gt_boxes.shape[0], 1).cuda()
        a, b = h.bbox_head.dn_loss_single(torch.from_numpy(s['dn_cls'][l]).cuda(), torch.from_numpy(s['dn_box'][l]).cuda(), known,
                                          torch.from_numpy(s['dn_labels']).cuda(), pad, None, float(g['dn_split']), neg_bbox_loss=False)
        close(a, g['dn_loss_cls'][l], 1e-6, 2e-5)
        close(b, g['dn_loss_bbox'][l], 1e-6, 2e-5)


def test_detector_shell_with_neck_matches_oracle(state_dicts):
    """DETECTORS.build(cfg.model) with the config's one-level FPN neck: a stub 2D detector hands over its five FPN
    levels + detections, MV2D.simple_test runs neck -> roi_head (detectors/mv2d.py:251-261, :122-127); compared with
    the oracle's neck + S forward + decode."""
    from mv2d_b200.config import Config
    from mv2d_b200.plugin.build import roi_head_cfg
    from mv2d_b200.registry import DETECTORS, build_from_cfg
    from oracle import mv2d_oracle as O
    cfg = Config.fromfile(CFG['S'])
    model = dict(cfg['model'])
    model['roi_head'] = roi_head_cfg(cfg)
    p4, boxes, metas = synth.make_sample(77, 6, 5)
    levels = [torch.zeros(6, 256, 8, 8), torch.zeros(6, 256, 8, 8), p4, torch.zeros(6, 256, 8, 8), torch.zeros(6, 256, 8, 8)]
    model['base_detector'] = lambda img, img_metas: ([l.cuda() for l in levels], [b.cuda() for b in boxes])
    det = build_from_cfg(model, DETECTORS).cuda()
    assert det.with_neck and type(det.neck).__name__ == 'FPN'
    sd, nsd = state_dicts(6), synth.make_neck_state_dict(9)
    det.roi_head.load_state_dict(sd, strict=True)
    det.neck.load_state_dict(nsd, strict=True)
    res, = det.simple_test(None, metas)
    got = res['pts_bbox']
    with torch.no_grad():
        cls, box = O.mv2d_s_forward(sd, O.fpn_neck(nsd, p4), boxes, metas, O.make_cfg('S'))
        rb, rs, rl = O.scene_nms(*O.nms_free_decode(cls[-1], box[-1], O.make_cfg('S')))
    # mv2d.py:266-287: grouped by class, descending score inside a class, on the host
    assert not got['scores_3d'].is_cuda and got['scores_3d'].shape == rs.shape
    assert torch.equal(got['labels_3d'], rl)
    close(got['scores_3d'], rs.numpy(), 1e-3, 1e-3)
    close(got['boxes_3d'], rb.numpy(), 2e-3, 2e-3)
    res = det.roi_head._bbox_forward([det.process_detector_feat([l.cuda() for l in levels])[0].permute(0, 3, 1, 2)],
                                     [x.cuda() for x in boxes], metas)
    close(torch.stack(res['cls_scores']), cls.numpy())
    close(torch.stack(res['bbox_preds']), box.numpy())


def test_scene_nms_matches_oracle(state_dicts):
    """mv2d_scene_nms (box3d_multiclass_nms at the configs' nms_thr = 1.0) incl. the > max_num branch and thresholds."""
    from mv2d_b200.engine import HotPath
    from oracle import mv2d_oracle as O
    eng = HotPath(state_dicts(6), mode='S')
    g = torch.Generator().manual_seed(3)
    for n, max_num, thr in ((300, 300, 0.0), (300, 120, 0.0), (57, 300, 0.4), (1, 300, 0.0), (0, 300, 0.0)):
        boxes, scores = torch.randn(n, 9, generator=g), torch.rand(n, generator=g)
        labels = torch.randint(0, 10, (n,), generator=g)
        b, s, l = eng.scene_nms(boxes.cuda(), scores.cuda(), labels.cuda(), score_thr=thr, max_num=max_num)
        rb, rs, rl = O.scene_nms(boxes, scores, labels, score_thr=thr, max_num=max_num)
        assert torch.equal(l.cpu(), rl) and torch.equal(s.cpu(), rs) and torch.equal(b.cpu(), rb)
    with pytest.raises(RuntimeError):
        eng.scene_nms(boxes.cuda(), scores.cuda(), labels.cuda(), nms_thr=0.5)
