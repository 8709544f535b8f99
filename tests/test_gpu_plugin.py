"""GPU: the reference-facing module interface (registry-built heads) against the golden vectors
of the reference's own run -- the tests read like calls into mmdet3d_plugin."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

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
    with torch.no_grad():           # inference: mv2d_fpn_neck (with gradients enabled the neck runs as torch convolutions)
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


def test_query_generator_forward_on_its_own(state_dicts):
    """QueryGenerator.forward with the reference's signature (utils/query_generator.py:343-350) against the oracle's
    stage outputs: RoI features, K', extrinsics and the intrinsics feature in, centre in lidar coordinates out."""
    from oracle import mv2d_oracle as O
    spec = synth.CASES['s_small']
    h = head('S', state_dicts)
    feat, boxes, metas = synth.case_inputs(spec)
    with torch.no_grad():
        _, _, st = O.mv2d_s_forward(state_dicts(6), feat, boxes, metas, O.make_cfg('S'), return_stages=True)
    qg = h.query_generator
    qg.return_cfg['enc'] = True
    try:
        center, feats = qg(st['roi_feat'].cuda(), st['intrinsics'].cuda(), st['extrinsics'].cuda(),
                           extra_feats=dict(intrinsic=st['intrins_feat'].cuda()))
    finally:
        qg.return_cfg.pop('enc')
    close(center, st['center_lidar'].numpy(), 1e-3, 1e-4)
    close(feats['enc'], st['enc'].numpy(), 1e-4, 1e-4)


def test_box_head_forward_single_frame_dense_interface(state_dicts):
    """CrossAttentionBoxHead.forward / MV2DTransformer.forward as MV2DSHead calls them (mv2d_s_head.py:184-198): the
    gathered RoI features [N,M,C,7,7], their masks and position embeddings, one query per batch entry."""
    from oracle import mv2d_oracle as O
    spec = synth.CASES['s_small']
    h = head('S', state_dicts)
    feat, boxes, metas = synth.case_inputs(spec)
    with torch.no_grad():
        cls, box, st = O.mv2d_s_forward(state_dicts(6), feat, boxes, metas, O.make_cfg('S'), return_stages=True)
    corr, cmask = st['corr'], st['corr_mask']
    N, M = corr.shape
    x = st['roi_feat'][corr].cuda()                                  # [N,M,C,7,7]
    pos = st['roi_pe'][corr].cuda()
    masks = (~cmask)[:, :, None, None].expand(N, M, 7, 7).cuda()
    ref = st['ref'][:, None].cuda()
    c, b = h.bbox_head(ref, x, masks, pos)
    assert c.shape == (6, N, 1, 10)
    close(c[:, :, 0], cls.numpy())
    close(b[:, :, 0], box.numpy())
    q = h.bbox_head.position_embedding(ref)
    out_dec, memory = h.bbox_head.transformer(x, masks, q, pos)
    close(out_dec[:, :, 0], st['outs_dec'].numpy(), 1e-3, 1e-3)
    assert memory.shape == x.shape


def test_box_head_forward_two_frame_dense_interface(state_dicts):
    """... and as MV2DTHead calls it (mv2d_t_head.py:100-109): the feature views as one batch entry, a dense per-query
    cross-attention mask; the velocity / dt rescaling happens in the head, after this call."""
    from oracle import mv2d_oracle as O
    spec = synth.CASES['t_small']
    h = head('T', state_dicts)
    feat, boxes, metas = synth.case_inputs(spec)
    with torch.no_grad():
        cls, box, st = O.mv2d_t_forward(state_dicts(6), feat, boxes, metas, O.make_cfg('T'), return_stages=True)
    V, C, hh, ww = feat.shape
    pad = O.feat_masks(metas, hh, ww)                                  # [1,V,h,w]
    c, b = h.bbox_head(st['ref'][None].cuda(), feat[None].cuda(), pad.cuda(), st['pe'][None].cuda(),
                       cross_attn_mask=(~st['key_mask']).cuda())
    close(c[:, 0], cls.numpy())
    close(b[:, 0, :, :8], box[..., :8].numpy())
    close(b[:, 0, :, 8:] / 0.5, box[..., 8:].numpy())                  # timestamps 0 / 0.5 s (mv2d_t_head.py:130-142)


def test_detections_handoff_on_the_device(state_dicts):
    """Next row f2 (detectors/mv2d.py:60-117): mv2d_handoff_2d against the goldens of the reference's own
    process_2d_detections / complement_2d_gt, all five cases as the views of one call, bit-exact and in order."""
    import glob
    h = head('S', state_dicts)
    eng = h.engine()
    cases = [np.load(p) for p in sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'f2_*.npz')))]
    assert len(cases) >= 5
    thr = {float(c['thr']) for c in cases}
    for t in thr:       # one call per threshold value (a call has one threshold)
        grp = [c for c in cases if float(c['thr']) == t]
        dets = [torch.from_numpy(c['det_in']) for c in grp]
        gts = [torch.cat([torch.from_numpy(c['gt_boxes']), torch.ones(len(c['gt_labels']), 1),
                          torch.from_numpy(c['gt_labels']).float()[:, None]], 1) for c in grp]
        out = eng.handoff_2d(dets, gts, float(grp[0]['min_size']), t)
        for o, c in zip(out, grp):
            assert np.array_equal(o.cpu().numpy(), c['out'])
        only_filter = eng.handoff_2d(dets, None, float(grp[0]['min_size']), -1.0)
        for o, c in zip(only_filter, grp):
            assert np.array_equal(o.cpu().numpy(), c['det_filtered'])


def test_detector_forward_train_shell(state_dicts):
    """MV2D.forward_train (detectors/mv2d.py:129-213) around the hot path: per-view metas / ground truth, the hand-off
    on the device, roi_head.forward_train.  The 2D detector is injected (torch side of the north star)."""
    from mv2d_b200.plugin.modules import MV2D
    spec = synth.CASES['s_small']
    h = head('S', state_dicts)
    feat, boxes, metas = synth.case_inputs(spec)
    gt_boxes, gt_labels, _ = synth.make_dn_inputs(dict(num_gt=6, seed=5))
    det = MV2D.__new__(MV2D)
    torch.nn.Module.__init__(det)
    det.roi_head, det.neck, det.base_detector = h, None, None
    det.train_cfg = dict(detection_proposal=dict(min_bbox_size=8), complement_2d_gt=0.4)
    det.test_cfg = None
    V = len(metas)
    img = torch.zeros(1, V, 3, 8, 8, device='cuda')
    meta = {k: [m[k] for m in metas] for k in metas[0] if k != 'num_views'}
    gt2d = [[b[:2, :4].cuda() for b in boxes]]              # two 2D ground-truth boxes per view (copies of detections)
    gl2d = [[b[:2, 5].long().cuda() for b in boxes]]
    to3d = [[torch.tensor([v % 6, -1]) for v in range(V)]]
    h.train()
    try:
        losses = det.forward_train(img, [meta], gt2d, gl2d, to3d, [gt_boxes.cuda()], [gt_labels.cuda()],
                                   detector_out=(feat.cuda(), [b.cuda() for b in boxes], dict(loss_rpn=torch.zeros(()))))
    finally:
        h.eval()
    assert 'det_loss_rpn' in losses and all(f'l{i}.loss_cls' in losses and f'l{i}.loss_bbox' in losses for i in range(6))
    assert all(torch.isfinite(v).all() for v in losses.values())


def test_detector_shell_trains_through_the_neck(state_dicts):
    """With the one-level FPN neck in front (configs/mv2d/exp/*.py:32-39) MV2D.forward_train keeps the autograd chain:
    d loss / d feat of the CUDA backward flows through the neck (two torch convolutions in training mode) into the neck's
    parameters and the backbone's feature map; in eval mode the same module runs mv2d_fpn_neck and agrees with it."""
    from mv2d_b200.plugin.modules import MV2D, FPN
    spec = synth.CASES['s_small']
    h = head('S', state_dicts)
    feat, boxes, metas = synth.case_inputs(spec)
    gt_boxes, gt_labels, _ = synth.make_dn_inputs(dict(num_gt=6, seed=5))
    torch.manual_seed(3)
    neck = FPN([256], 256, 1).cuda()
    with torch.no_grad():
        for prm in neck.parameters():
            prm.mul_(0.5)
    det = MV2D.__new__(MV2D)
    torch.nn.Module.__init__(det)
    det.roi_head, det.neck, det.base_detector = h, neck, None
    det.train_cfg = dict(detection_proposal=dict(min_bbox_size=8), complement_2d_gt=0.4)
    det.test_cfg = None
    V = len(metas)
    x = feat.cuda().clone().requires_grad_(True)            # the backbone's P4 map
    monkey_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False                 # fp32 torch convolutions: comparable with the 3xTF32 kernel at 1e-3
    with torch.no_grad():
        cl_eval, _ = det.process_detector_feat(x)           # eval path: mv2d_fpn_neck
    cl_train, _ = det.process_detector_feat(x)              # training path: autograd
    assert cl_train.requires_grad
    assert torch.allclose(cl_train.detach(), cl_eval, atol=1e-3, rtol=1e-3)
    img = torch.zeros(1, V, 3, 8, 8, device='cuda')
    meta = {k: [m[k] for m in metas] for k in metas[0] if k != 'num_views'}
    gt2d = [[b[:2, :4].cuda() for b in boxes]]
    gl2d = [[b[:2, 5].long().cuda() for b in boxes]]
    to3d = [[torch.tensor([v % 6, -1]) for v in range(V)]]
    h.train()
    try:
        losses = det.forward_train(img, [meta], gt2d, gl2d, to3d, [gt_boxes.cuda()], [gt_labels.cuda()],
                                   detector_out=(x, [b.cuda() for b in boxes], None))
        sum(v.sum() for v in losses.values()).backward()
    finally:
        h.eval()
    for name, prm in neck.named_parameters():
        assert prm.grad is not None and torch.isfinite(prm.grad).all() and prm.grad.abs().max() > 0, name
    assert x.grad is not None and torch.isfinite(x.grad).all() and x.grad.abs().max() > 0
    # after an in-place parameter update the eval path re-packs the neck weights
    with torch.no_grad():
        for prm in neck.parameters():
            prm.add_(0.01)
        cl2, _ = det.process_detector_feat(x)
        ref = F.conv2d(F.conv2d(x, neck.lateral_convs[0].conv.weight, neck.lateral_convs[0].conv.bias),
                       neck.fpn_convs[0].conv.weight, neck.fpn_convs[0].conv.bias, padding=1).permute(0, 2, 3, 1)
    torch.backends.cudnn.allow_tf32 = monkey_tf32
    assert torch.allclose(cl2, ref, atol=1e-3, rtol=1e-3)


def test_two_frame_head_forward_train_runs_the_backward(state_dicts, monkeypatch):
    """MV2DTHead.forward_train (mv2d_s_head.py:236-307 as the two-frame head inherits it, denoising queries on): the loss
    dict of the reference's golden run and gradients through ``loss.backward()`` for every hot-path Parameter."""
    import json
    g = dict(np.load(golden_path('grad_t_dn')))
    spec = json.loads(bytes(g['spec']).decode())
    from mv2d_b200.plugin.build import build_roi_head
    h = build_roi_head(CFG['T'], device='cuda')
    sd = synth.make_state_dict(0, num_layers=6)
    h.load_state_dict(sd, strict=True)
    feat, boxes, metas = synth.case_inputs(spec)
    gt_boxes, gt_labels, rand = synth.make_dn_inputs(spec['dn'])
    monkeypatch.setattr(torch, 'rand', lambda *a, **k: rand.clone().to(k.get('device', 'cpu')))
    h.train()
    try:
        f = feat.cuda().requires_grad_(True)
        losses = h.forward_train([f], metas, [b.cuda() for b in boxes], None, None, None, None, [gt_boxes.cuda()], [gt_labels.cuda()])
        sum(losses.values()).backward()
    finally:
        h.eval()
    assert {f'l{i}.{k}' for i in range(6) for k in ('loss_cls', 'loss_bbox', 'dn_loss_cls', 'dn_loss_bbox')} == set(losses)
    assert all(torch.isfinite(v).all() for v in losses.values())
    assert f.grad is not None and torch.isfinite(f.grad).all() and float(f.grad.abs().max()) > 0
    named = dict(h.named_parameters())
    missing = [n for n in h.trainer().table if named[n].grad is None or not torch.isfinite(named[n].grad).all()]
    assert not missing, missing
