"""Batches as a segment dimension (the reference asserts B == 1: detectors/mv2d.py:143, roi_heads/mv2d_head.py:210,251).

B samples go through ONE kernel chain (``HotPath.forward_batch``); every sample must reproduce its own golden vector
(written by the reference's unmodified Python, one sample at a time) within the parity gate, and the integer work
(RoI match lists, per-query key masks) must stay bit-exact.
"""
import numpy as np
import pytest
import torch

from mv2d_b200 import synth
from test_gpu_parity import assert_close, engine, load_golden

pytestmark = pytest.mark.gpu

GROUPS = {
    's2': ['s_small', 's_cfg2'],                          # ragged: 27 and 300 queries, an empty view
    's4': ['s_cfg2', 's_small', 's_small', 's_cfg2'],
    's8': ['s_cfg2', 's_small', 's_cfg2', 's_cfg2', 's_small', 's_cfg2', 's_cfg2', 's_cfg2'],     # the benched batch size: 8 x <= 300 queries (persistent GEMMs, one-stream front end)
    's3_masks': ['s_empty', 's_pad', 's_one'],            # zero detections, padded images (general sine branch), one query
    's2_pad': ['s_pad', 's_pad'],                         # identical padding masks: shared sine branch on the general path
    't2': ['t_small', 't_cfg3'],                          # BASELINE configs[2]: the two-frame head at bs = 2
    't2_pad': ['t_pad', 't_pad'],
}


def run_group(names, state_dicts, **kw):
    specs = [load_golden(n) for n in names]
    mode, L = specs[0][0]['mode'], specs[0][0]['num_layers']
    assert all(s['mode'] == mode and s['num_layers'] == L for s, _ in specs)
    eng = engine(mode, L, state_dicts)
    ins = [synth.case_inputs(s) for s, _ in specs]
    feats = torch.stack([i[0] for i in ins], 0).cuda()
    out = eng.forward_batch(feats, [i[1] for i in ins], [i[2] for i in ins], **kw)
    torch.cuda.synchronize()
    return eng, specs, ins, out


@pytest.mark.parametrize('group', list(GROUPS))
def test_batch_reproduces_per_sample_goldens(group, state_dicts):
    eng, specs, ins, out = run_group(GROUPS[group], state_dicts)
    Np = out['Np']
    for b, ((spec, g), smp) in enumerate(zip(specs, out['samples'])):
        n = g['rois'].shape[0]
        assert smp['N'] == n
        r = smp['rois'].cpu().numpy()
        assert np.array_equal(r[:, 1:], g['rois'][:, 1:]) and np.array_equal(r[:, 0] - b * len(ins[b][2]), g['rois'][:, 0])
        assert_close(smp['cls_scores'], g['cls_scores'], what=f'{group}[{b}] cls_scores')
        assert_close(smp['bbox_preds'], g['bbox_preds'], what=f'{group}[{b}] bbox_preds')
        assert_close(smp['center_lidar'], g['center_lidar'], 1e-3, 1e-4, 'center_lidar')
        if spec['mode'] == 'S':
            m, c = smp['match'].cpu().numpy(), smp['match_cnt'].cpu().numpy()
            got = [set(int(x) - b * Np for x in m[i, :c[i]]) for i in range(n)]
            ref = [set(int(cc) for cc, mm in zip(cr, mr) if mm) for cr, mr in zip(g['corr'], g['corr_mask'])]
            assert got == ref, 'RoI match lists differ'
        else:
            from mv2d_b200.engine import feat_pad_mask
            words = out['keymask'][b * Np:b * Np + n].cpu().numpy().view(np.uint32)
            bits = np.unpackbits(words.view(np.uint8), axis=1, bitorder='little')[:, :g['key_mask_packed'].shape[1] * 8]
            ref_bits = np.unpackbits(g['key_mask_packed'], axis=1)
            h, w = ins[b][0].shape[-2:]
            keep = 1 - feat_pad_mask(ins[b][2], h, w).reshape(1, -1)
            assert np.array_equal(bits[:, :ref_bits.shape[1]], ref_bits * keep), 'per-query key masks differ'


@pytest.mark.parametrize('mode', ['S', 'T'])
def test_batch_equals_one_at_a_time(mode, state_dicts):
    """Fresh jittered-camera samples, B = 4 (S) / 2 (T): the batch against the same engine fed one sample at a time."""
    eng = engine(mode, 6, state_dicts)
    V, B = (6, 4) if mode == 'S' else (12, 2)
    ins = [synth.make_sample(60 + i, V, [3 + i, 5, 0, 4, 2 + 2 * i, 6] * (V // 6), cam_jitter_deg=3.0) for i in range(B)]
    singles = []
    for f, boxes, metas in ins:
        o = eng.forward(f.cuda(), boxes, metas)
        singles.append((o['cls_scores'].clone(), o['bbox_preds'].clone(), o['ref'].clone()))
    out = eng.forward_batch(torch.stack([i[0] for i in ins], 0).cuda(), [i[1] for i in ins], [i[2] for i in ins])
    for smp, (cls, box, ref) in zip(out['samples'], singles):
        assert_close(smp['ref'], ref, 1e-6, 1e-6, 'ref')
        assert_close(smp['cls_scores'], cls, 1e-4, 1e-4, 'cls_scores')
        assert_close(smp['bbox_preds'], box, 1e-4, 1e-4, 'bbox_preds')


def test_batch_graph_replay_matches_eager(state_dicts):
    eng, specs, ins, out = run_group(GROUPS['s2'], state_dicts)
    eager = [(s['cls_scores'].clone(), s['bbox_preds'].clone()) for s in out['samples']]
    feats = torch.stack([i[0] for i in ins], 0).cuda()
    for _ in range(2):      # capture, then a pure replay
        o = eng.forward_batch(feats, [i[1] for i in ins], [i[2] for i in ins], use_graph=True)
    torch.cuda.synchronize()
    for smp, (cls, box) in zip(o['samples'], eager):
        assert torch.equal(smp['cls_scores'], cls) and torch.equal(smp['bbox_preds'], box)


def test_bucketed_graphs_serve_a_stream_of_changing_detection_counts(state_dicts):
    """real detections give a different N nearly every sample.  With bucket = 32 the query rows are
    padded to a multiple of 32 (device-side count), so N in [50, 450] needs at most 14 captured graphs (the buckets
    64, 96, ..., 480) instead of up to 40, and every result equals the eager, unpadded one."""
    from mv2d_b200.engine import HotPath
    eng = HotPath(state_dicts(2), mode='S')
    ref = engine('S', 2, state_dicts)
    rng = np.random.Generator(np.random.PCG64(5))
    for i in range(40):
        n = int(rng.integers(50, 451))
        per = [n // 6 + (1 if v < n % 6 else 0) for v in range(6)]
        f, boxes, metas = synth.make_sample(200 + i, 6, per)
        o = eng.forward(f.cuda(), boxes, metas, use_graph=True, bucket=32)
        e = ref.forward(f.cuda(), boxes, metas)
        assert o['N'] == n
        assert_close(o['cls_scores'], e['cls_scores'], 1e-4, 1e-4, 'cls_scores')
        assert_close(o['bbox_preds'], e['bbox_preds'], 1e-4, 1e-4, 'bbox_preds')
    assert len(eng._graphs_b) <= 14, len(eng._graphs_b)
