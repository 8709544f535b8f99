"""Row f3 on the GPU: mv2d_loss (cost matrix, device-side linear sum assignment, focal / L1 / denoising losses)
against the golden vectors written by the reference's own assigner and loss code, and against the oracle."""
import numpy as np
import pytest
import torch

from mv2d_b200 import synth
from test_loss_oracle_golden import LOSS_CASES, load_loss_case

pytestmark = pytest.mark.gpu
RTOL = 2e-5


def _engine(mode, state_dicts, L):
    from mv2d_b200.engine import HotPath
    return HotPath(state_dicts(L), mode=mode)


def close(a, b, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.all(np.abs(a - b) <= RTOL * np.abs(b) + 1e-6), f'{what}: {a} vs {b}'


@pytest.mark.parametrize('name', LOSS_CASES)
def test_loss_matches_reference_golden(name, state_dicts):
    g, s, gt_boxes, gt_labels = load_loss_case(name)
    cls, box = torch.from_numpy(s['cls_scores']).cuda(), torch.from_numpy(s['bbox_preds']).cuda()
    mode = 'T' if '_t_' in name else 'S'
    eng = _engine(mode, state_dicts, cls.shape[0])
    kw = {}
    if 'dn_loss_cls' in g:
        kw = dict(dn_cls=torch.from_numpy(s['dn_cls']).cuda(), dn_box=torch.from_numpy(s['dn_box']).cuda(),
                  dn_labels=torch.from_numpy(s['dn_labels']).cuda())
    out = eng.loss(cls, box, gt_boxes, gt_labels, **kw)
    torch.cuda.synchronize()
    assert np.array_equal(out['assigned'].cpu().numpy(), g['assigned']), 'Hungarian assignment differs from the reference'
    close(out['loss_cls'].cpu(), g['loss_cls'], 'loss_cls')
    close(out['loss_bbox'].cpu(), g['loss_bbox'], 'loss_bbox')
    if 'dn_loss_cls' in g:
        close(out['dn_loss_cls'].cpu(), g['dn_loss_cls'], 'dn_loss_cls')
        close(out['dn_loss_bbox'].cpu(), g['dn_loss_bbox'], 'dn_loss_bbox')


@pytest.mark.parametrize('N,G,seed', [(300, 100, 1), (900, 64, 2), (37, 90, 3), (300, 0, 4), (1, 1, 5)])
def test_assignment_and_loss_match_oracle_on_random_predictions(N, G, seed, state_dicts):
    """Sizes beyond the goldens (N up to 900, G up to 100, G > N, no GT): device LSA vs scipy through the oracle."""
    from oracle import mv2d_oracle as O
    g = torch.Generator().manual_seed(seed)
    L = 3
    cls = torch.randn(L, N, 10, generator=g) * 2 - 2
    box = torch.randn(L, N, 10, generator=g)
    gt = torch.randn(G, 9, generator=g)
    gt[:, 3:6] = gt[:, 3:6].abs() + 0.3
    labels = torch.randint(0, 10, (G,), generator=g)
    eng = _engine('S', state_dicts, 6)
    big = torch.zeros(L, N + 5, 10).cuda()          # strided view: rows contiguous, layer stride larger
    big[:, 5:] = cls.cuda()
    out = eng.loss(big[:, 5:], box.cuda(), gt, labels)
    torch.cuda.synchronize()
    for l in range(L):
        lc, lb, asg = O.loss_single(cls[l], box[l], gt, labels)
        assert np.array_equal(out['assigned'][l].cpu().numpy(), asg.numpy()), f'layer {l}: assignment differs'
        close(out['loss_cls'][l].cpu(), float(lc), 'loss_cls')
        close(out['loss_bbox'][l].cpu(), float(lb), 'loss_bbox')


@pytest.mark.parametrize('name', ['s_dn', 't_dn'])
def test_forward_losses_end_to_end(name, state_dicts):
    """Training-mode forward (denoising queries) + losses on the device against the reference goldens: the CUDA
    decoder outputs differ from the reference's by ~1e-4, so the losses agree to ~1e-3 relative; the Hungarian
    assignment must still be identical."""
    import json
    from conftest import golden_path
    s = dict(np.load(golden_path(name)))
    g = dict(np.load(golden_path('loss_' + name)))
    spec = json.loads(bytes(s['spec']).decode())
    eng = _engine(spec['mode'], state_dicts, spec['num_layers'])
    feat, boxes, metas = synth.case_inputs(spec)
    gt_boxes, gt_labels, rand = synth.make_dn_inputs(spec['dn'])
    losses, out, raw = eng.forward_losses(feat.cuda(), [b.cuda() for b in boxes], metas, gt_boxes, gt_labels, rand=rand,
                                          use_denoise=True)
    torch.cuda.synchronize()
    assert np.array_equal(raw['assigned'].cpu().numpy(), g['assigned'])
    for k, ref in (('loss_cls', g['loss_cls']), ('loss_bbox', g['loss_bbox']), ('dn_loss_cls', g['dn_loss_cls']),
                   ('dn_loss_bbox', g['dn_loss_bbox'])):
        got = raw[k].cpu().numpy().astype(np.float64)
        assert np.all(np.abs(got - ref) <= 2e-3 * np.abs(ref) + 1e-5), (k, got, ref)
    assert set(losses) == {f'l{i}.{k}' for i in range(spec['num_layers']) for k in ('loss_cls', 'loss_bbox', 'dn_loss_cls', 'dn_loss_bbox')}
    assert abs(float(losses['l0.loss_cls']) - 0.1 * g['loss_cls'][0]) <= 2e-3 * abs(0.1 * g['loss_cls'][0])
