"""GPU parity tests proper: the CUDA path, called through the C ABI (mv2d_b200.engine ->
ctypes -> libmv2d_b200.so), against (1) the golden vectors written by the reference's own
Python and (2) the oracle on fresh seeded inputs.

Tolerance (north star): |out - ref| <= 1e-3 + 1e-3*|ref| on cls_scores / bbox_preds of all
layers.  Integer / index work (RoI match lists, key masks) must be bit-exact.
"""
import json

import numpy as np
import pytest
import torch

from conftest import golden_path
from mv2d_b200 import synth

pytestmark = pytest.mark.gpu
PE_SUB = 251
ATOL = RTOL = 1e-3


def load_golden(name):
    g = dict(np.load(golden_path(name)))
    spec = json.loads(bytes(g.pop('spec')).decode())
    return spec, g


def assert_close(a, b, atol=ATOL, rtol=RTOL, what=''):
    a = np.asarray(a.detach().cpu() if torch.is_tensor(a) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu() if torch.is_tensor(b) else b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.isfinite(a).all(), f'{what}: non-finite values'
    bad = np.abs(a - b) > atol + rtol * np.abs(b)
    assert not bad.any(), f'{what}: {bad.sum()}/{bad.size} out of tol, max |d| = {np.abs(a - b).max():.3e}'


_ENGINES = {}


def engine(mode, num_layers, state_dicts):
    from mv2d_b200.engine import HotPath
    key = (mode, num_layers)
    if key not in _ENGINES:
        _ENGINES[key] = HotPath(state_dicts(num_layers), mode=mode)
    return _ENGINES[key]


def match_sets(out):
    m, c = out['match'].cpu().numpy(), out['match_cnt'].cpu().numpy()
    return [set(int(x) for x in m[i, :c[i]]) for i in range(len(c))]


@pytest.mark.parametrize('name', list(synth.CASES))
def test_cuda_path_matches_reference_golden(name, state_dicts):
    spec, g = load_golden(name)
    eng = engine(spec['mode'], spec['num_layers'], state_dicts)
    feat, boxes, metas = synth.case_inputs(spec)
    dn = None
    if 'dn' in spec:     # row a20: training-mode forward, denoising queries prepended
        gt_boxes, gt_labels, rand = synth.make_dn_inputs(spec['dn'])
        dn = dict(gt_boxes=gt_boxes, gt_labels=gt_labels, rand=rand)
    out = eng.forward(feat.cuda(), [b.cuda() for b in boxes], metas, dn=dn)
    torch.cuda.synchronize()
    N = out['N']
    assert N == g['rois'].shape[0]
    if dn is not None:
        from oracle import mv2d_oracle as O
        pad = int(g['dn_pad'])
        assert out['dn_pad'] == pad
        assert np.array_equal(out['dn_labels'].cpu().numpy(), g['dn_labels'])
        ref_all, sa_mask, _, _ = O.prepare_for_dn(out['ref'].cpu(), gt_boxes, gt_labels, rand,
                                                  O.make_cfg(spec['mode']))
        assert np.array_equal(out['dn_attn_mask'].cpu().numpy().astype(bool), sa_mask.numpy())
        assert_close(out['dn_ref'], ref_all[:pad], 1e-6, 0, 'dn reference points')
        assert_close(out['dn_cls_scores'], g['dn_cls'], what='dn cls_scores')
        assert_close(out['dn_bbox_preds'], g['dn_box'], what='dn bbox_preds')
    assert np.array_equal(out['rois'].cpu().numpy(), g['rois'])
    # ---- stage level
    pe = out['pe'].permute(0, 3, 1, 2).contiguous()          # back to NCHW for the comparison
    assert_close(pe.flatten()[::PE_SUB], g['pe_sub'], 3e-3, 1e-3, 'pe')     # TF32-tolerant stage
    tok = out['tok_feat'].view(N, 7, 7, 256).permute(0, 3, 1, 2).contiguous()
    assert_close(tok.flatten()[::PE_SUB], g['roi_feat_sub'], 1e-5, 1e-5, 'roi_feat')
    assert_close(out['center_lidar'], g['center_lidar'], 1e-3, 1e-4, 'center_lidar')
    if 'intrinsics' in g:
        assert_close(out['roi_intrinsics'].view(N, 4, 4), g['intrinsics'], 1e-9, 1e-12, 'roi_intrinsics')
        assert_close(out['query_pos'], g['query_pos'], 1e-3, 1e-3, 'query_pos')
        assert_close(out['outs_dec'], g['outs_dec'], 1e-3, 1e-3, 'outs_dec')
    if spec['mode'] == 'S':
        ref_sets = [set(int(c) for c, m in zip(cr, mr) if m) for cr, mr in zip(g['corr'], g['corr_mask'])]
        assert match_sets(out) == ref_sets, 'RoI match lists differ'
    else:
        words = out['keymask'].cpu().numpy().view(np.uint32)
        bits = np.unpackbits(words.view(np.uint8), axis=1, bitorder='little')[:, :g['key_mask_packed'].shape[1] * 8]
        ref_bits = np.unpackbits(g['key_mask_packed'], axis=1)
        # the library folds the key_padding_mask (padded-out cells, mv2d_t_head.py:68-76) into
        # the per-query mask; the reference applies it separately inside MultiheadAttention
        from mv2d_b200.engine import feat_pad_mask
        h, w = feat.shape[-2:]
        keep = 1 - feat_pad_mask(metas, h, w).reshape(1, -1)
        assert np.array_equal(bits[:, :ref_bits.shape[1]], ref_bits * keep), 'per-query key masks differ'
    # ---- the parity gate
    assert_close(out['cls_scores'], g['cls_scores'], what='cls_scores')
    assert_close(out['bbox_preds'], g['bbox_preds'], what='bbox_preds')
    # ---- next row f1: decode on the device
    b, s, l = eng.decode(torch.from_numpy(g['cls_scores'][-1]).cuda(), torch.from_numpy(g['bbox_preds'][-1]).cuda())
    assert_close(s, g['dec_scores'], 1e-6, 0, 'decode scores')
    assert_close(b, g['dec_boxes'], 1e-4, 1e-5, 'decode boxes')
    assert np.array_equal(l.cpu().numpy(), g['dec_labels'])


@pytest.mark.parametrize('mode,V,n,seed', [('S', 6, 7, 31), ('S', 6, 12, 32), ('T', 12, 5, 33)])
def test_cuda_path_matches_oracle_on_fresh_inputs(mode, V, n, seed, state_dicts):
    """Jittered cameras, different seeds: CUDA vs the oracle run here on the host CPU."""
    from oracle import mv2d_oracle as O
    eng = engine(mode, 6, state_dicts)
    feat, boxes, metas = synth.make_sample(seed, V, n, cam_jitter_deg=4.0)
    fn = O.mv2d_s_forward if mode == 'S' else O.mv2d_t_forward
    with torch.no_grad():
        cls, box, st = fn(state_dicts(6), feat, boxes, metas, O.make_cfg(mode), return_stages=True)
    out = eng.forward(feat.cuda(), boxes, metas)      # host-resident boxes: the other upload path
    assert_close(out['ref'], st['ref'], 1e-5, 1e-4, 'ref')
    assert_close(out['cls_scores'], cls, what='cls_scores')
    assert_close(out['bbox_preds'], box, what='bbox_preds')


@pytest.mark.parametrize('name', ['s_small', 't_small'])
def test_persistent_decoder_matches_reference_golden(name, state_dicts):
    """The single-launch (cooperative, device-wide barriers) decoder against the same goldens."""
    from mv2d_b200.engine import HotPath
    spec, g = load_golden(name)
    eng = HotPath(state_dicts(spec['num_layers']), mode=spec['mode'], persistent_decoder=True)
    feat, boxes, metas = synth.case_inputs(spec)
    before = eng.launch_count()
    out = eng.forward(feat.cuda(), boxes, metas)
    torch.cuda.synchronize()
    assert_close(out['cls_scores'], g['cls_scores'], what='cls_scores')
    assert_close(out['bbox_preds'], g['bbox_preds'], what='bbox_preds')
    assert eng.launch_count() - before < 40   # PE + query generator + ONE decoder launch


@pytest.mark.parametrize('mode,V,per_view', [('S', 6, 75), ('T', 12, 75)])
def test_maximum_sizes_match_oracle(mode, V, per_view, state_dicts):
    """The 2D detector's cap is 75 boxes per view (exp configs, max_per_img=75): 450 / 900 queries."""
    from oracle import mv2d_oracle as O
    eng = engine(mode, 6, state_dicts)
    feat, boxes, metas = synth.make_sample(41, V, per_view)
    out = eng.forward(feat.cuda(), boxes, metas)
    torch.cuda.synchronize()
    assert out['N'] == V * per_view
    fn = O.mv2d_s_forward if mode == 'S' else O.mv2d_t_forward
    with torch.no_grad():
        cls, box = fn(state_dicts(6), feat, boxes, metas, O.make_cfg(mode))
    assert_close(out['cls_scores'], cls, what='cls_scores')
    assert_close(out['bbox_preds'], box, what='bbox_preds')


def test_gemm_kernel_vs_torch_fp32():
    """The FFMA GEMM against torch fp32 (fp64-accumulated reference for the bound)."""
    import ctypes as C
    from mv2d_b200 import lib
    h = lib.load()
    g = torch.Generator().manual_seed(0)
    for (M, N, K, flags) in [(300, 256, 256, 0), (300, 2048, 256, 1), (37, 512, 1040, 1), (14700, 256, 2048, 0),
                             (1, 256, 384, 1), (16896, 1024, 192, 1), (129, 768, 256, 0)]:
        A = torch.randn(M, K, generator=g).cuda()
        W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
        b = torch.randn(N, generator=g).cuda()
        Cc = torch.empty(M, N, device='cuda')
        lib.check(h.mv2d_gemm(A.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), Cc.data_ptr(), N, M, N, K, flags,
                              lib.stream_ptr()), 'mv2d_gemm')
        ref = A.double() @ W.double().T + b.double()
        if flags & 1:
            ref = ref.relu()
        assert_close(Cc, ref, 2e-5, 2e-5, f'gemm {M}x{N}x{K}')


def test_tcgen05_gemm_single_pass_tf32():
    """tcgen05 kind::tf32 kernel: with TF32-representable operands every product is exact, so the
    result must match an fp64 matmul of the same operands to fp32-accumulation accuracy."""
    from mv2d_b200 import lib
    from mv2d_b200.pack import round_tf32
    h = lib.load()
    g = torch.Generator().manual_seed(1)
    for (M, N, K, flags) in [(128, 128, 32, 16), (256, 128, 64, 16), (16896, 1024, 192, 16 | 1), (1000, 256, 1024, 16),
                             (16896, 256, 1024, 16), (300, 256, 256, 16 | 1)]:
        A = round_tf32(torch.randn(M, K, generator=g)).cuda()
        W = round_tf32(torch.randn(N, K, generator=g) / K ** 0.5).cuda()
        b = torch.randn(N, generator=g).cuda()
        Cc = torch.full((M, N), float('nan'), device='cuda')
        lib.check(h.mv2d_gemm(A.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), Cc.data_ptr(), N, M, N, K, flags,
                              lib.stream_ptr()), 'mv2d_gemm(tc)')
        ref = A.double() @ W.double().T + b.double()
        if flags & 1:
            ref = ref.relu()
        assert_close(Cc, ref, 2e-5, 2e-5, f'gemm_tc {M}x{N}x{K}')


def test_first_layer_fold_equals_unfolded(state_dicts):
    """Layer 0's self-attention as a packed constant (default) against the decoder that runs it."""
    from mv2d_b200.engine import HotPath
    for name in ('s_small', 't_dn'):
        spec = synth.CASES[name]
        feat, boxes, metas = synth.case_inputs(spec)
        dn = None
        if 'dn' in spec:
            gt_boxes, gt_labels, rand = synth.make_dn_inputs(spec['dn'])
            dn = dict(gt_boxes=gt_boxes, gt_labels=gt_labels, rand=rand)
        outs = []
        for fold in (True, False):
            eng = HotPath(state_dicts(spec['num_layers']), mode=spec['mode'], fold_first_self_attn=fold)
            n0 = eng.launch_count()
            o = eng.forward(feat.cuda(), boxes, metas, dn=dn)
            torch.cuda.synchronize()
            outs.append((o['cls_scores'].clone(), o['bbox_preds'].clone(), eng.launch_count() - n0))
        assert outs[1][2] - outs[0][2] == 3          # in_proj, attention, out_proj launches dropped
        assert_close(outs[0][0], outs[1][0], 2e-5, 1e-5, 'cls fold/unfold')
        assert_close(outs[0][1], outs[1][1], 2e-5, 1e-5, 'box fold/unfold')


def test_cluster_multicast_gemm_variant():
    """MV2D_TC_MULTICAST=1 (read once per process, hence the subprocess) routes the decoder's BN=64 3xTF32 GEMMs
    through the 4-CTA-cluster kernel whose A tile is TMA-multicast; same goldens, same gate."""
    import os
    import subprocess
    import sys
    code = (
        "import json, numpy as np, torch\n"
        "from mv2d_b200 import synth\n"
        "from mv2d_b200.engine import HotPath\n"
        "g = dict(np.load('tests/golden/s_cfg2.npz')); spec = json.loads(bytes(g.pop('spec')).decode())\n"
        "eng = HotPath(synth.make_state_dict(0, num_layers=spec['num_layers']), mode='S')\n"
        "feat, boxes, metas = synth.case_inputs(spec)\n"
        "out = eng.forward(feat.cuda(), boxes, metas); torch.cuda.synchronize()\n"
        "for k in ('cls_scores', 'bbox_preds'):\n"
        "    a, b = out[k].cpu().numpy().astype(np.float64), g[k].astype(np.float64)\n"
        "    assert np.isfinite(a).all() and (np.abs(a - b) <= 1e-3 + 1e-3 * np.abs(b)).all(), k\n"
        "print('MC_OK')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-c', code], cwd=root, env=dict(os.environ, MV2D_TC_MULTICAST='1'),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'MC_OK' in r.stdout, r.stderr[-2000:]


def test_fused_projection_layernorm_variant():
    """MV2D_GEMM_LN=1 (opt-in, read once per process): out_proj / FFN-2 + residual + LayerNorm as ONE cluster launch --
    split-K over the cluster, partial tiles summed over distributed shared memory (csrc/gemm_ln.cu).  Same goldens, same
    gate, for the S head (cluster of 8 and 2), the T head's key-stationary form and a batch."""
    import os
    import subprocess
    import sys
    code = (
        "import json, numpy as np, torch\n"
        "from mv2d_b200 import synth\n"
        "from mv2d_b200.engine import HotPath\n"
        "for name in ('s_cfg2', 's_small', 't_cfg3'):\n"
        "    g = dict(np.load(f'tests/golden/{name}.npz')); spec = json.loads(bytes(g.pop('spec')).decode())\n"
        "    eng = HotPath(synth.make_state_dict(0, num_layers=spec['num_layers']), mode=spec['mode'])\n"
        "    feat, boxes, metas = synth.case_inputs(spec)\n"
        "    n0 = eng.launch_count()\n"
        "    out = eng.forward(feat.cuda(), boxes, metas); torch.cuda.synchronize()\n"
        "    print(name, 'launches', eng.launch_count() - n0)\n"
        "    for k in ('cls_scores', 'bbox_preds'):\n"
        "        a, b = out[k].cpu().numpy().astype(np.float64), g[k].astype(np.float64)\n"
        "        assert np.isfinite(a).all() and (np.abs(a - b) <= 1e-3 + 1e-3 * np.abs(b)).all(), (name, k)\n"
        "    if name == 's_cfg2':\n"
        "        outs = eng.forward_batch(torch.stack([feat, feat], 0).cuda(), [boxes, boxes], [metas, metas]); torch.cuda.synchronize()\n"
        "        for o in outs['samples']:\n"
        "            for k in ('cls_scores', 'bbox_preds'):\n"
        "                a, b = o[k].cpu().numpy().astype(np.float64), g[k].astype(np.float64)\n"
        "                assert (np.abs(a - b) <= 1e-3 + 1e-3 * np.abs(b)).all(), ('batch', k)\n"
        "print('GL_OK')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-c', code], cwd=root, env=dict(os.environ, MV2D_GEMM_LN='1'),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'GL_OK' in r.stdout, (r.stdout[-500:], r.stderr[-2000:])


def test_tcgen05_gemm_3xtf32_and_im2col(state_dicts):
    """Error-compensated 3xTF32: fp32-grade on arbitrary fp32 operands; and the TMA-im2col form
    against torch conv2d (fp64)."""
    import torch.nn.functional as F
    from mv2d_b200 import lib
    h = lib.load()
    g = torch.Generator().manual_seed(2)

    def split(x):
        hi, lo = torch.empty_like(x), torch.empty_like(x)
        lib.check(h.mv2d_split_tf32(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), lib.stream_ptr()), 'split')
        return hi, lo

    for (M, N, K) in [(128, 128, 32), (777, 256, 2304), (14700, 256, 512)]:
        A = torch.randn(M, K, generator=g).cuda()
        W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
        b = torch.randn(N, generator=g).cuda()
        (ah, al), (wh, wl) = split(A), split(W)
        Cc = torch.full((M, N), float('nan'), device='cuda')
        lib.check(h.mv2d_gemm_3xtf32(ah.data_ptr(), al.data_ptr(), K, wh.data_ptr(), wl.data_ptr(), K, b.data_ptr(),
                                     Cc.data_ptr(), N, M, N, K, 0, lib.stream_ptr()), 'gemm_3xtf32')
        # fp32 accumulation inside the tensor core (K up to 2304, three products per step)
        assert_close(Cc, A.double() @ W.double().T + b.double(), 1e-4, 1e-4, f'3xtf32 {M}x{N}x{K}')
        # plain fp32 operands, split into hi/lo inside the kernel (lo pointers NULL): same arithmetic, same bits
        Cr = torch.full((M, N), float('nan'), device='cuda')
        lib.check(h.mv2d_gemm_3xtf32(A.data_ptr(), None, K, W.data_ptr(), None, K, b.data_ptr(),
                                     Cr.data_ptr(), N, M, N, K, 0, lib.stream_ptr()), 'gemm_3xtf32(raw)')
        torch.cuda.synchronize()
        assert torch.equal(Cr, Cc), f'in-kernel split differs from pre-split operands, max |d| = {(Cr - Cc).abs().max().item():.3e}'
    for n_rois in (1, 2, 7, 300):
        x = torch.randn(n_rois, 256, 7, 7, generator=g).cuda()
        w = (torch.randn(256, 256, 3, 3, generator=g) / 48).cuda()
        b = torch.randn(256, generator=g).cuda()
        tok = x.permute(0, 2, 3, 1).contiguous()
        wk = w.permute(0, 2, 3, 1).reshape(256, -1).contiguous()
        (th, tl), (wh, wl) = split(tok), split(wk)
        out = torch.full((n_rois * 49, 256), float('nan'), device='cuda')
        lib.check(h.mv2d_gemm_3xtf32(th.data_ptr(), tl.data_ptr(), 256, wh.data_ptr(), wl.data_ptr(), 2304,
                                     b.data_ptr(), out.data_ptr(), 256, n_rois * 49, 256, 2304, 1 | 128,
                                     lib.stream_ptr()), 'gemm_3xtf32(im2col)')
        # fp64 reference as an explicit im2col matmul (no cuDNN involved)
        xp = F.pad(x.double(), (1, 1, 1, 1)).permute(0, 2, 3, 1)                      # [n, 9, 9, c]
        cols = torch.stack([xp[:, ky:ky + 7, kx:kx + 7] for ky in range(3) for kx in range(3)], 3)
        ref = (cols.reshape(-1, 9 * 256) @ wk.double().T + b.double()).relu()
        assert_close(out, ref, 1e-4, 1e-4, f'im2col conv n={n_rois}')


def test_full_size_properties(state_dicts):
    """Size-independent properties at BASELINE config 2 (N=300): determinism (bitwise), and
    equivariance under a permutation of the proposals inside each view."""
    eng = engine('S', 6, state_dicts)
    spec = synth.CASES['s_cfg2']
    feat, boxes, metas = synth.case_inputs(spec)
    featc = feat.cuda()
    o1 = eng.forward(featc, boxes, metas)
    c1, b1 = o1['cls_scores'].clone(), o1['bbox_preds'].clone()
    o2 = eng.forward(featc, boxes, metas)
    assert torch.equal(c1, o2['cls_scores']) and torch.equal(b1, o2['bbox_preds']), 'non-deterministic'
    rng = np.random.default_rng(0)
    perms = [torch.from_numpy(rng.permutation(len(b))) for b in boxes]
    o3 = eng.forward(featc, [b[p] for b, p in zip(boxes, perms)], metas)
    starts = np.concatenate([[0], np.cumsum([len(b) for b in boxes])])
    gperm = torch.cat([p + int(s) for p, s in zip(perms, starts[:-1])]).cuda()
    assert_close(o3['cls_scores'], c1[:, gperm], 2e-4, 2e-4, 'permuted cls')
    assert_close(o3['bbox_preds'], b1[:, gperm], 2e-4, 2e-4, 'permuted box')
    # CUDA-graph replay of the whole path is bitwise identical to the eager launches
    o4 = eng.forward(featc, boxes, metas, use_graph=True)
    o4 = eng.forward(featc, boxes, metas, use_graph=True)
    assert torch.equal(c1, o4['cls_scores']) and torch.equal(b1, o4['bbox_preds']), 'graph replay differs'
    # every query attends at least to its own RoI
    assert int(o1['match_cnt'].min()) >= 1 and bool((o1['match'][:, 0].cpu() == torch.arange(300)).all())


@pytest.mark.parametrize('name', ['t_small', 't_cfg3', 't_pad'])
def test_two_frame_query_stationary_form_matches_golden(name, state_dicts):
    """xa_form 0 (every query streams its own raw key rows, absorbed projections) stays available beside the
    default key-stationary form (xa_form 1, projected K/V tiles); both must meet the gate, and agree closely."""
    from mv2d_b200.engine import HotPath
    spec, g = load_golden(name)
    feat, boxes, metas = synth.case_inputs(spec)
    outs = {}
    for form in (0, 1):
        eng = HotPath(state_dicts(spec['num_layers']), mode='T', xa_form=form)
        assert eng.xa_form == form
        out = eng.forward(feat.cuda(), [b.cuda() for b in boxes], metas)
        torch.cuda.synchronize()
        assert_close(out['cls_scores'], g['cls_scores'], what=f'cls_scores (xa_form {form})')
        assert_close(out['bbox_preds'], g['bbox_preds'], what=f'bbox_preds (xa_form {form})')
        outs[form] = (out['cls_scores'].clone(), out['bbox_preds'].clone())
    assert_close(outs[1][0], outs[0][0], 2e-4, 2e-4, 'cls_scores form 1 vs form 0')
    assert_close(outs[1][1], outs[0][1], 2e-4, 2e-4, 'bbox_preds form 1 vs form 0')


@pytest.mark.parametrize('mode,case', [('S', 's_small'), ('T', 't_small')])
def test_pipeline_lanes_match_single_engine(mode, case, state_dicts):
    """mv2d_b200.pipeline.Pipeline (several samples in flight on independent lanes that share one set of packed
    weights) returns, for every sample, exactly what a single HotPath returns."""
    from mv2d_b200.engine import HotPath
    from mv2d_b200.pipeline import Pipeline
    spec = synth.CASES[case]
    sd = state_dicts(spec['num_layers'])
    samples = [synth.case_inputs(dict(spec, seed=100 + i)) for i in range(5)]
    eng = HotPath(sd, mode=mode)
    want = []
    for f, b, m in samples:
        o = eng.forward(f.cuda(), b, m, use_graph=True)
        torch.cuda.synchronize()
        want.append((o['cls_scores'].clone(), o['bbox_preds'].clone()))
    pipe = Pipeline(sd, mode=mode, depth=3)
    feats = [s[0].pin_memory() for s in samples]
    for rnd in range(2):            # second round replays the captured graphs
        got = []
        for i, (f, b, m) in enumerate(samples):
            t, o = pipe.submit(feats[i], b, m, to_host=True)
            pipe.wait(t)
            torch.cuda.synchronize()
            got.append((o['host_cls'].clone(), o['host_box'].clone()))
        for (wc, wb), (gc, gb) in zip(want, got):
            assert torch.equal(wc.cpu(), gc) and torch.equal(wb.cpu(), gb)
    # everything in flight at once, results read after the join
    outs = [pipe.submit(feats[i], samples[i][1], samples[i][2])[1] for i in range(3)]
    pipe.join()
    torch.cuda.synchronize()
    for i, o in enumerate(outs):
        assert torch.equal(o['cls_scores'], want[i][0]) and torch.equal(o['bbox_preds'], want[i][1])


@pytest.mark.parametrize('mode,V,n,seed', [('T', 12, 3, 41), ('S', 6, 5, 42)])
def test_odd_feature_grid_matches_oracle(mode, V, n, seed, state_dicts):
    """480x1360 images -> a 30x85 feature grid: the cell count is no multiple of 128 (GEMM row tiles), of 32 (mask
    words) or of 8 (the 8x8 key tiles of the two-frame head have ragged right / bottom edges)."""
    from oracle import mv2d_oracle as O
    eng = engine(mode, 6, state_dicts)
    shape = (480, 1360, 3)
    feat, boxes, metas = synth.make_sample(seed, V, n, img_shape=shape, pad_shape=shape, cam_jitter_deg=2.0)
    assert tuple(feat.shape[-2:]) == (30, 85)
    fn = O.mv2d_s_forward if mode == 'S' else O.mv2d_t_forward
    with torch.no_grad():
        cls, box = fn(state_dicts(6), feat, boxes, metas, O.make_cfg(mode))
    out = eng.forward(feat.cuda(), boxes, metas)
    torch.cuda.synchronize()
    assert_close(out['cls_scores'], cls, what='cls_scores')
    assert_close(out['bbox_preds'], box, what='bbox_preds')


def test_two_frame_kv_projection_skips_unused_row_tiles(state_dicts):
    """The K/V projection leaves 128-row tiles without any key untouched.  Poisoning the K/V buffers with NaN before
    the run shows that those rows are never read by the attention, and that some tiles really are skipped."""
    from mv2d_b200.engine import HotPath
    spec, g = load_golden('t_small')
    feat, boxes, metas = synth.case_inputs(spec)
    eng = HotPath(state_dicts(spec['num_layers']), mode='T')
    out = eng.forward(feat.cuda(), [b.cuda() for b in boxes], metas)
    torch.cuda.synchronize()
    want = (out['cls_scores'].clone(), out['bbox_preds'].clone())
    live = out['row_tile_live'].cpu().numpy()
    assert 0 < live.sum() < live.size, 'this case has both used and unused row tiles'
    eng._buf['kp'].fill_(float('nan'))
    eng._buf['vp'].fill_(float('nan'))
    out = eng.forward(feat.cuda(), [b.cuda() for b in boxes], metas)
    torch.cuda.synchronize()
    assert torch.equal(out['cls_scores'], want[0]) and torch.equal(out['bbox_preds'], want[1])
    R = eng._buf['kp'].numel() // (eng.L * 256)
    kp = eng._buf['kp'][:eng.L * R * 256].view(eng.L, R, 256)
    rows_nan = torch.isnan(kp[0]).any(dim=1).view(-1).cpu().numpy()
    for t in range(live.size):
        seg = rows_nan[t * 128:(t + 1) * 128]
        assert seg.all() if live[t] == 0 else not seg.any()
    assert_close(out['cls_scores'], g['cls_scores'], what='cls_scores')


@pytest.mark.parametrize('name', ['s_cfg2', 's_pad', 't_small'])
def test_fused_pe_mlps_match_the_layer_by_layer_gemms(name, state_dicts):
    """csrc/mlp2.cu (hidden activations in TMEM / shared memory, the sine branch's gather as a one-hot GEMM) against one
    tcgen05 GEMM per layer: both single-pass TF32, so they agree far inside the PE stage tolerance; 12 views exercise the
    wider one-hot operand (132 table rows), the padded case the general sine branch beside the fused position MLP."""
    spec, g = load_golden(name)
    eng = engine(spec['mode'], spec['num_layers'], state_dicts)
    feat, boxes, metas = synth.case_inputs(spec)
    outs = []
    for unfused in (False, True):
        eng.pe_unfused = unfused
        try:
            o = eng.forward(feat.cuda(), boxes, metas)
            outs.append((o['pe'].clone(), o['cls_scores'].clone(), o['bbox_preds'].clone()))
        finally:
            eng.pe_unfused = False
    assert_close(outs[0][0], outs[1][0], 1e-3, 1e-3, "pe fused vs unfused")   # a ReLU / TF32 rounding flip moves single entries by ~2e-4
    pe = outs[0][0].permute(0, 3, 1, 2).contiguous()
    assert_close(pe.flatten()[::PE_SUB], g['pe_sub'], 3e-3, 1e-3, 'pe')
    assert_close(outs[0][1], g['cls_scores'], what='cls_scores')
    assert_close(outs[0][2], g['bbox_preds'], what='bbox_preds')


def test_layout_pass_with_tf32_halves_is_bit_exact():
    """mv2d_nchw_to_nhwc_split: the channels-last copy, its TF32 hi half and lo half (operands of the 3xTF32 K / V projection)
    against the integer rounding the host packer uses -- bit for bit, on an odd grid (ragged 64-pixel tiles)."""
    from mv2d_b200 import lib
    h = lib.load()
    torch.manual_seed(11)
    V, C, H, W = 3, 256, 30, 85
    x = (torch.randn(V, C, H, W, device='cuda') * 3).contiguous()
    out = torch.empty(V, H, W, C, device='cuda')
    hi, lo = torch.empty_like(out), torch.empty_like(out)
    lib.check(h.mv2d_nchw_to_nhwc_split(x.data_ptr(), out.data_ptr(), hi.data_ptr(), lo.data_ptr(), V, C, H * W, lib.stream_ptr()),
              'mv2d_nchw_to_nhwc_split')
    torch.cuda.synchronize()

    def rnd(t):        # (bits + 0x1000) & ~0x1fff: cvt.rna.tf32.f32 on the bit pattern (csrc/pack.cpp, gemm_tc.cuh)
        b = t.contiguous().view(torch.int32)
        return ((b + 0x1000) & ~0x1FFF).view(torch.float32)
    ref = x.permute(0, 2, 3, 1).contiguous()
    assert torch.equal(out, ref)
    assert torch.equal(hi, rnd(ref))
    assert torch.equal(lo, rnd(ref - rnd(ref)))
    # the plain entry writes the same copy and hi half
    out2, hi2 = torch.empty_like(out), torch.empty_like(out)
    lib.check(h.mv2d_nchw_to_nhwc(x.data_ptr(), out2.data_ptr(), hi2.data_ptr(), V, C, H * W, lib.stream_ptr()), 'mv2d_nchw_to_nhwc')
    torch.cuda.synchronize()
    assert torch.equal(out2, ref) and torch.equal(hi2, hi)
