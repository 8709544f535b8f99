"""Device time of every stage when replayed from its own CUDA graph (no CPU launch gaps; L2 flushed before every
replay), plus the whole path.  usage: python tools/graph_stages.py [S|T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mv2d_b200 import synth
from mv2d_b200.engine import HotPath

mode = sys.argv[1] if len(sys.argv) > 1 else 'S'
sd = synth.make_state_dict(0)
feat, boxes, metas = synth.case_inputs(synth.CASES['s_cfg2' if mode == 'S' else 't_cfg3'])
eng = HotPath(sd, mode=mode)
featc = feat.cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
V, _, h, w = feat.shape


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


def graphed(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.synchronize()
    return timed(g.replay)


out = eng.forward(featc, boxes, metas)
torch.cuda.synchronize()
N = out['N']
cams, rois, roi_start, counts, _ = eng._upload_meta(boxes, metas)
f, f32r = eng.to_nhwc(featc)
i2l, trans = eng.geom_prep(cams)
pe, kin = eng.pe3d(f, i2l, metas, f32r)
qg = eng.roi_align_qg(rois, cams, f, pe, N)
corr = eng.box_corr(rois, roi_start, trans, N, V, metas, h, w)
torch.cuda.synchronize()
res = {}
res['nchw_to_nhwc'] = graphed(lambda: eng.to_nhwc(featc))
res['geom_prep'] = graphed(lambda: eng.geom_prep(cams))
res['pe3d'] = graphed(lambda: eng.pe3d(f, i2l, metas, f32r))
res['pe3d_phase1'] = graphed(lambda: eng.pe3d(f, i2l, metas, None, phase=1))
res['pe3d_phase2'] = graphed(lambda: eng.pe3d(f, i2l, metas, f32r, phase=2))
res['qg_phase1'] = graphed(lambda: eng.roi_align_qg(rois, cams, f, None, N, phase=1))
if mode == 'S':
    res['qg_phase2'] = graphed(lambda: eng.roi_align_qg(rois, cams, f, pe, N, phase=2))
res['box_corr'] = graphed(lambda: eng.box_corr(rois, roi_start, trans, N, V, metas, h, w))
if mode == 'S':
    res['decoder'] = graphed(lambda: eng.decoder(qg, corr, qg['tok_kin'].view(-1, 256), qg['tok_feat'].view(-1, 256), N))
else:
    kr, mr = kin.view(-1, 256), f.view(-1, 256)
    if eng.xa_form == 1:
        res['kv_project'] = graphed(lambda: eng.kv_project(kr, mr))
        kv = eng.kv_project(kr, mr)
        res['decoder_without_kv'] = graphed(lambda: eng.decoder(qg, corr, kr, mr, N, vel_dt=0.5, kv=kv, grid=(h, w)))
    else:
        res['decoder'] = graphed(lambda: eng.decoder(qg, corr, kr, mr, N, vel_dt=0.5))
res['whole_path'] = timed(lambda: eng.forward(featc, boxes, metas, use_graph=True))
print(mode, {k: round(v, 1) for k, v in res.items()})
