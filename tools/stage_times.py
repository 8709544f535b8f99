"""Per-stage device times of the hot path (CUDA events on the launching stream)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mv2d_b200 import synth
from mv2d_b200.engine import HotPath

mode = sys.argv[1] if len(sys.argv) > 1 else 'S'
case = 's_cfg2' if mode == 'S' else 't_cfg3'
sd = synth.make_state_dict(0)
feat, boxes, metas = synth.case_inputs(synth.CASES[case])
eng = HotPath(sd, mode=mode)
featc = feat.cuda()
V = len(metas)
_, _, h, w = feat.shape
for _ in range(3):
    out = eng.forward(featc, boxes, metas)
torch.cuda.synchronize()
N = out['N']
ev = lambda: torch.cuda.Event(enable_timing=True)
names, evs = [], [ev()]
def mark(n):
    names.append(n); e = ev(); e.record(); evs.append(e)
reps = 20
acc = {}
for _ in range(reps):
    names.clear(); evs[:] = [ev()]; evs[0].record()
    f, f32r = eng.to_nhwc(featc); mark('nchw_to_nhwc')
    cams, rois, roi_start, counts, N = eng._upload_meta(boxes, metas); mark('uploads')
    i2l, trans = eng.geom_prep(cams); mark('geom_prep')
    pe, kin = eng.pe3d(f, i2l, metas, f32r); mark('pe3d')
    qg = eng.roi_align_qg(rois, cams, f, pe, N); mark('roi_align_qg')
    corr = eng.box_corr(rois, roi_start, trans, N, V, metas, h, w); mark('box_corr')
    if mode == 'S':
        eng.decoder(qg, corr, qg['tok_kin'].view(-1, 256), qg['tok_feat'].view(-1, 256), N)
    else:
        eng.decoder(qg, corr, kin.view(-1, 256), f.view(-1, 256), N, vel_dt=0.5)
    mark('decoder')
    torch.cuda.synchronize()
    for i, n in enumerate(names):
        acc[n] = acc.get(n, 0.0) + evs[i].elapsed_time(evs[i + 1])
# whole path, eager vs CUDA graph
def timed(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
t_eager = timed(lambda: eng.forward(featc, boxes, metas))
t_graph = timed(lambda: eng.forward(featc, boxes, metas, use_graph=True))
og = eng.forward(featc, boxes, metas, use_graph=True); oe_cls = eng.forward(featc, boxes, metas)['cls_scores'].clone()
print(f'whole path: eager {t_eager:.1f} us   graph {t_graph:.1f} us   graph==eager: {torch.equal(og["cls_scores"], oe_cls)}')
# SM clock seen by kernels right after a burst of path replays
from mv2d_b200 import lib as L
probe = torch.zeros(2, dtype=torch.int64, device='cuda')
for _ in range(20): eng.forward(featc, boxes, metas, use_graph=True)
L.check(eng.lib.mv2d_debug_clock_probe(2_000_000, probe.data_ptr(), L.stream_ptr()), 'probe')
torch.cuda.synchronize()
print(f'SM clock under this load: {probe[1].item() / probe[0].item() * 1e3:.0f} MHz')
tot = sum(acc.values())
for n, t in acc.items():
    print(f'{n:16s} {t / reps * 1e3:9.1f} us  {100 * t / tot:5.1f} %')
print(f'{"total":16s} {tot / reps * 1e3:9.1f} us   N={N} mode={mode}')
if mode == 'T':
    print('keys/query mean', out['key_cnt'].float().mean().item(), 'max', out['key_cnt'].max().item())
else:
    print('matches/query mean', out['match_cnt'].float().mean().item())
