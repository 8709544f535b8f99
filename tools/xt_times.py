"""Two-frame head: device time of the K/V projection and of the decoder in both cross-attention forms."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mv2d_b200 import synth
from mv2d_b200.engine import HotPath

sd = synth.make_state_dict(0)
feat, boxes, metas = synth.case_inputs(synth.CASES['t_cfg3'])
featc = feat.cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


res = {}
for form in (0, 1):
    eng = HotPath(sd, mode='T', xa_form=form)
    for _ in range(2):
        out = eng.forward(featc, boxes, metas)
    torch.cuda.synchronize()
    N = out['N']
    qg = {k: out[k] for k in ('query_pos', 'ref')}
    corr = {k: out[k] for k in ('keymask', 'mask_words', 'key_list', 'key_cnt')}
    mem = out['feat_nhwc'].view(-1, 256)
    kin = eng._buf['kin'][:mem.numel()].view(-1, 256)
    t_dec = timed(lambda: eng.decoder(qg, corr, kin, mem, N, vel_dt=0.5))
    line = f'xa_form {form}: decoder stage {t_dec:8.1f} us'
    if form == 1:
        t_kv = timed(lambda: eng.kv_project(kin, mem))
        kv = eng.kv_project(kin, mem)
        t_only = timed(lambda: eng.decoder(qg, corr, kin, mem, N, vel_dt=0.5, kv=kv, grid=eng._last_grid))
        line += f'  (kv_project {t_kv:.1f} us, decoder without it {t_only:.1f} us)'
    t_all = timed(lambda: eng.forward(featc, boxes, metas, use_graph=True), reps=20)
    line += f'   whole path (graph) {t_all:8.1f} us'
    print(line)
    res[form] = out['cls_scores'].clone()
print('max |cls form1 - form0| =', (res[1] - res[0]).abs().max().item())
