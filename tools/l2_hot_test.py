"""Is the decoder bound by weights streaming in from HBM?  Time a 1-layer decoder with the L2 hot vs flushed."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mv2d_b200 import synth
from mv2d_b200.engine import HotPath
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for L in (1, 6):
    sd = synth.make_state_dict(0, num_layers=L)
    feat, boxes, metas = synth.case_inputs(dict(synth.CASES['s_cfg2'], num_layers=L))
    eng = HotPath(sd, mode='S')
    out = eng.forward(feat.cuda(), boxes, metas)
    torch.cuda.synchronize()
    N = out['N']
    qg = {k: out[k] for k in ('query_pos', 'ref', 'tok_feat', 'tok_kin')}
    corr = dict(match=out['match'], match_cnt=out['match_cnt'], max_match=out['max_match'])
    kin, mem = out['tok_kin'].view(-1, 256), out['tok_feat'].view(-1, 256)
    g = torch.cuda.CUDAGraph()
    eng.decoder(qg, corr, kin, mem, N); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        eng.decoder(qg, corr, kin, mem, N)
    for mode in ('hot', 'flushed'):
        ts = []
        for _ in range(30):
            if mode == 'flushed': flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        print(f'L={L} decoder graph replay, L2 {mode}: {statistics.median(ts):.1f} us')
