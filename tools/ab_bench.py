"""A/B helper: whole-path and decoder-only CUDA-graph replay times (L2 flushed before every replay) for the
configuration selected through the environment (MV2D_TC_MULTICAST, MV2D_FOLD0, MV2D_NO_PDL, MV2D_DECODER)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mv2d_b200 import synth
from mv2d_b200.engine import HotPath

mode = sys.argv[1] if len(sys.argv) > 1 else 'S'
case = 's_cfg2' if mode == 'S' else 't_cfg3'
sd = synth.make_state_dict(0)
feat, boxes, metas = synth.case_inputs(synth.CASES[case])
eng = HotPath(sd, mode=mode, fold_first_self_attn=os.environ.get('MV2D_FOLD0', '1') != '0')
featc = feat.cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def timed(fn, reps=40):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


whole = timed(lambda: eng.forward(featc, boxes, metas, use_graph=True))
# decoder alone in its own graph, inputs from one eager run
out = eng.forward(featc, boxes, metas)
torch.cuda.synchronize()
N = out['N']
qg = {k: out[k] for k in ('query_pos', 'ref', 'tok_kin', 'tok_feat') if k in out}
corr = {k: out[k] for k in ('match', 'match_cnt', 'max_match', 'keymask', 'key_list', 'key_cnt', 'mask_words') if k in out}
if mode == 'S':
    kin, mem = out['tok_kin'].view(-1, 256), out['tok_feat'].view(-1, 256)
else:
    kin, mem = eng._buf['kin'].view(-1, 256) if 'kin' in eng._buf else None, out['feat_nhwc'].view(-1, 256)
res = {'whole_med_us': whole[0], 'whole_min_us': whole[1]}
if kin is not None:
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        eng.decoder(qg, corr, kin, mem, N)
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            eng.decoder(qg, corr, kin, mem, N)
    torch.cuda.synchronize()
    dec = timed(g.replay)
    res.update(decoder_med_us=dec[0], decoder_min_us=dec[1])
print({k: round(v, 1) for k, v in res.items()}, {k: os.environ.get(k) for k in ('MV2D_TC_MULTICAST', 'MV2D_FOLD0', 'MV2D_NO_PDL')})
