"""Debug helper: compare the decoder workspace of the staged and the persistent decoder after a 1-layer run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mv2d_b200 import synth
from mv2d_b200.engine import HotPath

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1
spec = dict(synth.CASES['s_small'], num_layers=L)
sd = synth.make_state_dict(0, num_layers=L)
feat, boxes, metas = synth.case_inputs(spec)
res = {}
for name, kw in (('staged', dict(persistent_decoder=False, fold_first_self_attn=False)),
                 ('persist', dict(persistent_decoder=True, fold_first_self_attn=False))):
    eng = HotPath(sd, mode='S', **kw)
    out = eng.forward(feat.cuda(), boxes, metas)
    torch.cuda.synchronize()
    N = out['N']
    ws = eng._buf['dec_ws'].clone()
    res[name] = (ws, out['cls_scores'].clone(), out['outs_dec'].clone())
C = 256
names = [('x', N * C), ('xq', N * C), ('x1', N * C), ('x1q', N * C), ('x2', N * C), ('x1q_hi', N * C), ('x1q_lo', N * C),
         ('x2_hi', N * C), ('x2_lo', N * C), ('qkv', N * 768), ('sa', N * C), ('qt', N * 2048), ('ctx', N * 2048),
         ('ctx_lo', N * 2048), ('hdn', N * 2048), ('hdn_lo', N * 2048), ('part', 8 * N * C)]
off = 0
a, b = res['staged'][0], res['persist'][0]
for n, sz in names:
    d = (a[off:off + sz] - b[off:off + sz]).abs()
    print(f'{n:8s} max|d| = {d.max().item():.3e}   finite: {torch.isfinite(b[off:off + sz]).all().item()}')
    off += sz
print('outs_dec', (res['staged'][2] - res['persist'][2]).abs().max().item(), 'cls', (res['staged'][1] - res['persist'][1]).abs().max().item())
off = 4 * N * C
d = (a[off:off + N * C] - b[off:off + N * C]).abs().view(N, C)
print('x2 diff per 64-col block:', [round(d[:, i * 64:(i + 1) * 64].max().item(), 4) for i in range(4)])
print('x2 diff per row (first 27):', [round(v, 3) for v in d.max(1).values.tolist()[:27]])
# partial sums: compare the sum over the 8 split-K partials (part holds the LAST split-K GEMM = ffn2)
