"""Phase timeline of the persistent decoder kernel (globaltimer stamps taken by CTA 0 at every barrier)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mv2d_b200 import synth
from mv2d_b200.engine import HotPath
mode = sys.argv[1] if len(sys.argv) > 1 else 'S'
sd = synth.make_state_dict(0)
feat, boxes, metas = synth.case_inputs(synth.CASES['s_cfg2' if mode == 'S' else 't_cfg3'])
eng = HotPath(sd, mode=mode, persistent_decoder=True)
for _ in range(5):
    out = eng.forward(feat.cuda(), boxes, metas)
torch.cuda.synchronize()
ws = eng._buf['dec_ws']
N, L = out['N'], eng.L
nbytes = eng.lib.mv2d_decoder_workspace_bytes(N, L)
tail = ws[: nbytes // 4].view(torch.uint8)[-4096 - 64:].cpu().numpy()
# find the barrier word: aligned to 64 B; scan for the counter value 148*gen
raw = ws[: nbytes // 4].view(torch.uint8).cpu().numpy()
names = ['inproj', 'self_attn', 'outproj', 'LN1', 'q~ tc', 'cross_attn', 'ctxout tc', 'LN2', 'ffn1 tc', 'ffn2 tc', 'LN3']
best = None
for off in range(len(raw) - 4096 - 64, len(raw) - 64, 64):
    cnt = raw[off:off + 4].view(np.uint32)[0]
    if cnt == 148 * (11 * L + 4):
        best = off
        break
assert best is not None, 'barrier word not found'
ts = raw[best + 64: best + 64 + 8 * (11 * L + 5)].view(np.uint64).astype(np.float64)
d = np.diff(ts) / 1e3
print('total (first barrier -> last barrier): %.1f us' % ((ts[11 * L + 3] - ts[0]) / 1e3))
for l in range(L):
    seg = d[11 * l: 11 * l + 11] if l < L - 1 else d[11 * l: 11 * l + 11]
    print(f'layer {l}: ' + '  '.join(f'{n} {v:.1f}' for n, v in zip(names[1:] + ['inproj(next)'], seg)))
print('branches:', '  '.join(f'{v:.1f}' for v in d[11 * L:11 * L + 4]))

tc = raw[best + 64 + 400 * 8: best + 64 + 400 * 8 + 4 * 8 * 8].view(np.uint64).astype(np.float64).reshape(4, 8)
for gi, nm in enumerate(['q~', 'ctxout', 'ffn1', 'ffn2']):
    t = tc[gi]
    print(f'tc tile {nm}: start->first stage landed {(t[1]-t[0])/1e3:.2f} us, ->last stage landed {(t[2]-t[0])/1e3:.2f}, ->accumulator done {(t[3]-t[0])/1e3:.2f}, ->tile end {(t[4]-t[0])/1e3:.2f}')
