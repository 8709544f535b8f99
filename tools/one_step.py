"""Run the hot path eagerly a few times (for ncu launch lists): python tools/one_step.py [S|T] [reps] [batch]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mv2d_b200 import synth
from mv2d_b200.engine import HotPath

mode = sys.argv[1] if len(sys.argv) > 1 else 'S'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sd = synth.make_state_dict(0)
feat, boxes, metas = synth.case_inputs(synth.CASES['s_cfg2' if mode == 'S' else 't_cfg3'])
eng = HotPath(sd, mode=mode)
featc = feat.cuda()
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if batch:
    case = synth.CASES['s_cfg2' if mode == 'S' else 't_cfg3']
    ins = [synth.case_inputs(dict(case, seed=i)) for i in range(batch)]
    feats = torch.stack([i[0] for i in ins], 0).cuda()
for _ in range(reps):
    if batch:
        out = eng.forward_batch(feats, [i[1] for i in ins], [i[2] for i in ins])
    else:
        out = eng.forward(featc, boxes, metas)
    torch.cuda.synchronize()
print('launches per step:', eng.launch_count() // reps)
