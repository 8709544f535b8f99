"""One training step, eagerly (for ncu launch lists): python tools/train_one_step.py [S|T] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mv2d_b200 import synth                         # noqa: E402
from mv2d_b200.train import HotPathTrainer          # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'S'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sd = synth.make_state_dict(0)
feat, boxes, metas = synth.case_inputs(synth.CASES['s_cfg2' if mode == 'S' else 't_cfg3'])
gt_boxes, gt_labels, _ = synth.make_dn_inputs(dict(num_gt=30, seed=300))
tr = HotPathTrainer(sd, mode=mode)
for _ in range(reps):
    tr.zero_grad()
    out = tr.forward(feat.cuda(), boxes, metas, gt_boxes.cuda(), gt_labels.cuda())
    tr.backward()
    torch.cuda.synchronize()
print('loss', float(out['loss']))
