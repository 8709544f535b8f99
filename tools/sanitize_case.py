"""Small end-to-end runs for compute-sanitizer: python tools/sanitize_case.py  (two-frame head with the key-stationary
cross-attention, S head, training-mode forward + losses, neck, scene NMS) on the small golden cases."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mv2d_b200 import synth
from mv2d_b200.engine import HotPath

for mode, case in (('T', 't_small'), ('S', 's_small'), ('T', 't_dn'), ('S', 's_dn')):
    spec = synth.CASES[case]
    sd = synth.make_state_dict(0, num_layers=spec['num_layers'])
    eng = HotPath(sd, mode=mode)
    feat, boxes, metas = synth.case_inputs(spec)
    if 'dn' in spec:
        gt, lab, rand = synth.make_dn_inputs(spec['dn'])
        losses, out, raw = eng.forward_losses(feat.cuda(), boxes, metas, gt, lab, rand=rand, use_denoise=True)
        print(case, 'losses', float(raw['loss_cls'][-1]), float(raw['loss_bbox'][-1]))
    else:
        out = eng.forward(feat.cuda(), boxes, metas)
        b, s, l = eng.decode(out['cls_scores'][-1], out['bbox_preds'][-1])
        b, s, l = eng.scene_nms(b, s, l)
        print(case, 'N', out['N'], 'boxes', b.shape[0])
    torch.cuda.synchronize()
# batches as a segment dimension: the kernels only batches reach (tensor-core self-attention, persistent GEMMs, one-pass RoIAlign)
for mode, names in (('S', ['s_small', 's_cfg2', 's_small', 's_small']), ('T', ['t_small', 't_small'])):
    specs = [synth.CASES[n] for n in names]
    sd = synth.make_state_dict(0, num_layers=specs[0]['num_layers'])
    eng = HotPath(sd, mode=mode)
    ins = [synth.case_inputs(sp) for sp in specs]
    out = eng.forward_batch(torch.stack([i[0] for i in ins], 0).cuda(), [i[1] for i in ins], [i[2] for i in ins])
    torch.cuda.synchronize()
    print('batch', mode, [smp['N'] for smp in out['samples']])
nsd = synth.make_neck_state_dict(0)
f, _ = eng.neck(torch.randn(2, 256, 30, 85).cuda(), nsd)
torch.cuda.synchronize()
print('neck', tuple(f.shape), 'done')
