"""SURVEY 8d asks for the metric "also end-to-end incl. torch R50+FPN".  The reference's 2D detector is mmdet's Faster
R-CNN (ResNet-50 with DCNv2 in stages 3-4, 5-level FPN, RPN, RoI head), which this image does not have; this tool puts
the closest torch stand-in in front of the hot path -- torchvision ResNet-50 + FeaturePyramidNetwork, random weights,
6 x 3 x 512 x 1408 images, cuDNN with TF32 allowed, channels-last -- so the torch side here is a LOWER bound on the
reference's (no deformable convs, no RPN / RoI head).  It answers one question: what share of a full sample is the hot
path.   python tools/e2e_backbone.py"""
import json
import os
import sys
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mv2d_b200 import synth  # noqa: E402
from mv2d_b200.engine import HotPath  # noqa: E402
from mv2d_b200.pack import PackedNeck  # noqa: E402


def main():
    from torchvision.models import resnet50
    from torchvision.ops import FeaturePyramidNetwork
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    dev = torch.device('cuda')
    r = resnet50(weights=None).to(dev).eval().to(memory_format=torch.channels_last)
    fpn = FeaturePyramidNetwork([256, 512, 1024, 2048], 256).to(dev).eval()
    sd, nsd = synth.make_state_dict(0), synth.make_neck_state_dict(0)
    eng = HotPath(sd, mode='S')
    neck = PackedNeck(nsd, dev)
    _, boxes, metas = synth.case_inputs(synth.CASES['s_cfg2'])
    img = torch.randn(6, 3, 512, 1408, device=dev).contiguous(memory_format=torch.channels_last)

    @torch.no_grad()
    def backbone():
        x = r.maxpool(r.relu(r.bn1(r.conv1(img))))
        c2 = r.layer1(x); c3 = r.layer2(c2); c4 = r.layer3(c3); c5 = r.layer4(c4)
        return fpn(OrderedDict([('0', c2), ('1', c3), ('2', c4), ('3', c5)]))['2']      # P4, stride 16: [6,256,32,88]

    @torch.no_grad()
    def hot(p4):
        feat, _ = eng.neck(p4, neck)
        out = eng.forward(feat, boxes, metas, feat_is_nhwc=True)
        return eng.decode(out['cls_scores'][-1], out['bbox_preds'][-1])

    for _ in range(3):
        hot(backbone())
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tb, th = [], []
    for _ in range(10):
        ev[0].record(); p4 = backbone(); ev[1].record(); hot(p4); ev[2].record()
        torch.cuda.synchronize()
        tb.append(ev[0].elapsed_time(ev[1])); th.append(ev[1].elapsed_time(ev[2]))
    tb, th = sorted(tb)[len(tb) // 2], sorted(th)[len(th) // 2]
    line = dict(backbone_ms=tb, neck_hot_path_decode_ms=th, total_ms=tb + th, samples_per_s=1e3 / (tb + th),
                hot_path_share=th / (tb + th),
                note='torchvision ResNet-50 + FPN stand-in (TF32 cuDNN, channels-last, random weights) -> mv2d_fpn_neck -> hot path '
                     '(eager, one sample) -> NMS-free decode; the reference detector (DCNv2, RPN, RoI head) costs more on the torch side')
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'e2e_backbone.json'), 'w') as f:
        f.write(json.dumps(line) + '\n')


if __name__ == '__main__':
    main()
