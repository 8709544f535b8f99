"""Row f4: device time of the MV2D neck (1x1 lateral + 3x3 output conv as 3xTF32 tcgen05 GEMMs), replayed from a
CUDA graph with L2 flushed, beside torch conv2d on the same GPU (cuDNN, TF32 allowed / not allowed) and on the host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
from mv2d_b200 import synth
from mv2d_b200.engine import HotPath
from mv2d_b200.pack import PackedNeck

sd, nsd = synth.make_state_dict(0), synth.make_neck_state_dict(0)
eng = HotPath(sd, mode='S')
pn = PackedNeck(nsd, 'cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


for V in (6, 12):
    x = torch.randn(V, 256, 32, 88).cuda()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        eng.neck(x, pn); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            eng.neck(x, pn)
    torch.cuda.synchronize()
    t = timed(g.replay)
    P = V * 32 * 88
    flops = 2.0 * P * 256 * 256 + 2.0 * P * 2304 * 256
    wl, bl = nsd['lateral_convs.0.conv.weight'].cuda(), nsd['lateral_convs.0.conv.bias'].cuda()
    wf, bf = nsd['fpn_convs.0.conv.weight'].cuda(), nsd['fpn_convs.0.conv.bias'].cuda()
    ref = lambda: F.conv2d(F.conv2d(x, wl, bl), wf, bf, padding=1)
    torch.backends.cudnn.allow_tf32 = True
    t_tf32 = timed(ref)
    torch.backends.cudnn.allow_tf32 = False
    t_fp32 = timed(ref)
    xc = x.cpu(); cw = [t.cpu() for t in (wl, bl, wf, bf)]
    t0 = time.perf_counter()
    for _ in range(3):
        F.conv2d(F.conv2d(xc, cw[0], cw[1]), cw[2], cw[3], padding=1)
    t_cpu = (time.perf_counter() - t0) / 3 * 1e3
    print(f'V={V}: mv2d_fpn_neck {t:7.1f} us = {flops / t / 1e6:6.1f} TFLOP/s (fp32-grade 3xTF32; {3 * flops / t / 1e6:6.1f} TF32 issued)   '
          f'torch cuDNN tf32 {t_tf32:7.1f} us, fp32 {t_fp32:7.1f} us (NCHW out, needs a transpose after)   torch CPU {t_cpu:7.1f} ms')
