"""Row f3: device time of mv2d_loss (cost + device LSA + losses for all 6 layers, CUDA-graph replay) beside the
reference's way (per layer: cost on the device would be followed by .cpu() + scipy; here the whole oracle on the
host cores)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mv2d_b200 import synth
from mv2d_b200.engine import HotPath
from oracle import mv2d_oracle as O

sd = synth.make_state_dict(0)
eng = HotPath(sd, mode='S')
for N, G in ((300, 30), (300, 100), (900, 100)):
    g = torch.Generator().manual_seed(N + G)
    cls, box = torch.randn(6, N, 10, generator=g) * 2 - 2, torch.randn(6, N, 10, generator=g)
    gt = torch.randn(G, 9, generator=g); gt[:, 3:6] = gt[:, 3:6].abs() + 0.3
    lab = torch.randint(0, 10, (G,), generator=g)
    c, b = cls.cuda(), box.cuda()
    gtc, labc = gt.cuda(), lab.cuda()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        eng.loss(c, b, gtc, labc); torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            out = eng.loss(c, b, gtc, labc)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); graph.replay(); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e) * 1e3)
    t0 = time.perf_counter()
    for l in range(6):
        O.loss_single(cls[l], box[l], gt, lab)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    print(f'N={N} G={G}: mv2d_loss (6 layers) {np.median(ts):8.1f} us on the device;  oracle (torch CPU + scipy) {cpu_ms:7.2f} ms')
