"""Throughput of the hot path with 1..3 samples in flight (mv2d_b200.pipeline.Pipeline), device-resident inputs.
8 distinct samples rotate (8 x 17.3 MB of feature maps > the 126 MB L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mv2d_b200 import synth
from mv2d_b200.pipeline import Pipeline

mode = sys.argv[1] if len(sys.argv) > 1 else 'S'
case = synth.CASES['s_cfg2' if mode == 'S' else 't_cfg3']
sd = synth.make_state_dict(0)
samples = [synth.case_inputs(dict(case, seed=i)) for i in range(8)]
feats = [s[0].cuda() for s in samples]
pins = [s[0].pin_memory() for s in samples]
steps = 64
for depth in [int(x) for x in (sys.argv[2].split(',') if len(sys.argv) > 2 else '1,2,3'.split(','))]:
    pipe = Pipeline(sd, mode=mode, depth=depth)
    for host in (False, True):
        src = pins if host else feats
        for i in range(2 * depth + 8):
            pipe.submit(src[i % 8], samples[i % 8][1], samples[i % 8][2], to_host=host)
        pipe.join(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            pipe.submit(src[i % 8], samples[i % 8][1], samples[i % 8][2], to_host=host)
        pipe.join()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print(f'{mode} depth {depth} {"e2e (host buffers)" if host else "resident"}: {ms / steps * 1e3:8.1f} us/sample  {steps / ms * 1e3:8.1f} samples/s')
    del pipe
