"""Throughput of the batched hot path: samples/s for (batch B, lanes in flight) with device-resident and host inputs.
   python tools/batch_sweep.py [--mode S|T] [--steps 40]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mv2d_b200 import synth                     # noqa: E402
from mv2d_b200.pipeline import Pipeline         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', default='S')
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--grid', default='1x1,1x4,2x1,2x2,4x1,4x2,8x1,8x2')
    ap.add_argument('--overlap', type=int, default=1, help='0: the whole front end on one stream (batches of the single-frame head do that by themselves since the A/B run recorded in engine._enqueue_post; the flag still matters for bs = 1 and the two-frame head)')
    args = ap.parse_args()
    sd = synth.make_state_dict(0)
    case = synth.CASES['s_cfg2' if args.mode == 'S' else 't_cfg3']
    n_var = 16
    samples = [synth.case_inputs(dict(case, seed=i)) for i in range(n_var)]
    for cell in args.grid.split(','):
        B, depth = [int(x) for x in cell.split('x')]
        pipe = Pipeline(sd, mode=args.mode, depth=depth, overlap=bool(args.overlap))
        batches = []
        for i in range(0, n_var, B):
            grp = [samples[(i + j) % n_var] for j in range(B)]
            batches.append((torch.stack([g[0] for g in grp], 0).cuda(), torch.stack([g[0] for g in grp], 0).pin_memory(),
                            [g[1] for g in grp], [g[2] for g in grp]))
        res = {}
        for host in (False, True):
            def sub(i):
                bt = batches[i % len(batches)]
                if B == 1:
                    pipe.submit((bt[1] if host else bt[0])[0], bt[2][0], bt[3][0], to_host=host)
                else:
                    pipe.submit_batch(bt[1] if host else bt[0], bt[2], bt[3], to_host=host)
            for i in range(3 * depth + 3):
                sub(i)
            pipe.join()
            torch.cuda.synchronize()
            l0 = pipe.launch_count()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(args.steps):
                sub(i)
            pipe.join()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            res['e2e' if host else 'resident'] = dict(samples_per_s=args.steps * B / ms * 1e3, ms_per_sample=ms / (args.steps * B),
                                                     launches_per_sample=(pipe.launch_count() - l0) / (args.steps * B))
        print(json.dumps(dict(mode=args.mode, batch=B, lanes=depth, overlap=args.overlap, **res)), flush=True)
        del pipe
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
