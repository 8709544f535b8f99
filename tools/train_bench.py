"""Hot-path training step, timed (BASELINE configs[3] shape: MV2D-S, 300 queries, 6 layers, 2 samples per GPU):
per step and rank  zero_grad -> for each local sample [forward with saved activations (position encoding, RoIAlign,
query generator, decoder) + Hungarian targets / losses -> backward down to d feat] -> ONE NCCL sum all-reduce of the
flat gradient buffer (all 14.0 M hot-path parameters) -> fused AdamW.  The torch backbone is outside (north star).
--slice times the decoder slice alone (mv2d_decoder_train_*), its inputs prepared once outside the timed region.

    python tools/train_bench.py [--per-view 50] [--samples 2] [--steps 10] [--slice]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_bench.py
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mv2d_b200 import dist as D  # noqa: E402
from mv2d_b200 import synth  # noqa: E402
from mv2d_b200.engine import HotPath  # noqa: E402
from mv2d_b200.train import DecoderTrainer, HotPathTrainer, TrainStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--per-view', type=int, default=50)
    ap.add_argument('--samples', type=int, default=2, help='samples per rank and step')
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--slice', action='store_true', help='decoder slice only')
    ap.add_argument('--lanes', type=int, default=1, help='samples in flight per GPU (TrainStep); 1 = one after the other')
    a = ap.parse_args()
    rank, local_rank, world = D.env_rank()
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    D.init('nccl', dev)
    sd = synth.make_state_dict(0, num_layers=6)
    if a.slice:
        eng = HotPath(sd, mode='S', device=dev)
        tr = DecoderTrainer(sd, device=dev)
    elif a.lanes > 1:
        pipe = TrainStep(sd, device=dev, lanes=a.lanes)
        tr = pipe.main
    else:
        tr = HotPathTrainer(sd, device=dev)
    samples = []
    for s in D.shard_samples(a.samples * world, rank, world):
        spec = dict(mode='S', seed=100 + s, num_views=6, boxes_per_view=a.per_view, num_layers=6)
        feat, boxes, metas = synth.case_inputs(spec)
        gt_boxes, gt_labels, _ = synth.make_dn_inputs(dict(num_gt=30, seed=200 + s))
        if a.slice:
            o = eng.forward(feat.to(dev), boxes, metas)
            samples.append(tuple(t.clone() for t in (o['ref'], o['tok_kin'], o['tok_feat'], o['match'], o['match_cnt'])) +
                           (gt_boxes.to(dev), gt_labels.to(dev)))
        else:
            samples.append((feat.to(dev), boxes, metas, gt_boxes.to(dev), gt_labels.to(dev)))
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    rows, losses = [], []
    for it in range(a.warmup + a.steps):
        D.barrier()
        torch.cuda.synchronize()
        tr.zero_grad()
        t_f = t_b = 0.0
        ev[0].record()
        loss = 0.0
        pairs = []
        if a.lanes > 1 and not a.slice:       # samples in flight: one timed region for the whole step
            loss = pipe.step(samples, world=world) * a.samples
            ev[1].record(); ev[2].record(); ev[3].record()
            torch.cuda.synchronize()
            if it >= a.warmup:
                rows.append([ev[0].elapsed_time(ev[3]), 0.0, 0.0, 0.0, 0.0])
                losses.append(float(loss) / a.samples)
            continue
        for smp in samples:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            out = tr.forward(*smp)
            e[1].record()
            tr.backward()
            e[2].record()
            pairs.append(e)
            loss = loss + out['loss']
        ev[1].record()
        tr.all_reduce_grads()
        ev[2].record()
        tr.adamw_step(grad_scale=1.0 / (world * a.samples))
        ev[3].record()
        torch.cuda.synchronize()
        for e in pairs:
            t_f += e[0].elapsed_time(e[1])
            t_b += e[1].elapsed_time(e[2])
        if it >= a.warmup:
            rows.append([ev[0].elapsed_time(ev[3]), t_f, t_b, ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])])
            losses.append(float(loss) / a.samples)
    t = torch.tensor(rows, dtype=torch.float64)
    total_ms = D.max_over_ranks([float(t[:, 0].sum())], device=dev)[0]
    med = t.median(0).values.tolist()
    if rank == 0:
        what = 'decoder-slice' if a.slice else 'hot-path'
        line = dict(metric=f'samples/sec (MV2D-S {what} training step: fwd + targets/losses + bwd + grad all-reduce + AdamW)',
                    value=world * a.samples * a.steps / (total_ms * 1e-3), unit='samples/s', n_gpus=world, steps=a.steps,
                    warmup=a.warmup, samples_per_gpu=a.samples, lanes=a.lanes, N=6 * a.per_view, L=6,
                    step_ms=med[0], fwd_ms=med[1], bwd_ms=med[2], allreduce_ms=med[3], adamw_ms=med[4],
                    grad_bytes=tr.total * 4, loss_first=losses[0], loss_last=losses[-1],
                    scope='decoder slice only (rows a12-a18 + f3)' if a.slice else
                    'rows a1-a18 + f3: every hot-path parameter and d feat; the torch backbone is outside')
        print(json.dumps(line))
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', f'train_bench_{"slice_" if a.slice else ""}{world}gpu{"_lanes%d" % a.lanes if a.lanes > 1 else ""}.json'), 'w') as f:
            f.write(json.dumps(line) + '\n')


if __name__ == '__main__':
    main()
