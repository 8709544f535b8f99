#!/usr/bin/env python
"""bench.py -- headline benchmark of the MV2D decoder hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode S|T]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): MV2D-S, R50 single frame 1408x512, 6 cameras,
300 queries (50 2D boxes / view), 6 decoder layers, bs = 1.  A "step" is one pass of the hot
path -- (FPN P4 feature [6,256,32,88], per-view 2D boxes, img_metas) -> (cls_scores,
bbox_preds) of all 6 layers -- over one sample.  Metric: samples/sec, whole job.

  value  : whole-job throughput, inputs already resident in HBM.  Each step = one CUDA-graph replay of the
           whole path for ONE sample (bs = 1); mv2d_b200.pipeline.Pipeline keeps --depth (default 4) samples in
           flight on independent lanes, so the GPU-filling front end of sample i+1 runs under the latency-bound
           decoder of sample i.  The K steps are bracketed by one pair of CUDA events (+ barrier and
           synchronize on both sides); 8 distinct samples rotate (138 MB of feature maps > the 126 MB L2).
  e2e    : the same through the public API with HOST (pinned) buffers: H2D of the feature map, boxes and
           camera matrices, the path, D2H of cls_scores/bbox_preds -- all inside the timed region.
  serial : (extra object) one sample at a time, nothing else in flight, L2 flushed (256 MB memset) before every
           step, each step timed with its own pair of CUDA events: the per-sample latency view of both numbers.
  N > 1  : one process per GPU, independent replicas on different samples (the decoder is
           per-sample: no data-path collective); NCCL only for the barrier and the max-over-ranks.
  --impl reference : the CPU restatement of the reference (oracle/, kind "port": the reference's
           own Python needs mmcv/mmdet and /root/reference, neither exists on the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'samples/sec (MV2D-S decoder hot path, 6-cam 1408x512, 300 queries, 6 layers)'
METRIC_T = 'samples/sec (MV2D-T decoder hot path, two frames = 12 views 1408x512, 300 queries, 6 layers)'
UNIT = 'samples/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mode', default='S', choices=['S', 'T'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train', action='store_true', help='skip the extra training-step section')
    ap.add_argument('--depth', type=int, default=4, help='samples in flight (inter-sample pipelining); 1 = serial')
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d['hbm_gbs'], bf16_tflops=d['bf16_tflops'], source='measured (MEASURED_PEAKS.json)')
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------- workload
def make_inputs(mode, seed):
    from mv2d_b200 import synth
    case = dict(synth.CASES['s_cfg2' if mode == 'S' else 't_cfg3'], seed=seed)
    return synth.case_inputs(case)


def algorithmic_bytes_attention(N, match_cnt_mean, mode, keys_mean=None):
    """SURVEY.md section 8d, per decoder cross-attention layer (fp32 = 4 B).
    S: unique RoI tokens (key-input + memory rows, each read once) + Q/out + absorbed weights + mask."""
    C = 256
    if mode == 'S':
        n_k = 49 * N
        mask = N * match_cnt_mean * 49
    else:
        n_k = keys_mean[1]          # union of keys
        mask = N * n_k              # the reference's 1 B / element bool mask
    return 4 * (2 * n_k * C + 2 * N * C + 4 * C * C + 4 * C) + mask


def run_reference(args, rank):
    """CPU arm: the oracle port of the reference on the host cores, bounded sample."""
    import torch
    from mv2d_b200 import synth
    from oracle import mv2d_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.make_state_dict(0)
    fn = O.mv2d_s_forward if args.mode == 'S' else O.mv2d_t_forward
    cfg = O.make_cfg(args.mode)
    steps, warm = min(args.steps, 8), min(args.warmup, 1)
    times = []
    with torch.no_grad():
        for i in range(warm + steps):
            feat, boxes, metas = make_inputs(args.mode, seed=i % 4)
            t0 = time.perf_counter()
            fn(sd, feat, boxes, metas, cfg)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = 1e3 / ms
    sample = f'{steps} samples of the full workload after {warm} warm-up, torch {torch.__version__} CPU fp32'
    line = dict(metric=METRIC if args.mode == 'S' else METRIC_T, value=val, unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warm, ms_per_step=ms,
                higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                impl='reference',
                config=dict(workload=f'MV2D-{args.mode} R50 1408x512 V={6 if args.mode == "S" else 12} N=300 L=6 bs=1',
                            arm='CPU restatement of the reference (oracle/mv2d_oracle.py)'),
                cpu_baseline=dict(value=val, unit=UNIT, cores=cores, kind='port', sample=sample),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def main():
    args = parse()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        if rank == 0:
            run_reference(args, rank)
        return
    import torch
    import torch.distributed as dist
    from mv2d_b200 import dist as D
    from mv2d_b200 import synth
    from mv2d_b200.engine import HotPath
    assert torch.cuda.is_available(), 'bench.py needs a GPU (mv2d_b200 has no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
        os.environ['NCCL_DEBUG'] = 'WARN'       # keep stdout to the one JSON line (NCCL prints its version there)
    D.init('nccl', dev)
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world}'
    peaks = measured_peaks()

    mode = args.mode
    sd = synth.make_state_dict(0)
    from mv2d_b200.pipeline import Pipeline
    pipe = Pipeline(sd, mode=mode, device=dev, depth=max(args.depth, 1))
    eng = pipe.lanes[0]
    # distinct samples per rank (different seeds per rank: replicas work on different data); 8 feature maps
    # of the S head are 138 MB > the 126 MB L2
    n_var = 8
    samples = [make_inputs(mode, seed=i) for i in D.shard_samples(n_var * world, rank, world)]
    feats_dev = [s[0].to(dev) for s in samples]
    feats_pin = [s[0].pin_memory() for s in samples]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        f, (_, boxes, metas) = feats_dev[i % n_var], samples[i % n_var]
        return eng.forward(f, boxes, metas, use_graph=True)

    host_out = {}

    def step_e2e(i):
        f, (_, boxes, metas) = feats_pin[i % n_var], samples[i % n_var]
        out = eng.forward(f, boxes, metas, use_graph=True)          # H2D of feat/boxes/cams inside
        if 'cls' not in host_out:
            host_out['cls'] = torch.empty(out['cls_scores'].shape, dtype=torch.float32).pin_memory()
            host_out['box'] = torch.empty(out['bbox_preds'].shape, dtype=torch.float32).pin_memory()
        host_out['cls'].copy_(out['cls_scores'], non_blocking=True)  # D2H of the step's result
        host_out['box'].copy_(out['bbox_preds'], non_blocking=True)
        return host_out

    def timed(step, steps, warmup):
        for i in range(warmup):
            step(i)
        barrier()
        evs = []
        launches0 = eng.launch_count()
        t_wall = time.perf_counter()
        for i in range(steps):
            flush.zero_()                                            # flush L2 (not timed)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step(warmup + i)
            b.record()
            evs.append((a, b))
        barrier()
        wall = time.perf_counter() - t_wall
        ms = [a.elapsed_time(b) for a, b in evs]
        return sum(ms), ms, eng.launch_count() - launches0, wall

    def timed_pipe(host, steps, warmup):
        """K samples through the pipeline, ONE pair of events around all of them (device time, launching stream)."""
        src = feats_pin if host else feats_dev
        for i in range(warmup + 2 * pipe.depth):
            pipe.submit(src[i % n_var], samples[i % n_var][1], samples[i % n_var][2], to_host=host)
        pipe.join()
        barrier()
        launches0 = pipe.launch_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall = time.perf_counter()
        a.record()
        for i in range(steps):
            pipe.submit(src[i % n_var], samples[i % n_var][1], samples[i % n_var][2], to_host=host)
        pipe.join()
        b.record()
        barrier()
        return a.elapsed_time(b), pipe.launch_count() - launches0, time.perf_counter() - t_wall

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    W = max(args.warmup, 3)
    ser_total_ms, per_step, _, _ = timed(step_resident, args.steps, W)
    ser_e2e_total_ms, _, _, _ = timed(step_e2e, args.steps, W)
    total_ms, launches, wall = timed_pipe(False, args.steps, W)
    e2e_total_ms, _, _ = timed_pipe(True, args.steps, W)
    clocks = sampler.stop() if rank == 0 else None

    train = None
    if mode == 'S' and not args.no_train:
        try:
            train = train_step_section(sd, dev, samples, world, barrier)
        except Exception as e:      # the auxiliary section must not take the headline line down with it; the error is reported
            train = dict(error=f'{type(e).__name__}: {e}'[:400])

    # max over ranks of the device time
    total_ms, e2e_total_ms, ser_total_ms, ser_e2e_total_ms = D.max_over_ranks(
        [total_ms, e2e_total_ms, ser_total_ms, ser_e2e_total_ms], device=dev)
    ms_per_step = total_ms / args.steps
    value = D.aggregate_throughput(world, args.steps, 1, total_ms)
    e2e_value = D.aggregate_throughput(world, args.steps, 1, e2e_total_ms)
    serial = dict(value=D.aggregate_throughput(world, args.steps, 1, ser_total_ms), unit=UNIT,
                  ms_per_step=ser_total_ms / args.steps, ms_min=min(per_step), ms_median=statistics.median(per_step),
                  e2e_value=D.aggregate_throughput(world, args.steps, 1, ser_e2e_total_ms),
                  e2e_ms_per_step=ser_e2e_total_ms / args.steps,
                  note='one sample at a time, L2 flushed (256 MB memset) before every step, per-step CUDA events')

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the attention path (the north star's named roofline) and of the dominant kernel:
    # the kernels are re-issued alone, L2 flushed, CUDA events on the launching stream.
    out = step_resident(0)
    torch.cuda.synchronize()
    N = out['N']
    out['_inputs'] = (feats_dev[0], samples[0][1], samples[0][2])
    out['_metas'] = samples[0][2]
    roof = roofline_section(eng, out, mode, N, flush, peaks)

    feat, boxes, metas = samples[0]
    h2d = feat.numel() * 4 + N * 5 * 4 + (len(metas) + 1) * 4 + 3 * len(metas) * 16 * 8
    d2h = 2 * eng.L * N * 10 * 4
    line = dict(
        metric=METRIC if mode == 'S' else METRIC_T, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
        ms_per_step=ms_per_step, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
        data='synthetic',
        config=dict(workload=f'MV2D-{mode} R50 1408x512 V={len(metas)} N={N} L={eng.L} bs=1 per GPU (BASELINE configs[{1 if mode == "S" else 2}])',
                    precision='fp32 storage; single-pass TF32 tcgen05 in the PE MLPs, 3xTF32 tcgen05 in the QG conv and the four wide decoder GEMMs, fp32 FFMA elsewhere, fp64 geometry',
                    schedule=f'{pipe.depth} samples in flight per GPU (inter-sample pipelining on independent lanes, '
                             'each sample processed at bs=1; mv2d_b200/pipeline.py); "serial" holds the one-at-a-time numbers',
                    l2=f'{n_var} distinct samples rotate: {n_var * samples[0][0].numel() * 4 / 1e6:.0f} MB of feature maps > 126 MB L2 '
                       '(weights stay L2-resident, as in serving); the serial numbers flush L2 before every step',
                    timing='one pair of CUDA events on the launching stream around the K steps, barrier + synchronize on both sides, max over ranks',
                    launch='one CUDA-graph replay of the whole path per step',
                    sine_branch='recomputed every step (not cached)', wall_s=wall),
        e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                 ms_per_step=e2e_total_ms / args.steps),
        gpu_launches=launches, clocks=clocks, roofline=roof['roofline'], attention_roofline=roof['attention'],
        stage_us=roof['stage_us'], peaks=peaks, serial=serial)
    if train is not None:
        line['train_step'] = train
    if not args.no_cpu_baseline and world == 1:
        line['cpu_baseline'] = cpu_baseline(mode)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def train_step_section(sd, dev, samples, world, barrier, per_rank=2, steps=8, warmup=3):
    """Extra object `train_step` (BASELINE configs[3] shape: MV2D-S, 2 samples per GPU): per step and rank, for each local
    sample the hot-path training forward (saved activations) + Hungarian targets / losses + backward down to d feat (the
    two samples in flight on their own streams, `lanes`), then ONE NCCL sum all-reduce of the flat gradient buffer (all
    14.0 M hot-path parameters) and a fused AdamW pass.
    Device time by CUDA events, max over ranks; the torch backbone is outside the hot path."""
    import torch
    from mv2d_b200 import dist as D
    from mv2d_b200 import synth
    from mv2d_b200.train import TrainStep
    lanes = int(os.environ.get('MV2D_TRAIN_LANES', '2'))     # samples in flight per GPU (mv2d_b200.train.TrainStep)
    pipe = TrainStep(sd, device=dev, lanes=lanes)
    tr = pipe.main
    before = tr.lib.mv2d_launch_count()
    batch = []
    for i in range(per_rank):
        feat, boxes, metas = samples[i % len(samples)]
        gt_boxes, gt_labels, _ = synth.make_dn_inputs(dict(num_gt=30, seed=300 + i))
        batch.append((feat.to(dev), boxes, metas, gt_boxes.to(dev), gt_labels.to(dev)))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    rows, losses = [], []
    for it in range(warmup + steps):
        barrier()
        if lanes > 1:
            ev[0].record()
            loss = pipe.step(batch, world=world) * per_rank
            ev[1].record(); ev[2].record(); ev[3].record()
            torch.cuda.synchronize()
            if it >= warmup:
                rows.append([ev[0].elapsed_time(ev[3]), ev[0].elapsed_time(ev[3]), 0.0, 0.0])
                losses.append(float(loss) / per_rank)
            continue
        tr.zero_grad()
        ev[0].record()
        loss = 0.0
        for smp in batch:
            loss = loss + tr.forward(*smp)['loss']
            tr.backward()
        ev[1].record()
        tr.all_reduce_grads()
        ev[2].record()
        tr.adamw_step(grad_scale=1.0 / (world * per_rank))
        ev[3].record()
        torch.cuda.synchronize()
        if it >= warmup:
            rows.append([ev[0].elapsed_time(ev[3]), ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])])
            losses.append(float(loss) / per_rank)
    t = torch.tensor(rows, dtype=torch.float64)
    total_ms = D.max_over_ranks([float(t[:, 0].sum())], device=dev)[0]
    med = t.median(0).values.tolist()
    return dict(value=world * per_rank * steps / (total_ms * 1e-3), unit='samples/s', samples_per_gpu=per_rank, lanes=lanes, steps=steps, warmup=warmup,
                step_ms=med[0], fwd_bwd_ms=med[1], allreduce_ms=med[2], adamw_ms=med[3], grad_bytes=tr.total * 4,
                launches_per_step=int(tr.lib.mv2d_launch_count() - before) // (warmup + steps),
                loss_first=losses[0], loss_last=losses[-1],
                scope='rows a1-a18 + f3 forward and backward: every hot-path parameter gradient and d loss / d feat; 3xTF32 tcgen05 for the '
                      'GPU-filling contractions, fp32 FFMA elsewhere; the torch backbone is outside',
                collective='one NCCL sum all-reduce of the flat gradient buffer per step' if world > 1 else 'none (1 GPU)')


def roofline_section(eng, out, mode, N, flush, peaks):
    """Dominant kernel re-issued alone (L2 flushed, CUDA events on the launching stream), the
    attention path's HBM roofline, and the device time of each stage."""
    import numpy as np
    import torch
    from mv2d_b200 import lib as L
    h, W = eng.lib, eng.w

    def time_fn(fn, reps=20):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        return statistics.median(ts)

    # --- dominant kernel: the tcgen05 3xTF32 TMA-im2col GEMM of the query-generator 3x3 conv
    ws, n_tok = eng._buf['qg_ws'], N * 49
    conv_out, thi, tlo = ws[:n_tok * 256], ws[n_tok * 256: 2 * n_tok * 256], ws[2 * n_tok * 256: 3 * n_tok * 256]

    def conv():
        L.check(h.mv2d_gemm_3xtf32(thi.data_ptr(), tlo.data_ptr(), 256, W.p('w_conv'), W.p('w_conv_lo'), 2304,
                                   W.p('b_conv'), conv_out.data_ptr(), 256, n_tok, 256, 2304, 1 | 128,
                                   L.stream_ptr()), 'conv')
    t_conv = time_fn(conv)
    flops = 2.0 * n_tok * 256 * 2304
    achieved = flops / (t_conv * 1e-6) / 1e12
    roofline = dict(kernel='gemm_tc_kernel<128,3,im2col,3> (query-generator 3x3 conv, 3xTF32 tcgen05 + 4-D TMA)',
                    bound='tensor', achieved=achieved, peak=peaks['bf16_tflops'], unit='TFLOP/s',
                    frac=achieved / peaks['bf16_tflops'],
                    # dram__bytes_read.sum + dram__bytes_write.sum of this launch at N = 300 from the committed
                    # `ncu --set full` capture (profiles/r01_ncu_full_tc_kernels.csv: 34.88 MB + 0.07 MB); the
                    # algorithmic bytes are 2 x 15.05 MB of hi/lo tokens + 4.7 MB of weights + 15.05 MB out (L2-resident)
                    traffic=(34.882816e6 + 0.07168e6) if N == 300 else None, us_per_launch=t_conv,
                    algorithmic_flops=flops, tf32_flops_issued=3 * flops * 128.0 / 98.0,
                    note='achieved counts the conv flops once (17.3 GFLOP at N=300); the kernel issues 3 TF32 MMAs '
                         'per product (error compensation) on 128-row tiles that hold 98 real rows; peak = measured '
                         f'dense bf16 ({peaks["source"]}), the TF32 pipe is nominally half of it; kernel timed alone')

    # --- attention path: one decoder layer's cross-attention, algorithmic bytes / time
    qg = {k: out[k] for k in ('query_pos', 'ref', 'tok_feat', 'tok_kin')}
    if mode == 'S':
        corr = dict(match=out['match'], match_cnt=out['match_cnt'], max_match=out['max_match'])
        kin_rows, mem_rows = out['tok_kin'].view(-1, 256), out['tok_feat'].view(-1, 256)
        mc = float(out['match_cnt'].float().mean())
        bytes_layer = algorithmic_bytes_attention(N, mc, 'S')
        extra = dict(matches_per_query=mc)
    else:
        corr = dict(keymask=out['keymask'], mask_words=out['mask_words'], key_list=out['key_list'], key_cnt=out['key_cnt'])
        mem_rows = out['feat_nhwc'].view(-1, 256)
        kin_rows = eng._buf['kin'][:mem_rows.numel()].view(-1, 256)
        km = out['keymask'].cpu().numpy().view(np.uint32)
        n_union = int(np.unpackbits(np.bitwise_or.reduce(km, axis=0).view(np.uint8)).sum())
        kmean = float(out['key_cnt'].float().mean())
        bytes_layer = algorithmic_bytes_attention(N, 0, 'T', (kmean, n_union))
        extra = dict(keys_per_query=kmean, union_keys=n_union)
    vel = eng._vel_dt(out['_metas'])
    t_dec = time_fn(lambda: eng.decoder(qg, corr, kin_rows, mem_rows, N, vel_dt=vel), reps=10)
    per_layer = t_dec / eng.L
    ach = bytes_layer / (per_layer * 1e-6) / 1e9
    attention = dict(bound='hbm', achieved=ach, peak=peaks['hbm_gbs'], unit='GB/s', frac=ach / peaks['hbm_gbs'],
                     algorithmic_bytes_per_layer=bytes_layer, us_per_layer=per_layer, decoder_stage_us=t_dec,
                     note='achieved = SURVEY 8d algorithmic bytes of one cross-attention layer / (decoder stage '
                          'time / L), i.e. a whole decoder layer (self-attn + sparse cross-attn + FFN, 11 launches) '
                          'is charged to the attention bytes; at N=300 the layer is launch/latency bound', **extra)

    # --- per-stage device time (eager launches, L2 flushed before each stage)
    feat_nchw, boxes, metas = out['_inputs']
    V, _, hh, ww = feat_nchw.shape
    cams, rois, roi_start, counts, _ = eng._upload_meta(boxes, metas)
    st = {}
    st['nchw_to_nhwc'] = time_fn(lambda: eng.to_nhwc(feat_nchw))
    f, f32r = eng.to_nhwc(feat_nchw)
    st['geom_prep'] = time_fn(lambda: eng.geom_prep(cams))
    i2l, trans = eng.geom_prep(cams)
    st['pe3d'] = time_fn(lambda: eng.pe3d(f, i2l, metas, f32r))
    pe, kin = eng.pe3d(f, i2l, metas, f32r)
    st['roi_align_qg'] = time_fn(lambda: eng.roi_align_qg(rois, cams, f, pe, N))
    st['box_corr'] = time_fn(lambda: eng.box_corr(rois, roi_start, trans, N, V, metas, hh, ww))
    st['decoder'] = t_dec
    return dict(roofline=roofline, attention=attention, stage_us=st)


def cpu_baseline(mode):
    """The oracle port timed on this box's host cores, bounded sample (rank 0, N=1 only)."""
    import torch
    from mv2d_b200 import synth
    from oracle import mv2d_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.make_state_dict(0)
    fn = O.mv2d_s_forward if mode == 'S' else O.mv2d_t_forward
    cfg = O.make_cfg(mode)
    n, times = (6 if mode == 'S' else 3), []
    with torch.no_grad():
        for i in range(n + 1):
            feat, boxes, metas = make_inputs(mode, seed=i)
            t0 = time.perf_counter()
            fn(sd, feat, boxes, metas, cfg)
            if i > 0:
                times.append(time.perf_counter() - t0)
    v = len(times) / sum(times)
    return dict(value=v, unit=UNIT, cores=cores, kind='port',
                sample=f'{n} samples of the same workload after 1 warm-up (oracle/mv2d_oracle.py, torch CPU fp32, '
                       f'{cores} threads)')


if __name__ == '__main__':
    main()
