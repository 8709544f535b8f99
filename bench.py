#!/usr/bin/env python
"""bench.py -- headline benchmark of the MV2D decoder hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode S|T]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): MV2D-S, R50 single frame 1408x512, 6 cameras,
300 queries (50 2D boxes / view), 6 decoder layers, bs = 1.  A "step" is one pass of the hot
path -- (FPN P4 feature [6,256,32,88], per-view 2D boxes, img_metas) -> (cls_scores,
bbox_preds) of all 6 layers -- over one sample.  Metric: samples/sec, whole job.

  value  : whole-job throughput in samples/s, inputs already resident in HBM.  A step = one pass of the hot path over
           one BATCH of --batch (default 8) samples: one CUDA-graph replay of ONE kernel chain in which the batch is a
           segment dimension (HotPath.forward_batch; every sample's result is the one it gets at bs = 1, checked by
           tests/test_gpu_batch.py against the per-sample goldens).  mv2d_b200.pipeline.Pipeline keeps --depth
           (default 3) batches in flight on independent lanes.  The K steps are bracketed by one pair of CUDA events
           (+ barrier and synchronize on both sides); 2 x batch distinct samples rotate (277 MB of feature maps at
           batch 8 > the 126 MB L2).
  e2e    : the same through the public API with HOST (pinned) buffers: H2D of the feature map, boxes and
           camera matrices, the path, D2H of cls_scores/bbox_preds -- all inside the timed region.
  serial : (extra object) bs = 1, one sample at a time, nothing else in flight, L2 flushed (256 MB memset) before
           every step, each step timed with its own pair of CUDA events: the per-sample latency view of both numbers.
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
    ap.add_argument('--batch', type=int, default=None, help='samples per step (a segment dimension through one kernel chain); default 8 (S) / 2 (T)')
    ap.add_argument('--depth', type=int, default=0, help='batches in flight (pipelining on independent lanes); 1 = one batch at a time; 0 = 4 for the single-frame head, 3 for the two-frame head (tools/batch_sweep.py)')
    ap.add_argument('--no-two-frame', action='store_true', help='skip the extra MV2D-T (BASELINE configs[2]) section')
    ap.add_argument('--no-gpu-torch-baseline', action='store_true')
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 8 if args.mode == 'S' else 2
    return args


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
def workload_name(mode):
    return (f'MV2D-{mode} R50 1408x512 V={6 if mode == "S" else 12} N=300 L=6 '
            f'(BASELINE configs[{1 if mode == "S" else 2}])')


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
    steps, warm = max(args.steps, 1), max(args.warmup, 0)     # ~0.45 s per S sample on 16 cores: 50 + 5 steps take ~25 s
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
                config=dict(workload=workload_name(args.mode),
                            arm='CPU restatement of the reference (oracle/mv2d_oracle.py), one sample per step (the reference asserts bs = 1)'),
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
    numa = D.bind_to_gpu_numa_node(local_rank)      # before any pinned allocation
    if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
        os.environ['NCCL_DEBUG'] = 'WARN'       # keep stdout to the one JSON line (NCCL prints its version there)
    D.init('nccl', dev)
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world}'
    peaks = measured_peaks()

    mode = args.mode
    B = max(args.batch, 1)
    sd = synth.make_state_dict(0)
    from mv2d_b200.pipeline import Pipeline
    pipe = Pipeline(sd, mode=mode, device=dev, depth=args.depth if args.depth > 0 else (4 if mode == 'S' else 3))
    eng = pipe.lanes[0]
    # distinct samples per rank (different seeds per rank: replicas work on different data): 2 x B samples = two
    # batches whose feature maps (2 x B x 17.3 MB) exceed the 126 MB L2 for B >= 4
    n_var = max(2 * B, 8)
    samples = [make_inputs(mode, seed=i) for i in D.shard_samples(n_var * world, rank, world)]
    feats_dev = [s[0].to(dev) for s in samples]
    wc = os.environ.get('MV2D_BENCH_WC', '0') == '1'      # write-combined staging buffers (dist.pin_host): measured, no difference (tools/h2d_probe.py)
    feats_pin = [D.pin_host(s[0], wc) for s in samples]
    n_batches = n_var // B
    batches_dev = [torch.stack([feats_dev[k * B + j] for j in range(B)], 0) for k in range(n_batches)]
    batches_pin = [D.pin_host(torch.stack([samples[k * B + j][0] for j in range(B)], 0), wc) for k in range(n_batches)]
    batch_boxes = [[samples[k * B + j][1] for j in range(B)] for k in range(n_batches)]
    batch_metas = [[samples[k * B + j][2] for j in range(B)] for k in range(n_batches)]
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

    def submit(pp, host, i, nb=n_batches, bd=batches_dev, bp=batches_pin, bb=batch_boxes, bm=batch_metas):
        k = i % nb
        if bd[k].shape[0] == 1:
            return pp.submit((bp if host else bd)[k][0], bb[k][0], bm[k][0], to_host=host)
        return pp.submit_batch((bp if host else bd)[k], bb[k], bm[k], to_host=host)

    def timed_pipe(host, steps, warmup, pp=pipe, sub=submit):
        """K steps (batches) through the pipeline, ONE pair of events around all of them (device time, launching stream)."""
        for i in range(warmup + 2 * pp.depth):
            sub(pp, host, i)
        pp.join()
        barrier()
        launches0 = pp.launch_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall = time.perf_counter()
        a.record()
        for i in range(steps):
            sub(pp, host, i)
        pp.join()
        b.record()
        barrier()
        return a.elapsed_time(b), pp.launch_count() - launches0, time.perf_counter() - t_wall

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    W = max(args.warmup, 3)
    ser_total_ms, per_step, _, _ = timed(step_resident, args.steps, W)
    ser_e2e_total_ms, _, _, _ = timed(step_e2e, args.steps, W)
    total_ms, launches, wall = timed_pipe(False, args.steps, W)
    e2e_total_ms, _, _ = timed_pipe(True, args.steps, W)
    clocks = sampler.stop() if rank == 0 else None

    two_frame = None
    if mode == 'S' and not args.no_two_frame:
        try:
            two_frame = two_frame_section(sd, dev, world, rank, barrier, max(min(args.steps, 40), 4), W)
        except Exception as e:
            two_frame = dict(error=f'{type(e).__name__}: {e}'[:400])

    train = train_t = None
    if mode == 'S' and not args.no_train:
        try:
            train = train_step_section(sd, dev, samples, world, barrier)
        except Exception as e:      # the auxiliary section must not take the headline line down with it; the error is reported
            train = dict(error=f'{type(e).__name__}: {e}'[:400])
        try:        # the two-frame head with denoising queries (the configuration the reference trains MV2D-T with)
            smp_t = [make_inputs('T', seed=i) for i in D.shard_samples(2 * world, rank, world)]
            train_t = train_step_section(sd, dev, smp_t, world, barrier, steps=4, warmup=2, mode='T')
            train_t['scope'] = ('MV2D-T (12 views) with denoising queries: front end, key masks, 10 x G denoising queries, decoder over the '
                                'feature cells as keys, Hungarian + denoising losses, backward to every parameter and d feat')
        except Exception as e:
            train_t = dict(error=f'{type(e).__name__}: {e}'[:400])

    # max over ranks of the device time
    total_ms, e2e_total_ms, ser_total_ms, ser_e2e_total_ms = D.max_over_ranks(
        [total_ms, e2e_total_ms, ser_total_ms, ser_e2e_total_ms], device=dev)
    ms_per_step = total_ms / args.steps
    value = D.aggregate_throughput(world, args.steps, B, total_ms)
    e2e_value = D.aggregate_throughput(world, args.steps, B, e2e_total_ms)
    serial = dict(value=D.aggregate_throughput(world, args.steps, 1, ser_total_ms), unit=UNIT,
                  ms_per_step=ser_total_ms / args.steps, ms_min=min(per_step), ms_median=statistics.median(per_step),
                  e2e_value=D.aggregate_throughput(world, args.steps, 1, ser_e2e_total_ms),
                  e2e_ms_per_step=ser_e2e_total_ms / args.steps,
                  note='bs = 1: one sample at a time, L2 flushed (256 MB memset) before every step, per-step CUDA events')
    if two_frame is not None and 'total_ms' in two_frame:
        t_ms, t_e2e_ms = D.max_over_ranks([two_frame.pop('total_ms'), two_frame.pop('e2e_total_ms')], device=dev)
        two_frame['value'] = D.aggregate_throughput(world, two_frame['steps'], two_frame['batch'], t_ms)
        two_frame['e2e_value'] = D.aggregate_throughput(world, two_frame['steps'], two_frame['batch'], t_e2e_ms)
        two_frame['ms_per_step'] = t_ms / two_frame['steps']

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the attention path (the north star's named roofline) and of the dominant kernel:
    # the kernels are re-issued alone, L2 flushed, CUDA events on the launching stream.
    roof = roofline_section(eng, (batches_dev[0], batch_boxes[0], batch_metas[0]), mode, flush, peaks)
    N = roof.pop('N')

    feat, boxes, metas = samples[0]
    h2d = B * (feat.numel() * 4 + N * 5 * 4 + (len(metas) + 1) * 4 + 3 * len(metas) * 16 * 8)
    d2h = B * 2 * eng.L * N * 10 * 4
    line = dict(
        metric=METRIC if mode == 'S' else METRIC_T, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
        ms_per_step=ms_per_step, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
        data='synthetic',
        config=dict(workload=workload_name(mode), batch=B, samples_per_step=B,
                    batching=f'{B} samples per step per GPU travel through ONE kernel chain as a segment dimension (the reference asserts '
                             'bs = 1: detectors/mv2d.py:143); every sample gets the result it gets alone (tests/test_gpu_batch.py); '
                             '"serial" holds the bs = 1 one-at-a-time numbers',
                    precision='fp32 storage; single-pass TF32 tcgen05 in the PE MLPs, 3xTF32 tcgen05 in the QG conv, the QG FC chain and every '
                              'decoder GEMM (fp32 FFMA for those below 512 rows), fp32 FFMA attention, fp64 geometry',
                    schedule=f'{pipe.depth} batches in flight per GPU (independent lanes, mv2d_b200/pipeline.py)',
                    l2=f'{n_var} distinct samples rotate: {n_var * samples[0][0].numel() * 4 / 1e6:.0f} MB of feature maps > 126 MB L2 '
                       '(weights stay L2-resident, as in serving); the serial numbers flush L2 before every step',
                    timing='one pair of CUDA events on the launching stream around the K steps, barrier + synchronize on both sides, max over ranks',
                    launch='one CUDA-graph replay of the whole path per step',
                    sine_branch='recomputed every step (not cached); inside a step it is evaluated once for the whole batch when all samples '
                                'share the padding masks (a common subexpression: it does not depend on the inputs)', wall_s=wall),
        e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                 ms_per_step=e2e_total_ms / args.steps,
                 h2d_gb_per_s_per_gpu=h2d / (e2e_total_ms / args.steps) * 1e-6,
                 h2d_gb_per_s_all_gpus=world * h2d / (e2e_total_ms / args.steps) * 1e-6,
                 note='every sample carries its 17.3 MB fp32 feature map over PCIe inside the timed region; one GPU moves ~42 GB/s of '
                      'its x16 link (55 GB/s raw), eight ranks together saturate the host: plain pinned cudaMemcpyAsync from all eight '
                      'ranks at once tops out at 186 GB/s = 23 GB/s per GPU on this box (tools/h2d_probe.py), which is the ceiling of '
                      'the 8-GPU end-to-end number, not a collective or a kernel'),
        gpu_launches=launches, launches_per_sample=launches / (args.steps * B), clocks=clocks, numa=numa, roofline=roof['roofline'],
        attention_roofline=roof['attention'], stage_us=roof['stage_us'], peaks=peaks, serial=serial)
    if two_frame is not None:
        line['two_frame'] = two_frame
    if train is not None:
        line['train_step'] = train
    if train_t is not None:
        line['train_step_two_frame'] = train_t
    if not args.no_gpu_torch_baseline and world == 1 and mode == 'S':
        try:
            from oracle import gpu_baseline as GB
            line['gpu_torch_baseline'] = GB.time_s(sd, samples[:4], dev, warmup=5, iters=10)
        except Exception as e:
            line['gpu_torch_baseline'] = dict(error=f'{type(e).__name__}: {e}'[:400])
    if not args.no_cpu_baseline and world == 1:
        line['cpu_baseline'] = cpu_baseline(mode)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def two_frame_section(sd, dev, world, rank, barrier, steps, warmup, batch=2, depth=4):
    """Extra object `two_frame` -- BASELINE configs[2]: MV2D-T, 12 feature views, 300 queries, bs = 2 as a real batch (both
    samples in one kernel chain), `depth` batches in flight.  Same timing protocol as the headline."""
    import torch
    from mv2d_b200 import dist as D
    from mv2d_b200.pipeline import Pipeline
    pp = Pipeline(sd, mode='T', device=dev, depth=depth)
    n_var = 4 * batch
    smp = [make_inputs('T', seed=i) for i in D.shard_samples(n_var * world, rank, world)]
    nb = n_var // batch
    bd = [torch.stack([smp[k * batch + j][0] for j in range(batch)], 0).to(dev) for k in range(nb)]
    wc = os.environ.get('MV2D_BENCH_WC', '0') == '1'
    bp = [D.pin_host(torch.stack([smp[k * batch + j][0] for j in range(batch)], 0), wc) for k in range(nb)]
    bb = [[smp[k * batch + j][1] for j in range(batch)] for k in range(nb)]
    bm = [[smp[k * batch + j][2] for j in range(batch)] for k in range(nb)]
    res = {}
    for host in (False, True):
        for i in range(warmup + 2 * depth):
            pp.submit_batch((bp if host else bd)[i % nb], bb[i % nb], bm[i % nb], to_host=host)
        pp.join()
        barrier()
        l0 = pp.launch_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            pp.submit_batch((bp if host else bd)[i % nb], bb[i % nb], bm[i % nb], to_host=host)
        pp.join()
        b.record()
        barrier()
        res['e2e_total_ms' if host else 'total_ms'] = a.elapsed_time(b)
        if not host:
            res['gpu_launches'] = pp.launch_count() - l0
    res.update(metric=METRIC_T, unit=UNIT, workload=workload_name('T'), batch=batch, depth=depth, steps=steps, warmup=warmup,
               l2=f'{n_var} distinct samples rotate: {n_var * smp[0][0].numel() * 4 / 1e6:.0f} MB of feature maps > 126 MB L2',
               note='bs = 2 in one kernel chain; K/V projections of all cells and layers by one persistent 3xTF32 tcgen05 launch (csrc/kvproj.cu), key-stationary cross-attention on TF32 tensor cores (csrc/xa_tile.cuh)')
    del pp
    torch.cuda.empty_cache()
    return res


def train_step_section(sd, dev, samples, world, barrier, per_rank=2, steps=8, warmup=3, mode='S'):
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
    pipe = TrainStep(sd, device=dev, lanes=lanes, mode=mode)
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
                tm = pipe.timing      # CUDA events inside TrainStep.step: forward+backward | gradient all-reduce | AdamW
                if tm is not None:
                    rows.append([ev[0].elapsed_time(ev[3]), tm[0].elapsed_time(tm[1]), tm[1].elapsed_time(tm[2]), tm[2].elapsed_time(tm[3])])
                else:
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
                cuda_graphs=bool(getattr(pipe, 'use_graphs', False)),
                loss_first=losses[0], loss_last=losses[-1],
                scope='rows a1-a18 + f3 forward and backward: every hot-path parameter gradient and d loss / d feat; 3xTF32 tcgen05 for the '
                      'GPU-filling contractions, fp32 FFMA elsewhere; the torch backbone is outside',
                collective=('none (1 GPU)' if world == 1 else
                            'NCCL sum all-reduce of the flat gradient buffer in two buckets: the decoder slice (80 % of the bytes) on a '
                            'communication stream under the front-end half of the backward, the front-end slice after it; allreduce_ms is '
                            'the exposed part' if (mode == 'S' and getattr(pipe, 'overlap_all_reduce', False)) else
                            'one NCCL sum all-reduce of the flat gradient buffer per step'))


def roofline_section(eng, batch_in, mode, flush, peaks):
    """Dominant kernel re-issued alone (L2 flushed, CUDA events on the launching stream), the attention path's HBM
    roofline (a whole decoder layer AND the sparse cross-attention core alone), and the device time of each stage --
    all on one batch of the headline workload (eager launches of the stage entries)."""
    import numpy as np
    import torch
    from mv2d_b200 import lib as L
    h, W = eng.lib, eng.w
    feats, boxes_l, metas_l = batch_in
    B, V, _, hh, ww = feats.shape

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

    out = eng.forward_batch(feats, boxes_l, metas_l) if B > 1 else eng.forward(feats[0], boxes_l[0], metas_l[0])
    torch.cuda.synchronize()
    Np = out['Np'] if B > 1 else out['N']
    N = B * Np                                    # query rows of the batch
    n_per = out['n_b'][0] if B > 1 else out['N']

    # --- dominant kernel: the tcgen05 3xTF32 TMA-im2col GEMM of the query-generator 3x3 conv (all RoIs of the batch)
    ws, n_tok = eng._buf['qg_ws'], N * 49
    conv_out, thi, tlo = ws[:n_tok * 256], ws[n_tok * 256: 2 * n_tok * 256], ws[2 * n_tok * 256: 3 * n_tok * 256]

    def conv():
        L.check(h.mv2d_gemm_3xtf32(thi.data_ptr(), tlo.data_ptr(), 256, W.p('w_conv'), W.p('w_conv_lo'), 2304,
                                   W.p('b_conv'), conv_out.data_ptr(), 256, n_tok, 256, 2304, 1 | 128,
                                   L.stream_ptr()), 'conv')
    t_conv = time_fn(conv)
    flops = 2.0 * n_tok * 256 * 2304
    achieved = flops / (t_conv * 1e-6) / 1e12
    # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the committed `ncu --set full` capture
    # (profiles/r02_ncu_full_kernels.csv: B = 8, N = 300)
    traffic = {(8, 300): 257.363712e6 + 95.28448e6}.get((B, n_per))     # profiles/r02_ncu_full_kernels.csv
    roofline = dict(kernel='gemm_tc_kernel<128,3,im2col=3,2> (query-generator 3x3 conv, 3xTF32 tcgen05 + 4-D TMA, five RoIs per 256-row tile)',
                    bound='tensor', achieved=achieved, peak=peaks['bf16_tflops'], unit='TFLOP/s',
                    frac=achieved / peaks['bf16_tflops'], traffic=traffic, us_per_launch=t_conv, rois_per_launch=N,
                    algorithmic_flops=flops, tf32_flops_issued=3 * flops * 256.0 / 245.0,
                    note='achieved counts the conv flops once (17.3 GFLOP per 300 RoIs); the kernel issues 3 TF32 MMAs '
                         'per product (error compensation) on 256-row tiles that hold 245 real rows (ncu: tensor pipe 78 % '
                         'active, DRAM traffic = the operand + output bytes); peak = measured '
                         f'dense bf16 ({peaks["source"]}), the TF32 pipe is nominally half of it; kernel timed alone')

    # --- attention path: one decoder layer, algorithmic bytes (per sample x B) / time
    bt = None
    if B > 1:
        cams, rois, roi_start, bt = eng._upload_meta_batch(boxes_l, metas_l)
    qg = {k: out[k] for k in ('query_pos', 'ref', 'tok_feat', 'tok_kin')}
    grid = (hh, ww)
    kv = None
    if mode == 'S':
        corr = dict(match=out['match'], match_cnt=out['match_cnt'], max_match=out['max_match'])
        kin_rows, mem_rows = out['tok_kin'].view(-1, 256), out['tok_feat'].view(-1, 256)
        mc = float(out['match_cnt'].float().mean())
        bytes_layer = B * algorithmic_bytes_attention(Np, mc, 'S')
        extra = dict(matches_per_query=mc)
        q_core = torch.randn(N, 2048, device=feats.device)
    else:
        corr = {k: out.get(k) for k in ('keymask', 'mask_words', 'key_list', 'key_cnt', 'xa_prepared_for', 'row_tile_live')}
        mem_rows = out['feat_nhwc'].view(-1, 256)
        kin_rows = eng._buf['kin'][:mem_rows.numel()].view(-1, 256)
        km = out['keymask'].cpu().numpy().view(np.uint32).reshape(B, Np, -1)
        n_union = [int(np.unpackbits(np.bitwise_or.reduce(km[b], axis=0).view(np.uint8)).sum()) for b in range(B)]
        kmean = float(out['key_cnt'].float().mean())
        bytes_layer = sum(algorithmic_bytes_attention(Np, 0, 'T', (kmean, u)) for u in n_union)
        extra = dict(keys_per_query=kmean, union_keys=n_union)
        kv = (eng._buf['kp'].view(eng.L, -1, 256)[:, :mem_rows.shape[0]], eng._buf['vp'].view(eng.L, -1, 256)[:, :mem_rows.shape[0]])
        q_core = torch.randn(N, 256, device=feats.device)
    vel = eng._vel_dt(metas_l[0])
    t_dec = time_fn(lambda: eng.decoder(qg, corr, kin_rows, mem_rows, N, vel_dt=vel, batch=bt,
                                        kv=kv, grid=grid if kv is not None else None), reps=10)
    per_layer = t_dec / eng.L
    ach = bytes_layer / (per_layer * 1e-6) / 1e9
    attention = dict(bound='hbm', achieved=ach, peak=peaks['hbm_gbs'], unit='GB/s', frac=ach / peaks['hbm_gbs'],
                     algorithmic_bytes_per_layer=bytes_layer, us_per_layer=per_layer, decoder_stage_us=t_dec, batch=B,
                     note='achieved = SURVEY 8d algorithmic bytes of one cross-attention layer (x samples of the batch) / (decoder '
                          'stage time / L), i.e. a WHOLE decoder layer (self-attn + sparse cross-attn + FFN + LayerNorms) is '
                          'charged to the attention bytes; `core` below isolates the cross-attention kernel(s)', **extra)
    try:    # the sparse cross-attention core alone (mv2d_cross_attention_core: xa_roi / xt_attn + xt_merge)
        t_core = time_fn(lambda: eng.cross_attention_core(qg, corr, kin_rows, mem_rows, N, q_core, layer=0, kv=kv,
                                                          grid=grid if kv is not None else None, batch=bt))
        # S: every (query, matched RoI) unit streams the RoI's 49 key-input + 49 value rows; T: projected K / V of the union
        core_bytes = (float(out['match_cnt'].sum()) * 49 * 2 * 1024) if mode == 'S' else bytes_layer
        attention['core'] = dict(us=t_core, bytes=core_bytes, achieved=core_bytes / t_core / 1e3,
                                 frac=core_bytes / t_core / 1e3 / peaks['hbm_gbs'],
                                 # dram__bytes_read + write of the kernel(s) in the committed ncu captures (S: profiles/r02_ncu_full_kernels_b.csv, T: r02_ncu_full_kernels.csv)
                                 traffic={('S', 8, 300): 260.83328e6 + 37.432576e6, ('T', 2, 300): 131.600896e6 + 21.562368e6}.get((mode, B, n_per)),
                                 note='cross-attention kernel(s) alone, L2 flushed; bytes = key / value rows the kernel has to stream '
                                      '(S: 100 KB per (query, RoI) unit; T: the SURVEY 8d bytes); ncu dram bytes: profiles/')
    except Exception as e:
        attention['core'] = dict(error=f'{type(e).__name__}: {e}'[:300])

    # --- per-stage device time (eager launches, L2 flushed before each stage), whole batch
    feat_nchw = feats.view(B * V, 256, hh, ww)
    st = {}
    st['nchw_to_nhwc'] = time_fn(lambda: eng.to_nhwc(feat_nchw))
    f, f32r = eng.to_nhwc(feat_nchw)
    if B > 1:
        i2l = eng._get('img2lidar', (B * V, 16), torch.float64)
        trans = eng._get('trans', (B, V, V, 16), torch.float64)
        gp = lambda: L.check(h.mv2d_geom_prep_batch(cams[0].data_ptr(), B, V, i2l.data_ptr(), trans.data_ptr(), L.stream_ptr()), 'geom')
        metas0 = metas_l[0]
    else:
        cams, rois, roi_start, counts, _ = eng._upload_meta(boxes_l[0], metas_l[0])
        gp = lambda: eng.geom_prep(cams)
        i2l, trans = eng.geom_prep(cams)
        metas0 = metas_l[0]
    st['geom_prep'] = time_fn(gp)
    st['pe3d'] = time_fn(lambda: eng.pe3d(f, i2l, metas0, f32r, batch=bt))
    pe, kin = eng.pe3d(f, i2l, metas0, f32r, batch=bt)
    st['roi_align_qg'] = time_fn(lambda: eng.roi_align_qg(rois, cams, f, pe, N))
    st['box_corr'] = time_fn(lambda: eng.box_corr(rois, roi_start, trans, N, V, metas0, hh, ww, batch=bt))
    st['decoder'] = t_dec
    st['batch'] = B
    return dict(roofline=roofline, attention=attention, stage_us=st, N=n_per)


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
