"""Hardware multi-GPU tests (SURVEY.md section 4 "Distributed"): need >= 2 GPUs, skipped otherwise -- run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`; the record of that run is in profiles/.

1. N-GPU sharded step == 1-GPU step on the concatenated batch: two ranks with one sample each, NCCL sum all-reduce of
   the flat gradient buffer and the tiny all-reduce of the loss normaliser, against one process that steps over both
   samples.  The post-all-reduce gradients (all 14.0 M entries) must agree to fp32 summation order.
2. The plugin head under torch DistributedDataParallel with find_unused_parameters=False (what the reference's
   MMDistributedDataParallel does): every hot-path Parameter reaches its AccumulateGrad node, two iterations run, the
   averaged gradients are identical on both ranks and equal the flat-buffer result."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _samples():
    from mv2d_b200 import synth
    out = []
    for i, spec in enumerate(('s_small', 's_cfg1')):
        feat, boxes, metas = synth.case_inputs(dict(synth.CASES[spec], num_layers=2))
        gt_boxes, gt_labels, _ = synth.make_dn_inputs(dict(num_gt=4 + 5 * i, seed=40 + i))
        out.append((feat, boxes, metas, gt_boxes, gt_labels))
    return out


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from mv2d_b200 import synth
    from mv2d_b200.train import TrainStep
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    sd = synth.make_state_dict(0, num_layers=2)
    smp = _samples()[rank]
    step = TrainStep(sd, device=dev, lanes=1)
    loss = step.step([(smp[0].to(dev), smp[1], smp[2], smp[3].to(dev), smp[4].to(dev))], optimize=False, world=world)
    torch.cuda.synchronize()
    q.put((rank, step.main.grads.cpu(), float(loss)))
    dist.barrier()
    dist.destroy_process_group()


def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from conftest import ROOT
    from mv2d_b200 import synth
    from mv2d_b200.plugin.build import build_roi_head
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    head = build_roi_head(os.path.join(ROOT, 'configs', 'mv2d_b200', 'mv2d_s_r50_1408x512.py'), device=dev)
    head.load_state_dict(synth.make_state_dict(0), strict=True)
    head.train()
    head.trainer()                                   # Parameters become views of the flat buffer before DDP buckets them

    class Wrap(torch.nn.Module):
        def __init__(self, h):
            super().__init__()
            self.h = h

        def forward(self, feat, boxes, metas, gt_boxes, gt_labels):
            losses = self.h.forward_train([feat], metas, boxes, None, None, None, None, [gt_boxes], [gt_labels])
            return sum(losses.values())
    ddp = DDP(Wrap(head), device_ids=[rank], find_unused_parameters=False)
    feat, boxes, metas = synth.case_inputs(dict(synth.CASES['s_small' if rank == 0 else 's_cfg1'], num_layers=6))
    gt_boxes, gt_labels, _ = synth.make_dn_inputs(dict(num_gt=4 + 5 * rank, seed=40 + rank))
    for it in range(2):                              # the second iteration is the one an incomplete reduction breaks
        for prm in head.parameters():
            prm.grad = None
        loss = ddp(feat.to(dev).requires_grad_(True), [b.to(dev) for b in boxes], metas, gt_boxes.to(dev), gt_labels.to(dev))
        loss.backward()
    torch.cuda.synchronize()
    named = dict(head.named_parameters())
    tr = head.trainer()
    flat = torch.cat([named[n].grad.reshape(-1).cpu() for n in tr.table])
    q.put((rank, flat, float(loss)))
    dist.barrier()
    dist.destroy_process_group()


def _spawn(fn, world=2):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=fn, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return res


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_two_gpu_sharded_step_equals_one_gpu_step_on_the_concatenated_batch():
    from mv2d_b200 import synth
    from mv2d_b200.train import TrainStep
    res = _spawn(_worker)
    assert torch.equal(res[0][1], res[1][1]), 'the all-reduced gradient buffers differ between the ranks'
    sd = synth.make_state_dict(0, num_layers=2)
    one = TrainStep(sd, device='cuda:0', lanes=2)
    smp = [(s[0].cuda(), s[1], s[2], s[3].cuda(), s[4].cuda()) for s in _samples()]
    loss = one.step(smp, optimize=False)
    torch.cuda.synchronize()
    a, b = res[0][1].double(), one.main.grads.cpu().double()
    rel = float((a - b).norm() / b.norm())
    assert rel < 1e-5, f'2-GPU vs 1-GPU gradient: relative L2 difference {rel:.3e}'
    assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max()) + 1e-7
    assert abs(0.5 * (res[0][2] + res[1][2]) - float(loss)) < 1e-4 * max(abs(float(loss)), 1.0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_plugin_head_under_ddp_without_unused_parameter_search():
    res = _spawn(_ddp_worker)
    assert torch.isfinite(res[0][1]).all() and float(res[0][1].abs().max()) > 0
    assert torch.equal(res[0][1], res[1][1]), 'DDP-averaged gradients differ between the ranks'
