"""Multi-GPU plumbing of the hot path.  The decoder is per-sample (the reference asserts B == 1,
detectors/mv2d.py:143), so ranks are independent replicas on disjoint samples: no data-path
collective.  torch.distributed is used only to rendezvous, to barrier around the timed region
and to take the max over ranks of the device time (bench.py)."""
import os

import torch
import torch.distributed as dist


def env_rank():
    return (int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0')),
            int(os.environ.get('WORLD_SIZE', '1')))


def init(backend, device=None):
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        kw = dict(device_id=device) if (backend == 'nccl' and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_samples(num_samples, rank, world):
    """Sample ids of this rank: contiguous shards [r*B/G, (r+1)*B/G)."""
    per = num_samples // world
    assert per * world == num_samples, 'global batch must divide by the number of ranks'
    return list(range(rank * per, (rank + 1) * per))


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(values, device='cpu'):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def aggregate_throughput(world, steps, samples_per_step, max_total_ms):
    """Whole-job samples/s: all ranks' samples / the slowest rank's time."""
    return world * steps * samples_per_step / (max_total_ms * 1e-3)
