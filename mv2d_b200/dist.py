"""Multi-GPU plumbing of the hot path.  The decoder is per-sample (the reference asserts B == 1,
detectors/mv2d.py:143), so ranks are independent replicas on disjoint samples: no data-path
collective in the forward.  torch.distributed is used to rendezvous, to barrier around the timed region,
to take the max over ranks of the device time (bench.py) and -- in the training step, the one real exchange
the path has -- to sum the flat gradient buffer over ranks (``all_reduce_sum``)."""
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


def all_reduce_sum(flat, group=None):
    """Sum a flat tensor over ranks in place (NCCL on GPUs, gloo in the CPU tests); returns the world size.
    With the parameters in one flat buffer this is the whole gradient exchange of a data-parallel step."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        return dist.get_world_size(group)
    return 1


def aggregate_throughput(world, steps, samples_per_step, max_total_ms):
    """Whole-job samples/s: all ranks' samples / the slowest rank's time."""
    return world * steps * samples_per_step / (max_total_ms * 1e-3)
