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


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (and, with first-touch allocation, its pinned
    staging buffers to that node's memory): with 8 ranks each pushing a 17 MB feature map per sample over PCIe, the
    host-side copies otherwise cross the socket interconnect.  Best effort: returns a small dict for the bench record."""
    info = dict(node=None, cpus=None)
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        if hasattr(pr, 'pci_bus_id'):
            pci = f'{getattr(pr, "pci_domain_id", 0):04x}:{pr.pci_bus_id:02x}:{getattr(pr, "pci_device_id", 0):02x}.0'
        else:
            import subprocess
            pci = subprocess.run(['nvidia-smi', '-i', str(local_rank), '--query-gpu=pci.bus_id', '--format=csv,noheader'],
                                 capture_output=True, text=True, timeout=10).stdout.strip().lower()
            if len(pci.split(':')[0]) == 8:      # nvidia-smi prints an 8-digit domain
                pci = pci[4:]
        with open(f'/sys/bus/pci/devices/{pci}/numa_node') as f:
            node = int(f.read().strip())
        if node < 0:
            return info
        with open(f'/sys/devices/system/node/node{node}/cpulist') as f:
            cpus = set()
            for part in f.read().strip().split(','):
                a, _, b = part.partition('-')
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(node=node, cpus=len(allowed))
    except Exception as e:      # containers without /sys access, non-Linux hosts
        info['error'] = f'{type(e).__name__}: {e}'[:120]
    return info


_WC_KEEP = []


def pin_host(t, write_combined=True):
    """A pinned host copy of ``t`` for asynchronous H2D copies.  write_combined: cudaHostAllocWriteCombined memory -- the CPU
    writes it once, the GPU's DMA reads it without snooping the CPU caches, which is what a staging buffer for feature maps is;
    with eight ranks pulling ~180 GB/s out of one host that path is the end-to-end ceiling.  Falls back
    to torch's pinned allocator if the runtime call is unavailable."""
    if write_combined:
        try:
            import ctypes
            rt = None
            for name in ('libcudart.so.12', 'libcudart.so'):
                try:
                    rt = ctypes.CDLL(name)
                    break
                except OSError:
                    continue
            if rt is None:
                raise OSError('libcudart not found')
            ptr = ctypes.c_void_p()
            nbytes = t.numel() * t.element_size()
            rc = rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04 | 0x01))   # WriteCombined | Portable
            if rc != 0 or not ptr.value:
                raise RuntimeError(f'cudaHostAlloc failed ({rc})')
            buf = (ctypes.c_uint8 * nbytes).from_address(ptr.value)
            out = torch.frombuffer(buf, dtype=t.dtype).view(t.shape)
            out.copy_(t)
            _WC_KEEP.append((rt, ptr, buf))          # lives for the process
            if out.is_pinned():
                return out
        except Exception:
            pass
    return t.pin_memory()
