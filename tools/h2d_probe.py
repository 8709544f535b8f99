"""Host-to-device copy bandwidth of the e2e staging buffers, per rank and over all ranks at once:
   python tools/h2d_probe.py            |  torchrun --nproc-per-node N tools/h2d_probe.py
Compares torch's pinned allocator with write-combined cudaHostAlloc memory (mv2d_b200.dist.pin_host)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mv2d_b200 import dist as D  # noqa: E402


def main():
    rank, local_rank, world = D.env_rank()
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    D.init('nccl', dev)
    n = 8 * 6 * 256 * 32 * 88            # one batch of 8 feature maps: 138 MB
    src = torch.randn(n)
    dst = torch.empty(n, device=dev)
    res = {}
    for kind in ('pinned', 'write_combined'):
        h = D.pin_host(src, kind == 'write_combined')
        for _ in range(3):
            dst.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        D.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            dst.copy_(h, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        ms = D.max_over_ranks([a.elapsed_time(b)], device=dev)[0]
        res[kind] = dict(gb_per_s_per_gpu=20 * n * 4 / ms * 1e-6, gb_per_s_all=world * 20 * n * 4 / ms * 1e-6, is_pinned=bool(h.is_pinned()))
        assert torch.equal(dst.cpu(), src)
    if rank == 0:
        print(json.dumps(dict(world=world, **res)))


if __name__ == '__main__':
    main()
