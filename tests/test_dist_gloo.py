"""CPU, world_size 2, gloo: the N > 1 host logic of bench.py (sharding, barrier, max over ranks,
throughput aggregation) and of the training step (flat gradient buffer, sum all-reduce).  The forward has no
data-path collective (replicas); the gradient all-reduce is the one exchange of the path."""
import os
import socket
import subprocess
import sys

from conftest import ROOT

WORKER = r'''
import json, os, sys
sys.path.insert(0, %r)
from mv2d_b200 import dist as D
rank, local_rank, world = D.init('gloo')
shard = D.shard_samples(8, rank, world)
D.barrier()
tmax = D.max_over_ranks([10.0 + 5.0 * rank, 3.0 - rank])
val = D.aggregate_throughput(world, 4, 1, tmax[0])
# training step: flat parameter / gradient buffers on the CPU (layout from the library), gradient all-reduce
import torch
from mv2d_b200 import synth
from mv2d_b200.train import DecoderTrainer
tr = DecoderTrainer(synth.make_state_dict(0, num_layers=1), device='cpu')
tr.grads.fill_(float(rank + 1))
tr.grad('bbox_head.cls_branches.0.6.bias').fill_(10.0 * (rank + 1))
n = tr.all_reduce_grads()
back = tr.state_dict()
same = all(torch.equal(back[k], v) for k, v in synth.make_state_dict(0, num_layers=1).items() if k in back)
print(json.dumps(dict(rank=rank, world=world, shard=shard, tmax=tmax, value=val, ar_world=n,
                      g_first=float(tr.grads[0]), g_bias=tr.grad('bbox_head.cls_branches.0.6.bias').tolist(),
                      n_params=tr.total, roundtrip=bool(same), n_tensors=len(back))))
'''


def test_two_rank_gloo_plumbing(tmp_path):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % ROOT)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1',
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, text=True))
    import json
    outs = []
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0
        outs.append(json.loads(out.strip().splitlines()[-1]))
    outs.sort(key=lambda o: o['rank'])
    assert outs[0]['shard'] == [0, 1, 2, 3] and outs[1]['shard'] == [4, 5, 6, 7]
    for o in outs:
        assert o['world'] == 2 and o['tmax'] == [15.0, 3.0]
        assert abs(o['value'] - 2 * 4 / 0.015) < 1e-6
        # the data-parallel exchange of the training step: one sum over the flat gradient buffer
        assert o['ar_world'] == 2 and o['g_first'] == 3.0 and o['g_bias'] == [30.0] * 10
        assert o['roundtrip'] and o['n_tensors'] == 6 + 34 + 22 and o['n_params'] >= 1_800_000
