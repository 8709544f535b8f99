"""CPU, world_size 2, gloo: the N > 1 host logic of bench.py (sharding, barrier, max over ranks,
throughput aggregation).  No data-path collective exists on this path (replicas)."""
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
print(json.dumps(dict(rank=rank, world=world, shard=shard, tmax=tmax, value=val)))
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
