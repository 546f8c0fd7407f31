"""The N > 1 path on the CPU: world_size 2 over gloo. Each rank scores its interleaved share of the
candidate list (here with the host simulation standing in for the GPU) and the slabs are all-gathered;
the result must be bit-identical to the single-process cache."""
import os
import socket
import subprocess
import sys

import numpy as np

from abn_b200 import shard

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {here!r})
import numpy as np, torch.distributed as dist
import helpers as H
from abn_b200 import prep, synth, shard
from abn_b200 import cache as abncache
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
names, data, dists, mp, _ = synth.make_config(4, n=600, n_nodes=6)
d = len(names)
P = prep.prepare(data, dists, "mle")
ch, par = abncache.enumerate_candidates(d, np.zeros((d, d), int), np.zeros((d, d), int), 3)
ctrl = H.control("mle", d)
def score(c, p):
    r = H.sim_score(P, c, p, ctrl)
    return {{k: r[k] for k in ("mlik", "aic", "bic", "mdl", "status", "iters")}}
full = shard.score_sharded(score, ch, par, world, rank, block=16)
if rank == 0:
    single = score(ch, par)
    for k in single:
        assert np.array_equal(full[k], single[k], equal_nan=True), k
    print("OK", len(ch))
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_partition_is_a_disjoint_cover_with_balanced_mix():
    rows = 10_000
    for world in (1, 2, 4, 8):
        parts = [shard.partition(rows, world, r, block=64) for r in range(world)]
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(rows))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 64
        assert all(len(p) <= shard.slab_len(rows, world, 64) for p in parts)


def test_world_size_2_gloo_gather_matches_single_process(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, here=HERE))
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "OK" in outs[0]
