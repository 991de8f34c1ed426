"""world_size-2 `gloo` test of the pair sharding (CPU, no GPU): two processes each align their
share of the pairs with the oracle standing in for the device library, results are gathered on
rank 0 and must equal the single-process result bit for bit."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np
import torch.distributed as dist
from cvo_slam_b200 import capi, parallel
from oracle import oracle

def run_pairs(orc, clouds, pairs, idx):
    cal = capi.TUM1_CALIB()
    res = np.zeros(len(idx), dtype=capi.RESULT_DTYPE)
    h = orc.create(cal)
    for k, p in enumerate(idx):
        fi, mi = pairs[p]
        orc.set_cloud(h, 0, *clouds[fi]); orc.set_cloud(h, 1, *clouds[mi])
        orc.set_ell(h, 0.15); orc.set_RT(h, np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
        r, _ = orc.align(h)
        res[k]["transform"] = np.array(r.transform); res[k]["iterations"] = r.iterations
        res[k]["A_nonzero"] = r.A_nonzero; res[k]["ell"] = r.ell
    orc.destroy(h)
    return res

def make_clouds(n, seed):
    rng = np.random.default_rng(seed)
    base = rng.uniform(-0.5, 0.5, (300, 3)).astype(np.float32); base[:, 2] += 2.0
    feat = rng.uniform(0, 255, (300, 5)).astype(np.float32)
    out = []
    for k in range(n):
        sh = rng.normal(0, 0.004, 3).astype(np.float32)
        out.append((base + sh, feat))
    return out

if __name__ == "__main__":
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    orc = oracle.load(kd=False)
    orc.set_num_threads(1)
    clouds = make_clouds(4, 3)
    pairs = np.array([(i, (i + o) % 4) for i in range(4) for o in (1, 2)])
    idx = parallel.partition(len(pairs), rank, world)
    assert set(parallel.frames_needed(pairs, idx)) <= set(range(4))
    local = run_pairs(orc, clouds, pairs, idx)
    full = parallel.gather_results(local, idx, len(pairs))
    if rank == 0:
        ref = run_pairs(orc, clouds, pairs, np.arange(len(pairs)))
        assert np.array_equal(full["transform"], ref["transform"])
        assert np.array_equal(full["iterations"], ref["iterations"])
        np.save(sys.argv[2], full["iterations"])
    dist.barrier()
    dist.destroy_process_group()
'''


def test_partition_covers_everything_once():
    from cvo_slam_b200 import parallel
    for n, w in ((8192, 8), (10, 4), (3, 8), (0, 2)):
        parts = [parallel.partition(n, r, w) for r in range(w)]
        allp = np.sort(np.concatenate(parts)) if n else np.zeros(0, np.int64)
        assert np.array_equal(allp, np.arange(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_two_rank_gloo_sharding(tmp_path):
    import subprocess
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = tmp_path / "iters.npy"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), ROOT, str(out)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    it = np.load(out)
    assert len(it) == 8 and (it >= 1).all()
