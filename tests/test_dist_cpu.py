"""CPU tier, world_size 2 over gloo: the sharding contract of the multi-GPU path.

Reads are split into contiguous blocks; each rank computes the E-step and the per-locus M-step sums of ITS block
only, the K-length sums are all-reduced once per iteration, and every rank then holds identical pi/theta.  That is
exactly what libtelescope_b200 does over NCCL; here the per-rank arithmetic is the CPU oracle so the contract
(shard boundaries, which quantities are reduced, max_score and init totals being global) is tested without GPUs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_rows, n_cols, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle.em_numpy import EMOracle, q_lut
    from telescope_b200.synthetic import shard_bounds, synth_csr
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n_rows, world)[rank]
    ip, ix, raw = synth_csr(n_rows, n_cols, 8, True, 4242, lo, hi)
    ms = torch.tensor([int(raw.max())])
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)                       # global max score before the Q table
    o = EMOracle(ip, ix, raw, n_cols, max_score=int(ms.item()), lut=q_lut(int(ms.item())))

    def allsum(x):
        t = torch.from_numpy(np.atleast_1d(np.asarray(x, dtype=np.float64)).copy())
        dist.all_reduce(t)
        return t.numpy()
    total_wt, ambig_wt = allsum(o.total_wt)[0], allsum(o.ambig_wt)[0]
    wmax = torch.tensor([o.weights.max()], dtype=torch.float64)
    dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
    pisum0 = allsum(o.pisum0)
    tpw = 200000 * wmax.item()
    pi = theta = np.repeat(1.0 / n_cols, n_cols)
    for _ in range(6):
        z = o.estep(pi, theta)
        local = np.bincount(o.indices, weights=(z * o.weights[o.row]) * o.Y[o.row], minlength=n_cols)
        thetasum = allsum(local)                                    # the one collective per iteration
        theta = (thetasum + tpw) / (ambig_wt + tpw * n_cols)
        pi = (pisum0 + thetasum) / total_wt
    np.save(out_path % rank, np.stack([pi, theta]))
    dist.destroy_process_group()


def test_two_rank_sharded_em_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    from oracle.em_numpy import EMOracle
    sys.path.insert(0, ROOT)
    from telescope_b200.synthetic import synth_csr
    n_rows, n_cols = 30000, 400
    port = 29500 + (os.getpid() % 2000)
    out = str(tmp_path / "rank%d.npy")
    mp.spawn(_worker, args=(2, port, n_rows, n_cols, out), nprocs=2, join=True)
    a, b = np.load(out % 0), np.load(out % 1)
    assert np.array_equal(a, b), "all ranks must hold bit-identical parameters after the all-reduce"
    ip, ix, raw = synth_csr(n_rows, n_cols, 8, True, 4242)
    o = EMOracle(ip, ix, raw, n_cols, em_epsilon=-1, max_iter=6).em()
    assert np.max(np.abs(a[0] - o.pi) / np.maximum(o.pi, 1e-300)) < 1e-9
    assert np.max(np.abs(a[1] - o.theta) / o.theta) < 1e-9


def test_shard_bounds_cover_all_reads():
    sys.path.insert(0, ROOT)
    from telescope_b200.synthetic import shard_bounds
    for n, w in [(10, 1), (10, 3), (50_000_000, 8), (7, 8)]:
        b = shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))


_RDZV_CHILD = """
import sys, time
sys.path.insert(0, %r)
from telescope_b200 import dist
d = dist.rendezvous(timeout=60, transport="nccl")
print("RDZV %%d %%d %%s" %% (d.proc_rank, d.n_procs, d.nccl_id.hex()))
# a second rendezvous in the same launch must not see the first one's blobs
blobs = dist.allgather("ipc", bytes([d.proc_rank]) * 64, timeout=60)
again = dist.allgather("ipc", bytes([d.proc_rank + 10]) * 64, timeout=60)
print("GATHER %%d %%s %%s" %% (d.proc_rank, "".join("%%d" %% b[0] for b in blobs), ",".join("%%d" %% b[0] for b in again)))
dist.file_barrier("end")
dist.cleanup()
"""


def _run_children(extra_env=None):
    import subprocess
    procs = []
    for r in (1, 0):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_PORT="29%03d" % (os.getpid() % 1000))
        env.update(extra_env or {})
        procs.append(subprocess.Popen([sys.executable, "-c", _RDZV_CHILD % ROOT], env=env, stdout=subprocess.PIPE,
                                      universal_newlines=True, cwd=ROOT))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    return outs


def test_torch_free_rendezvous_hands_rank0s_nccl_id_to_every_rank():
    """Two children of this process (as under torchrun, ranks share a parent) agree on one 128-byte NCCL id, and the
    blob all-gather (CUDA IPC handles in the peer transport) returns every rank's blob in rank order, call by call."""
    outs = _run_children()
    got = sorted(l.split() for o in outs for l in o.splitlines() if l.startswith("RDZV"))
    assert [g[1] for g in got] == ["0", "1"] and got[0][2] == got[1][2] == "2"
    assert got[0][3] == got[1][3] and len(got[0][3]) == 256
    gat = sorted(l.split() for o in outs for l in o.splitlines() if l.startswith("GATHER"))
    assert [g[2] for g in gat] == ["01", "01"] and [g[3] for g in gat] == ["10,11", "10,11"]


def test_rendezvous_key_is_unique_per_restart_attempt():
    """A restarted worker group (TORCHELASTIC_RESTART_COUNT bumped) uses fresh files: ids never leak across attempts."""
    a = _run_children({"TORCHELASTIC_RESTART_COUNT": "0", "TORCHELASTIC_RUN_ID": "t"})
    b = _run_children({"TORCHELASTIC_RESTART_COUNT": "1", "TORCHELASTIC_RUN_ID": "t"})
    ida = {l.split()[3] for o in a for l in o.splitlines() if l.startswith("RDZV")}
    idb = {l.split()[3] for o in b for l in o.splitlines() if l.startswith("RDZV")}
    assert len(ida) == 1 and len(idb) == 1 and ida != idb


def test_multi_process_model_requires_the_global_max_score():
    import pytest
    import scipy.sparse as sp
    sys.path.insert(0, ROOT)
    from telescope_b200.likelihood import DistInfo, TelescopeLikelihood

    class O(object):
        em_epsilon, max_iter, pi_prior, theta_prior = 1e-7, 5, 0, 200000
    m = sp.csr_matrix(np.array([[3, 5], [0, 7]], dtype=np.uint16))
    with pytest.raises(ValueError, match="max_score"):
        TelescopeLikelihood(m, O, dist=DistInfo(2, 0, b"\0" * 128))
