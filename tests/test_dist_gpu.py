"""GPU tests of the multi-GPU layer: the peer-memory exchange and the native level
loop (gb200_dist_bfs) against the oracle.  World size 1 runs on any GPU box (the
owner stores into its own replica); the 2-rank test needs two GPUs."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle_binding as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(scale, dev):
    import graphblast_b200 as gb
    from graphblast_b200 import dist as gdist
    rp, ci = orc.rmat_csr(scale)
    n = len(rp) - 1
    rowptr = torch.from_numpy(rp).to(dev)
    colind = torch.from_numpy(ci).to(dev)
    rp_l, ci_l, colptr, rowind = gdist.local_slice(rowptr, colind, 0, n, n)
    desc = gb.Descriptor(mxvmode=0, struconly=1, opreuse=0, earlyexit=1)
    ops = gdist.GpuLocalOps(gb, n, 0, n, rp_l, ci_l, colptr, rowind, desc)
    comm = gdist.Comm([0, n], dev)
    return gb, gdist, rp, ci, n, ops, comm


@pytest.mark.gpu
@pytest.mark.parametrize("scale", [10, 14, 17])
def test_native_level_loop_world1(scale):
    dev = torch.device("cuda", 0)
    gb, gdist, rp, ci, n, ops, comm = _setup(scale, dev)
    x = gdist.PeerExchange(gb, comm, dev)
    try:
        deg = np.diff(rp)
        for source in (int(np.argmax(deg)), int(np.nonzero(deg)[0][-1])):
            want = orc.bfs(rp, ci, source)
            for _ in range(2):                      # state is reset per traversal
                levels = x.bfs(ops, n, source)
                got = ops.levels().astype(np.int32)
                assert np.array_equal(got, want)
                assert levels == int(want.max())
    finally:
        x.close()


@pytest.mark.gpu
def test_python_and_native_loops_agree():
    dev = torch.device("cuda", 0)
    gb, gdist, rp, ci, n, ops, comm = _setup(13, dev)
    source = int(np.argmax(np.diff(rp)))
    gdist.run_bfs(ops, comm, source)
    a = ops.levels().copy()
    x = gdist.PeerExchange(gb, comm, dev)
    try:
        x.bfs(ops, n, source)
        b = ops.levels().copy()
    finally:
        x.close()
    assert np.array_equal(a, b)
    assert np.array_equal(a.astype(np.int32), orc.bfs(rp, ci, source))


@pytest.mark.gpu
def test_two_ranks_peer_exchange():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, GB200_BENCH_SCALE="18")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "bench.py"),
           "--gpus", "2", "--steps", "3", "--warmup", "1"]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    res = json.loads(lines[0])
    assert res["n_gpus"] == 2
    assert res["parity_vs_cpu_reference"] is True
    assert "peer-memory" in res["config"]["exchange"]
