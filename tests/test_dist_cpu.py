"""world_size-2 CPU tests (gloo) of the multi-GPU host logic in
graphblast_b200/dist.py: nnz-balanced partition, slice extraction, the bitmap
exchange record format and the level loop's termination — with a host-side
stand-in for the per-rank GraphBLAS operations (the real ones need a GPU and are
covered by the gpu-marked tests and bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_binding as orc
from graphblast_b200 import dist as gdist


class CpuLocalOps(object):
    """Test double: the owned slice's share of one BFS level with numpy."""

    def __init__(self, n, lo, hi, rp_local, ci_local):
        self.n, self.lo, self.hi = n, lo, hi
        self.nl = hi - lo
        self.rp = rp_local.numpy()
        self.ci = ci_local.numpy()
        self.v = np.zeros(self.nl, dtype=np.float32)
        self.nwords = gdist.words_of(lo, hi)

    def reset(self):
        self.v[:] = 0

    @staticmethod
    def _unpack(words, count):
        bits = np.unpackbits(words.view(np.uint8), bitorder="little")
        return bits[:count].astype(bool)

    def assign_level(self, gbits, word_lo, level):
        words = gbits[word_lo:word_lo + self.nwords].numpy().astype(np.int32)
        own = self._unpack(words, self.nl)
        self.v[own] = level

    def expand(self, gbits, gcount, comm):
        frontier = self._unpack(gbits.numpy().astype(np.int32), self.n)
        out = np.zeros(self.nwords * 32, dtype=bool)
        for r in range(self.nl):
            if self.v[r] == 0:
                nbrs = self.ci[self.rp[r]:self.rp[r + 1]]
                if frontier[nbrs].any():
                    out[r] = True
        words = np.packbits(out, bitorder="little").view(np.int32)
        comm.set_host_record(torch.from_numpy(words.copy()), int(out.sum()))

    def levels(self):
        return self.v


def _worker(rank, world, port, scale, source, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rp, ci = orc.rmat_csr(scale)
    n = len(rp) - 1
    rowptr = torch.from_numpy(rp)
    colind = torch.from_numpy(ci)
    bounds = gdist.partition_bounds(rowptr, world, align=64)
    lo, hi = bounds[rank], bounds[rank + 1]
    rp_l, ci_l, colptr, rowind = gdist.local_slice(rowptr, colind, lo, hi, n)
    # CSC of the slice is the transpose of its CSR
    dense = np.zeros((hi - lo, n), dtype=bool)
    for r in range(hi - lo):
        dense[r, ci_l.numpy()[rp_l.numpy()[r]:rp_l.numpy()[r + 1]]] = True
    cp, ri = colptr.numpy(), rowind.numpy()
    for c in range(0, n, max(n // 50, 1)):
        assert sorted(np.nonzero(dense[:, c])[0].tolist()) == ri[cp[c]:cp[c + 1]].tolist()
    ops = CpuLocalOps(n, lo, hi, rp_l, ci_l)
    comm = gdist.Comm(bounds, "cpu")
    levels = gdist.run_bfs(ops, comm, source)
    levels2 = gdist.run_bfs(ops, comm, source)       # state is reset per run
    assert levels == levels2
    np.save(os.path.join(result_dir, "levels_%d.npy" % rank), ops.levels())
    np.save(os.path.join(result_dir, "bounds_%d.npy" % rank), np.array(bounds))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_bfs_matches_oracle(tmp_path, world):
    scale = 9
    rp, ci = orc.rmat_csr(scale)
    source = int(np.argmax(np.diff(rp)))
    mp.spawn(_worker, args=(world, _free_port(), scale, source, str(tmp_path)),
             nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "levels_%d.npy" % r))
             for r in range(world)]
    got = np.concatenate(parts).astype(np.int32)
    assert np.array_equal(got, orc.bfs(rp, ci, source))


def test_partition_bounds_properties():
    rp, ci = orc.rmat_csr(12)
    n = len(rp) - 1
    for world in (1, 2, 4, 8):
        b = gdist.partition_bounds(rp, world, align=32)
        assert b[0] == 0 and b[-1] == n and len(b) == world + 1
        assert all(x % 32 == 0 for x in b)
        assert all(b[i] <= b[i + 1] for i in range(world))
        if world > 1:
            share = np.diff(rp[np.array(b)]) / float(rp[-1])
            # RMAT rows are heavily skewed towards low ids; equal vertex ranges
            # would give rank 0 ~40% of the entries at world 8
            assert share.max() < 2.0 / world + 0.05


def test_exchange_record_roundtrip_single_rank():
    bounds = [0, 96]
    comm = gdist.Comm(bounds, "cpu")
    words = torch.tensor([5, -2 ** 31, 7], dtype=torch.int32)
    comm.set_host_record(words, 2 ** 33 + 5)
    g, total = comm.exchange()
    assert g[:3].tolist() == [5, -2 ** 31, 7]
    assert total == 2 ** 33 + 5


def test_local_slice_permutation_carries_values():
    """local_slice(return_order=True): val[order] are the CSC values of the owned
    rows, i.e. the (colptr, rowind, val[order]) triple is the transpose of
    (rp_local, ci_local, val) — what weighted_local_matrix hands to the library."""
    rp, ci = orc.rmat_csr(9)
    n = len(rp) - 1
    rowptr = torch.from_numpy(rp)
    colind = torch.from_numpy(ci)
    rng = np.random.RandomState(3)
    val = torch.from_numpy(rng.randint(1, 65, size=len(ci)).astype(np.float32))
    for lo, hi in ((0, n), (64, 320), (n - 128, n)):
        rp_l, ci_l, colptr, rowind, order = gdist.local_slice(
            rowptr, colind, lo, hi, n, return_order=True)
        e0, e1 = int(rp[lo]), int(rp[hi])
        val_l = val[e0:e1]
        cval = val_l[order]
        dense = np.zeros((hi - lo, n), dtype=np.float32)
        rp_n, ci_n = rp_l.numpy(), ci_l.numpy()
        for r in range(hi - lo):
            dense[r, ci_n[rp_n[r]:rp_n[r + 1]]] = val_l.numpy()[rp_n[r]:rp_n[r + 1]]
        cp, ri, cv = colptr.numpy(), rowind.numpy(), cval.numpy()
        assert cp[-1] == e1 - e0
        for c in range(n):
            rows = ri[cp[c]:cp[c + 1]]
            assert np.all(np.diff(rows) > 0)                 # sorted, unique
            assert np.array_equal(dense[rows, c], cv[cp[c]:cp[c + 1]])
        assert np.count_nonzero(dense) == e1 - e0
