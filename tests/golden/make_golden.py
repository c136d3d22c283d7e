#!/usr/bin/env python
"""Generates tests/golden/golden.json from the REFERENCE's own CPU code
(oracle/_ref/libgbref.so, compiled from /root/reference by oracle/Makefile).
Run in a container that mounts /root/reference; the JSON is committed because the
GPU box has neither the reference nor (necessarily) _ref.

Contents (all on the reference loader's CSR of the bundled graphs, --directed 2):
  chesapeake : BFS levels from 0, SSSP distances from 0 with the reference's
               weight stream (seed 1, uniform_int[1,64]), PageRank (alpha .85,
               10 iterations), triangle count of tril
  test_cc    : row sums (reference test/greduce.cu:65), BFS levels from 0
  weights    : first 64 draws of the weight stream for seeds 1 and 7
  rmat10     : BFS level histogram / SSSP checksum / triangle count of an R-MAT
               scale-10 graph from the oracle generator (graph rebuilt in the test)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_binding as orc  # noqa: E402


def main():
    assert orc.ref() is not None, "oracle/_ref/libgbref.so missing: make -C oracle ref"
    out = {}
    path = os.path.join(HERE, "chesapeake.mtx")
    rp, ci, _ = orc.ref_load_mtx(path, 2)
    w = orc.ref_uniform_weights(1, 1, 64, len(ci))
    lr, lc = orc.tril(rp, ci)
    out["chesapeake"] = {
        "n": int(len(rp) - 1), "nnz": int(len(ci)),
        "rowptr": rp.tolist(), "colind": ci.tolist(),
        "bfs_levels_src0": orc.ref_bfs(rp, ci, 0).tolist(),
        "sssp_weights_seed1": w.tolist(),
        "sssp_dist_src0": orc.ref_sssp(rp, ci, w, 0).tolist(),
        "pagerank_a085_it10": [float(x) for x in orc.ref_pr(rp, ci, 0.85, 1e-8, 10)],
        "triangles_tril": int(orc.ref_tc(lr, lc)),
    }
    path = os.path.join(HERE, "test_cc.mtx")
    rp, ci, val = orc.ref_load_mtx(path, 0)
    out["test_cc"] = {
        "n": int(len(rp) - 1), "nnz": int(len(ci)),
        "rowptr": rp.tolist(), "colind": ci.tolist(),
        "row_sums": orc.reduce_rows(rp, val).tolist(),
        "bfs_levels_src0": orc.ref_bfs(rp, ci, 0).tolist(),
    }
    out["weights"] = {
        "seed1_first64": orc.ref_uniform_weights(1, 1, 64, 64).tolist(),
        "seed7_first64": orc.ref_uniform_weights(7, 1, 64, 64).tolist(),
    }
    rp, ci = orc.rmat_csr(10)
    w = orc.ref_uniform_weights(1, 1, 64, len(ci))
    lv = orc.ref_bfs(rp, ci, 0)
    d = orc.ref_sssp(rp, ci, w, 0)
    lr, lc = orc.tril(rp, ci)
    finite = d[d < orc.FLT_MAX]
    out["rmat10"] = {
        "n": int(len(rp) - 1), "nnz": int(len(ci)),
        "colind_checksum": int(np.sum(ci.astype(np.int64) * (np.arange(len(ci)) % 97 + 1))),
        "bfs_level_hist_src0": np.bincount(lv).tolist(),
        "sssp_sum_finite_src0": float(finite.astype(np.float64).sum()),
        "sssp_reached_src0": int(len(finite)),
        "triangles_tril": int(orc.ref_tc(lr, lc)),
    }
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote golden.json")


if __name__ == "__main__":
    main()
