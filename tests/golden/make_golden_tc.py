#!/usr/bin/env python
"""Generates tests/golden/tc_golden.json: triangle counts of tril(A) for the R-MAT
bench graphs (oracle generator, edge factor 16, seed 1, symmetrised, no self-loops /
duplicates), counted ONCE by the reference's own CPU code SimpleReferenceTc
(oracle/_ref/libgbref.so, built from /root/reference by oracle/Makefile).

The reference CPU count is sequential: scale 18 takes ~30 s, scale 20 minutes,
scale 22 about an hour on one core, which is why the bench and the GPU tests
compare against these committed numbers instead of recounting on every run.
Usage: make_golden_tc.py SCALE [SCALE ...]   (merges into the existing file)
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_binding as orc  # noqa: E402

PATH = os.environ.get("GOLDEN_OUT", os.path.join(HERE, "tc_golden.json"))


def main():
    assert orc.ref() is not None, "oracle/_ref/libgbref.so missing: make -C oracle ref"
    out = json.load(open(PATH)) if os.path.exists(PATH) else {}
    for scale in [int(a) for a in sys.argv[1:]]:
        rp, ci = orc.rmat_csr(scale)
        lr, lc = orc.tril(rp, ci)
        t0 = time.perf_counter()
        count = int(orc.ref_tc(lr, lc))
        dt = time.perf_counter() - t0
        out["rmat%d" % scale] = {
            "scale": scale, "edgefactor": 16, "seed": 1,
            "n": int(len(rp) - 1), "nnz": int(len(ci)), "nnz_tril": int(len(lc)),
            "colind_checksum": int(np.sum(ci.astype(np.int64) *
                                          (np.arange(len(ci), dtype=np.int64) % 97 + 1))),
            "triangles_tril": count,
            "counted_by": "SimpleReferenceTc (reference graphblas/algorithm/"
                          "test_tc.hpp:15-85) via oracle/_ref",
            "cpu_seconds": round(dt, 2),
        }
        with open(PATH, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)
        print("rmat%d: %d triangles in %.1f s" % (scale, count, dt), flush=True)


if __name__ == "__main__":
    main()
