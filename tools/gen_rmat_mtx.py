#!/usr/bin/env python
"""Write a small R-MAT graph as a Matrix Market pattern file (numpy; smoke use only).

The product generator is the device kernel behind gb200_rmat_edges (include/graphblast_b200.h);
this script only feeds the unchanged reference drivers (build/dropin/g*) a mid-size .mtx.
"""
import sys
import numpy as np


def rmat_edges(scale, edgefactor, seed=1, a=0.57, b=0.19, c=0.19):
    rng = np.random.default_rng(seed)
    m = edgefactor << scale
    src = np.zeros(m, dtype=np.int64)
    dst = np.zeros(m, dtype=np.int64)
    for _ in range(scale):
        r = rng.random(m)
        src_bit = r >= (a + b)
        dst_bit = ((r >= a) & (r < a + b)) | (r >= a + b + c)
        src = (src << 1) | src_bit
        dst = (dst << 1) | dst_bit
    return src, dst


def main():
    scale = int(sys.argv[1])
    ef = int(sys.argv[2])
    out = sys.argv[3]
    src, dst = rmat_edges(scale, ef)
    n = 1 << scale
    with open(out, "w") as f:
        f.write("%%MatrixMarket matrix coordinate pattern general\n")
        f.write("%d %d %d\n" % (n, n, len(src)))
        np.savetxt(f, np.stack([src + 1, dst + 1], axis=1), fmt="%d")


if __name__ == "__main__":
    main()
