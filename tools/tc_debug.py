"""Per-entry comparison of the hash and the search formulation of the masked mxm."""
import os, sys, subprocess, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def run(scale, hash_on):
    os.environ["GB200_SPGEMM_HASH"] = str(hash_on)
    import graphblast_b200 as gb
    from graphblast_b200 import algorithm, graphs
    n = 1 << scale
    src, dst = graphs.rmat_edges(scale, 16, seed=1)
    rowptr, colind = graphs.build_csr(n, src, dst, undirected=True)
    desc = gb.Descriptor(mxvmode=0)
    A = graphs.matrix_from_csr(n, rowptr, colind, dtype=gb.api.INT32, symmetric=True)
    A.tril(desc)
    B = gb.Matrix(n, n, dtype=gb.api.INT32)
    ntris, ms = algorithm.tc(A, B, desc)
    rp, ci, val = B.extract_csr()
    lrp, lci, lval = A.extract_csr()
    np.savez("/tmp/tc_dbg_%d_%d.npz" % (scale, hash_on), rp=rp, ci=ci, val=val, lrp=lrp, lci=lci)
    print("scale", scale, "hash", hash_on, "ntris", int(ntris), "ms", ms)

if __name__ == "__main__":
    if len(sys.argv) > 2:
        run(int(sys.argv[1]), int(sys.argv[2])); sys.exit(0)
    for scale in (10, 14):
        for h in (0, 1):
            subprocess.run([sys.executable, __file__, str(scale), str(h)], check=False)
        a = np.load("/tmp/tc_dbg_%d_0.npz" % scale); b = np.load("/tmp/tc_dbg_%d_1.npz" % scale)
        assert np.array_equal(a["ci"], b["ci"]) and np.array_equal(a["rp"], b["rp"])
        bad = np.nonzero(a["val"] != b["val"])[0]
        print("scale", scale, "entries", len(a["val"]), "mismatching", len(bad))
        if len(bad):
            rp = a["lrp"].astype(np.int64); ln = np.diff(rp)
            rows = np.repeat(np.arange(len(ln)), np.diff(a["rp"].astype(np.int64)))
            for e in bad[:20]:
                i = rows[e]; j = a["ci"][e]
                print("  entry", e, "row", i, "col", j, "len_i", ln[i], "len_j", ln[j], "search", a["val"][e], "hash", b["val"][e])
            li = ln[rows[bad]]; lj = ln[a["ci"][bad]]
            print("  owner=row (len_j<=len_i):", int((lj <= li).sum()), " owner=col:", int((lj > li).sum()))
            print("  max len hist of bad:", np.histogram(np.maximum(li, lj), bins=[0,1,2,33,129,2049,12289,1<<30])[0])
