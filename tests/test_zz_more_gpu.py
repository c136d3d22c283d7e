"""GPU tests written after the round's GPU budget was spent (so they have not run on a
device yet); kept in a file that sorts last so that, should one of them uncover a bug,
`pytest -x` has already gone through the rest of the suite."""
import numpy as np
import pytest

import oracle_binding as orc
from test_parity_gpu import make_matrix, ragged_graph

# Non-strict xfail until their first run on a device has been looked at: a pass is
# reported as XPASS, a failure as XFAIL, and the suite's verdict stays what the
# verified tests say.  (r02: drop the mark once these show XPASS.)
pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False,
                                reason="written after the r01 GPU budget was spent; "
                                       "never executed on a device yet")]


@pytest.fixture(scope="module")
def gb():
    import graphblast_b200 as g
    g.init(0)
    return g


@pytest.mark.parametrize("name", ["PlusMultiplies", "MinimumPlus", "MaximumMultiplies"])
def test_pull_accum_combines_with_the_semiring_add(gb, name):
    """reference spmv.hpp:213-219: with an accumulator the pull result is combined
    into the old w with the SEMIRING's add (the accum functor itself is ignored)."""
    rp, ci = ragged_graph()
    n = len(rp) - 1
    rng = np.random.RandomState(5)
    val = (2.0 ** rng.randint(0, 3, len(ci))).astype(np.float32)
    A = make_matrix(gb, rp, ci, val, symmetric=False)
    sem = getattr(gb.Semiring, name)
    u_h = (2.0 ** rng.randint(0, 2, n)).astype(np.float32)
    w_old = (2.0 ** rng.randint(0, 4, n)).astype(np.float32)
    desc = gb.Descriptor(mxvmode=2)
    u = gb.Vector(n)
    u.build(u_h)
    w = gb.Vector(n)
    w.build(w_old)
    gb.vxm(w, None, "accum", sem, u, A, desc)
    res, _ = orc.vxm(int(sem), rp, ci, val, u_h)
    add = {"PlusMultiplies": np.add, "MinimumPlus": np.minimum,
           "MaximumMultiplies": np.maximum}[name]
    assert np.array_equal(w.extractTuples(), add(w_old, res))


def test_edge_share_check_does_not_change_results(gb):
    """GrB_PUSHPULL may hand a push back to the pull direction when the frontier
    owns more than a third of the stored entries (spmspv.hpp); a star graph whose
    hub is in a 4096+-entry frontier forces exactly that, and the result must equal
    the oracle's (and the forced-push result)."""
    n = 20000
    hub = 0
    src = np.zeros(n - 1, dtype=np.int32) + hub
    dst = np.arange(1, n, dtype=np.int32)
    rp, ci = orc.build_csr(n, src, dst, True)
    A = make_matrix(gb, rp, ci)
    sem = gb.Semiring.LogicalOrAnd
    f_ind = np.arange(0, 5000, dtype=np.int32)            # contains the hub
    f_val = np.ones(len(f_ind), dtype=np.float32)
    out = {}
    for mode in (0, 1):                                   # automatic, forced push
        desc = gb.Descriptor(mxvmode=mode, switchpoint=0.9)
        u = gb.Vector(n)
        u.build(f_ind, f_val)
        w = gb.Vector(n)
        gb.vxm(w, None, None, sem, u, A, desc)
        out[mode] = (w.extractTuples() != 0)
        if mode == 0:
            assert desc.lastmxv == gb.Desc_value.GrB_PULLONLY
        else:
            assert desc.lastmxv == gb.Desc_value.GrB_PUSHONLY
    up = np.zeros(n, np.uint8)
    uu = np.zeros(n, np.float32)
    up[f_ind] = 1
    uu[f_ind] = 1
    val = np.ones(len(ci), dtype=np.float32)
    want, wp = orc.vxm(int(sem), rp, ci, val, uu, u_present=up)
    assert np.array_equal(out[0], want != 0)
    assert np.array_equal(out[1], want != 0)


@pytest.mark.parametrize("name", ["PlusMultiplies", "MinimumPlus"])
def test_column_relabelling_is_bit_identical(gb, name):
    """GB200_SPMV_RELABEL=1 (experimental, off by default): the generic pull SpMV
    gathers through relabelled column indices from a permuted copy of u; products
    and their order are unchanged, so the result must not differ in a single bit
    (random float values on purpose)."""
    import os
    rp, ci = orc.rmat_csr(12)
    n = len(rp) - 1
    rng = np.random.RandomState(9)
    val = rng.rand(len(ci)).astype(np.float32) + 0.5
    u_h = rng.rand(n).astype(np.float32) + 0.5
    sem = getattr(gb.Semiring, name)
    out = {}
    for flag in ("0", "1", "1"):                   # second "1" reuses the cached copy
        os.environ["GB200_SPMV_RELABEL"] = flag
        try:
            A = make_matrix(gb, rp, ci, val, symmetric=False)
            desc = gb.Descriptor(mxvmode=2)
            u = gb.Vector(n)
            u.build(u_h)
            w = gb.Vector(n)
            gb.vxm(w, None, None, sem, u, A, desc)
            gb.vxm(w, None, None, sem, u, A, desc)
            out.setdefault(flag, []).append(w.extractTuples().copy())
        finally:
            os.environ["GB200_SPMV_RELABEL"] = "0"
    assert np.array_equal(out["0"][0].view(np.uint32), out["1"][0].view(np.uint32))
    assert np.array_equal(out["0"][0].view(np.uint32), out["1"][1].view(np.uint32))
