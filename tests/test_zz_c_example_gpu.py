"""examples/capi_bfs.c on a GPU: plain C through the C ABI (the binary is built by
__graft_entry__.build(); its one manual B200 run in r01 printed depth 3 on the bundled
graph, which is what this test asserts against the oracle)."""
import os

import pytest


@pytest.mark.gpu
def test_c_example_runs_bfs_through_the_c_abi():
    """examples/capi_bfs.c (plain C, built by __graft_entry__.build) on the bundled
    graph: depth must equal the oracle's."""
    import subprocess
    import oracle_binding as orc
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "build", "capi_bfs")
    if not os.path.exists(exe):
        pytest.skip("build/capi_bfs not built")
    mtx = os.path.join(root, "tests", "golden", "chesapeake.mtx")
    run = subprocess.run([exe, mtx, "0"], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stderr
    import graphblast_b200 as gb
    gb.init(0)
    A = gb.Matrix.from_mtx(mtx, directed=2)
    rp, ci, _ = A.extract_csr()
    depth = int(orc.bfs(rp, ci, 0).max())
    assert ("depth = %d," % depth) in run.stdout, run.stdout
