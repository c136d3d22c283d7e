"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every
symbol include/graphblast_b200.h declares; the Python mirror's enums equal the
header's; compute entry points refuse to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import graphblast_b200 as gb
from graphblast_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "graphblast_b200.h")).read()


def declared_symbols():
    return sorted(set(re.findall(r"\b(gb200_[a-z0-9_]+)\s*\(", HEADER)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 55
    for name in names:
        assert hasattr(lib, name), "missing export: " + name


def test_python_binding_covers_every_declared_symbol():
    bound = {s[0] for s in _lib.SIGNATURES}
    assert bound == set(declared_symbols())


def test_semiring_and_monoid_order_matches_header():
    block = HEADER[HEADER.index("GB200_LOGICAL_OR_AND"):HEADER.index("GB200_NSEMIRINGS")]
    names = re.findall(r"GB200_([A-Z_]+)", block)
    want = [re.sub(r"(?<!^)(?=[A-Z])", "_", s.name).upper() for s in gb.Semiring]
    assert names == want
    block = HEADER[HEADER.index("GB200_PLUS_MONOID"):HEADER.index("GB200_NMONOIDS")]
    names = [n[:-len("_MONOID")] for n in re.findall(r"GB200_([A-Z_]+)", block)]
    want = [re.sub(r"(?<!^)(?=[A-Z])", "_", m.name).upper() for m in gb.Monoid]
    assert names == want


def test_descriptor_enum_values_match_reference_layout():
    # GrB_SCMP=0, GrB_REPLACE=1, GrB_TRAN=2 is what Descriptor::toggle relies on
    assert int(gb.Desc_value.GrB_SCMP) == 0
    assert int(gb.Desc_value.GrB_REPLACE) == 1
    assert int(gb.Desc_value.GrB_TRAN) == 2
    assert int(gb.Desc_value.GrB_DEFAULT) == 3
    assert [int(v) for v in (gb.Desc_value.GrB_PUSHPULL,
                             gb.Desc_value.GrB_PUSHONLY,
                             gb.Desc_value.GrB_PULLONLY)] == [10, 11, 12]
    assert int(gb.Info.GrB_NOT_IMPLEMENTED) == 9 and int(gb.Info.GrB_PANIC) == 14


def test_descriptor_host_logic_without_device():
    """Descriptor set/get/toggle and the flag defaults are host-only."""
    d = gb.Descriptor()
    assert d.get(gb.Desc_field.GrB_MASK) == gb.Desc_value.GrB_DEFAULT
    d.toggle(gb.Desc_field.GrB_MASK)
    assert d.get(gb.Desc_field.GrB_MASK) == gb.Desc_value.GrB_SCMP
    d.toggle(gb.Desc_field.GrB_MASK)
    assert d.get(gb.Desc_field.GrB_MASK) == gb.Desc_value.GrB_DEFAULT
    d.toggle(gb.Desc_field.GrB_INP1)
    assert d.get(gb.Desc_field.GrB_INP1) == gb.Desc_value.GrB_TRAN
    d.toggle(gb.Desc_field.GrB_OUTP)
    assert d.get(gb.Desc_field.GrB_OUTP) == gb.Desc_value.GrB_REPLACE
    # parseArgs defaults (reference util.hpp:39-132)
    assert d.get_knob("mxvmode") == 1
    assert d.get(gb.Desc_field.GrB_MXVMODE) == gb.Desc_value.GrB_PUSHONLY
    assert d.get_knob("switchpoint") == pytest.approx(0.01)
    assert d.get_knob("earlyexit") == 1 and d.get_knob("fusedmask") == 1
    assert d.get_knob("sort") == 1 and d.get_knob("struconly") == 0
    assert d.get_knob("max_niter") == 10000
    d.set_knob("mxvmode", 0)
    assert d.get(gb.Desc_field.GrB_MXVMODE) == gb.Desc_value.GrB_PUSHPULL
    with pytest.raises(gb.GraphBLASError):
        d.set_knob("mxvmode", 7)
    with pytest.raises(gb.GraphBLASError):
        d.set_knob("no_such_knob", 1)
    # every numeric command-line knob goes through the same enumeration
    # (backend descriptor.hpp: eachKnob), whatever its type
    for name, value in [("niter", 3), ("timing", 2), ("directed", 1), ("ta", 8),
                        ("memusage", 0.5), ("switchpoint", 0.25), ("opreuse", 1),
                        ("dirinfo", 1), ("atomic", 1), ("debug", 0), ("ndevice", 1)]:
        d.set_knob(name, value)
        assert d.get_knob(name) == pytest.approx(value), name
    assert d.get_knob("opreuse") == 1 and d.get_knob("struconly") == 0
    d.set_knob("nthread", 256)                       # GrB_NT follows a valid thread count
    assert int(d.get(gb.Desc_field.GrB_NT)) == 256         # Desc_value GrB_256
    assert d.get_knob("mxvmode") == 0                # the rejected value left no trace
    with pytest.raises(gb.GraphBLASError):
        d.set_knob("mode", 1)                        # a string knob is not settable by number


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(gb.GraphBLASError) as e:
        gb.Vector(8)
    assert e.value.info == gb.Info.GrB_PANIC
    with pytest.raises(gb.GraphBLASError):
        gb.Matrix(4, 4)
    with pytest.raises(gb.GraphBLASError):
        gb.init(0)


def test_missing_extension_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.ExtensionMissing):
        _lib.load()


def test_header_is_plain_c():
    """include/graphblast_b200.h must compile as C99 (it is what cgo / JNI / Rust
    FFI bindings of this path would include)."""
    import subprocess
    import tempfile
    src = '#include "graphblast_b200.h"\nint main(void) { return 0; }\n'
    with tempfile.NamedTemporaryFile("w", suffix=".c", delete=False) as f:
        f.write(src)
        path = f.name
    try:
        out = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic",
                              "-Werror", "-I", os.path.join(ROOT, "include"),
                              "-fsyntax-only", path],
                             capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
    finally:
        os.unlink(path)


def test_c_example_links_and_fails_loudly_without_a_device():
    """examples/capi_bfs.c builds with gcc alone against the shared library; on a
    box without a GPU it must refuse to run (no CPU fallback), not compute."""
    import subprocess
    import torch
    out = os.path.join(ROOT, "build", "capi_bfs_test")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    lib_dir = os.path.join(ROOT, "graphblast_b200", "lib")
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I",
                         os.path.join(ROOT, "include"),
                         os.path.join(ROOT, "examples", "capi_bfs.c"), "-L", lib_dir,
                         "-lgraphblast_b200", "-Wl,-rpath," + lib_dir, "-o", out],
                        capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    if torch.cuda.is_available():
        return                       # the GPU run is tests/test_zz_c_example_gpu.py's job
    run = subprocess.run([out, os.path.join(ROOT, "tests", "golden", "chesapeake.mtx")],
                         capture_output=True, text=True)
    assert run.returncode != 0
    assert "no CUDA device" in run.stderr
