"""The C-ABI library loads and exports every symbol include/fuelgpu.h declares; the ctypes
mirrors of its structs have the C compiler's layout; without a GPU the product fails loudly
(no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fuelgpu.h")).read()
    return sorted(set(re.findall(r"FUELGPU_API\s+[\w\s\*]+?\b(fuelgpu_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(fuel):
    from fuel_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 25
    L = C.CDLL(_lib.SO)
    for n in names:
        assert hasattr(L, n), "libfuelgpu.so does not export %s" % n
        assert n in _lib.SIGNATURES, "python binding misses %s" % n
    assert sorted(_lib.SIGNATURES) == names
    assert b"sm_100a" in _lib.lib().fuelgpu_version()


def test_struct_layouts_match_the_c_compiler(tmp_path):
    from fuel_b200 import _lib
    import oracle
    prog = tmp_path / "layout.c"
    prog.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "fuelgpu.h"
#include "fuel_oracle.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\\n", sizeof(FuelGridDesc), sizeof(FuelFrontierParams), sizeof(FuelOptParams),
         sizeof(FuelTrajConst), sizeof(FuelSolveParams));
  printf("%zu %zu %zu %zu\\n", offsetof(FuelTrajConst, n_end), offsetof(FuelTrajConst, guide),
         offsetof(FuelTrajConst, waypt_idx), offsetof(FuelOptParams, order));
  printf("%zu %zu %zu %zu %zu\\n", sizeof(OrcGrid), sizeof(OrcFrontierParams), sizeof(OrcOptParams),
         sizeof(OrcTrajConst), sizeof(OrcSolveParams));
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"),
                           str(prog), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    got = [int(v) for v in out]
    T = _lib.FuelTrajConst
    exp = [C.sizeof(_lib.FuelGridDesc), C.sizeof(_lib.FuelFrontierParams), C.sizeof(_lib.FuelOptParams),
           C.sizeof(T), C.sizeof(_lib.FuelSolveParams), T.n_end.offset, T.guide.offset, T.waypt_idx.offset,
           _lib.FuelOptParams.order.offset, C.sizeof(oracle.OrcGrid), C.sizeof(oracle.OrcFrontierParams),
           C.sizeof(oracle.OrcOptParams), C.sizeof(oracle.OrcTrajConst), C.sizeof(oracle.OrcSolveParams)]
    assert got == exp


def test_no_cpu_fallback(fuel):
    """On a box without an sm_100 device every entry point refuses to run."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fuel.FuelGpuError) as e:
        fuel.SDFMap((8, 8, 8), 0.1, (0, 0, 0))
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    """fuel_b200/ must never import, link or call oracle/ (the parity checker)."""
    pkg = os.path.join(ROOT, "fuel_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "fuel_oracle" not in src, f
    code = "import sys; import fuel_b200; import fuel_b200.workloads; assert 'oracle' not in sys.modules"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)


def test_workloads_are_deterministic():
    import numpy as np
    from fuel_b200 import workloads as W
    g, inflate = W.office_map()
    assert g.n == (200, 120, 40) and int(inflate.sum()) == 57661
    t1 = W.office_known(g, inflate)
    t2 = W.office_known(g, inflate)
    assert np.array_equal(t1, t2) and set(np.unique(t1)) == {0, 1, 2}
    a = W.make_trajectories(g, inflate, B=8)
    b = W.make_trajectories(g, inflate, B=8)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert np.all(a["ctrl"] >= g.box_min + 0.1 - 1e-12) and np.all(a["ctrl"] <= g.box_max - 0.1 + 1e-12)
