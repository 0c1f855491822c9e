"""The C++ shim (include/fuelgpu_shim.hpp: SDFMap / EDTEnvironment / FrontierFinder /
BsplineOptimizer with the reference's names) compiles against the C ABI and links libfuelgpu.so.
On a box without a GPU the program must stop in initMap with FUELGPU_ENODEVICE (no fallback);
on the GPU box its results are compared with the oracle (tests/test_gpu_shim below)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_shim(tmp_path):
    from fuel_b200 import _lib
    _lib.lib()
    exe = str(tmp_path / "shim_smoke")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "shim_smoke.cpp"), "-o", exe,
                           "-L", os.path.join(ROOT, "fuel_b200"), "-lfuelgpu",
                           "-Wl,-rpath," + os.path.join(ROOT, "fuel_b200")])
    return exe


def test_shim_compiles_and_refuses_without_gpu(tmp_path):
    import torch
    exe = build_shim(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_shim_matches_oracle")
    r = subprocess.run([exe, str(tmp_path / "out.txt")], capture_output=True, text=True)
    assert r.returncode == 42, r.stdout + r.stderr
    assert "no CPU fallback" in r.stdout


def scene():
    n = (48, 40, 24)
    x, y, z = np.meshgrid(np.arange(48), np.arange(40), np.arange(24), indexing="ij")
    known = (x >= 4) & (x < 44) & (y >= 4) & (y < 36) & (z >= 2) & (z < 22)
    ball = (x - 24) ** 2 + (y - 20) ** 2 + 2 * (z - 12) ** 2 < 81
    wall = (x >= 12) & (x <= 13) & (y >= 8) & (y < 30) & (z < 18)
    tri = np.zeros(n, dtype=np.uint8)
    tri[known & ~ball] = 1
    tri[known & ~ball & wall] = 2
    inflate = (known & ~ball & wall).astype(np.int8)
    return n, tri, inflate


@pytest.mark.gpu
def test_shim_matches_oracle(tmp_path, orc):
    exe = build_shim(tmp_path)
    out = tmp_path / "out.txt"
    r = subprocess.run([exe, str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = open(out).read().strip().split("\n")
    n, tri, inflate = scene()
    g = orc.make_grid(n, 0.1, (-2.4, -2.0, -0.5), (-2.2, -1.8, -0.3), (2.2, 1.8, 1.7))
    d = orc.update_esdf3d(g, inflate, tri, [0, 0, 0], np.array(n) - 1, True, False)
    got = [float(v) for v in lines[0].split()[1:]]
    probes = [(5, 5, 3), (20, 20, 5), (30, 10, 15), (13, 15, 10), (40, 30, 20), (0, 0, 0)]
    for v, p in zip(got, probes):
        assert abs(v - d[p]) <= 1e-4 * abs(d[p]) + 1e-12
    s = [float(v) for v in lines[1].split()[1:]]
    dd, gg = orc.dist_with_grad(g, d, np.array([[0.513, -0.377, 0.642]]))
    assert np.allclose(s, [dd[0], *gg[0]], rtol=1e-4, atol=1e-6)
    fl = np.zeros(n, dtype=np.int8)
    ref = orc.frontier_search(g, tri, fl, (-2.4, -2.0, -0.5), (2.4, 2.0, 1.9),
                              orc.frontier_params(cluster_min=20, cluster_size_xy=1.0, cell_order=1))
    assert int(lines[2].split()[1]) == len(ref) and len(ref) >= 2
    for ln, c in zip(lines[3:3 + len(ref)], ref):
        t = ln.split()
        assert int(t[1]) == len(c["addr"]) and int(t[2]) == len(c["filtered"])
        assert np.allclose([float(v) for v in t[3:6]], c["average"], rtol=1e-12)
        h = 0
        for a in c["addr"]:
            h = (h * 1000003 + int(a)) % 2147483647
        assert int(t[6]) == h
    # cost at the initial point and the solver result
    pts = np.array([[-1.9 + 0.3 * i, -1.2 + 0.18 * i + ((i % 3) - 1) * 0.1, 0.6 + 0.03 * i] for i in range(12)])
    start = np.array([(pts[0] + 4 * pts[1] + pts[2]) / 6, [1.0, 0.6, 0.1], [0, 0, 0]])
    tcs = orc.traj_consts(1)
    orc.fill_traj_const(tcs[0], orc.pt_dist(pts), 0.2, start, np.array([[1.4, 0.8, 0.93]]))
    mask = orc.NORMAL_PHASE | orc.MINTIME
    x = np.concatenate([pts.reshape(-1), [0.2]])[None, :]
    f0, g0 = orc.combine_cost_batch(g, d, orc.opt_params(), tcs, 12, mask, x)
    t = lines[3 + len(ref)].split()
    assert abs(float(t[1]) - f0[0]) <= 1e-4 * abs(f0[0])
    assert np.allclose([float(t[3]), float(t[4]), float(t[5]), float(t[7])], [g0[0, 0], g0[0, 16], g0[0, 35], g0[0, 36]],
                       rtol=1e-4, atol=1e-4 * np.abs(g0).max())
    xb, fb, ne = orc.optimize_batch(g, d, orc.opt_params(), tcs, 12, mask, x, max_eval=40)
    t = lines[4 + len(ref)].split()
    assert int(t[1]) <= 40 and float(t[2]) <= f0[0]
    assert abs(float(t[2]) - fb[0]) <= 0.05 * abs(fb[0])
