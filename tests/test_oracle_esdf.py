"""Pins the ESDF oracle (oracle/fuel_oracle.c, restating plan_env/src/sdf_map.cpp:116-241 and
:497-536) against independent ground truths: brute-force nearest-site distance, scipy's
exact EDT, a numpy trilinear interpolator.  The reference ships no golden vectors for this
path, so these are what the oracle is pinned on ("parity unpinned")."""
import numpy as np
import pytest
from scipy import ndimage

from fuel_b200 import workloads as W
from tests.helpers import orc_grid, random_scene

SENT = 1e150


def brute(site, res, bmin, bmax):
    """distance to the nearest site INSIDE the box, for voxels of the box."""
    sl = tuple(slice(bmin[i], bmax[i] + 1) for i in range(3))
    sub = site[sl]
    pts = np.argwhere(sub)
    out = np.full(sub.shape, np.inf)
    if len(pts):
        idx = np.argwhere(np.ones(sub.shape, dtype=bool))
        d2 = ((idx[:, None, :] - pts[None, :, :]) ** 2).sum(-1).min(1)
        out = (res * np.sqrt(d2.astype(np.float64))).reshape(sub.shape)
    return out


@pytest.mark.parametrize("n,seed", [((9, 7, 11), 0), ((14, 13, 6), 1), ((5, 1, 9), 2), ((1, 1, 1), 3)])
@pytest.mark.parametrize("optimistic", [True, False])
def test_matches_bruteforce(orc, n, seed, optimistic):
    g = W.Grid(n, (0.3, -0.2, 0.1), 0.1)
    inflate, tri = random_scene(n, seed, p_site=0.03, p_unknown=0.4, blobs=2)
    rng = np.random.default_rng(seed)
    for _ in range(3):
        bmin = np.array([rng.integers(0, k) for k in n])
        bmax = np.array([rng.integers(bmin[i], n[i]) for i in range(3)])
        d = orc.update_esdf3d(orc_grid(orc, g), inflate, tri, bmin, bmax, optimistic, False)
        site = (inflate == 1) if optimistic else ((inflate == 1) | (tri == W.UNKNOWN))
        ref = brute(site, g.res, bmin, bmax)
        got = d[tuple(slice(bmin[i], bmax[i] + 1) for i in range(3))]
        fin = np.isfinite(ref)
        assert np.all(got[~fin] > SENT)  # no site in the box: resolution*sqrt(DBL_MAX)
        assert np.allclose(got[fin], ref[fin], rtol=1e-14, atol=0)
        outside = np.ones(n, dtype=bool)
        outside[tuple(slice(bmin[i], bmax[i] + 1) for i in range(3))] = False
        assert np.all(d[outside] == 0.0)  # untouched (default_dist 0.0)


def test_office_matches_scipy_and_threads(orc):
    g, inflate = W.office_map()
    tri = W.office_known(g, inflate)
    og = orc_grid(orc, g)
    full = ([0, 0, 0], np.array(g.n) - 1)
    d1 = orc.update_esdf3d(og, inflate, tri, *full, True, False, threads=1)
    d8 = orc.update_esdf3d(og, inflate, tri, *full, True, False, threads=8)
    assert np.array_equal(d1, d8)
    sc = ndimage.distance_transform_edt(inflate == 0, sampling=g.res)
    assert np.allclose(d1, sc, rtol=1e-13, atol=1e-13)
    dn = orc.update_esdf3d(og, inflate, tri, *full, False, False, threads=8)
    sc = ndimage.distance_transform_edt(~((inflate == 1) | (tri == W.UNKNOWN)), sampling=g.res)
    assert np.allclose(dn, sc, rtol=1e-13, atol=1e-13)


def test_signed_distance(orc):
    n = (24, 20, 16)
    g = W.Grid(n, (0, 0, 0), 0.1)
    inflate = np.zeros(n, dtype=np.int8)
    inflate[8:15, 6:13, 4:11] = 1
    tri = np.full(n, W.FREE, dtype=np.uint8)
    d = orc.update_esdf3d(orc_grid(orc, g), inflate, tri, [0, 0, 0], np.array(n) - 1, True, True)
    pos = ndimage.distance_transform_edt(inflate == 0, sampling=g.res)
    neg = ndimage.distance_transform_edt(inflate == 1, sampling=g.res)
    ref = pos.copy()
    ref[neg > 0] += -neg[neg > 0] + g.res  # sdf_map.cpp:232-239
    assert np.allclose(d, ref, rtol=1e-13, atol=1e-13)
    assert d[11, 9, 7] < 0 and d[0, 0, 0] > 0


def trilinear_numpy(g, dist, pos):
    """independent restatement of getDistWithGrad's value for interior points"""
    res = g.res
    pm = pos - 0.5 * res
    idx = np.floor((pm - g.origin) / res).astype(int)
    ip = (idx + 0.5) * res + g.origin
    diff = (pos - ip) / res
    v = 0.0
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                w = ((diff[0] if dx else 1 - diff[0]) * (diff[1] if dy else 1 - diff[1]) *
                     (diff[2] if dz else 1 - diff[2]))
                v += w * dist[idx[0] + dx, idx[1] + dy, idx[2] + dz]
    return v


def test_dist_with_grad(orc):
    n = (20, 18, 16)
    g = W.Grid(n, (-1.0, -0.9, -0.3), 0.1)
    rng = np.random.default_rng(3)
    dist = rng.uniform(0, 3, size=n)
    og = orc_grid(orc, g)
    pos = rng.uniform(g.origin + 0.2, g.map_max - 0.2, size=(200, 3))
    d, gr = orc.dist_with_grad(og, dist, pos)
    for i in range(50):
        assert abs(d[i] - trilinear_numpy(g, dist, pos[i])) < 1e-12
    # gradient = derivative of the trilinear form (exact within a cell)
    eps = 1e-6
    for k in range(3):
        p2 = pos.copy()
        p2[:, k] += eps
        d2, _ = orc.dist_with_grad(og, dist, p2)
        same_cell = np.floor((p2 - 0.5 * g.res - g.origin) / g.res)[:, k] == np.floor(
            (pos - 0.5 * g.res - g.origin) / g.res)[:, k]
        assert np.allclose(((d2 - d) / eps)[same_cell], gr[same_cell, k], atol=1e-4)
    # outside the map (1e-4 margin, sdf_map.h:153-161): (0, 0-grad)
    out = np.array([g.origin - 0.05, g.map_max + 0.01, g.origin + [5e-5, 0.5, 0.5]])
    d, gr = orc.dist_with_grad(og, dist, out)
    assert np.all(d == 0) and np.all(gr == 0)
    # stencil poking outside the map reads -1 (sdf_map.h:228-231) and interpolates it as data
    p = np.array([[g.origin[0] + 0.02, g.origin[1] + 0.5, g.origin[2] + 0.5]])
    d, gr = orc.dist_with_grad(og, np.ones(n), p)
    assert d[0] < 1.0 and gr[0, 0] > 0


def test_tristate_thresholds(orc):
    import math
    cmin, pocc = math.log(0.12 / 0.88), math.log(0.8 / 0.2)
    lo = np.array([cmin - 0.01, cmin - 1e-3 - 1e-9, cmin - 1e-3 + 1e-9, cmin, 0.0, pocc, pocc + 1e-9, 3.0])
    t = orc.tristate_from_logodds(lo, cmin, pocc)
    assert list(t) == [0, 0, 1, 1, 1, 1, 2, 2]  # sdf_map.h:196-199
