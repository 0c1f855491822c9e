"""Pins the B-spline cost oracle (restating bspline_opt/src/bspline_optimizer.cpp:255-647) with
central finite differences -- only where the reference gradient IS the gradient of the
reference cost: position gradients of smoothness / feasibility / start / end /
guide / waypoint terms, dt-gradient of feasibility, the velocity boundary terms and the
time term.  The normalised distance gradient, the 4x-off acceleration-boundary dt-gradient
and the zero smoothness dt-gradient are reference behaviour and are checked as such."""
import numpy as np
import pytest

from fuel_b200 import workloads as W
from tests.helpers import orc_grid


@pytest.fixture(scope="module")
def env(orc):
    n = (60, 50, 30)
    g0 = W.Grid(n, (-3.0, -2.5, -0.5), 0.1)
    g = W.Grid(n, g0.origin, 0.1, box_min=g0.origin + 0.3, box_max=g0.map_max - 0.3)
    rng = np.random.default_rng(1)
    inflate = (rng.random(n) < 0.002).astype(np.int8)
    og = orc_grid(orc, g)
    d = orc.update_esdf3d(og, inflate, None, [0, 0, 0], np.array(n) - 1, True, False, threads=4)
    return dict(g=g, og=og, dist=d, inflate=inflate)


def batch(orc, env, B, N, seed, n_end=1, time_lb=None, guide=False, waypt=False):
    tr = W.make_trajectories(env["g"], env["inflate"], B=B, n_pts=N, seed=seed, sigma=0.15, spacing=0.2)
    rng = np.random.default_rng(seed)
    tcs = orc.traj_consts(B)
    for b in range(B):
        end = np.zeros((n_end, 3))
        end[0] = tr["end_pos"][b]
        if n_end > 1:
            end[1:] = rng.normal(size=(n_end - 1, 3))
        orc.fill_traj_const(tcs[b], tr["pt_dist"][b], tr["dt"][b], tr["start"][b], end,
                            -1.0 if time_lb is None else time_lb[b],
                            tr["ctrl"][b, 3:N - 3] + 0.1 if guide else None,
                            tr["ctrl"][b, [1, 5, N - 2]] + 0.05 if waypt else None, [0, 4, N - 3] if waypt else None)
    return tr, tcs


def fd(orc, env, tcs, N, mask, x, cols, eps=1e-6):
    g = np.zeros((x.shape[0], len(cols)))
    for j, c in enumerate(cols):
        xp, xm = x.copy(), x.copy()
        xp[:, c] += eps
        xm[:, c] -= eps
        fp, _ = orc.combine_cost_batch(env["og"], env["dist"], orc.opt_params(), tcs, N, mask, xp)
        fm, _ = orc.combine_cost_batch(env["og"], env["dist"], orc.opt_params(), tcs, N, mask, xm)
        g[:, j] = (fp - fm) / (2 * eps)
    return g


@pytest.mark.parametrize("term", ["SMOOTHNESS", "FEASIBILITY", "START", "END", "GUIDE", "WAYPOINTS"])
def test_position_gradient_fd(orc, env, term):
    B, N = 6, 12
    tr, tcs = batch(orc, env, B, N, 3, n_end=3, guide=True, waypt=True)
    mask = getattr(orc, term)
    x = W.pack_x(tr["ctrl"], tr["dt"] * 0.6, mintime=False)
    for b in range(B):
        tcs[b].knot_span = tr["dt"][b] * 0.6  # make the feasibility hinge bite
    _, gr = orc.combine_cost_batch(env["og"], env["dist"], orc.opt_params(), tcs, N, mask, x)
    cols = list(range(3 * N))
    num = fd(orc, env, tcs, N, mask, x, cols)
    scale = np.maximum(np.abs(gr).max(), 1.0)
    assert np.allclose(num, gr, atol=2e-5 * scale), np.abs(num - gr).max()


def test_dt_gradient_fd_exact_terms(orc, env):
    """feasibility + velocity boundary (n_end = 2, start acc target chosen so the acc terms vanish)
    + time term with an active lower bound."""
    B, N = 6, 12
    tr, tcs = batch(orc, env, B, N, 5, n_end=2, time_lb=np.full(6, 10.0))
    mask = orc.FEASIBILITY | orc.MINTIME
    x = W.pack_x(tr["ctrl"], tr["dt"] * 0.6, mintime=True)
    _, gr = orc.combine_cost_batch(env["og"], env["dist"], orc.opt_params(), tcs, N, mask, x)
    num = fd(orc, env, tcs, N, mask, x, [3 * N])
    assert np.allclose(num[:, 0], gr[:, -1], rtol=1e-5, atol=1e-4)


def test_reference_quirks_are_reproduced(orc, env):
    B, N = 4, 12
    tr, tcs = batch(orc, env, B, N, 7, n_end=3)
    x = W.pack_x(tr["ctrl"], tr["dt"], mintime=True)
    # (iii) smoothness has no dt-gradient (:279-280 commented out)
    _, gr = orc.combine_cost_batch(env["og"], env["dist"], orc.opt_params(), tcs, N, orc.SMOOTHNESS | orc.MINTIME, x)
    assert np.allclose(gr[:, -1], orc.opt_params().ld_time * (N - 3))
    # (ii) start: analytic dt-gradient of the acceleration term is 1/4 of the true one (:390)
    p = orc.opt_params(ld_time=0.0)
    _, g1 = orc.combine_cost_batch(env["og"], env["dist"], p, tcs, N, orc.START | orc.MINTIME, x)
    eps = 1e-7
    xp, xm = x.copy(), x.copy()
    xp[:, -1] += eps
    xm[:, -1] -= eps
    fp, _ = orc.combine_cost_batch(env["og"], env["dist"], p, tcs, N, orc.START | orc.MINTIME, xp)
    fm, _ = orc.combine_cost_batch(env["og"], env["dist"], p, tcs, N, orc.START | orc.MINTIME, xm)
    assert not np.allclose((fp - fm) / (2 * eps), g1[:, -1], rtol=1e-3)  # reference gradient != d f/d dt
    # (i) distance term uses the normalised ESDF gradient when |grad| > 1e-4 (:294-295)
    q = tr["ctrl"].reshape(-1, 3)
    d, dg = orc.dist_with_grad(env["og"], env["dist"], q)
    nrm = np.linalg.norm(dg, axis=1)
    dgn = np.where(nrm[:, None] > 1e-4, dg / np.maximum(nrm, 1e-300)[:, None], dg)
    man = np.where((d < 0.7)[:, None], 2.0 * (d - 0.7)[:, None] * dgn, 0.0).reshape(B, N, 3)
    xs = W.pack_x(tr["ctrl"], tr["dt"], mintime=False)
    f, gd = orc.combine_cost_batch(env["og"], env["dist"], orc.opt_params(), tcs, N, orc.DISTANCE, xs)
    assert np.allclose(gd.reshape(B, N, 3), 10.0 * man, rtol=1e-12, atol=1e-12)
    assert np.allclose(f, 10.0 * np.where(d < 0.7, (d - 0.7) ** 2, 0).reshape(B, N).sum(1), rtol=1e-12)


def test_pt_dist_divides_by_point_count(orc):
    ctrl = np.array([[0, 0, 0], [1, 0, 0], [1, 2, 0], [1, 2, 2.0]])
    assert orc.pt_dist(ctrl) == pytest.approx(5.0 / 4.0)  # :136-140, not /3


def test_cpu_twin_optimizer_descends(orc, env):
    B, N = 16, 16
    tr, tcs = batch(orc, env, B, N, 11)
    mask = orc.NORMAL_PHASE | orc.MINTIME
    x = W.pack_x(tr["ctrl"], tr["dt"])
    f0, _ = orc.combine_cost_batch(env["og"], env["dist"], orc.opt_params(), tcs, N, mask, x)
    xb, fb, ne = orc.optimize_batch(env["og"], env["dist"], orc.opt_params(), tcs, N, mask, x, max_eval=40)
    assert np.all(fb <= f0) and np.all(ne <= 40) and np.mean(fb) < 0.2 * np.mean(f0)
    fchk, _ = orc.combine_cost_batch(env["og"], env["dist"], orc.opt_params(), tcs, N, mask, xb)
    assert np.allclose(fchk, fb, rtol=1e-12)
    x1, f1, n1 = orc.optimize_batch(env["og"], env["dist"], orc.opt_params(), tcs, N, mask, x, max_eval=1)
    assert np.all(n1 == 1) and np.allclose(f1, f0)
