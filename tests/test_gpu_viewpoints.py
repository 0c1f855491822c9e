"""Viewpoint sampling parity: fuelgpu_frontier_sample_viewpoints / _changed_counts (C ABI) vs the CPU oracle of
FrontierFinder::sampleViewpoints / countVisibleCells / isFrontierCovered
(active_perception/src/frontier_finder.cpp:662-755,697-719).
Bar: candidate positions bit-exact (host libm on both sides), yaw <= 1e-9 rad (device vs host acos/atan2 ulps;
NaN where the reference is NaN), visible-cell counts IDENTICAL for every candidate: the device raises a flag when a
FOV plane test comes within 1e-9 of zero and those candidates are redone with the yaw from the host's libm."""
import numpy as np
import pytest

from fuel_b200 import workloads as W
from tests.helpers import make_sdf_map, orc_grid, random_scene

pytestmark = pytest.mark.gpu


def compare(orc, og, tri, inflate, vp, ftrs, pos, yaw, vis):
    n_checked = n_border = 0
    for i, f in enumerate(ftrs):
        r = orc.sample_viewpoints(og, tri, inflate, vp, f.average_, f.filtered_cells_)
        assert np.array_equal(pos[i], r["pos"]), "cluster %d candidate positions differ" % i
        assert np.array_equal(vis[i] < 0, r["visib"] < 0), "cluster %d rejected candidates differ" % i
        ok = r["visib"] >= 0
        assert np.array_equal(np.isnan(yaw[i][ok]), np.isnan(r["yaw"][ok]))
        fin = ok & ~np.isnan(r["yaw"])
        d = np.angle(np.exp(1j * (yaw[i][fin] - r["yaw"][fin])))
        assert np.all(np.abs(d) < 1e-9), "cluster %d yaw differs by %g" % (i, np.abs(d).max())
        assert np.array_equal(vis[i][ok], r["visib"][ok]), "cluster %d visible counts differ" % i
        n_checked += int(ok.sum())
        n_border += int((ok & (r["border"] != 0)).sum())
    return n_checked, n_border


def test_office_clusters(fuel, orc):
    """The replan sequence of fast_exploration_manager.cpp:91-105 on the office map: searchFrontiers then
    computeFrontiersToVisit."""
    g, inflate = W.office_map()
    tri = W.office_known(g, inflate)
    m = make_sdf_map(fuel, g, inflate, tri)
    env = fuel.EDTEnvironment()
    env.setMap(m)
    ff = fuel.FrontierFinder(env)
    ftrs = ff.search_box(g.box_min, g.box_max)
    assert len(ftrs) >= 8
    pos, yaw, vis = ff.sampleViewpointsRaw(ftrs)
    vp = orc.view_params()
    n_checked, n_border = compare(orc, orc_grid(orc, g), tri, inflate, vp, ftrs, pos, yaw, vis)
    assert n_checked > 200
    # bookkeeping of computeFrontiersToVisit (:392-423)
    ff.tmp_frontiers_ = ftrs
    ff.computeFrontiersToVisit()
    assert len(ff.frontiers_) + len(ff.dormant_frontiers_) == len(ftrs)
    assert len(ff.frontiers_) >= 1
    for k, f in enumerate(ff.frontiers_):
        assert f.id_ == k and f.viewpoints_
        v = [x[2] for x in f.viewpoints_]
        assert v == sorted(v, reverse=True) and v[-1] > 15
    for f in ff.dormant_frontiers_:
        assert not f.viewpoints_
    pts, yaws, avgs = ff.getTopViewpointsInfo(np.array([0.0, 0.0, 1.0]))
    assert len(pts) == len(ff.frontiers_)
    m.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_scenes_small_fov_and_clearance(fuel, orc, seed):
    """Cluttered random scenes, non-default parameters (short range, narrow FOV, 3-voxel clearance), synthetic clusters."""
    rng = np.random.default_rng(seed)
    n = (80, 72, 30)
    g = W.Grid(n, (-4.0, -3.6, -0.5), 0.1, box_min=(-3.7, -3.3, -0.4), box_max=(3.7, 3.3, 2.3))
    inflate, tri = random_scene(n, seed, p_site=0.004, p_unknown=0.3, blobs=5)
    m = make_sdf_map(fuel, g, inflate, tri)
    env = fuel.EDTEnvironment()
    env.setMap(m)
    ff = fuel.FrontierFinder(env)
    kw = dict(candidate_rmin=0.8, candidate_rmax=2.0, candidate_rnum=4, candidate_dphi=0.3, min_candidate_clearance=0.31,
              top_angle=0.4, left_angle=0.5, right_angle=0.45, max_dist=2.2)
    ff.setViewParams(**kw)
    ftrs = []
    for k in range(12):
        c = np.array([rng.uniform(-2.5, 2.5), rng.uniform(-2.2, 2.2), rng.uniform(0.3, 1.8)])
        cells = c + rng.normal(size=(int(rng.integers(1, 300)), 3)) * np.array([0.5, 0.5, 0.3])
        ftrs.append(fuel.Frontier(m, np.zeros(0, np.int32), cells, cells.mean(axis=0), cells.min(axis=0),
                                                  cells.max(axis=0)))
    pos, yaw, vis = ff.sampleViewpointsRaw(ftrs)
    n_checked, _ = compare(orc, orc_grid(orc, g), tri, inflate, orc.view_params(**kw), ftrs, pos, yaw, vis)
    assert n_checked > 100 and (vis > 0).sum() > 50
    m.close()


def test_changed_counts_and_is_frontier_covered(fuel, orc):
    g, inflate = W.office_map()
    tri = W.office_known(g, inflate)
    m = make_sdf_map(fuel, g, inflate, tri)
    env = fuel.EDTEnvironment()
    env.setMap(m)
    ff = fuel.FrontierFinder(env)
    ff.tmp_frontiers_ = ff.search_box(g.box_min, g.box_max)
    ff.computeFrontiersToVisit()
    m.update_min_, m.update_max_ = np.array(g.box_min, float), np.array(g.box_max, float)
    assert not ff.isFrontierCovered()
    # the robot "sees" around the first cluster: its unknown neighbours become free
    f0 = ff.frontiers_[0]
    tri2 = tri.copy()
    idx = np.stack(np.unravel_index(f0.cells_addr_, g.n), axis=1)
    for d in range(-2, 3):
        for ax in range(3):
            j = idx.copy()
            j[:, ax] = np.clip(j[:, ax] + d, 0, g.n[ax] - 1)
            sel = tri2[j[:, 0], j[:, 1], j[:, 2]] == W.UNKNOWN
            tri2[j[sel, 0], j[sel, 1], j[sel, 2]] = W.FREE
    m.setOccupancyBuffer(tristate=tri2)
    m.upload()
    from fuel_b200._lib import check, lib, ptr
    allf = ff.frontiers_ + ff.dormant_frontiers_
    offs = np.zeros(len(allf) + 1, np.int32)
    offs[1:] = np.cumsum([f.cells_addr_.size for f in allf])
    addr = np.ascontiguousarray(np.concatenate([f.cells_addr_ for f in allf]).astype(np.int32))
    counts = np.zeros(len(allf), np.int32)
    check(lib().fuelgpu_frontier_changed_counts(m.handle, len(allf), ptr(offs), ptr(addr), ptr(counts)), m.handle)
    og = orc_grid(orc, g)
    want = [orc.frontier_changed_count(og, tri2, f.cells_addr_) for f in allf]
    assert list(counts) == want and counts[0] == f0.cells_addr_.size
    assert ff.isFrontierCovered()
    m.close()
