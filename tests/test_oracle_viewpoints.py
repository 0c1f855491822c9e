"""Oracle self-checks for the viewpoint-sampling restatement (oracle/fuel_oracle_viewpoints.c;
FrontierFinder::sampleViewpoints / countVisibleCells, active_perception/src/frontier_finder.cpp:662-755)
against geometry that can be worked out by hand."""
import numpy as np

import oracle as O
from fuel_b200 import workloads as W


def scene(n=(120, 120, 30), origin=(-6.0, -6.0, -0.5), res=0.1):
    g = O.make_grid(n, res, origin)
    tri = np.full(n, W.FREE, np.uint8)
    inflate = np.zeros(n, np.int8)
    return g, tri, inflate


def wall_cells():
    # a small vertical patch of "frontier cells" around the origin, facing +x/-x
    ys, zs = np.meshgrid(np.arange(-0.45, 0.5, 0.3), np.arange(0.55, 1.5, 0.3))
    return np.stack([np.zeros(ys.size), ys.ravel(), zs.ravel()], axis=1)


def test_candidate_grid_matches_the_reference_loops():
    g, tri, inflate = scene()
    vp = O.view_params()
    r = O.sample_viewpoints(g, tri, inflate, vp, [0.0, 0.0, 1.0], wall_cells())
    assert r["pos"].shape == (100, 3)          # 4 radii x 25 angles (-pi + 24*dphi < pi because dphi uses 3.1415926)
    rad = np.linalg.norm(r["pos"][:, :2], axis=1)
    assert np.allclose(rad.reshape(4, 25), np.array([1.5, 1.5 + 1 / 3, 1.5 + 2 / 3, 2.5])[:, None])
    assert np.allclose(r["pos"][:, 2], 1.0)
    ang = np.arctan2(r["pos"][:25, 1], r["pos"][:25, 0])
    assert np.isclose(ang[0], -np.pi) or np.isclose(ang[0], np.pi)
    assert np.allclose(np.diff(np.unwrap(ang)), 15 * 3.1415926 / 180.0)


def test_free_space_sees_every_cell_and_yaw_points_at_the_cluster():
    g, tri, inflate = scene()
    vp = O.view_params()
    cells = wall_cells()
    r = O.sample_viewpoints(g, tri, inflate, vp, cells.mean(axis=0), cells)
    assert np.all(r["visib"] == len(cells))      # nothing occludes, patch well inside the FOV at >= 1.5 m
    # the "average yaw" is the mean of signed 3-D angles to the first cell (:675-684); for cells at the viewpoint's own
    # height those are planar bearings, so the result is the mean bearing of the cells
    flat = np.stack([np.zeros(7), np.linspace(-0.45, 0.45, 7), np.full(7, 1.0)], axis=1)
    r = O.sample_viewpoints(g, tri, inflate, vp, flat.mean(axis=0), flat)
    # viewpoints in line with the cells give dot products a hair above 1 -> acos = NaN in the reference too
    assert np.isnan(r["yaw"]).sum() <= 8 and np.all(np.abs(r["pos"][np.isnan(r["yaw"]), 0]) < 1e-6)
    for c in np.nonzero(~np.isnan(r["yaw"]))[0]:
        b = np.arctan2(flat[:, 1] - r["pos"][c, 1], flat[:, 0] - r["pos"][c, 0])
        rel = np.angle(np.exp(1j * (b - b[0])))
        want = np.angle(np.exp(1j * (rel.sum() / len(flat) + b[0])))
        assert abs(np.angle(np.exp(1j * (r["yaw"][c] - want)))) < 1e-6  # acos near 1 is ill-conditioned


def test_occluder_unknown_and_rejections():
    g, tri, inflate = scene()
    vp = O.view_params()
    cells = wall_cells()
    avg = cells.mean(axis=0)
    # an inflated wall at x = +0.5 m hides the patch from every candidate with x > 0.5 whose rays cross it
    ix = O.pos_to_index(g, np.array([0.5, 0.0, 0.0]))[0]
    inflate[ix, :, :] = 1
    r = O.sample_viewpoints(g, tri, inflate, vp, avg, cells)
    right = r["pos"][:, 0] > 0.7
    left = r["pos"][:, 0] < -0.7
    assert np.all(r["visib"][right] == 0) and np.all(r["visib"][left] == len(cells))
    # unknown space hides too
    inflate[...] = 0
    tri[ix, :, :] = W.UNKNOWN
    r2 = O.sample_viewpoints(g, tri, inflate, vp, avg, cells)
    assert np.all(r2["visib"][right & (r2["visib"] >= 0)] == 0) and np.all(r2["visib"][left] == len(cells))
    # a candidate standing within min_candidate_clearance (2 voxels) of unknown, or in an inflated voxel, or outside
    # the box is rejected (-1)
    tri[...] = W.FREE
    c7 = r["pos"][7]
    i7 = O.pos_to_index(g, c7)
    tri[i7[0] + 2, i7[1], i7[2]] = W.UNKNOWN
    inflate[tuple(O.pos_to_index(g, r["pos"][40]))] = 1
    r3 = O.sample_viewpoints(g, tri, inflate, vp, avg, cells)
    assert r3["visib"][7] == -1 and r3["visib"][40] == -1
    assert (r3["visib"] == -1).sum() < 12
    g2 = O.make_grid((120, 120, 30), 0.1, (-6.0, -6.0, -0.5), box_mind=(-1.0, -6.0, -0.5), box_maxd=(6.0, 6.0, 2.5))
    r4 = O.sample_viewpoints(g2, np.full((120, 120, 30), W.FREE, np.uint8), np.zeros((120, 120, 30), np.int8), vp, avg, cells)
    assert np.all((r4["visib"] == -1) == (r4["pos"][:, 0] <= -1.0))


def test_fov_limits():
    """A cell straight above the viewpoint's optical axis beyond the vertical half-angle is not counted."""
    g, tri, inflate = scene()
    vp = O.view_params()
    base = wall_cells()
    # add cells high above: elevation from 1.5 m away = atan(1.2/1.5) = 0.675 rad > top_angle 0.56125
    high = np.array([[0.0, 0.0, 1.0 + 1.2], [0.0, 0.1, 1.0 + 1.25]])
    cells = np.concatenate([base, high])
    r = O.sample_viewpoints(g, tri, inflate, vp, [0.0, 0.0, 1.0], cells)
    near = np.isclose(np.linalg.norm(r["pos"][:, :2], axis=1), 1.5)
    assert np.all(r["visib"][near] == len(base))
    # range limit: max_dist 1.0 sees nothing from >= 1.5 m
    r = O.sample_viewpoints(g, tri, inflate, O.view_params(max_dist=1.0), [0.0, 0.0, 1.0], cells)
    assert np.all(r["visib"] == 0)


def test_changed_count():
    g = O.make_grid((20, 20, 10), 0.1, (0.0, 0.0, 0.0))
    tri = np.full((20, 20, 10), W.FREE, np.uint8)
    tri[10:, :, :] = W.UNKNOWN
    addr = np.array([(9 * 20 + y) * 10 + 5 for y in range(20)], np.int32)  # the free layer touching unknown
    assert O.frontier_changed_count(g, tri, addr) == 0
    tri[10, :7, 5] = W.FREE  # 7 of them lose their unknown neighbour
    assert O.frontier_changed_count(g, tri, addr) == 7
