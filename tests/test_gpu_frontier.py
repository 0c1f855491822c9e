"""Frontier parity: libfuelgpu vs the CPU oracle of FrontierFinder::searchFrontiers /
expandFrontier / splitLargeFrontiers (active_perception/src/frontier_finder.cpp:54-242).
Bar (north_star): bit-exact voxel indices and cluster IDs.  Cell ORDER inside a cluster is
canonical (ascending address) on the device; the oracle is run in that mode
(cell_order=1) for the bit-exact check and in the reference's BFS order (cell_order=0)
for the set-level check (DESIGN.md "frontier cell order")."""
import numpy as np
import pytest

from fuel_b200 import workloads as W
from tests.helpers import make_sdf_map, orc_grid, random_scene

pytestmark = pytest.mark.gpu


def run_gpu(fuel, g, inflate, tri, upd_min, upd_max, flags=None, **kw):
    m = make_sdf_map(fuel, g, inflate, tri)
    env = fuel.EDTEnvironment()
    env.setMap(m)
    ff = fuel.FrontierFinder(env, **kw)
    if flags is not None:
        ff.upload_flags(flags)
    out = ff.search_box(upd_min, upd_max)
    fl = ff.download_flags()
    m.close()
    return out, fl


def run_orc(orc, g, tri, upd_min, upd_max, flags=None, cell_order=1, **kw):
    fl = np.zeros(g.n, dtype=np.int8) if flags is None else flags.copy()
    p = orc.frontier_params(cell_order=cell_order, **kw)
    out = orc.frontier_search(orc_grid(orc, g), tri, fl, upd_min, upd_max, p)
    return out, fl


def assert_same(gpu, ref, exact_order=True):
    assert len(gpu) == len(ref), "cluster count %d vs %d" % (len(gpu), len(ref))
    for i, (a, b) in enumerate(zip(gpu, ref)):
        if exact_order:
            assert np.array_equal(a.cells_addr_, b["addr"]), "cluster %d cells differ" % i
        else:
            assert np.array_equal(np.sort(a.cells_addr_), np.sort(b["addr"])), "cluster %d cell set differs" % i
        assert np.allclose(a.average_, b["average"], rtol=1e-12, atol=1e-12)
        assert np.allclose(a.box_min_, b["box_min"], rtol=0, atol=1e-12)
        assert np.allclose(a.box_max_, b["box_max"], rtol=0, atol=1e-12)
        if exact_order:
            # VoxelGrid centroids: float accumulation in the same order -> identical
            assert a.filtered_cells_.shape == b["filtered"].shape
            assert np.array_equal(a.filtered_cells_, b["filtered"]), "cluster %d filtered cells differ" % i


CASES = [
    # n, seed, box margin (voxels), cluster_min, size_xy, update box (fraction of map), min_z
    ((40, 36, 30), 1, 2, 5, 2.0, (0.0, 1.0), 0.4),
    ((40, 36, 30), 2, 2, 0, 0.8, (0.2, 0.7), 0.4),
    ((64, 50, 24), 3, 1, 20, 1.0, (0.3, 0.6), 0.4),
    ((30, 30, 30), 4, 3, 3, 0.6, (0.0, 1.0), -10.0),
    ((72, 64, 20), 5, 2, 10, 1.5, (0.1, 0.5), 0.9),
    # nz % 32 == 0: the 32-voxels-per-thread sweep (classify_words_kernel / compact_words_kernel)
    ((40, 36, 32), 6, 2, 5, 2.0, (0.0, 1.0), 0.4),
    ((33, 29, 64), 7, 1, 3, 0.8, (0.2, 0.7), 0.4),
    ((24, 20, 96), 8, 3, 0, 0.6, (0.0, 1.0), 3.3),
    ((50, 41, 64), 9, 1, 8, 1.0, (0.1, 0.9), -10.0),
]


@pytest.mark.parametrize("n,seed,margin,cmin,sxy,upd,min_z", CASES)
def test_random_scene(fuel, orc, n, seed, margin, cmin, sxy, upd, min_z):
    res = 0.1
    origin = np.array([-1.0, -2.0, -0.5])
    g0 = W.Grid(n, origin, res)
    g = W.Grid(n, origin, res, box_min=origin + margin * res, box_max=g0.map_max - margin * res)
    inflate, tri = random_scene(n, seed, p_site=0.01, p_unknown=0.5, blobs=7)
    ext = g0.map_max - origin
    upd_min = origin + upd[0] * ext
    upd_max = origin + upd[1] * ext
    kw = dict(cluster_min=cmin, cluster_size_xy=sxy, down_sample=3, min_z=min_z)
    gpu, gfl = run_gpu(fuel, g, inflate, tri, upd_min, upd_max, **kw)
    ref, rfl = run_orc(orc, g, tri, upd_min, upd_max, cell_order=1, **kw)
    assert len(ref) > 0
    assert_same(gpu, ref, exact_order=True)
    assert np.array_equal(gfl, rfl), "frontier_flag_ differs"
    # the reference's own BFS cell order gives the same clusters as sets
    ref_bfs, rfl2 = run_orc(orc, g, tri, upd_min, upd_max, cell_order=0, **kw)
    assert_same(gpu, ref_bfs, exact_order=False)
    assert np.array_equal(gfl, rfl2)
    # FUELGPU_CELLS_BFS: the reference's order itself, with average_ and filtered_cells_ recomputed in it
    gpu_bfs, gfl3 = run_gpu(fuel, g, inflate, tri, upd_min, upd_max, cell_order="bfs", **kw)
    assert len(gpu_bfs) == len(ref_bfs) and np.array_equal(gfl3, rfl2)
    for i, (a, b) in enumerate(zip(gpu_bfs, ref_bfs)):
        assert np.array_equal(a.cells_addr_, b["addr"]), "cluster %d BFS order differs" % i
        assert np.array_equal(a.average_, b["average"]), "cluster %d average_ not bit-exact" % i
        assert np.array_equal(a.filtered_cells_, b["filtered"]), "cluster %d filtered_cells_ not bit-exact" % i


def test_preexisting_flags_and_second_sweep(fuel, orc):
    """Stateful use: cells flagged by an earlier sweep are not re-clustered; dropped small
    clusters stay flagged (frontier_finder.cpp:157)."""
    n = (48, 44, 26)
    origin = np.array([0.0, 0.0, -0.2])
    g0 = W.Grid(n, origin, 0.1)
    g = W.Grid(n, origin, 0.1, box_min=origin + 0.2, box_max=g0.map_max - 0.2)
    inflate, tri = random_scene(n, 21, p_site=0.01, p_unknown=0.5, blobs=8)
    kw = dict(cluster_min=15, cluster_size_xy=1.0, down_sample=3, min_z=0.4)
    u1 = (origin + [0.0, 0.0, 0.0], origin + [2.0, 2.0, 2.6])
    u2 = (origin + [1.0, 1.0, 0.0], origin + [4.8, 4.4, 2.6])
    ref1, fl1 = run_orc(orc, g, tri, *u1, **kw)
    ref2, fl2 = run_orc(orc, g, tri, *u2, flags=fl1, **kw)
    m = make_sdf_map(fuel, g, inflate, tri)
    env = fuel.EDTEnvironment()
    env.setMap(m)
    ff = fuel.FrontierFinder(env, **kw)
    gpu1 = ff.search_box(*u1)
    assert_same(gpu1, ref1)
    assert np.array_equal(ff.download_flags(), fl1)
    gpu2 = ff.search_box(*u2)
    assert_same(gpu2, ref2)
    assert np.array_equal(ff.download_flags(), fl2)
    m.close()


def test_search_frontiers_removes_changed(fuel, orc):
    """searchFrontiers bookkeeping (:65-92): stored clusters overlapping the updated box whose
    cells stopped being frontier are dropped and their flags cleared, then re-found."""
    n = (48, 44, 26)
    origin = np.array([0.0, 0.0, -0.2])
    g0 = W.Grid(n, origin, 0.1)
    g = W.Grid(n, origin, 0.1, box_min=origin + 0.2, box_max=g0.map_max - 0.2)
    inflate, tri = random_scene(n, 33, p_site=0.0, p_unknown=0.5, blobs=8)
    kw = dict(cluster_min=10, cluster_size_xy=1.5, down_sample=3, min_z=0.1)
    m = make_sdf_map(fuel, g, inflate, tri)
    env = fuel.EDTEnvironment()
    env.setMap(m)
    ff = fuel.FrontierFinder(env, **kw)
    m.update_min_, m.update_max_ = origin.copy(), g0.map_max.copy()
    first = ff.searchFrontiers()
    assert len(first) >= 2
    ff.frontiers_ = list(first)
    # explore: turn the unknown neighbours of cluster 0 into free space
    tri2 = tri.copy()
    a = first[0].cells_addr_.astype(np.int64)
    nyz = n[1] * n[2]
    idx = np.stack([a // nyz, (a % nyz) // n[2], a % n[2]], axis=1)
    lo = np.maximum(idx.min(axis=0) - 2, 0)
    hi = np.minimum(idx.max(axis=0) + 3, n)
    sub = tri2[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]
    sub[sub == W.UNKNOWN] = W.FREE
    m.setOccupancyBuffer(tristate=tri2)
    m.upload()
    m.update_min_ = g0.index_to_pos(lo)
    m.update_max_ = g0.index_to_pos(hi - 1)
    # oracle: same bookkeeping with its own predicate
    og = orc_grid(orc, g)
    changed_ref = [bool(orc.lib().orc_frontier_is_changed(__import__("ctypes").byref(og),
                                                          tri2.ctypes.data_as(__import__("ctypes").c_void_p),
                                                          f.cells_addr_.ctypes.data_as(__import__("ctypes").c_void_p),
                                                          f.cells_addr_.size)) for f in first]
    fl = ff.download_flags()
    overl = [ff.haveOverlap(f.box_min_, f.box_max_, m.update_min_, m.update_max_) for f in first]
    for f, ch, ov in zip(first, changed_ref, overl):
        if ch and ov:
            fl.reshape(-1)[f.cells_addr_] = 0
    ref2, fl_ref = run_orc(orc, g, tri2, m.update_min_, m.update_max_, flags=fl, **kw)
    second = ff.searchFrontiers()
    assert changed_ref[0]
    assert 0 in ff.removed_ids_
    assert len(ff.frontiers_) == len(first) - sum(1 for ch, ov in zip(changed_ref, overl) if ch and ov)
    assert_same(second, ref2)
    assert np.array_equal(ff.download_flags(), fl_ref)
    m.close()


def test_office_fixture(fuel, orc):
    """BASELINE config 1/2: office.pcd map with a seeded known region, FUEL's parameters."""
    g, inflate = W.office_map()
    tri = W.office_known(g, inflate)
    kw = dict(cluster_min=100, cluster_size_xy=2.0, down_sample=3, min_z=0.4)
    gpu, gfl = run_gpu(fuel, g, inflate, tri, g.origin, g.map_max, **kw)
    ref, rfl = run_orc(orc, g, tri, g.origin, g.map_max, cell_order=1, **kw)
    assert len(ref) >= 3
    assert_same(gpu, ref)
    assert np.array_equal(gfl, rfl)
    ref_bfs, _ = run_orc(orc, g, tri, g.origin, g.map_max, cell_order=0, **kw)
    assert_same(gpu, ref_bfs, exact_order=False)


def test_empty_and_all_unknown(fuel, orc):
    n = (20, 20, 20)
    g = W.Grid(n, (0, 0, 0), 0.1, box_min=(0.2, 0.2, 0.2), box_max=(1.8, 1.8, 1.8))
    inflate = np.zeros(n, dtype=np.int8)
    for tri in (np.zeros(n, dtype=np.uint8), np.full(n, W.FREE, dtype=np.uint8)):
        gpu, gfl = run_gpu(fuel, g, inflate, tri, (0, 0, 0), (2, 2, 2), cluster_min=0)
        assert gpu == [] and not gfl.any()


def test_large_scene_uses_multi_kernel_path(fuel, orc):
    """More than 32768 candidate cells: the single-CTA small path overflows and the multi-kernel
    path takes over; results must be identical to the oracle either way."""
    n = (192, 160, 48)
    origin = np.array([0.0, 0.0, 0.0])
    g0 = W.Grid(n, origin, 0.1)
    g = W.Grid(n, origin, 0.1, box_min=origin + 0.3, box_max=g0.map_max - 0.3)
    inflate, tri = random_scene(n, 99, p_site=0.002, p_unknown=0.5, blobs=90)
    kw = dict(cluster_min=30, cluster_size_xy=1.5, down_sample=3, min_z=0.4)
    gpu, gfl = run_gpu(fuel, g, inflate, tri, origin, g0.map_max, **kw)
    ref, rfl = run_orc(orc, g, tri, origin, g0.map_max, cell_order=1, **kw)
    assert int(rfl.sum()) > 32768, int(rfl.sum())
    assert_same(gpu, ref)
    assert np.array_equal(gfl, rfl)


@pytest.mark.parametrize("nz", [48, 64])
def test_large_scene_levels_without_host(fuel, orc, nz):
    """The multi-kernel path keeps its counts on the device (LevelCtl): many split levels (small cluster_size_xy), both
    sweep layouts (nz = 64: 32 voxels per thread), and a second search on the flags the first one left."""
    n = (160, 144, nz)
    origin = np.array([-3.0, 1.0, -0.4])
    g0 = W.Grid(n, origin, 0.1)
    g = W.Grid(n, origin, 0.1, box_min=origin + 0.2, box_max=g0.map_max - 0.2)
    inflate, tri = random_scene(n, 123 + nz, p_site=0.002, p_unknown=0.5, blobs=80)
    kw = dict(cluster_min=12, cluster_size_xy=0.45, down_sample=3, min_z=0.4)
    m = make_sdf_map(fuel, g, inflate, tri)
    env = fuel.EDTEnvironment()
    env.setMap(m)
    ff = fuel.FrontierFinder(env, **kw)
    ext = g0.map_max - origin
    u1 = (origin + 0.0 * ext, origin + 0.55 * ext)
    gpu1 = ff.search_box(*u1)
    ref1, fl1 = run_orc(orc, g, tri, *u1, cell_order=1, **kw)
    assert int(fl1.sum()) > 32768, int(fl1.sum())
    assert_same(gpu1, ref1)
    assert np.array_equal(ff.download_flags(), fl1)
    gpu2 = ff.search_box(origin, g0.map_max)
    ref2, fl2 = run_orc(orc, g, tri, origin, g0.map_max, flags=fl1, cell_order=1, **kw)
    assert_same(gpu2, ref2)
    assert np.array_equal(ff.download_flags(), fl2)
    m.close()


@pytest.mark.parametrize("slabs,nz", [(2, 24), (3, 24), (5, 24), (2, 64), (3, 64), (5, 96)])
def test_sharded_sweep_equals_whole_search(fuel, orc, slabs, nz):
    """SURVEY 8e row 2 on one GPU: the sweep cut into z slabs (fuelgpu_frontier_candidates per slab), the candidate lists
    merged by address, the clustering run on the union (fuelgpu_frontier_search_from_candidates) == the whole search,
    bit for bit: clusters, order, cells, average_, filtered_cells_, frontier_flag_.  Also after a first search left flags."""
    from fuel_b200.dist import merge_candidates
    n = (64, 50, nz)
    origin = np.array([-1.0, -2.0, -0.5])
    g0 = W.Grid(n, origin, 0.1)
    g = W.Grid(n, origin, 0.1, box_min=origin + 0.1, box_max=g0.map_max - 0.1)
    inflate, tri = random_scene(n, 3, p_site=0.01, p_unknown=0.5, blobs=7)
    kw = dict(cluster_min=8, cluster_size_xy=1.0, down_sample=3, min_z=0.4)
    ext = g0.map_max - origin
    boxes = [(origin + 0.3 * ext, origin + 0.6 * ext), (origin, g0.map_max)]
    m1, m2 = make_sdf_map(fuel, g, inflate, tri), make_sdf_map(fuel, g, inflate, tri)
    ffs = []
    for m in (m1, m2):
        env = fuel.EDTEnvironment()
        env.setMap(m)
        ffs.append(fuel.FrontierFinder(env, **kw))
    cuts = np.linspace(0, n[2], slabs + 1).astype(int)
    for it, (umin, umax) in enumerate(boxes):  # the second search runs on the flags the first one left
        if it == 1:  # clusters 0 and 2 are "removed" (resetFlag, :62-69) on both maps: their cells can be found again
            for ff in ffs:
                ff._clear_flags(np.ascontiguousarray(np.concatenate([whole[0].cells_addr_, whole[2].cells_addr_])))
        whole = ffs[0].search_box(umin, umax)
        parts = [ffs[1].candidates(umin, umax, cuts[i], cuts[i + 1] - 1) for i in range(slabs)]
        assert sum(a.size for a, _ in parts) > 0
        addr, cls = merge_candidates(parts)
        got = ffs[1].search_from_candidates(umin, umax, addr, cls)
        assert len(got) == len(whole) and len(whole) > 0
        for a, b in zip(got, whole):
            assert np.array_equal(a.cells_addr_, b.cells_addr_)
            assert np.array_equal(a.average_, b.average_) and np.array_equal(a.filtered_cells_, b.filtered_cells_)
            assert np.array_equal(a.box_min_, b.box_min_) and np.array_equal(a.box_max_, b.box_max_)
        assert np.array_equal(ffs[0].download_flags(), ffs[1].download_flags())
    m1.close()
    m2.close()
