"""The oracle against the committed golden OUTPUTS of the reference's own code (tests/golden/ref_outputs.npz, written by
tools/make_ref_golden.py from oracle/_ref in the build container).  Runs anywhere -- no /root/reference needed."""
import os

import numpy as np
import pytest

import oracle as O
from fuel_b200 import workloads as W

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_outputs.npz")


@pytest.fixture(scope="module")
def gold():
    d = dict(np.load(GOLD))
    n = tuple(int(v) for v in d["n"])
    d["shape"] = n
    d["inflate"] = np.unpackbits(d["inflate_bits"])[:int(np.prod(n))].astype(np.int8).reshape(n)
    mp = dict(zip(d["map_keys"], d["map_vals"]))
    d["map_size"] = [mp["map_size_" + a] for a in "xyz"]
    d["grid"] = O.make_grid(n, float(d["res"]), d["origin"], [mp["box_min_" + a] for a in "xyz"], [mp["box_max_" + a] for a in "xyz"],
                            map_size=d["map_size"])
    d["mp"] = mp
    return d


@pytest.mark.parametrize("name,opt,sgn", [("opt", 1, 0), ("nonopt", 0, 0), ("signed", 1, 1)])
def test_esdf(gold, name, opt, sgn):
    lo, hi = gold["esdf_lo"], gold["esdf_hi"]
    got = O.update_esdf3d(gold["grid"], gold["inflate"], gold["tri"], lo, hi, opt, sgn)
    got = got[lo[0]:hi[0] + 1, lo[1]:hi[1] + 1, lo[2]:hi[2] + 1]
    want = gold["esdf_" + name]
    assert np.array_equal(np.where(got > 1e150, np.inf, got).astype(np.float32), want)


def test_frontier_and_viewpoints(gold):
    ffp = dict(zip(gold["ff_keys"], gold["ff_vals"]))
    flag = np.zeros(gold["shape"], np.int8)
    p = O.frontier_params(cluster_min=int(ffp["cluster_min"]), cluster_size_xy=ffp["cluster_size_xy"],
                          down_sample=int(ffp["down_sample"]), cell_order=0)
    got = O.frontier_search(gold["grid"], gold["tri"], flag, gold["upd_min"], gold["upd_max"], p)
    off, foff = gold["fr_offsets"], gold["fr_foffsets"]
    assert len(got) == len(off) - 1 >= 4
    for i, c in enumerate(got):
        assert np.array_equal(c["addr"], gold["fr_addr"][off[i]:off[i + 1]])
        assert np.array_equal(c["filtered"], gold["fr_filtered"][foff[i]:foff[i + 1]])
        assert np.array_equal(c["average"], gold["fr_average"][i])
        assert np.array_equal(c["box_min"], gold["fr_box_min"][i]) and np.array_equal(c["box_max"], gold["fr_box_max"][i])
    assert np.array_equal(np.packbits(flag.astype(np.uint8)), gold["fr_flags_bits"])
    pu = dict(zip(gold["pu_keys"], gold["pu_vals"]))
    vp = O.view_params(candidate_rmin=ffp["candidate_rmin"], candidate_rmax=ffp["candidate_rmax"],
                       candidate_rnum=int(ffp["candidate_rnum"]), candidate_dphi=ffp["candidate_dphi"],
                       min_candidate_clearance=ffp["min_candidate_clearance"], top_angle=pu["top_angle"],
                       left_angle=pu["left_angle"], right_angle=pu["right_angle"], max_dist=pu["max_dist"])
    voff = gold["vp_offsets"]
    kept = []
    for i, c in enumerate(got):
        r = O.sample_viewpoints(gold["grid"], gold["tri"], gold["inflate"], vp, c["average"], c["filtered"])
        keep = np.nonzero(r["visib"] > int(ffp["min_visib_num"]))[0]
        if len(keep):
            k = len(kept)
            kept.append(i)
            mine = sorted(zip(-r["visib"][keep], r["yaw"][keep], map(tuple, r["pos"][keep])))
            sl = slice(voff[k], voff[k + 1])
            theirs = sorted(zip(-gold["vp_visib"][sl], gold["vp_yaw"][sl], map(tuple, gold["vp_pos"][sl])))
            assert mine == theirs
    assert kept == list(gold["vp_cluster"])


def test_fusion_and_inflation(gold):
    mp = gold["mp"]
    f = O.Fusion(gold["grid"], O.fusion_params(max_ray_length=mp["max_ray_length"]))
    for pts, cam in zip(gold["fus_points"], gold["fus_cams"]):
        lo, hi = f.input_point_cloud(pts, cam)
    assert np.array_equal(f.logodds, gold["fus_logodds"])
    assert np.array_equal(lo, gold["fus_local_lo"]) and np.array_equal(hi, gold["fus_local_hi"])
    a, b = f.updated_box()
    assert np.array_equal(a, gold["fus_upd_min"]) and np.array_equal(b, gold["fus_upd_max"])
    tri = f.tristate().reshape(gold["shape"]).copy()
    inf = np.zeros(gold["shape"], np.int8)
    O.clear_and_inflate(gold["grid"], tri, inf, lo, hi, 2, -1)
    assert np.array_equal(np.packbits(inf.astype(np.uint8)), gold["fus_inflate_bits"])


def test_combine_cost(gold):
    n = gold["shape"]
    dist = O.update_esdf3d(gold["grid"], gold["inflate"], gold["tri"], [0, 0, 0], np.array(n) - 1, 1, 0)
    p = O.opt_params(ld_waypt=0.3)
    for b in range(gold["bs_ctrl"].shape[0]):
        tc = O.traj_consts(1)
        O.fill_traj_const(tc[0], gold["bs_pt_dist"][b], gold["bs_dt"][b], gold["bs_start"][b], gold["bs_end"][b][None, :])
        f, g = O.combine_cost_batch(gold["grid"], dist, p, (type(tc[0]) * 4)(*[tc[0]] * 4), 20, int(gold["bs_mask"]), gold["bs_x"][b])
        assert np.array_equal(f, gold["bs_f"][b]) and np.array_equal(g, gold["bs_grad"][b])
