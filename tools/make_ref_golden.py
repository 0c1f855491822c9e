#!/usr/bin/env python
"""Golden OUTPUT vectors from the reference's own code -> tests/golden/ref_outputs.npz.

Runs ONLY in the build container: it drives oracle/_ref/libfuel_ref.so, i.e. the reference's plan_env/src/sdf_map.cpp,
raycast.cpp, bspline_opt/src/bspline_optimizer.cpp, active_perception/src/frontier_finder.cpp and perception_utils.cpp
compiled UNMODIFIED from /root/reference (oracle/Makefile).  The file it writes holds small
seeded inputs and what the reference computes for them: ESDF in three modes, the frontier clusters / flags / viewpoints
of a partly known room, the log-odds map after a few fused point clouds, the inflated occupancy, and combineCost values
and gradients.  tests/test_gpu_golden_ref.py compares the CUDA path with these on the GPU box (where /root/reference
does not exist); tests/test_oracle_golden.py compares the oracle with them anywhere.
Not the reference's: pcl::VoxelGrid and Eigen::EigenSolver inside the frontier path (third party, absent; the oracle's
reconstructions stand in, see oracle/ref_standin) -- they influence `filtered` and the split direction only.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from fuel_b200 import workloads as W  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_outputs.npz")
MAP = dict(resolution=0.1, map_size_x=6.4, map_size_y=4.8, map_size_z=2.4, ground_height=-0.5, obstacles_inflation=0.199,
           local_bound_inflate=0.5, local_map_margin=50, default_dist=0.0, optimistic=0, signed_dist=0, p_hit=0.65,
           p_miss=0.35, p_min=0.12, p_max=0.90, p_occ=0.80, max_ray_length=2.5, virtual_ceil_height=-10.0,
           box_min_x=-2.9, box_min_y=-2.1, box_min_z=-0.3, box_max_x=2.9, box_max_y=2.1, box_max_z=1.7)
FF = dict(cluster_min=20, cluster_size_xy=1.0, cluster_size_z=10.0, min_candidate_dist=0.75, min_candidate_clearance=0.21,
          candidate_dphi=15 * 3.1415926 / 180.0, candidate_rmax=2.5, candidate_rmin=1.5, candidate_rnum=3, down_sample=3,
          min_visib_num=8, min_view_finish_fraction=0.2)
PU = dict(top_angle=0.56125, left_angle=0.69222, right_angle=0.68901, max_dist=4.5, vis_dist=1.0)
OPT = dict(ld_smooth=20.0, ld_dist=10.0, ld_feasi=2.0, ld_start=100.0, ld_end=0.5, ld_guide=1.5, ld_waypt=0.3, ld_view=0.0,
           ld_time=1.0, dist0=0.7, max_vel=2.0, max_acc=2.0, dlmin=0.0, wnl=0.0, max_iteration_num1=2, max_iteration_num2=2000,
           max_iteration_num3=200, max_iteration_num4=200, max_iteration_time1=0.0001, max_iteration_time2=0.005,
           max_iteration_time3=0.003, max_iteration_time4=0.003, algorithm1=15, algorithm2=11, bspline_degree=3)


def logit(p):
    return float(np.log(p / (1 - p)))


def scene(n, seed):
    rng = np.random.default_rng(seed)
    inflate = (rng.random(n) < 0.004).astype(np.int8)
    X, Y, Z = np.meshgrid(*[np.arange(k) for k in n], indexing="ij")
    known = np.zeros(n, bool)
    for _ in range(5):
        c = rng.uniform(0.2, 0.8, 3) * np.array(n)
        r = rng.uniform(0.22, 0.45) * min(n[0], n[1])
        known |= ((X - c[0]) ** 2 + (Y - c[1]) ** 2 + 4.0 * (Z - c[2]) ** 2) < r * r
    tri = np.where(known, W.FREE, W.UNKNOWN).astype(np.uint8)
    tri[known & (inflate == 1)] = W.OCCUPIED
    inflate[~known] = 0
    return inflate, tri


def main():
    O.build()
    assert O.ref_raycast() is not None, "oracle/_ref was not built (needs /root/reference)"
    out = {}
    ref = O.RefSDFMap(**MAP)
    n = ref.n
    out["map_keys"] = np.array(list(MAP.keys()))
    out["map_vals"] = np.array(list(MAP.values()), dtype=np.float64)
    out["n"], out["origin"], out["res"] = np.array(n, np.int32), ref.origin.copy(), np.float64(ref.res)
    inflate, tri = scene(n, 11)
    out["inflate_bits"], out["tri"] = np.packbits(inflate.astype(np.uint8)), tri

    def load_state():
        ref.inflate[:] = inflate.reshape(-1)
        ref.occupancy[:] = np.where(tri == W.UNKNOWN, logit(0.12) - 0.01,
                                    np.where(tri == W.OCCUPIED, logit(0.90), logit(0.12))).reshape(-1)

    # ---- ESDF: optimistic / non-optimistic / signed, on a sub-box ----
    lo, hi = np.array([4, 3, 1], np.int32), np.array([59, 44, 22], np.int32)
    out["esdf_lo"], out["esdf_hi"] = lo, hi
    for name, (opt, sgn) in dict(opt=(1, 0), nonopt=(0, 0), signed=(1, 1)).items():
        load_state()
        ref.distance[:] = 0.0
        ref.set_modes(opt, sgn)
        ref.set_local_bound(lo, hi)
        ref.update_esdf3d()
        d = ref.distance.reshape(n)[lo[0]:hi[0] + 1, lo[1]:hi[1] + 1, lo[2]:hi[2] + 1]
        out["esdf_" + name] = np.where(d > 1e150, np.inf, d).astype(np.float32)  # sentinel -> inf; bar is 1e-4 relative

    # ---- frontier search + viewpoints ----
    load_state()
    ff = O.RefFrontierFinder(ref, PU, **FF)
    upd = (np.array([-3.2, -2.4, -0.5]), np.array([3.2, 2.4, 1.9]))
    out["ff_keys"], out["ff_vals"] = np.array(list(FF.keys())), np.array(list(FF.values()), dtype=np.float64)
    out["pu_keys"], out["pu_vals"] = np.array(list(PU.keys())), np.array(list(PU.values()), dtype=np.float64)
    out["upd_min"], out["upd_max"] = upd
    tmp = ff.search(*upd)
    out["fr_offsets"] = np.cumsum([0] + [len(t["addr"]) for t in tmp]).astype(np.int32)
    out["fr_addr"] = np.concatenate([t["addr"] for t in tmp]).astype(np.int32)
    out["fr_foffsets"] = np.cumsum([0] + [len(t["filtered"]) for t in tmp]).astype(np.int32)
    out["fr_filtered"] = np.concatenate([t["filtered"] for t in tmp])
    out["fr_average"] = np.stack([t["average"] for t in tmp])
    out["fr_box_min"] = np.stack([t["box_min"] for t in tmp])
    out["fr_box_max"] = np.stack([t["box_max"] for t in tmp])
    out["fr_flags_bits"] = np.packbits(ff.flags.astype(np.uint8))
    visit, dormant = ff.compute_to_visit()
    # which tmp clusters were kept, and their viewpoints (sorted by the reference)
    first = [int(t["addr"][0]) for t in tmp]
    out["vp_cluster"] = np.array([first.index(int(v["addr"][0])) for v in visit], np.int32)
    out["vp_offsets"] = np.cumsum([0] + [len(v["view_yaw"]) for v in visit]).astype(np.int32)
    out["vp_pos"] = np.concatenate([v["view_pos"] for v in visit])
    out["vp_yaw"] = np.concatenate([v["view_yaw"] for v in visit])
    out["vp_visib"] = np.concatenate([v["view_visib"] for v in visit]).astype(np.int32)
    ff.close()
    print("frontier: %d clusters (%d to visit, %d dormant), %d cells" % (len(tmp), len(visit), len(dormant), len(out["fr_addr"])))

    # ---- fusion of three clouds into a fresh map, then inflation ----
    ref2 = O.RefSDFMap(**MAP)
    rng = np.random.default_rng(21)
    clouds, cams = [], []
    for k in range(3):
        cam = np.array([rng.uniform(-2, 2), rng.uniform(-1.5, 1.5), rng.uniform(0.3, 1.5)])
        pts = cam + rng.normal(size=(2500, 3)) * np.array([1.6, 1.6, 0.6])
        pts[:40] = np.round(pts[:40])
        pts[40:80] = pts[40]
        pts[80:110] *= 6.0
        pts = pts.astype(np.float32)
        ref2.input_point_cloud(pts, cam)
        clouds.append(pts)
        cams.append(cam)
    out["fus_points"], out["fus_cams"] = np.stack(clouds), np.stack(cams)
    out["fus_logodds"] = ref2.occupancy.copy()
    out["fus_local_lo"], out["fus_local_hi"] = ref2.get_local_bound()
    out["fus_upd_min"], out["fus_upd_max"] = ref2.updated_box()
    ref2.clear_and_inflate()
    out["fus_inflate_bits"] = np.packbits(ref2.inflate.astype(np.uint8))
    ref2.close()

    # ---- combineCost on the optimistic ESDF of the scene ----
    load_state()
    ref.distance[:] = 0.0
    ref.set_modes(1, 0)
    ref.set_local_bound((0, 0, 0), np.array(n) - 1)
    ref.update_esdf3d()
    opt = O.RefBsplineOptimizer(ref, **OPT)
    wg = W.Grid(n, tuple(ref.origin), ref.res)
    tr = W.make_trajectories(wg, inflate, B=8, n_pts=20, seed=13)
    mask = O.NORMAL_PHASE | O.MINTIME
    X, F, G = [], [], []
    prng = np.random.default_rng(4)
    for b in range(8):
        x_init = np.concatenate([tr["ctrl"][b].reshape(-1), [tr["dt"][b]]])
        probes = x_init + prng.normal(size=(3, 61)) * 0.2
        probes[:, -1] = np.abs(probes[:, -1]) + 0.05
        probes[0, -1] = 0.12
        r = opt.evaluate(tr["ctrl"][b], float(tr["dt"][b]), mask, tr["start"][b], tr["end_pos"][b][None, :], probes=probes)
        X.append(np.concatenate([r["x0"][None, :], probes]))
        F.append(r["f"])
        G.append(r["grad"])
    out["bs_keys"], out["bs_vals"] = np.array(list(OPT.keys())), np.array(list(OPT.values()), dtype=np.float64)
    out["bs_mask"] = np.int32(mask)
    out["bs_ctrl"], out["bs_dt"], out["bs_start"], out["bs_end"] = tr["ctrl"], tr["dt"], tr["start"], tr["end_pos"]
    out["bs_pt_dist"] = tr["pt_dist"]
    out["bs_x"], out["bs_f"], out["bs_grad"] = np.stack(X), np.stack(F), np.stack(G)
    opt.close()
    ref.close()
    np.savez_compressed(OUT, **out)
    print("wrote %s (%.0f KB)" % (OUT, os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
