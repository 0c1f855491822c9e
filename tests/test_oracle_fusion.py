"""Oracle self-checks for the occupancy fusion restatement (oracle/fuel_oracle_fusion.c;
SDFMap::inputPointCloud, plan_env/src/sdf_map.cpp:259-345) against independent ground truths:
an analytic supercover of the ray, hand-computed log-odds sequences, and box arithmetic."""
import numpy as np
import pytest

import oracle as O
from fuel_b200 import workloads as W


def logit(p):
    return np.log(p / (1 - p))


def make(n=(60, 50, 30), origin=(-3.0, -2.5, -0.5), res=0.1, **kw):
    g = O.make_grid(n, res, origin)
    return g, O.Fusion(g, O.fusion_params(**kw))


def test_single_ray_marks_segment_and_end():
    g, f = make()
    cam = np.array([0.03, 0.04, 1.02])
    pt = np.array([[1.57, 0.83, 1.46]], dtype=np.float32)
    lo, hi = f.input_point_cloud(pt, cam)
    L = f.logodds.reshape(60, 50, 30)
    cmin = logit(0.12)
    touched = np.argwhere(L > cmin - 1e-3)
    end = O.pos_to_index(g, pt[0].astype(np.float64))
    # the end voxel was unknown -> min_occupancy_log + hit, clamped to clamp_max
    assert np.isclose(L[tuple(end)], min(logit(0.80) + logit(0.65), logit(0.90)))
    # every other touched voxel: unknown -> min_occupancy_log + miss
    others = [tuple(t) for t in touched if tuple(t) != tuple(end)]
    assert all(np.isclose(L[t], logit(0.80) + logit(0.35)) for t in others)
    # geometry: all lie within one voxel diagonal of the segment and form a 6-connected chain whose
    # length is the Manhattan distance between the end voxels (minus both ends: the camera voxel is
    # never reported, raycast.cpp:374-381, the end voxel is the hit)
    p0, p1 = pt[0].astype(np.float64), cam
    d = p1 - p0
    for t in others:
        c = (np.array(t) + 0.5) * 0.1 + np.array([-3.0, -2.5, -0.5])
        s = np.clip(np.dot(c - p0, d) / np.dot(d, d), 0, 1)
        assert np.linalg.norm(c - (p0 + s * d)) < 0.1 * np.sqrt(3)
    cam_idx = O.pos_to_index(g, cam)
    assert len(others) == int(np.abs(cam_idx - end).sum()) - 1
    # local bound = box of {camera, point} inflated by 0.5 m in x,y (:313-318)
    exp_lo = O.pos_to_index(g, np.minimum(cam, p0) - np.array([0.5, 0.5, 0.0]))
    exp_hi = O.pos_to_index(g, np.maximum(cam, p0) + np.array([0.5, 0.5, 0.0]))
    assert np.array_equal(lo, exp_lo) and np.array_equal(hi, exp_hi)


def test_logodds_sequence_and_clamps():
    g, f = make()
    cam = np.array([0.0, 0.0, 1.0])
    pt = np.array([[1.0, 0.0, 1.0]], dtype=np.float32)
    end = tuple(O.pos_to_index(g, pt[0].astype(np.float64)))
    mid = tuple(O.pos_to_index(g, np.array([0.5, 0.0, 1.0])))
    exp_end, exp_mid = logit(0.80), logit(0.80)
    for it in range(12):
        f.input_point_cloud(pt, cam)
        exp_end = min(max(exp_end + logit(0.65), logit(0.12)), logit(0.90))
        exp_mid = min(max(exp_mid + logit(0.35), logit(0.12)), logit(0.90))
        L = f.logodds.reshape(60, 50, 30)
        assert L[end] == exp_end and L[mid] == exp_mid
    assert L[end] == logit(0.90) and L[mid] == logit(0.12)
    tri = f.tristate().reshape(60, 50, 30)
    assert tri[end] == W.OCCUPIED and tri[mid] == W.FREE and tri[0, 0, 0] == W.UNKNOWN


def test_hit_beats_miss_in_same_voxel():
    """count_hit >= count_miss (:329): a voxel that is the end point of one ray and traversed by another is a hit."""
    g, f = make()
    cam = np.array([0.0, 0.0, 1.0])
    pts = np.array([[1.0, 0.0, 1.0], [2.0, 0.0, 1.0]], dtype=np.float32)  # the 2nd ray passes through the 1st end voxel
    f.input_point_cloud(pts, cam)
    L = f.logodds.reshape(60, 50, 30)
    e1 = tuple(O.pos_to_index(g, np.array([1.0, 0.0, 1.0])))
    assert np.isclose(L[e1], min(logit(0.80) + logit(0.65), logit(0.90)))


def test_far_and_outside_points_are_free_endpoints():
    g, f = make(max_ray_length=1.5)
    cam = np.array([0.0, 0.0, 1.0])
    pts = np.array([[2.5, 0.0, 1.0],      # in map, beyond max_ray_length -> clipped to 1.5 m, flagged free
                    [0.0, 30.0, 1.0],     # outside the map -> closetPointInMap, then clipped
                    [0.0, 0.0, -5.0]],    # outside below -> clipped point has z < 0.2 -> skipped
                   dtype=np.float32)
    f.input_point_cloud(pts, cam)
    L = f.logodds.reshape(60, 50, 30)
    assert (L > logit(0.80)).sum() == 0  # nothing occupied
    e = tuple(O.pos_to_index(g, np.array([1.5 - 1e-9, 0.0, 1.0])))
    assert np.isclose(L[e], logit(0.80) + logit(0.35))
    e = tuple(O.pos_to_index(g, np.array([0.0, 1.5 - 1e-9, 1.0])))
    assert np.isclose(L[e], logit(0.80) + logit(0.35))
    below = L[:, :, : O.pos_to_index(g, np.array([0.0, 0.0, 0.2]))[2]]
    assert (below > logit(0.12) - 1e-3).sum() == 0
    # updated box accumulates until reset (:267-271, :321-324, :491-495)
    lo, hi = f.updated_box(reset=True)
    assert np.allclose(lo, [0.0, 0.0, 1.0]) and np.allclose(hi, [1.5, 1.5, 1.0])
    f.input_point_cloud(np.array([[-1.0, 0.0, 1.0]], dtype=np.float32), np.array([0.0, 0.0, 1.2]))
    lo, hi = f.updated_box()
    assert np.allclose(lo, [-1.0, 0.0, 1.0]) and np.allclose(hi, [0.0, 0.0, 1.2])


def test_empty_cloud_is_a_no_op():
    g, f = make()
    before = f.logodds.copy()
    f.input_point_cloud(np.zeros((0, 3), np.float32), np.array([0.0, 0.0, 1.0]))
    assert np.array_equal(before, f.logodds) and f.st.raycast_num == 0


def test_depth_frames_reconstruct_the_scene():
    """Fusing a few synthetic depth frames of the office map: occupied voxels are (almost) all true obstacles,
    and the space between camera and surfaces becomes FREE."""
    g, inflate = W.office_map()
    og = O.make_grid(g.n, g.res, g.origin, g.box_min, g.box_max)
    f = O.Fusion(og, O.fusion_params())
    cam = np.array([0.0, 0.0, 1.0])
    for yaw in (0.0, 1.6, 3.1, 4.7):
        pts = W.depth_frame(g, inflate, cam, yaw)
        assert pts.shape[0] > 50000
        f.input_point_cloud(pts, cam)
    tri = f.tristate().reshape(g.n)
    occ = tri == W.OCCUPIED
    assert occ.sum() > 500
    # an OCCUPIED voxel is a ground-truth obstacle or touches one (depth quantisation / ray-march step)
    from scipy.ndimage import binary_dilation
    near = binary_dilation(inflate != 0, iterations=1, structure=np.ones((3, 3, 3), bool))
    assert (occ & ~near).sum() <= 0.01 * occ.sum()
    assert (tri == W.FREE).sum() > 20 * occ.sum()
    ci = tuple(O.pos_to_index(og, cam + np.array([0.3, 0.0, 0.0])))
    assert tri[ci] == W.FREE


def test_depth_projection_matches_numpy_and_keeps_the_next_pixel_quirk():
    """orc_process_depth_image (map_ros.cpp:176-215) against an independent numpy pinhole projection."""
    g, inflate = W.office_map()
    cam = np.array([0.5, -0.3, 1.1])
    img, R = W.depth_image(g, inflate, cam, 0.9, pitch=-0.1)
    cp = O.camera_params()
    pts = O.process_depth_image(cp, img, R, cam)
    ref = W.depth_frame(g, inflate, cam, 0.9, pitch=-0.1)
    assert pts.shape == ref.shape and 70000 < pts.shape[0] <= 238 * 318
    assert np.max(np.abs(pts - ref)) < 2e-6
    # quirk: the zero test looks at pixel u+skip, the depth comes from pixel u
    img2 = np.full((480, 640), 1500, np.uint16)
    img2[100, 204] = 0            # makes the point of pixel (100, 202) "no return" (depth 5.0), not its own
    p2 = O.process_depth_image(cp, img2, np.eye(3), np.zeros(3))
    k = 49 * 318 + 100            # v = 100 -> row 49, u = 202 -> column 100
    assert np.isclose(p2[k, 2], 5.0) and np.isclose(p2[k - 1, 2], 1.5)
    # ... and pixel (100, 204) itself (depth 0, right neighbour valid) is dropped as "too close"
    assert p2.shape[0] == 238 * 318 - 1 and np.isclose(p2[k + 1, 2], 1.5)
    assert np.isclose(p2[k + 1, 0], (206 - cp.cx) * 1.5 / cp.fx)
    # depth below depth_filter_mindist is skipped (count drops by one), beyond maxdist is clamped
    img3 = np.full((480, 640), 1500, np.uint16)
    img3[10, 10] = 100
    img3[10, 12] = 9000
    p3 = O.process_depth_image(cp, img3, np.eye(3), np.zeros(3))
    assert p3.shape[0] == 238 * 318 - 1 and np.isclose(p3[:, 2].max(), 5.0)
