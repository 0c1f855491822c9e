"""Deterministic workloads shared by tests/ and bench.py.

Inputs only; everything is derived from the committed fixtures in tests/golden/ (voxelised
reference .pcd maps, see tools/make_fixtures.py) or from seeded numpy PCG64 generators.
Nothing here touches /root/reference or the oracle.
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

UNKNOWN, FREE, OCCUPIED = 0, 1, 2


class Grid:
    def __init__(self, voxel_num, origin, resolution=0.1, box_min=None, box_max=None):
        self.n = tuple(int(v) for v in voxel_num)
        self.origin = np.asarray(origin, dtype=np.float64)
        self.res = float(resolution)
        self.map_max = self.origin + np.asarray(self.n) * self.res
        self.box_min = self.origin.copy() if box_min is None else np.asarray(box_min, dtype=np.float64)
        self.box_max = self.map_max.copy() if box_max is None else np.asarray(box_max, dtype=np.float64)

    @property
    def nvox(self):
        return self.n[0] * self.n[1] * self.n[2]

    def pos_to_index(self, pos):
        return np.floor((np.asarray(pos, dtype=np.float64) - self.origin) * (1 / self.res)).astype(np.int64)

    def index_to_pos(self, idx):
        return (np.asarray(idx) + 0.5) * self.res + self.origin


def load_occupancy(name):
    """-> (Grid on the fixture's natural extents, inflate int8 [nx,ny,nz])."""
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    n = tuple(int(v) for v in d["voxel_num"])
    inflate = np.zeros(n[0] * n[1] * n[2], dtype=np.int8)
    inflate[d["addr"]] = 1
    return Grid(n, d["origin"], float(d["resolution"])), inflate.reshape(n)


def office_map(interior_box=True):
    """Config 1/2: office.pcd on 200x120x40 @0.1 m, origin (-10,-6,-1).  The exploration box
    is kept strictly inside the map, as the reference's launch files do."""
    g, inflate = load_occupancy("office_200x120x40")
    if interior_box:
        g = Grid(g.n, g.origin, g.res, box_min=(-9.0, -5.0, -0.8), box_max=(9.0, 5.0, 2.0))
    return g, inflate


def office3_map():
    """Config 5: office3.pcd on 200x300x40."""
    g, inflate = load_occupancy("office3_200x300x40")
    return Grid(g.n, g.origin, g.res, box_min=(-9.0, -14.0, -0.8), box_max=(9.0, 14.0, 2.0)), inflate


def pillar_map(variant="V1"):
    """Config 3: pillar.pcd on 512^3 @0.1 m, origin (-25.6,-25.6,-1).
    V0 = file as is (mostly empty cube); V1 = the occupied voxels tiled with periods
    (150, 280, 40) voxels = (15, 28, 4) m so that the cube is filled."""
    g, inflate = load_occupancy("pillar_512")
    if variant == "V1":
        idx = np.argwhere(inflate == 1)
        lo = idx.min(axis=0)
        rel = idx - lo
        out = np.zeros_like(inflate)
        per = (150, 280, 40)
        for ox in range(-4, 5):
            for oy in range(-3, 4):
                for oz in range(-13, 14):
                    sh = rel + lo + np.array([ox * per[0], oy * per[1], oz * per[2]])
                    ok = np.all(sh >= 0, axis=1) & np.all(sh < np.array(g.n), axis=1)
                    s = sh[ok]
                    out[s[:, 0], s[:, 1], s[:, 2]] = 1
        inflate = out
    elif variant != "V0":
        raise ValueError(variant)
    g = Grid(g.n, g.origin, g.res, box_min=g.origin + 0.5, box_max=g.map_max - 0.5)
    return g, inflate


def random_boxes_map(n=(1024, 1024, 256), seed=11, n_boxes=4096, ground_idx=10, origin=None):
    """Config 4: synthetic map, axis-aligned boxes with side U[0.3,3] m + a ground plane."""
    rng = np.random.default_rng(seed)
    res = 0.1
    if origin is None:
        origin = (-n[0] * res / 2, -n[1] * res / 2, -1.0)
    g = Grid(n, origin, res)
    inflate = np.zeros(n, dtype=np.int8)
    side = rng.uniform(0.3, 3.0, size=(n_boxes, 3))
    ctr = rng.uniform(0, 1, size=(n_boxes, 3)) * (np.array(n) * res)
    lo = np.clip(np.floor((ctr - side / 2) / res).astype(np.int64), 0, np.array(n) - 1)
    hi = np.clip(np.floor((ctr + side / 2) / res).astype(np.int64), 0, np.array(n) - 1)
    for a, b in zip(lo, hi):
        inflate[a[0]:b[0] + 1, a[1]:b[1] + 1, a[2]:b[2] + 1] = 1
    if 0 <= ground_idx < n[2]:
        inflate[:, :, ground_idx] = 1
    return g, inflate


def known_region(g, inflate, seed=7, n_poses=64, radius=4.5, z_range=(0.5, 2.0)):
    """Tri-state occupancy for the frontier sweep: FREE = voxels within `radius` metres
    (max_ray_length, algorithm.xml:50) of seeded camera poses in free space and not
    occupied; OCCUPIED = occupied voxels inside the same balls; everything else UNKNOWN."""
    rng = np.random.default_rng(seed)
    n = np.array(g.n)
    tri = np.zeros(g.n, dtype=np.uint8)
    r = int(np.ceil(radius / g.res))
    ax = np.arange(-r, r + 1)
    ball = (ax[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2) * g.res ** 2 <= radius ** 2
    placed = 0
    tries = 0
    lo_b = g.pos_to_index(g.box_min + 0.3)
    hi_b = g.pos_to_index(g.box_max - 0.3)
    zlo = max(lo_b[2], int(g.pos_to_index([0, 0, z_range[0]])[2]))
    zhi = min(hi_b[2], int(g.pos_to_index([0, 0, z_range[1]])[2]))
    while placed < n_poses and tries < 100 * n_poses:
        tries += 1
        c = np.array([rng.integers(lo_b[0], hi_b[0] + 1), rng.integers(lo_b[1], hi_b[1] + 1),
                      rng.integers(zlo, max(zlo, zhi) + 1)])
        if inflate[c[0], c[1], c[2]]:
            continue
        placed += 1
        a0 = np.maximum(c - r, 0)
        a1 = np.minimum(c + r + 1, n)
        b0 = a0 - (c - r)
        b1 = b0 + (a1 - a0)
        sub = tri[a0[0]:a1[0], a0[1]:a1[1], a0[2]:a1[2]]
        sub[ball[b0[0]:b1[0], b0[1]:b1[1], b0[2]:b1[2]]] = FREE
    tri[(tri == FREE) & (inflate == 1)] = OCCUPIED
    return tri


def office_known(g, inflate):
    """The known region used with the office maps (configs 1, 2, 5): 8 camera balls of 2.5 m,
    which leaves ~2/3 of the map unknown and a dozen frontier clusters above cluster_min."""
    return known_region(g, inflate, seed=7, n_poses=8, radius=2.5)


def cubic_boundary_states(ctrl, dt):
    """Uniform cubic B-spline boundary maps (bspline_optimizer.cpp:367-389, 405-427)."""
    q = ctrl
    start = np.stack([(q[0] + 4 * q[1] + q[2]) / 6.0, (q[2] - q[0]) / (2 * dt), (q[0] - 2 * q[1] + q[2]) / (dt * dt)])
    end_pos = (q[-1] + 4 * q[-2] + q[-3]) / 6.0
    return start, end_pos


def make_trajectories(g, inflate, B=1024, n_pts=20, seed=20260922, sigma=0.3, spacing=0.35, max_vel=2.0):
    """Config 2 trajectory batch.  Straight lines between seeded start/goal voxels in the free
    space of the box interior, control points spaced ~ctrl_pt_dist 0.35 m (algorithm.xml:140),
    interior control points perturbed by N(0, sigma) so a realistic fraction lies inside
    dist0; boundary states from the unperturbed spline; time_lb = -1.
    Returns dict(ctrl [B,N,3], dt [B], start [B,3,3], end_pos [B,3], pt_dist [B])."""
    rng = np.random.default_rng(seed)
    lo = g.box_min + 0.15
    hi = g.box_max - 0.15
    length = spacing * (n_pts - 1)
    ctrl = np.zeros((B, n_pts, 3))
    dts = np.zeros(B)
    starts = np.zeros((B, 3, 3))
    ends = np.zeros((B, 3))
    ptd = np.zeros(B)
    b = 0
    while b < B:
        s = rng.uniform(lo, hi)
        si = g.pos_to_index(s)
        if inflate[si[0], si[1], si[2]]:
            continue
        d = rng.normal(size=3)
        d[2] *= 0.15
        d /= np.linalg.norm(d)
        L = length * rng.uniform(0.7, 1.15)
        e = s + d * L
        if np.any(e < lo) or np.any(e > hi):
            continue
        ei = g.pos_to_index(e)
        if inflate[ei[0], ei[1], ei[2]]:
            continue
        t = np.linspace(0.0, 1.0, n_pts)[:, None]
        line = s[None, :] * (1 - t) + e[None, :] * t
        dt = (L / (n_pts - 1)) / (max_vel * rng.uniform(0.55, 0.95))
        st, en = cubic_boundary_states(line, dt)
        pert = line.copy()
        pert[3:n_pts - 3] += rng.normal(scale=sigma, size=(n_pts - 6, 3))
        pert = np.minimum(np.maximum(pert, g.box_min + 0.1), g.box_max - 0.1)  # optimize() clamp :196-204
        ctrl[b] = pert
        dts[b] = dt
        starts[b] = st
        ends[b] = en
        # pt_dist_ is frozen from the initial control points (:136-140)
        seg = np.sqrt(np.sum((pert[1:] - pert[:-1]) ** 2, axis=1))
        acc = 0.0
        for v in seg:
            acc += float(v)
        ptd[b] = acc / float(n_pts)
        b += 1
    return dict(ctrl=ctrl, dt=dts, start=starts, end_pos=ends, pt_dist=ptd)


def pack_x(ctrl, dt, mintime=True):
    B = ctrl.shape[0]
    x = ctrl.reshape(B, -1)
    if mintime:
        x = np.concatenate([x, dt[:, None]], axis=1)
    return np.ascontiguousarray(x, dtype=np.float64)


def _camera_rotation(yaw, pitch=0.0):
    """camera -> world rotation: camera z = forward, x = right, y = down"""
    cyw, syw, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    fwd = np.array([cyw * cp, syw * cp, sp])
    right = np.array([syw, -cyw, 0.0])
    down = np.cross(fwd, right)
    return np.stack([right, down, fwd], axis=1)


def depth_image(g, inflate, cam_pos, yaw, pitch=0.0, width=640, height=480, fx=387.229248046875, fy=387.229248046875,
                cx=321.04638671875, cy=243.44969177246094, margin=2, skip=2, maxdist=5.0, mindist=0.2):
    """Synthetic sensor frame: a pinhole depth camera (intrinsics of exploration.launch:38-41) at cam_pos looking
    along `yaw`, ray-marched against the ground-truth occupancy `inflate`; uint16 millimetres, 0 = no return.  Rays
    are marched for the pixels MapROS samples (margin + k*skip) and replicated to their neighbours.
    -> (image uint16 [height,width], R [3,3] camera->world)"""
    cam_pos = np.asarray(cam_pos, dtype=np.float64)
    off = margin % skip
    us = np.arange(off, width, skip)
    vs = np.arange(off, height, skip)
    U, V = np.meshgrid(us, vs)
    dirs_c = np.stack([(U - cx) / fx, (V - cy) / fy, np.ones_like(U, dtype=np.float64)], axis=-1).reshape(-1, 3)
    R = _camera_rotation(yaw, pitch)
    dirs_w = dirs_c @ R.T
    n = np.asarray(g.n)
    depth = np.zeros(dirs_c.shape[0])  # 0 = no return
    alive = np.ones(dirs_c.shape[0], dtype=bool)
    flat = np.ascontiguousarray(inflate).reshape(-1)
    for t in np.arange(mindist, maxdist + 0.3, 0.04):
        idx_alive = np.nonzero(alive)[0]
        if idx_alive.size == 0:
            break
        p = cam_pos + dirs_w[idx_alive] * t
        vi = np.floor((p - np.asarray(g.origin)) / g.res).astype(np.int64)
        inside = np.all((vi >= 0) & (vi < n), axis=1)
        adr = (np.clip(vi[:, 0], 0, n[0] - 1) * n[1] + np.clip(vi[:, 1], 0, n[1] - 1)) * n[2] + np.clip(vi[:, 2], 0, n[2] - 1)
        hit = inside & (flat[adr] != 0)
        depth[idx_alive[hit]] = t
        alive[idx_alive[hit]] = False
    low = np.round(depth * 1000.0).astype(np.uint16).reshape(len(vs), len(us))
    vi = np.clip((np.arange(height) - off) // skip, 0, len(vs) - 1)
    ui = np.clip((np.arange(width) - off) // skip, 0, len(us) - 1)
    return np.ascontiguousarray(low[vi][:, ui]), R


def depth_frame(g, inflate, cam_pos, yaw, pitch=0.0, width=640, height=480, fx=387.229248046875, fy=387.229248046875,
                cx=321.04638671875, cy=243.44969177246094, margin=2, skip=2, maxdist=5.0, mindist=0.2):
    """depth_image() projected to world points the way MapROS::proessDepthImage (plan_env/src/map_ros.cpp:176-215) hands
    them to inputPointCloud (no-return pixels at depth_filter_maxdist).  Input generation only (numpy); the parity-checked
    projection is fuelgpu_map_input_depth_image vs the oracle.  -> float32 [n,3] world points."""
    img, R = depth_image(g, inflate, cam_pos, yaw, pitch, width, height, fx, fy, cx, cy, margin, skip, maxdist, mindist)
    us = np.arange(margin, width - margin, skip)
    vs = np.arange(margin, height - margin, skip)
    U, V = np.meshgrid(us, vs)
    d16 = img[V, U].reshape(-1)
    d = d16 * (1.0 / 1000.0)
    # the reference's "no return" test looks at the pixel `skip` further along the row buffer (map_ros.cpp:190-198)
    flat = img.reshape(-1)
    nxt_at = (V * width + U + skip).reshape(-1)
    nxt = np.where(nxt_at < flat.size, flat[np.minimum(nxt_at, flat.size - 1)], 0)
    far = (nxt == 0) | (d > maxdist)
    keep = far | (d >= mindist)
    d = np.where(far, maxdist, d)
    dirs_c = np.stack([(U - cx) / fx, (V - cy) / fy, np.ones_like(U, dtype=np.float64)], axis=-1).reshape(-1, 3)
    pts = (dirs_c * d[:, None]) @ R.T + np.asarray(cam_pos, dtype=np.float64)
    return np.ascontiguousarray(pts[keep], dtype=np.float32)
