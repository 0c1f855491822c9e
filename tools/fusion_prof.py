"""Time fuelgpu_map_input_point_cloud on synthetic depth frames of the office map (76k points per frame,
the 640x480 / skip 2 image of exploration.launch) and time the oracle on the same frames."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import fuel_b200
from fuel_b200 import workloads as W


def main():
    g, inflate = W.office_map()
    m = fuel_b200.SDFMap(g.n, g.res, g.origin, g.box_min, g.box_max)
    m.setFusionParams()
    frames = []
    for i in range(8):
        cam = np.array([0.2 * i, 0.1 * i, 1.0])
        frames.append((W.depth_frame(g, inflate, cam, 0.7 * i), cam))
    for pts, cam in frames:  # warm-up (also allocates)
        m.inputPointCloud(pts, pts.shape[0], cam)
    torch.cuda.synchronize()
    reps = 20
    t0 = time.perf_counter()
    for r in range(reps):
        for pts, cam in frames:
            m.inputPointCloud(pts, pts.shape[0], cam)
    m.synchronize()
    dt = (time.perf_counter() - t0) / (reps * len(frames))
    npts = np.mean([p.shape[0] for p, _ in frames])
    out = {"frame_ms_e2e": dt * 1e3, "points_per_frame": npts, "Mpoints_per_s": npts / dt / 1e6}
    # depth-image entry: projection on the device, image pinned
    imgs = []
    for i in range(8):
        cam = np.array([0.2 * i, 0.1 * i, 1.0])
        img, R = W.depth_image(g, inflate, cam, 0.7 * i)
        m.pin(img)
        imgs.append((img, R, cam))
    for img, R, cam in imgs:
        m.inputDepthImage(img, R, cam)
    t0 = time.perf_counter()
    for r in range(reps):
        for img, R, cam in imgs:
            m.inputDepthImage(img, R, cam)
    m.synchronize()
    out["depth_frame_ms_e2e"] = (time.perf_counter() - t0) / (reps * len(imgs)) * 1e3
    if "--cpu" in sys.argv:
        import oracle as O
        f = O.Fusion(O.make_grid(g.n, g.res, g.origin, g.box_min, g.box_max), O.fusion_params())
        for pts, cam in frames:
            f.input_point_cloud(pts, cam)
        t0 = time.perf_counter()
        for pts, cam in frames:
            f.input_point_cloud(pts, cam)
        out["cpu_frame_ms"] = (time.perf_counter() - t0) / len(frames) * 1e3
        cp = O.camera_params()
        t0 = time.perf_counter()
        for img, R, cam in imgs:
            f.input_point_cloud(O.process_depth_image(cp, img, R, cam), cam)
        out["cpu_depth_frame_ms"] = (time.perf_counter() - t0) / len(imgs) * 1e3
    print(json.dumps(out))


if __name__ == "__main__":
    main()
