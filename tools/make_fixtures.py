#!/usr/bin/env python
"""Generate the committed input fixtures under tests/golden/ from the reference's .pcd maps.

Runs ONLY in the build container (reads /root/reference, which does not exist on the GPU
box).  The fixtures are inputs, not expected outputs: the reference ships no golden vectors
for this path.  What is stored is the result of the offline map-build
recipe of plan_manage/test/compare_topo.cpp:122-133 -- resetBuffer(); setOccupied(pt) for
every cloud point -- i.e. the sorted unique voxel addresses with
occupancy_buffer_inflate_ == 1 on the BASELINE grids.

PCD parsing follows uav_simulator/map_generator/src/map_publisher.cpp:15-64 (pcl::io::loadPCDFile
of an ASCII x,y,z float32 cloud); the publisher's ground plane is NOT added (stated).
"""
import os
import sys

import numpy as np

REF = "/root/reference/uav_simulator/map_generator/resource"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (pcd file, voxel_num, origin)   resolution 0.1 everywhere (algorithm.xml:33)
GRIDS = {
    "office_200x120x40": ("office.pcd", (200, 120, 40), (-10.0, -6.0, -1.0)),
    "office3_200x300x40": ("office3.pcd", (200, 300, 40), (-10.0, -15.0, -1.0)),
    "pillar_512": ("pillar.pcd", (512, 512, 512), (-25.6, -25.6, -1.0)),
}
RES = 0.1


def load_pcd_ascii(path):
    with open(path, "r") as f:
        n = None
        for line in f:
            if line.startswith("POINTS"):
                n = int(line.split()[1])
            if line.startswith("DATA"):
                assert line.split()[1] == "ascii"
                break
        pts = np.loadtxt(f, dtype=np.float32)  # PointXYZ is float32
    assert pts.shape == (n, 3)
    return pts


def voxelise(pts, n, origin):
    """setOccupied (sdf_map.h:210-215): isInMap(pos) with the 1e-4 margin, then posToIndex."""
    pos = pts.astype(np.float64)  # obs_pt(k) = pt.x (float -> double)
    origin = np.asarray(origin, dtype=np.float64)
    n = np.asarray(n)
    mx = origin + n * RES
    ok = np.all(pos >= origin + 1e-4, axis=1) & np.all(pos <= mx - 1e-4, axis=1)
    idx = np.floor((pos[ok] - origin) * (1 / RES)).astype(np.int64)
    addr = (idx[:, 0] * n[1] + idx[:, 1]) * n[2] + idx[:, 2]
    return np.unique(addr).astype(np.int32), int(ok.sum())


def main():
    if not os.path.isdir(REF):
        sys.exit("reference not mounted; fixtures can only be regenerated in the build container")
    os.makedirs(OUT, exist_ok=True)
    for name, (pcd, n, origin) in GRIDS.items():
        pts = load_pcd_ascii(os.path.join(REF, pcd))
        addr, kept = voxelise(pts, n, origin)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, addr=addr, voxel_num=np.asarray(n, dtype=np.int32),
                            origin=np.asarray(origin, dtype=np.float64), resolution=np.float64(RES),
                            n_points=np.int64(pts.shape[0]), n_points_in_map=np.int64(kept),
                            source=np.bytes_(pcd))
        print("%s: %d points, %d in map, %d occupied voxels -> %s (%d bytes)" %
              (pcd, pts.shape[0], kept, addr.size, path, os.path.getsize(path)))


if __name__ == "__main__":
    main()
