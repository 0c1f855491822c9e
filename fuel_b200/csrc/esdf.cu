// esdf.cu -- ESDF update entry (SDFMap::updateESDF3d, plan_env/src/sdf_map.cpp:152-241: unsigned / signed,
// optimistic / unknown-as-obstacle, restricted to [local_bound_min_, local_bound_max_]), the trilinear sampler
// (SDFMap::getDistWithGrad, :497-536) and obstacle inflation (clearAndInflateLocalMap, :364-472).  The distance
// transform itself lives in esdf_tile.cu.  "No site in the box" is +inf in the result (the reference ends with
// resolution*sqrt(DBL_MAX) there).
//
// Layout: address = (x*ny + y)*nz + z, z fastest (sdf_map.h:145-147).
#include "common.cuh"

namespace {

constexpr int INF_I = FUELGPU_EDT_INF_I;

struct Box {
  int lo[3], hi[3];  // inclusive
};

// site predicate of the z sweep.  mode 0: optimistic (sdf_map.cpp:156-166) inflate==1;
// mode 1: non-optimistic (:167-181) inflate==1 || unknown; mode 2: negative field (:203-214)
// inflate==0.
__device__ __forceinline__ bool is_site(uint8_t o, int mode) {
  const bool infl = (o & 4) != 0;
  if (mode == 0) return infl;
  if (mode == 1) return infl || ((o & 3) == FUELGPU_UNKNOWN);
  return !infl;
}

// merge of the negative field, sdf_map.cpp:232-239
__global__ void signed_merge_kernel(float* __restrict__ dist, const float* __restrict__ neg, int ny,
                                    int nz, Box b, float res) {
  const int nzb = b.hi[2] - b.lo[2] + 1, nyb = b.hi[1] - b.lo[1] + 1, nxb = b.hi[0] - b.lo[0] + 1;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)nzb * nyb * nxb) return;
  const int z = b.lo[2] + (int)(t % nzb);
  const int y = b.lo[1] + (int)((t / nzb) % nyb);
  const int x = b.lo[0] + (int)(t / ((int64_t)nzb * nyb));
  const int64_t a = ((int64_t)x * ny + y) * nz + z;
  const float ng = neg[a];
  if (ng > 0.0f) dist[a] += (-ng + res);
}

__global__ void sample_kernel(Geom g, const float* __restrict__ dist, int64_t n,
                              const double* __restrict__ pos, double* __restrict__ d,
                              double* __restrict__ grad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double p[3] = { pos[3 * i], pos[3 * i + 1], pos[3 * i + 2] };
  double gr[3];
  d[i] = dev_dist_with_grad(g, dist, p, gr);
  grad[3 * i] = gr[0];
  grad[3 * i + 1] = gr[1];
  grad[3 * i + 2] = gr[2];
}

}  // namespace

int esdf_update_impl(FuelMap* m, const int bmin[3], const int bmax[3], int flags) {
  Box b;
  for (int i = 0; i < 3; ++i) {
    b.lo[i] = bmin[i];
    b.hi[i] = bmax[i];
  }
  const int mode = (flags & FUELGPU_ESDF_OPTIMISTIC) ? 0 : 1;
  const float res = (float)m->g.res;
  int rc = esdf_tile_transform(m, b.lo, b.hi, mode, m->dist);
  if (rc) return rc;
  if (flags & FUELGPU_ESDF_SIGNED) {
    if (!m->dist_neg) FUEL_CUDA(m, cudaMalloc(&m->dist_neg, sizeof(float) * m->nvox));
    rc = esdf_tile_transform(m, b.lo, b.hi, 2, m->dist_neg);
    if (rc) return rc;
    const int64_t nb = (int64_t)(b.hi[0] - b.lo[0] + 1) * (b.hi[1] - b.lo[1] + 1) * (b.hi[2] - b.lo[2] + 1);
    signed_merge_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, m->stream>>>(m->dist, m->dist_neg,
                                                                          m->g.ny, m->g.nz, b, res);
    FUEL_LAUNCHES(m, 1);
  }
  FUEL_CUDA(m, cudaGetLastError());
  return 0;
}

// ---- obstacle inflation: SDFMap::clearAndInflateLocalMap (sdf_map.cpp:364-472) -----------------
namespace {
__global__ void inflate_clear_kernel(uint8_t* __restrict__ occ, int ny, int nz, Box b) {
  const int nzb = b.hi[2] - b.lo[2] + 1, nyb = b.hi[1] - b.lo[1] + 1, nxb = b.hi[0] - b.lo[0] + 1;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)nzb * nyb * nxb) return;
  const int z = b.lo[2] + (int)(t % nzb);
  const int y = b.lo[1] + (int)((t / nzb) % nyb);
  const int x = b.lo[0] + (int)(t / ((int64_t)nzb * nyb));
  const int64_t a = ((int64_t)x * ny + y) * nz + z;
  occ[a] &= (uint8_t)~4u;  // :440-444
}
// one thread per voxel of the box; an occupied voxel stamps its (2s+1)^3 neighbourhood.  The
// reference checks only the LINEAR address of a stamp cell (:452-458), so stamps wrap across rows
// at the map faces; the same arithmetic is used here.  All writers set the same bit of a byte.
__global__ void inflate_stamp_kernel(uint8_t* __restrict__ occ, int ny, int nz, int64_t nvox, Box b, int step) {
  const int nzb = b.hi[2] - b.lo[2] + 1, nyb = b.hi[1] - b.lo[1] + 1, nxb = b.hi[0] - b.lo[0] + 1;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)nzb * nyb * nxb) return;
  const int z = b.lo[2] + (int)(t % nzb);
  const int y = b.lo[1] + (int)((t / nzb) % nyb);
  const int x = b.lo[0] + (int)(t / ((int64_t)nzb * nyb));
  if ((occ[((int64_t)x * ny + y) * nz + z] & 3) != FUELGPU_OCCUPIED) return;
  for (int dx = -step; dx <= step; ++dx)
    for (int dy = -step; dy <= step; ++dy)
      for (int dz = -step; dz <= step; ++dz) {
        const int64_t a = ((int64_t)(x + dx) * ny + (y + dy)) * nz + (z + dz);
        if (a >= 0 && a < nvox) {
          const uint8_t o = occ[a];
          if (!(o & 4)) occ[a] = o | 4;
        }
      }
}
__global__ void ceiling_kernel(uint8_t* __restrict__ occ, double* __restrict__ logodds, double clamp_max_log, int ny,
                               int nz, Box b, int ceil_id) {
  const int nyb = b.hi[1] - b.lo[1] + 1, nxb = b.hi[0] - b.lo[0] + 1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nyb * nxb) return;
  const int y = b.lo[1] + t % nyb, x = b.lo[0] + t / nyb;
  const int64_t a = ((int64_t)x * ny + y) * nz + ceil_id;
  occ[a] = (uint8_t)((occ[a] & ~3u) | FUELGPU_OCCUPIED);  // occupancy_buffer_ = clamp_max_log_ (:463-470)
  // the fused log-odds volume holds the same value, so that the ceiling survives the misses of later frames
  if (logodds) logodds[a] = clamp_max_log;
}
}  // namespace

int map_inflate_impl(FuelMap* m, const int bmin[3], const int bmax[3], int step, int ceil_id) {
  Box b;
  for (int i = 0; i < 3; ++i) {
    b.lo[i] = bmin[i];
    b.hi[i] = bmax[i];
  }
  const int64_t nb = (int64_t)(b.hi[0] - b.lo[0] + 1) * (b.hi[1] - b.lo[1] + 1) * (b.hi[2] - b.lo[2] + 1);
  inflate_clear_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, m->stream>>>(m->occ, m->g.ny, m->g.nz, b);
  inflate_stamp_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, m->stream>>>(m->occ, m->g.ny, m->g.nz, m->nvox, b, step);
  FUEL_LAUNCHES(m, 2);
  if (ceil_id >= 0 && ceil_id < m->g.nz) {
    const int n2 = (b.hi[0] - b.lo[0] + 1) * (b.hi[1] - b.lo[1] + 1);
    double cmax = 0.0;
    double* lo = fusion_logodds_ptr(m, &cmax);
    ceiling_kernel<<<(n2 + 255) / 256, 256, 0, m->stream>>>(m->occ, lo, cmax, m->g.ny, m->g.nz, b, ceil_id);
    FUEL_LAUNCHES(m, 1);
  }
  FUEL_CUDA(m, cudaGetLastError());
  return 0;
}

int esdf_sample_impl(FuelMap* m, int64_t n, const double* pos, double* d, double* grad) {
  if (n <= 0) return 0;
  sample_kernel<<<(unsigned)((n + 127) / 128), 128, 0, m->stream>>>(m->g, m->dist, n, pos, d, grad);
  FUEL_LAUNCHES(m, 1);
  FUEL_CUDA(m, cudaGetLastError());
  return 0;
}
