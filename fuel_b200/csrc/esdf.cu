// esdf.cu -- exact Euclidean distance transform of the occupancy grid on sm_100a.
//
// Replaces SDFMap::updateESDF3d / fillESDF (plan_env/src/sdf_map.cpp:116-241) and
// SDFMap::getDistWithGrad (:497-536).  The reference runs three 1-D lower-envelope sweeps
// (z, y, x) in fp64 with DBL_MAX as "no site".  Every finite intermediate is an integer
// (squared voxel distance <= 3*(n-1)^2), so the device keeps the transform in exact int32
// arithmetic and only the last pass converts: dist = resolution * sqrt(d2) in fp32.
// "No site on this line / in this box" is FUELGPU_EDT_INF between passes and +inf in the
// result (the reference ends with resolution*sqrt(DBL_MAX) there).
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

// ---------------------------------------------------------------------------------------
// z sweep.  out = distance (in voxels, uint16) to the nearest site on the (x,y) line, 0xFFFF
// if the line has none.  The 1-D pass needs no envelope: nearest set bit on each side.
// ---------------------------------------------------------------------------------------
constexpr unsigned short INF16 = 0xffffu;
constexpr int BIG = 1 << 20;

// generic box variant: one warp per line, <= 1024 voxels, any z range.  The line's site bits are
// gathered with warp ballots (lane j keeps the 32-bit mask of chunk j), then every voxel finds
// its nearest set bit on both sides with clz/ffs.
__global__ void __launch_bounds__(256) zsweep_warp_kernel(const uint8_t* __restrict__ occ,
                                                          uint16_t* __restrict__ out, int ny, int nz,
                                                          Box b, int mode, int nlines) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= nlines) return;
  const int nyb = b.hi[1] - b.lo[1] + 1;
  const int x = b.lo[0] + warp / nyb;
  const int y = b.lo[1] + warp % nyb;
  const int z0 = b.lo[2];
  const int n = b.hi[2] - b.lo[2] + 1;
  const int64_t base = ((int64_t)x * ny + y) * nz + z0;
  const int nchunks = (n + 31) >> 5;

  unsigned mymask = 0;
  for (int i = 0; i < nchunks; ++i) {
    const int p = (i << 5) + lane;
    bool s = false;
    if (p < n) s = is_site(__ldg(occ + base + p), mode);
    const unsigned m = __ballot_sync(0xffffffffu, s);
    if (lane == i) mymask = m;
  }
  const unsigned nzb = __ballot_sync(0xffffffffu, mymask != 0);
  const int mylast = mymask ? 31 - __clz(mymask) : 0;
  const int myfirst = mymask ? __ffs(mymask) - 1 : 0;

  for (int i = 0; i < nchunks; ++i) {
    const unsigned m = __shfl_sync(0xffffffffu, mymask, i);
    const int pos = (i << 5) + lane;
    const unsigned prev = nzb & ((1u << i) - 1u);
    const int pc = prev ? 31 - __clz(prev) : 0;
    const int plast = __shfl_sync(0xffffffffu, mylast, pc);
    const unsigned next = nzb & ~((2u << i) - 1u);
    const int nc = next ? __ffs(next) - 1 : 0;
    const int nfirst = __shfl_sync(0xffffffffu, myfirst, nc);

    int d = BIG;
    const unsigned ml = m & (0xffffffffu >> (31 - lane));
    if (ml)
      d = lane - (31 - __clz(ml));
    else if (prev)
      d = pos - ((pc << 5) + plast);
    const unsigned mr = m & (0xffffffffu << lane);
    int dr = BIG;
    if (mr)
      dr = (__ffs(mr) - 1) - lane;
    else if (next)
      dr = ((nc << 5) + nfirst) - pos;
    d = min(d, dr);
    if (pos < n) out[base + pos] = (d >= BIG) ? INF16 : (uint16_t)d;
  }
}

// full-line variant (box spans the whole z axis, nz % VPL == 0, nz <= 32*VPL): one warp per
// line, lane L owns VPL consecutive voxels, loaded with ONE vector load per lane (the whole line
// is one coalesced request) and written with vector stores.  Nearest site outside the lane's
// own voxels comes from a ballot over "lane has a site" + two shuffles.
template <int VPL>
struct VecT;
template <>
struct VecT<16> { using T = uint4; };
template <>
struct VecT<8> { using T = uint2; };
template <>
struct VecT<4> { using T = uint32_t; };

template <int VPL>
__global__ void __launch_bounds__(256) zsweep_vec_kernel(const uint8_t* __restrict__ occ,
                                                         uint16_t* __restrict__ out, int ny, int nz,
                                                         Box b, int mode, int nlines) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= nlines) return;
  const int nyb = b.hi[1] - b.lo[1] + 1;
  const int x = b.lo[0] + warp / nyb;
  const int y = b.lo[1] + warp % nyb;
  const int64_t base = ((int64_t)x * ny + y) * nz;
  const bool have = lane * VPL < nz;
  using V = typename VecT<VPL>::T;
  union {
    V v;
    uint8_t b8[VPL];
  } u;
  unsigned mask = 0;
  if (have) {
    u.v = __ldg(reinterpret_cast<const V*>(occ + base + lane * VPL));
#pragma unroll
    for (int i = 0; i < VPL; ++i) mask |= is_site(u.b8[i], mode) ? (1u << i) : 0u;
  }
  const unsigned ball = __ballot_sync(0xffffffffu, mask != 0);
  const int mylast = mask ? 31 - __clz(mask) : 0;
  const int myfirst = mask ? __ffs(mask) - 1 : 0;
  const unsigned pm = ball & ((1u << lane) - 1u);
  const int pl = pm ? 31 - __clz(pm) : 0;
  const int plast = __shfl_sync(0xffffffffu, mylast, pl);
  const unsigned nm = ball & ~((2u << lane) - 1u);
  const int nl = nm ? __ffs(nm) - 1 : 0;
  const int nfirst = __shfl_sync(0xffffffffu, myfirst, nl);
  if (!have) return;
  // distance from "one before my first voxel" to the nearest site on the left, and from "one
  // past my last voxel" to the nearest on the right
  int dl = pm ? (VPL * lane - 1) - (VPL * pl + plast) : BIG;
  int dr = nm ? (VPL * nl + nfirst) - (VPL * lane + VPL) : BIG;
  int dleft[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    dl = ((mask >> i) & 1u) ? 0 : dl + 1;
    dleft[i] = dl;
  }
  union {
    uint4 q[VPL / 8 > 0 ? VPL / 8 : 1];
    uint16_t h[VPL < 8 ? 8 : VPL];
  } o;
#pragma unroll
  for (int i = VPL - 1; i >= 0; --i) {
    dr = ((mask >> i) & 1u) ? 0 : dr + 1;
    const int d = min(dleft[i], dr);
    o.h[i] = d >= BIG ? INF16 : (uint16_t)d;
  }
  uint16_t* dst = out + base + lane * VPL;
  if (VPL >= 8) {
#pragma unroll
    for (int j = 0; j < VPL / 8; ++j) reinterpret_cast<uint4*>(dst)[j] = o.q[j];
  } else {
    *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<uint2*>(o.h);
  }
}

// ---------------------------------------------------------------------------------------
// Envelope sweep along y or x: one thread per line, lanes along z so that every global
// access of a warp is a contiguous run.  Felzenszwalb-Huttenlocher lower envelope restated
// in exact integers: parabola of site v has height h(v) = f(v) + v^2; the abscissa where w
// overtakes u is (h(w)-h(u)) / (2(w-u)); all comparisons are cross-multiplied, no division.
// Loads of the line are issued U at a time before any of them is consumed, so each thread
// keeps U requests in flight (a single dependent load per warp is latency-, not
// bandwidth-bound).  The hull stack (slot k of a line at the line's k-th element) lives in
// `stk`; the x sweep aliases it on its own input, whose slots <= q are dead once read.
// The two topmost entries are cached in registers.
// Stack entry = v (10 bits) | h (22 bits): requires n <= 1024 per axis.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_vh(int v, int h) { return ((uint32_t)v << 22) + (uint32_t)h; }  // one LEA
__device__ __forceinline__ int unpack_v(uint32_t e) { return (int)(e >> 22); }
__device__ __forceinline__ int unpack_h(uint32_t e) { return (int)(e & 0x3fffffu); }

struct LineMap {
  int n;            // samples per line
  int64_t stride;   // elements between consecutive samples
  int nz_run;       // lines along z per row (fastest)
  int n_outer;      // rows of lines
  int64_t outer_stride;  // elements between rows of lines
  int64_t base;     // element offset of line (0,0), sample 0
};


__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));  // <= 1 ulp-ish; the bar is 1e-4 relative
  return r;
}

// IN16: input is the uint16 1-D distance of the z sweep (squared on load); else int32 squared.
// All addressing is by running byte pointers (one 64-bit add per step) -- the kernel is
// issue-bound, so index*stride multiplies in the inner loops are what it cannot afford.
// MINB: minimum resident CTAs per SM.  16 caps the kernel at 32 registers (a few spills) and wins
// ~10 % on large volumes by occupancy; small maps are latency-bound per line and prefer no cap.
template <bool IN16, bool FINAL, int ENV_U, int MINB = 1>
__global__ void __launch_bounds__(128, MINB) envelope_kernel(const void* inv, void* outv, uint32_t* stk,
                                                       LineMap lm, float res) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)lm.nz_run * lm.n_outer) return;
  const int zi = (int)(t % lm.nz_run);
  const int oi = (int)(t / lm.nz_run);
  const int64_t off0 = lm.base + (int64_t)oi * lm.outer_stride + zi;
  const int n = lm.n;
  constexpr int ISZ = IN16 ? 2 : 4;
  const int64_t SBi = lm.stride * ISZ;  // byte strides
  const int64_t SB = lm.stride * 4;
  const char* pin = (const char*)inv + off0 * ISZ;
  char* const stk0 = (char*)(stk + off0);  // slot 0
  char* ptop = stk0 - SB;                  // slot k (k = -1: empty)

  int k = -1;
  // top (v1,h1) and the differences to the entry below it: dv = v1-v0, dh = h1-h0
  int v1 = 0, h1 = 0, dv = 0, dh = 0;
  for (int q0 = 0; q0 < n; q0 += ENV_U) {
    int fb[ENV_U];
    {
      const char* pl = pin;
#pragma unroll
      for (int u = 0; u < ENV_U; ++u) {
        int f = INF_I;
        if (q0 + u < n) {
          if (IN16) {
            const int d = *(const uint16_t*)pl;
            f = d == INF16 ? INF_I : d * d;
          } else {
            f = *(const int32_t*)pl;
          }
        }
        fb[u] = f;
        pl += SBi;
      }
      pin = pl;
    }
#pragma unroll
    for (int u = 0; u < ENV_U; ++u) {
      const int q = q0 + u;
      const int f = fb[u];
      if (f < INF_I) {
        const int h = f + q * q;
        while (k >= 1) {
          // pop while  s(top,q) <= s(second,top):  (h-h1)*(v1-v0) <= (h1-h0)*(q-v1)
          const long long lhs = (long long)(h - h1) * (long long)dv;
          const long long rhs = (long long)dh * (long long)(q - v1);
          if (lhs > rhs) break;
          --k;
          v1 -= dv;  // = v0
          h1 -= dh;  // = h0
          ptop -= SB;
          if (k >= 1) {
            const uint32_t e = *(const uint32_t*)(ptop - SB);
            dv = v1 - unpack_v(e);
            dh = h1 - unpack_h(e);
          }
        }
        ++k;
        ptop += SB;
        *(uint32_t*)ptop = pack_vh(q, h);
        dv = q - v1;
        dh = h - h1;
        v1 = q;
        h1 = h;
      }
    }
  }
  const int kmax = k;

  char* pout = (char*)outv + off0 * 4;
  if (kmax < 0) {
    for (int q = 0; q < n; ++q) {
      if (FINAL)
        *(float*)pout = __int_as_float(0x7f800000);
      else
        *(int32_t*)pout = INF_I;
      pout += SB;
    }
    return;
  }
  // Query, driven by the hull instead of by q: entries are read ENV_U at a time (independent
  // loads), and each entry emits every q it owns before the next one takes over.  Parabola
  // `nxt` takes over from `cur` at the first integer q with (hn-hc) < 2q(vn-vc).
  // val(q) = (q-vc)^2 + f(vc) is carried incrementally: val(q+1) = val(q) + 2(q-vc) + 1.
  uint32_t e = *(const uint32_t*)stk0;
  int vc = unpack_v(e), hc = unpack_h(e);
  int q = 0;
  const char* ps = stk0 + SB;  // slot 1
  for (int kb = 1; kb <= kmax; kb += ENV_U) {
    uint32_t eb[ENV_U];
    {
      const char* pl = ps;
#pragma unroll
      for (int u = 0; u < ENV_U; ++u) {
        eb[u] = (kb + u <= kmax) ? *(const uint32_t*)pl : 0u;
        pl += SB;
      }
      ps = pl;
    }
#pragma unroll
    for (int u = 0; u < ENV_U; ++u) {
      if (kb + u <= kmax) {
        const int vn = unpack_v(eb[u]), hn = unpack_h(eb[u]);
        const int dh = hn - hc, dv2 = 2 * (vn - vc);
        int val = hc + q * (q - 2 * vc);
        int inc = 2 * (q - vc) + 1;
        int lim = q * dv2;
        while (q < n && dh >= lim) {
          if (FINAL)
            *(float*)pout = res * fast_sqrt((float)val);
          else
            *(int32_t*)pout = val;
          pout += SB;
          val += inc;
          inc += 2;
          lim += dv2;
          ++q;
        }
        vc = vn;
        hc = hn;
      }
    }
  }
  {
    int val = hc + q * (q - 2 * vc);
    int inc = 2 * (q - vc) + 1;
    for (; q < n; ++q) {
      if (FINAL)
        *(float*)pout = res * fast_sqrt((float)val);
      else
        *(int32_t*)pout = val;
      pout += SB;
      val += inc;
      inc += 2;
    }
  }
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
__global__ void ceiling_kernel(uint8_t* __restrict__ occ, int ny, int nz, Box b, int ceil_id) {
  const int nyb = b.hi[1] - b.lo[1] + 1, nxb = b.hi[0] - b.lo[0] + 1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nyb * nxb) return;
  const int y = b.lo[1] + t % nyb, x = b.lo[0] + t / nyb;
  const int64_t a = ((int64_t)x * ny + y) * nz + ceil_id;
  occ[a] = (uint8_t)((occ[a] & ~3u) | FUELGPU_OCCUPIED);  // occupancy_buffer_ = clamp_max_log_ (:463-470)
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
    ceiling_kernel<<<(n2 + 255) / 256, 256, 0, m->stream>>>(m->occ, m->g.ny, m->g.nz, b, ceil_id);
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

// ---- multi-GPU building blocks ----------------------------------------------------------
// z-sharded volume: each rank owns nzl planes of every (x,y) column.  The x and y sweeps
// never cross z, so they run on the local slab first (the transform is separable and exact
// in integers, so the sweep order is immaterial); after one all-to-all the z sweep sees whole
// columns assembled from G chunks.
namespace {

// first sweep along y straight from the occupancy byte: 1-D distance (uint16, voxels) to the
// nearest site on the (x,z) line.  Thread per line, lanes along z; rows are read U at a time so
// that every thread keeps U loads in flight (forward pass), then the backward pass folds in the
// nearest site above.
constexpr int YB_U = 8;
__global__ void __launch_bounds__(128) ysweep_binary_kernel(const uint8_t* __restrict__ occ,
                                                            uint16_t* __restrict__ out, int nx, int ny, int nz,
                                                            int mode) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)nx * nz) return;
  const int z = (int)(t % nz);
  const int x = (int)(t / nz);
  const int64_t off0 = (int64_t)x * ny * nz + z;
  int d = BIG;
  for (int y0 = 0; y0 < ny; y0 += YB_U) {
    uint8_t o[YB_U];
#pragma unroll
    for (int u = 0; u < YB_U; ++u) o[u] = (y0 + u < ny) ? occ[off0 + (int64_t)(y0 + u) * nz] : 0;
#pragma unroll
    for (int u = 0; u < YB_U; ++u) {
      if (y0 + u < ny) {
        d = is_site(o[u], mode) ? 0 : min(d + 1, BIG);
        out[off0 + (int64_t)(y0 + u) * nz] = d >= BIG ? INF16 : (uint16_t)d;
      }
    }
  }
  int dr = BIG;
  for (int y1 = ny - 1; y1 >= 0; y1 -= YB_U) {
    uint16_t v[YB_U];
#pragma unroll
    for (int u = 0; u < YB_U; ++u) v[u] = (y1 - u >= 0) ? out[off0 + (int64_t)(y1 - u) * nz] : INF16;
#pragma unroll
    for (int u = 0; u < YB_U; ++u) {
      if (y1 - u >= 0) {
        const int dl = v[u] == INF16 ? BIG : (int)v[u];
        dr = dl == 0 ? 0 : min(dr + 1, BIG);
        const int m = min(dl, dr);
        if (m != dl) out[off0 + (int64_t)(y1 - u) * nz] = m >= BIG ? INF16 : (uint16_t)m;
      }
    }
  }
}

// [G][nxl][ny][nzl] chunks (z fastest inside a chunk) -> [nxl][nz][ny] (y fastest), 32x32 tiles
// through shared memory so that both sides are coalesced.
__global__ void __launch_bounds__(256) chunks_to_zy_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                           int G, int nxl, int ny, int nzl) {
  __shared__ int32_t tile[32][33];
  const int nz = G * nzl;
  const int x = blockIdx.z;
  const int y0 = blockIdx.y * 32, z0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int64_t chunk = (int64_t)nxl * ny * nzl;
  for (int r = ty; r < 32; r += 8) {
    const int y = y0 + r, z = z0 + tx;
    if (y < ny && z < nz) tile[r][tx] = in[(int64_t)(z / nzl) * chunk + ((int64_t)x * ny + y) * nzl + (z % nzl)];
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int z = z0 + r, y = y0 + tx;
    if (y < ny && z < nz) out[((int64_t)x * nz + z) * ny + y] = tile[tx][r];
  }
}

// [nxl][nz][ny] (y fastest) -> [nxl][ny][nz] (z fastest, the reference layout)
__global__ void __launch_bounds__(256) zy_to_yz_kernel(const float* __restrict__ in, float* __restrict__ out, int nxl,
                                                       int ny, int nz) {
  __shared__ float tile[32][33];
  const int x = blockIdx.z;
  const int y0 = blockIdx.x * 32, z0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int z = z0 + r, y = y0 + tx;
    if (y < ny && z < nz) tile[r][tx] = in[((int64_t)x * nz + z) * ny + y];
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int y = y0 + r, z = z0 + tx;
    if (y < ny && z < nz) out[((int64_t)x * ny + y) * nz + z] = tile[tx][r];
  }
}

}  // namespace

int edt_xy_dev_impl(cudaStream_t s, const uint8_t* occ, int nx, int ny, int nzl, int flags,
                    int32_t* g2, int32_t* scratch) {
  const int mode = (flags & FUELGPU_ESDF_OPTIMISTIC) ? 0 : 1;
  // scratch (2 * nx*ny*nzl int32): first half = uint16 y distances, second half = hull stacks
  uint16_t* g1h = (uint16_t*)scratch;
  uint32_t* stk = (uint32_t*)(scratch + (int64_t)nx * ny * nzl);
  const int64_t nl = (int64_t)nx * nzl;
  ysweep_binary_kernel<<<(unsigned)((nl + 127) / 128), 128, 0, s>>>(occ, g1h, nx, ny, nzl, mode);
  LineMap lm;
  lm.n = nx;
  lm.stride = (int64_t)ny * nzl;
  lm.nz_run = nzl;
  lm.n_outer = ny;
  lm.outer_stride = nzl;
  lm.base = 0;
  const int64_t nl2 = (int64_t)nzl * ny;
  envelope_kernel<true, false, 16><<<(unsigned)((nl2 + 127) / 128), 128, 0, s>>>(g1h, g2, stk, lm, 0.f);
  return cudaGetLastError() == cudaSuccess ? 0 : FUELGPU_ECUDA;
}

int edt_z_chunks_dev_impl(cudaStream_t s, const int32_t* g2c, int G, int nxl, int ny, int nzl,
                          double res, float* out, int32_t* scratch) {
  // scratch (2 * nxl*ny*nz int32): T = the chunks transposed to [nxl][nz][ny] (y fastest, so the
  // z sweep has its lanes along a contiguous axis), T2 = the fp32 result in the same layout.
  const int nz = G * nzl;
  const int64_t voxl = (int64_t)nxl * ny * nz;
  int32_t* T = scratch;
  float* T2 = (float*)(scratch + voxl);
  dim3 g1((nz + 31) / 32, (ny + 31) / 32, nxl);
  chunks_to_zy_kernel<<<g1, 256, 0, s>>>(g2c, T, G, nxl, ny, nzl);
  LineMap lm;
  lm.n = nz;
  lm.stride = ny;
  lm.nz_run = ny;
  lm.n_outer = nxl;
  lm.outer_stride = (int64_t)nz * ny;
  lm.base = 0;
  const int64_t nl = (int64_t)ny * nxl;
  envelope_kernel<false, true, 8><<<(unsigned)((nl + 127) / 128), 128, 0, s>>>(T, T2, (uint32_t*)T, lm, (float)res);
  dim3 g2((ny + 31) / 32, (nz + 31) / 32, nxl);
  zy_to_yz_kernel<<<g2, 256, 0, s>>>(T2, out, nxl, ny, nz);
  return cudaGetLastError() == cudaSuccess ? 0 : FUELGPU_ECUDA;
}
