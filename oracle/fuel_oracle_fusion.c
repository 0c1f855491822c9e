/*
 * fuel_oracle_fusion.c -- CPU restatement of the occupancy fusion that feeds the hot path:
 * SDFMap::inputPointCloud (plan_env/src/sdf_map.cpp:243-345), setCacheOccupancy (:243-257),
 * closetPointInMap (:347-362) and RayCaster::input / nextId (plan_env/src/raycast.cpp:6-23,323-407).
 * SURVEY.md 8f rank 3.  TEST INFRASTRUCTURE ONLY (see fuel_oracle.h); parity unpinned by reference tests.
 * Citations are file:line under /root/reference/fuel_planner/.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fuel_oracle.h"

#define ORC_RAY_GUARD 4096

/* ---- RayCaster (raycast.cpp) ----------------------------------------------------------------- */
static int signum_i(int x) { return x == 0 ? 0 : x < 0 ? -1 : 1; }                     /* :6-8 */
static double mod_d(double value, double modulus) {                                      /* :10-12 */
  return fmod(fmod(value, modulus) + modulus, modulus);
}
static double intbound(double s, double ds) {                                            /* :14-23 */
  if (ds < 0) {
    return intbound(-s, -ds);
  } else {
    s = mod_d(s, 1);
    return (1 - s) / ds;
  }
}

typedef struct {
  int x, y, z, endX, endY, endZ, stepX, stepY, stepZ;
  double tMaxX, tMaxY, tMaxZ, tDeltaX, tDeltaY, tDeltaZ;
  double resolution, offset[3];
} RayCaster;

static void rc_set_params(RayCaster* r, double res, const double origin[3]) {           /* :323-327 */
  r->resolution = res;
  for (int i = 0; i < 3; ++i) r->offset[i] = 0.5 - origin[i] / res;
}

static int rc_input(RayCaster* r, const double start[3], const double end[3]) {          /* :329-372 */
  double s[3], e[3];
  for (int i = 0; i < 3; ++i) {
    s[i] = start[i] / r->resolution;
    e[i] = end[i] / r->resolution;
  }
  r->x = (int)floor(s[0]);
  r->y = (int)floor(s[1]);
  r->z = (int)floor(s[2]);
  r->endX = (int)floor(e[0]);
  r->endY = (int)floor(e[1]);
  r->endZ = (int)floor(e[2]);
  const double dx = r->endX - r->x, dy = r->endY - r->y, dz = r->endZ - r->z;
  r->stepX = signum_i((int)dx);
  r->stepY = signum_i((int)dy);
  r->stepZ = signum_i((int)dz);
  r->tMaxX = intbound(s[0], dx);
  r->tMaxY = intbound(s[1], dy);
  r->tMaxZ = intbound(s[2], dz);
  r->tDeltaX = ((double)r->stepX) / dx;
  r->tDeltaY = ((double)r->stepY) / dy;
  r->tDeltaZ = ((double)r->stepZ) / dz;
  return !(r->stepX == 0 && r->stepY == 0 && r->stepZ == 0);
}

static int rc_next_id(RayCaster* r, int32_t idx[3]) {                                     /* :374-407 */
  /* idx = (Vector3d(x,y,z) + offset_).cast<int>(): truncation toward zero */
  idx[0] = (int32_t)(r->x + r->offset[0]);
  idx[1] = (int32_t)(r->y + r->offset[1]);
  idx[2] = (int32_t)(r->z + r->offset[2]);
  if (r->x == r->endX && r->y == r->endY && r->z == r->endZ) return 0;
  if (r->tMaxX < r->tMaxY) {
    if (r->tMaxX < r->tMaxZ) {
      r->x += r->stepX;
      r->tMaxX += r->tDeltaX;
    } else {
      r->z += r->stepZ;
      r->tMaxZ += r->tDeltaZ;
    }
  } else {
    if (r->tMaxY < r->tMaxZ) {
      r->y += r->stepY;
      r->tMaxY += r->tDeltaY;
    } else {
      r->z += r->stepZ;
      r->tMaxZ += r->tDeltaZ;
    }
  }
  return 1;
}

/* ---- SDFMap fusion ---------------------------------------------------------------------------- */
static int in_map_pos(const OrcGrid* g, const double p[3]) { return orc_is_in_map_pos(g, p); }

static void closest_point_in_map(const OrcGrid* g, const double pt[3], const double cam[3], double out[3]) {
  /* closetPointInMap, sdf_map.cpp:347-362 */
  double diff[3], max_tc[3], min_tc[3];
  for (int i = 0; i < 3; ++i) {
    diff[i] = pt[i] - cam[i];
    max_tc[i] = (g->origin[i] + (g->map_size[i] > 0.0 ? g->map_size[i] : g->n[i] * g->res)) - cam[i];
    min_tc[i] = g->origin[i] - cam[i];
  }
  double min_t = 1000000;
  for (int i = 0; i < 3; ++i) {
    if (fabs(diff[i]) > 0) {
      double t1 = max_tc[i] / diff[i];
      if (t1 > 0 && t1 < min_t) min_t = t1;
      double t2 = min_tc[i] / diff[i];
      if (t2 > 0 && t2 < min_t) min_t = t2;
    }
  }
  for (int i = 0; i < 3; ++i) out[i] = cam[i] + (min_t - 1e-3) * diff[i];
}

void orc_fusion_state_init(OrcFusionState* st, int64_t nvox) {
  memset(st, 0, sizeof(*st));
  st->count_hit = (int16_t*)calloc(nvox, sizeof(int16_t));
  st->count_miss = (int16_t*)calloc(nvox, sizeof(int16_t));
  st->flag_rayend = (int8_t*)malloc(nvox);
  memset(st->flag_rayend, -1, nvox); /* sdf_map.cpp:71 */
  st->raycast_num = 0;
  st->reset_updated_box = 1;
}

void orc_fusion_state_free(OrcFusionState* st) {
  free(st->count_hit);
  free(st->count_miss);
  free(st->flag_rayend);
  memset(st, 0, sizeof(*st));
}

/* inputPointCloud, sdf_map.cpp:259-345.  points: float32 xyz (pcl::PointXYZ). */
void orc_input_point_cloud(const OrcGrid* g, const OrcFusionParams* fp, OrcFusionState* st, double* logodds,
                           const float* points, int32_t point_num, const double camera_pos[3],
                           int32_t local_bound_min[3], int32_t local_bound_max[3]) {
  if (point_num == 0) return;
  st->raycast_num = (int8_t)(st->raycast_num + 1); /* char counter, :263 */
  const int64_t nyz = (int64_t)g->n[1] * g->n[2];
  double update_min[3], update_max[3];
  for (int k = 0; k < 3; ++k) update_min[k] = update_max[k] = camera_pos[k];
  if (st->reset_updated_box) {
    for (int k = 0; k < 3; ++k) st->update_min[k] = st->update_max[k] = camera_pos[k];
    st->reset_updated_box = 0;
  }
  /* cache_voxel_ queue */
  int64_t qcap = 1 << 16, qn = 0;
  int32_t* queue = (int32_t*)malloc(sizeof(int32_t) * qcap);
  RayCaster caster;
  rc_set_params(&caster, g->res, g->origin);

#define SET_CACHE(adr, occ)                                                             \
  do { /* setCacheOccupancy, :243-257 */                                                \
    const int64_t a_ = (adr);                                                           \
    if (st->count_hit[a_] == 0 && st->count_miss[a_] == 0) {                            \
      if (qn == qcap) {                                                                 \
        qcap *= 2;                                                                      \
        queue = (int32_t*)realloc(queue, sizeof(int32_t) * qcap);                       \
      }                                                                                 \
      queue[qn++] = (int32_t)a_;                                                        \
    }                                                                                   \
    if ((occ) == 0)                                                                     \
      st->count_miss[a_] = 1;                                                           \
    else if ((occ) == 1)                                                                \
      st->count_hit[a_] += 1;                                                           \
  } while (0)

  for (int32_t i = 0; i < point_num; ++i) {
    double pt_w[3] = { (double)points[3 * i], (double)points[3 * i + 1], (double)points[3 * i + 2] };
    int tmp_flag;
    double length;
    if (!in_map_pos(g, pt_w)) {
      double c[3];
      closest_point_in_map(g, pt_w, camera_pos, c);
      memcpy(pt_w, c, sizeof(c));
      double d[3] = { pt_w[0] - camera_pos[0], pt_w[1] - camera_pos[1], pt_w[2] - camera_pos[2] };
      length = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      if (length > fp->max_ray_length)
        for (int k = 0; k < 3; ++k) pt_w[k] = d[k] / length * fp->max_ray_length + camera_pos[k];
      if (pt_w[2] < 0.2) continue;
      tmp_flag = 0;
    } else {
      double d[3] = { pt_w[0] - camera_pos[0], pt_w[1] - camera_pos[1], pt_w[2] - camera_pos[2] };
      length = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      if (length > fp->max_ray_length) {
        for (int k = 0; k < 3; ++k) pt_w[k] = d[k] / length * fp->max_ray_length + camera_pos[k];
        if (pt_w[2] < 0.2) continue;
        tmp_flag = 0;
      } else
        tmp_flag = 1;
    }
    int32_t idx[3];
    orc_pos_to_index(g, pt_w, idx);
    /* the reference indexes without a check; a camera outside the map could push this out of range
     * (MapROS returns early in that case, map_ros.cpp:126-127).  Skip instead of reading out of bounds. */
    if (idx[0] < 0 || idx[1] < 0 || idx[2] < 0 || idx[0] >= g->n[0] || idx[1] >= g->n[1] || idx[2] >= g->n[2]) continue;
    const int64_t vox_adr = idx[0] * nyz + (int64_t)idx[1] * g->n[2] + idx[2];
    SET_CACHE(vox_adr, tmp_flag);
    for (int k = 0; k < 3; ++k) {
      update_min[k] = update_min[k] < pt_w[k] ? update_min[k] : pt_w[k];
      update_max[k] = update_max[k] > pt_w[k] ? update_max[k] : pt_w[k];
    }
    /* one raycast per end voxel and frame, :303-306 */
    if (st->flag_rayend[vox_adr] == st->raycast_num)
      continue;
    else
      st->flag_rayend[vox_adr] = st->raycast_num;
    rc_input(&caster, pt_w, camera_pos);
    rc_next_id(&caster, idx);
    /* the reference loops until nextId() reports the camera voxel; ORC_RAY_GUARD (> the Manhattan length
     * of any ray in a <= 1024^3 map) only keeps a ray whose DDA steps past the end voxel from spinning. */
    int guard = 1;
    while (guard++ < ORC_RAY_GUARD && rc_next_id(&caster, idx)) {
      if (idx[0] < 0 || idx[1] < 0 || idx[2] < 0 || idx[0] >= g->n[0] || idx[1] >= g->n[1] || idx[2] >= g->n[2]) continue;
      SET_CACHE(idx[0] * nyz + (int64_t)idx[1] * g->n[2] + idx[2], 0);
    }
  }
#undef SET_CACHE

  /* local bound, :313-318 */
  double hi[3] = { update_max[0] + fp->local_bound_inflate, update_max[1] + fp->local_bound_inflate, update_max[2] };
  double lo[3] = { update_min[0] - fp->local_bound_inflate, update_min[1] - fp->local_bound_inflate, update_min[2] };
  orc_pos_to_index(g, hi, local_bound_max);
  orc_pos_to_index(g, lo, local_bound_min);
  for (int k = 0; k < 3; ++k) {
    local_bound_min[k] = local_bound_min[k] < g->n[k] - 1 ? local_bound_min[k] : g->n[k] - 1;
    local_bound_min[k] = local_bound_min[k] > 0 ? local_bound_min[k] : 0;
    local_bound_max[k] = local_bound_max[k] < g->n[k] - 1 ? local_bound_max[k] : g->n[k] - 1;
    local_bound_max[k] = local_bound_max[k] > 0 ? local_bound_max[k] : 0;
  }
  /* bounding box for subsequent updating, :321-324 */
  for (int k = 0; k < 3; ++k) {
    st->update_min[k] = update_min[k] < st->update_min[k] ? update_min[k] : st->update_min[k];
    st->update_max[k] = update_max[k] > st->update_max[k] ? update_max[k] : st->update_max[k];
  }
  /* log-odds update of the cached voxels, :326-344 */
  const double clamp_min = log(fp->p_min / (1 - fp->p_min)), clamp_max = log(fp->p_max / (1 - fp->p_max));
  const double hit = log(fp->p_hit / (1 - fp->p_hit)), miss = log(fp->p_miss / (1 - fp->p_miss));
  const double min_occ = log(fp->p_occ / (1 - fp->p_occ));
  for (int64_t q = 0; q < qn; ++q) {
    const int32_t adr = queue[q];
    const double upd = st->count_hit[adr] >= st->count_miss[adr] ? hit : miss;
    st->count_hit[adr] = st->count_miss[adr] = 0;
    if (logodds[adr] < clamp_min - 1e-3) logodds[adr] = min_occ;
    const double v = logodds[adr] + upd;
    logodds[adr] = fmin(fmax(v, clamp_min), clamp_max);
  }
  free(queue);
}

/* MapROS::proessDepthImage, map_ros.cpp:176-215 */
int32_t orc_process_depth_image(const OrcCameraParams* cp, const uint16_t* depth_image, int32_t rows, int32_t cols,
                                const double R[9], const double camera_pos[3], float* points_out) {
  int32_t cnt = 0;
  const double inv_factor = 1.0 / cp->k_depth_scaling_factor;
  const int m = cp->depth_filter_margin, skip = cp->skip_pixel;
  const uint16_t* const img_end = depth_image + (int64_t)rows * cols;
  for (int v = m; v < rows - m; v += skip) {
    const uint16_t* row_ptr = depth_image + (int64_t)v * cols + m;
    for (int u = m; u < cols - m; u += skip) {
      double depth = (*row_ptr) * inv_factor;
      row_ptr = row_ptr + skip;
      /* the reference dereferences the advanced pointer; with margin 0 the very last read would leave the
       * image -- treat that one as "no return" instead of reading out of bounds */
      const uint16_t nxt = row_ptr < img_end ? *row_ptr : 0;
      if (nxt == 0 || depth > cp->depth_filter_maxdist)
        depth = cp->depth_filter_maxdist;
      else if (depth < cp->depth_filter_mindist)
        continue;
      const double pc[3] = { (u - cp->cx) * depth / cp->fx, (v - cp->cy) * depth / cp->fy, depth };
      for (int k = 0; k < 3; ++k) {
        /* Eigen 3x3 * vector, coefficient order, then + camera_pos_ (:206) */
        const double w = R[3 * k] * pc[0] + R[3 * k + 1] * pc[1] + R[3 * k + 2] * pc[2] + camera_pos[k];
        points_out[3 * cnt + k] = (float)w; /* pt.x = pt_world[0] (float member) */
      }
      ++cnt;
    }
  }
  return cnt;
}

/* The RayCaster restatement on its own, for pinning against the reference's raycast.cpp (oracle/_ref): the voxel
 * indices nextId() reports between input(start, end) and the first `false`. */
int32_t orc_raycast_ids(const OrcGrid* g, const double start[3], const double end[3], int32_t* ids, int32_t max) {
  RayCaster r;
  rc_set_params(&r, g->res, g->origin);
  rc_input(&r, start, end);
  int32_t n = 0, idx[3];
  while (n < max && rc_next_id(&r, idx)) {
    ids[3 * n] = idx[0], ids[3 * n + 1] = idx[1], ids[3 * n + 2] = idx[2];
    ++n;
  }
  return n;
}
double orc_intbound(double s, double ds) { return intbound(s, ds); }
