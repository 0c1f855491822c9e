/*
 * fuel_oracle_viewpoints.c -- CPU restatement of the step right after the frontier path:
 * FrontierFinder::sampleViewpoints / countVisibleCells / isNearUnknown / wrapYaw
 * (active_perception/src/frontier_finder.cpp:662-695,721-755,776-781), PerceptionUtils::setPose / insideFOV
 * (active_perception/src/perception_utils.cpp:49-81) and the per-cluster change count of isFrontierCovered
 * (frontier_finder.cpp:697-719).  TEST INFRASTRUCTURE ONLY (see fuel_oracle.h); parity unpinned by reference tests.
 * Citations are file:line under /root/reference/fuel_planner/.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fuel_oracle.h"

#define ORC_RAY_GUARD 4096

static int in_map_idx(const OrcGrid* g, const int32_t id[3]) { /* sdf_map.h:163-169 */
  return !(id[0] < 0 || id[1] < 0 || id[2] < 0 || id[0] > g->n[0] - 1 || id[1] > g->n[1] - 1 || id[2] > g->n[2] - 1);
}
static int64_t to_adr(const OrcGrid* g, const int32_t id[3]) {
  return ((int64_t)id[0] * g->n[1] + id[1]) * g->n[2] + id[2];
}
/* getOccupancy(id) / getInflateOccupancy(id), sdf_map.h:196-226: -1 outside the map */
static int get_occ(const OrcGrid* g, const uint8_t* tri, const int32_t id[3]) {
  return in_map_idx(g, id) ? (int)tri[to_adr(g, id)] : -1;
}
static int get_inflate(const OrcGrid* g, const int8_t* inflate, const int32_t id[3]) {
  return in_map_idx(g, id) ? (int)inflate[to_adr(g, id)] : -1;
}

static int is_near_unknown(const OrcGrid* g, const uint8_t* tri, const double pos[3], double clearance) { /* :721-732 */
  const int vox_num = (int)floor(clearance / g->res);
  for (int x = -vox_num; x <= vox_num; ++x)
    for (int y = -vox_num; y <= vox_num; ++y)
      for (int z = -1; z <= 1; ++z) {
        const double vox[3] = { pos[0] + x * g->res, pos[1] + y * g->res, pos[2] + z * g->res };
        int32_t id[3];
        orc_pos_to_index(g, vox, id);
        if (get_occ(g, tri, id) == ORC_UNKNOWN) return 1;
      }
  return 0;
}

/* RayCaster as used by countVisibleCells (raycast.cpp:323-407); same arithmetic as fuel_oracle_fusion.c */
static double intbound_v(double s, double ds) {
  if (ds < 0) {
    s = -s;
    ds = -ds;
  }
  s = fmod(fmod(s, 1.0) + 1.0, 1.0);
  return (1 - s) / ds;
}

static int ray_is_clear(const OrcGrid* g, const uint8_t* tri, const int8_t* inflate, const double start[3],
                        const double end[3]) {
  const double res = g->res;
  int x = (int)floor(start[0] / res), y = (int)floor(start[1] / res), z = (int)floor(start[2] / res);
  const int ex = (int)floor(end[0] / res), ey = (int)floor(end[1] / res), ez = (int)floor(end[2] / res);
  const double dx = ex - x, dy = ey - y, dz = ez - z;
  const int sx = (int)dx == 0 ? 0 : ((int)dx < 0 ? -1 : 1), sy = (int)dy == 0 ? 0 : ((int)dy < 0 ? -1 : 1),
            sz = (int)dz == 0 ? 0 : ((int)dz < 0 ? -1 : 1);
  double tmx = intbound_v(start[0] / res, dx), tmy = intbound_v(start[1] / res, dy), tmz = intbound_v(start[2] / res, dz);
  const double tdx = ((double)sx) / dx, tdy = ((double)sy) / dy, tdz = ((double)sz) / dz;
  const double off[3] = { 0.5 - g->origin[0] / res, 0.5 - g->origin[1] / res, 0.5 - g->origin[2] / res };
  for (int guard = 0; guard < ORC_RAY_GUARD; ++guard) { /* while (raycaster_->nextId(idx)), :745 */
    const int32_t idx[3] = { (int32_t)(x + off[0]), (int32_t)(y + off[1]), (int32_t)(z + off[2]) };
    if (x == ex && y == ey && z == ez) return 1;
    if (tmx < tmy) {
      if (tmx < tmz) {
        x += sx;
        tmx += tdx;
      } else {
        z += sz;
        tmz += tdz;
      }
    } else {
      if (tmy < tmz) {
        y += sy;
        tmy += tdy;
      } else {
        z += sz;
        tmz += tdz;
      }
    }
    if (get_inflate(g, inflate, idx) == 1 || get_occ(g, tri, idx) == ORC_UNKNOWN) return 0; /* :746-750 */
  }
  return 1;
}

static void normalized3(const double v[3], double out[3]) { /* Eigen normalized(): v / sqrt(squaredNorm) if > 0 */
  const double z = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  if (z > 0) {
    const double n = sqrt(z);
    out[0] = v[0] / n, out[1] = v[1] / n, out[2] = v[2] / n;
  } else
    out[0] = v[0], out[1] = v[1], out[2] = v[2];
}

int32_t orc_viewpoint_candidates(const OrcViewParams* vp, double* off_xy /*[max][2] or NULL*/, int32_t max) {
  /* the two loops of sampleViewpoints, :664-667, with the host libm */
  int32_t n = 0;
  for (double rc = vp->candidate_rmin, dr = (vp->candidate_rmax - vp->candidate_rmin) / vp->candidate_rnum;
       rc <= vp->candidate_rmax + 1e-3; rc += dr)
    for (double phi = -M_PI; phi < M_PI; phi += vp->candidate_dphi) {
      if (off_xy && n < max) {
        off_xy[2 * n] = rc * cos(phi);
        off_xy[2 * n + 1] = rc * sin(phi);
      }
      ++n;
    }
  return n;
}

/* sampleViewpoints for one cluster, every candidate reported:
 *   cand_pos[c]   sample_pos
 *   cand_yaw[c]   avg_yaw (only meaningful when cand_visib[c] >= 0)
 *   cand_visib[c] -1 if the candidate is rejected by isInBox / getInflateOccupancy / isNearUnknown (:671-673),
 *                 else countVisibleCells; the caller keeps those with visib > min_visib_num (:688)
 *   cand_border[c] (may be NULL) 1 if some FOV / range test of this candidate is within 1e-9 of its threshold, i.e.
 *                 an implementation with a different libm (sin/cos/acos ulps) may legitimately count differently */
int32_t orc_sample_viewpoints(const OrcGrid* g, const uint8_t* tri, const int8_t* inflate, const OrcViewParams* vp,
                              const double average[3], const double* cells /*[n][3] filtered_cells_*/, int32_t n_cells,
                              double* cand_pos, double* cand_yaw, int32_t* cand_visib, uint8_t* cand_border) {
  const int32_t ncand = orc_viewpoint_candidates(vp, NULL, 0);
  double* off = (double*)malloc(sizeof(double) * 2 * ncand);
  orc_viewpoint_candidates(vp, off, ncand);
  /* PerceptionUtils constructor, perception_utils.cpp:13-17 */
  const double ta = sin(M_PI_2 - vp->top_angle), tb = cos(M_PI_2 - vp->top_angle);
  const double lc = sin(M_PI_2 - vp->left_angle), ld = cos(M_PI_2 - vp->left_angle);
  const double re = sin(M_PI_2 - vp->right_angle), rf = cos(M_PI_2 - vp->right_angle);
  for (int32_t c = 0; c < ncand; ++c) {
    const double pos[3] = { average[0] + off[2 * c], average[1] + off[2 * c + 1], average[2] + 0.0 };
    memcpy(cand_pos + 3 * c, pos, sizeof(pos));
    cand_yaw[c] = 0.0;
    cand_visib[c] = -1;
    if (cand_border) cand_border[c] = 0;
    /* isInBox(pos), sdf_map.h:180-187 */
    int inbox = 1;
    for (int i = 0; i < 3; ++i)
      if (pos[i] <= g->box_mind[i] || pos[i] >= g->box_maxd[i]) inbox = 0;
    int32_t pid[3];
    orc_pos_to_index(g, pos, pid);
    if (!inbox || get_inflate(g, inflate, pid) == 1 || is_near_unknown(g, tri, pos, vp->min_candidate_clearance)) continue;
    if (n_cells <= 0) continue;
    /* average yaw, :675-685 */
    double d0[3] = { cells[0] - pos[0], cells[1] - pos[1], cells[2] - pos[2] }, ref_dir[3];
    normalized3(d0, ref_dir);
    double avg_yaw = 0.0;
    for (int32_t i = 1; i < n_cells; ++i) {
      double d[3] = { cells[3 * i] - pos[0], cells[3 * i + 1] - pos[1], cells[3 * i + 2] - pos[2] }, dir[3];
      normalized3(d, dir);
      double yaw = acos(dir[0] * ref_dir[0] + dir[1] * ref_dir[1] + dir[2] * ref_dir[2]);
      if (ref_dir[0] * dir[1] - ref_dir[1] * dir[0] < 0) yaw = -yaw;
      avg_yaw += yaw;
    }
    avg_yaw = avg_yaw / n_cells + atan2(ref_dir[1], ref_dir[0]);
    while (avg_yaw < -M_PI) avg_yaw += 2 * M_PI; /* wrapYaw :776-781 */
    while (avg_yaw > M_PI) avg_yaw -= 2 * M_PI;
    cand_yaw[c] = avg_yaw;
    /* setPose, perception_utils.cpp:49-66: normals = R_wc * {n_top, n_bottom, n_left, n_right},
     * R_wc = R_wb(yaw) * R_bc, R_bc = [0 0 1; -1 0 0; 0 1 0] (T_cb_.inverse(), :18-19) */
    const double cy = cos(avg_yaw), sy = sin(avg_yaw);
    const double nrm[4][3] = { { cy * tb, sy * tb, ta },
                               { cy * tb, sy * tb, -ta },
                               { sy * lc + cy * ld, (-cy) * lc + sy * ld, 0.0 },
                               { sy * (-re) + cy * rf, (-cy) * (-re) + sy * rf, 0.0 } };
    /* countVisibleCells, :734-755 */
    int visib = 0;
    for (int32_t i = 0; i < n_cells; ++i) {
      const double* cell = cells + 3 * i;
      double dir[3] = { cell[0] - pos[0], cell[1] - pos[1], cell[2] - pos[2] };
      const double nn = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]); /* insideFOV :83-93 */
      if (cand_border && fabs(nn - vp->max_dist) < 1e-9) cand_border[c] = 1;
      if (nn > vp->max_dist) continue;
      double u[3];
      normalized3(dir, u);
      int inside = 1;
      for (int k = 0; k < 4; ++k) {
        const double dt = u[0] * nrm[k][0] + u[1] * nrm[k][1] + u[2] * nrm[k][2];
        if (cand_border && fabs(dt) < 1e-9) cand_border[c] = 1;
        if (inside && dt < 0.0) inside = 0;
      }
      if (!inside) continue;
      if (ray_is_clear(g, tri, inflate, cell, pos)) visib += 1;
    }
    cand_visib[c] = visib;
  }
  free(off);
  return ncand;
}

/* isFrontierCovered's inner count, frontier_finder.cpp:703-712: how many cells of a stored cluster are no longer
 * frontier cells (the reference returns as soon as the count reaches min_view_finish_fraction_ * size). */
int32_t orc_frontier_changed_count(const OrcGrid* g, const uint8_t* tri, const int32_t* addr, int32_t n) {
  int32_t cnt = 0;
  const int64_t nyz = (int64_t)g->n[1] * g->n[2];
  for (int32_t i = 0; i < n; ++i) {
    const int32_t id[3] = { (int32_t)(addr[i] / nyz), (int32_t)((addr[i] % nyz) / g->n[2]), (int32_t)(addr[i] % g->n[2]) };
    if (!orc_is_frontier_cell(g, tri, id)) ++cnt;
  }
  return cnt;
}
