/* TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
 *
 * CPU restatement of the point-cloud helpers behind the reference's `bruce_slam.pcl`
 * pybind module, bruce_slam/src/bruce_slam/cpp/pcl.cpp:
 *     remove_outlier :54-74    PCL RadiusOutlierRemoval<PointXYZ>
 *     downsample     :128-159  libpointmatcher OctreeGridDataPointsFilter (samplingMethod 3 = medoid)
 *     match          :161-174  libpointmatcher KDTreeMatcher (libnabo), knn nearest within maxDist
 * The arithmetic of these lives in third-party libraries that are NOT vendored in
 * /root/reference (libpointmatcher pinned at d478ef2e by README.md:51-55, libnabo unpinned,
 * PCL as shipped by ROS Noetic) and are not installed here:  **parity unpinned** -- what
 * follows restates their published algorithms in float32, the scalar
 * type the wrapper instantiates (`PointMatcher<float>`, pcl.cpp:12; pcl::PointXYZ).
 *
 * Conventions restated
 *   match        squared Euclidean distance dx*dx + dy*dy in float32; a reference point is
 *                accepted when dist <= maxDist^2 (libnabo `dist <= maxRadius2`); no neighbour ->
 *                id -1, dist +inf; ties -> lowest reference index.
 *   remove_outlier  keep a point iff at least (min_points + 1) points of the cloud, itself
 *                included, lie within float32 squared distance <= radius^2 (PCL's dense path:
 *                nearestKSearch(k = min_pts + 1) and `radius^2 < dist[k-1]` rejects); output
 *                keeps input order.
 *   downsample   quadtree over the cloud's bounding SQUARE (centre = bbox middle, half-size =
 *                larger half-extent); a node is split while its size (2*radius) > resolution AND
 *                it holds more than one point; children are indexed by bit0 = (x > cx), bit1 =
 *                (y > cy) and visited in index order; every non-empty leaf yields its medoid =
 *                the member minimising the float32 sum of Euclidean distances to all members
 *                (first one wins ties); output in visiting order.
 *
 * nn_grid_* is an exact nearest-neighbour search on a uniform grid (the CPU stand-in for the
 * KD-tree: same answers as brute force, see tests), shared with oracle/icp_ref.c.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ exact NN on a uniform grid */
typedef struct {
  int n, nx, ny;
  float ox, oy, cell, inv_cell;
  int *start;   /* [nx*ny + 1] */
  int *order;   /* point indices sorted by cell, ascending index inside a cell */
  const float *pts;
} nn_grid;

static int cell_of(const nn_grid *g, float v, float o, int n) {
  int c = (int)floorf((v - o) * g->inv_cell);
  return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

nn_grid *nn_grid_build(const float *pts, int n) {
  nn_grid *g = (nn_grid *)calloc(1, sizeof(nn_grid));
  g->n = n;
  g->pts = pts;
  float minx = FLT_MAX, miny = FLT_MAX, maxx = -FLT_MAX, maxy = -FLT_MAX;
  for (int i = 0; i < n; ++i) {
    const float x = pts[2 * i], y = pts[2 * i + 1];
    if (x < minx) minx = x;
    if (x > maxx) maxx = x;
    if (y < miny) miny = y;
    if (y > maxy) maxy = y;
  }
  if (n == 0) minx = miny = maxx = maxy = 0.f;
  const float w = maxx - minx, h = maxy - miny;
  float cell = sqrtf((w * h) / (float)(n > 0 ? n : 1));
  if (!(cell > 1e-3f)) cell = (w > h ? w : h) / 64.f;
  if (!(cell > 1e-6f)) cell = 1.f;
  g->cell = cell;
  g->inv_cell = 1.f / cell;
  g->ox = minx;
  g->oy = miny;
  g->nx = (int)(w * g->inv_cell) + 1;
  g->ny = (int)(h * g->inv_cell) + 1;
  if (g->nx > 2048) g->nx = 2048;
  if (g->ny > 2048) g->ny = 2048;
  const int nc = g->nx * g->ny;
  g->start = (int *)calloc((size_t)nc + 1, sizeof(int));
  g->order = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i)
    g->start[cell_of(g, pts[2 * i + 1], g->oy, g->ny) * g->nx + cell_of(g, pts[2 * i], g->ox, g->nx) + 1]++;
  for (int c = 0; c < nc; ++c) g->start[c + 1] += g->start[c];
  int *cur = (int *)malloc(sizeof(int) * (size_t)nc);
  memcpy(cur, g->start, sizeof(int) * (size_t)nc);
  for (int i = 0; i < n; ++i)
    g->order[cur[cell_of(g, pts[2 * i + 1], g->oy, g->ny) * g->nx + cell_of(g, pts[2 * i], g->ox, g->nx)]++] = i;
  free(cur);
  return g;
}

void nn_grid_free(nn_grid *g) {
  if (!g) return;
  free(g->start);
  free(g->order);
  free(g);
}

/* nearest point to (qx,qy) with float32 squared distance <= max_d2; returns index or -1 */
int nn_grid_query(const nn_grid *g, float qx, float qy, float max_d2, float *d2_out) {
  int best = -1;
  float best_d2 = INFINITY;
  if (g->n > 0) {
    const int cx = cell_of(g, qx, g->ox, g->nx), cy = cell_of(g, qy, g->oy, g->ny);
    const int kmax = (g->nx > g->ny ? g->nx : g->ny);
    for (int k = 0; k <= kmax; ++k) {
      const int x0 = cx - k, x1 = cx + k, y0 = cy - k, y1 = cy + k;
      for (int y = (y0 < 0 ? 0 : y0); y <= (y1 >= g->ny ? g->ny - 1 : y1); ++y) {
        const int edge_row = (y == y0 || y == y1);
        for (int x = (x0 < 0 ? 0 : x0); x <= (x1 >= g->nx ? g->nx - 1 : x1); ++x) {
          if (!edge_row && x != x0 && x != x1) continue; /* ring only */
          const int c = y * g->nx + x;
          for (int s = g->start[c]; s < g->start[c + 1]; ++s) {
            const int i = g->order[s];
            const float dx = qx - g->pts[2 * i], dy = qy - g->pts[2 * i + 1];
            const float d2 = dx * dx + dy * dy;
            if (d2 < best_d2 || (d2 == best_d2 && i < best)) {
              best_d2 = d2;
              best = i;
            }
          }
        }
      }
      /* everything not yet visited lies outside the (2k+1)^2 block around the query's cell */
      float bound = INFINITY;
      if (x0 > 0) bound = fminf(bound, qx - (g->ox + (float)x0 * g->cell));
      if (x1 < g->nx - 1) bound = fminf(bound, (g->ox + (float)(x1 + 1) * g->cell) - qx);
      if (y0 > 0) bound = fminf(bound, qy - (g->oy + (float)y0 * g->cell));
      if (y1 < g->ny - 1) bound = fminf(bound, (g->oy + (float)(y1 + 1) * g->cell) - qy);
      if (bound == INFINITY) break;
      if (bound > 0.f) {
        const double b = (double)bound * (1.0 - 1e-5) - 1e-6; /* conservative against float rounding */
        if (b > 0 && (b * b > (double)best_d2 || b * b > (double)max_d2)) break;
      }
    }
  }
  if (best >= 0 && !(best_d2 <= max_d2)) best = -1;
  if (d2_out) *d2_out = best >= 0 ? best_d2 : INFINITY;
  return best;
}

/* pcl.match(ref, in, 1, max_dist): ids[n_in], dists[n_in] (squared) */
void orc_match(const float *ref, int n_ref, const float *in, int n_in, float max_dist, int32_t *ids, float *dists) {
  nn_grid *g = nn_grid_build(ref, n_ref);
  const float max_d2 = max_dist * max_dist;
  for (int i = 0; i < n_in; ++i) ids[i] = nn_grid_query(g, in[2 * i], in[2 * i + 1], max_d2, &dists[i]);
  nn_grid_free(g);
}

/* brute-force twin of orc_match, for cross-checking the grid */
void orc_match_brute(const float *ref, int n_ref, const float *in, int n_in, float max_dist, int32_t *ids,
                     float *dists) {
  const float max_d2 = max_dist * max_dist;
  for (int i = 0; i < n_in; ++i) {
    int best = -1;
    float bd = INFINITY;
    for (int j = 0; j < n_ref; ++j) {
      const float dx = in[2 * i] - ref[2 * j], dy = in[2 * i + 1] - ref[2 * j + 1];
      const float d2 = dx * dx + dy * dy;
      if (d2 < bd) bd = d2, best = j;
    }
    if (best >= 0 && !(bd <= max_d2)) best = -1;
    ids[i] = best;
    dists[i] = best >= 0 ? bd : INFINITY;
  }
}

/* ------------------------------------------------------------------ remove_outlier */
/* pts: [n][dim] (dim 2 or 3); keep[n] <- 0/1; returns number kept */
int orc_remove_outlier(const float *pts, int n, int dim, double radius, int min_points, uint8_t *keep) {
  const double r2 = radius * radius;
  /* grid over x,y with cell = radius for the neighbour scan (exact: all candidates are re-tested) */
  float *xy = (float *)calloc(2 * (size_t)(n > 0 ? n : 1), sizeof(float));
  for (int i = 0; i < n; ++i) xy[2 * i] = pts[(size_t)i * dim], xy[2 * i + 1] = pts[(size_t)i * dim + 1];
  nn_grid *g = nn_grid_build(xy, n);
  int kept = 0;
  for (int i = 0; i < n; ++i) {
    const float qx = xy[2 * i], qy = xy[2 * i + 1], qz = dim == 3 ? pts[(size_t)i * 3 + 2] : 0.f;
    const int span = (int)ceil(radius * (double)g->inv_cell) + 1;
    const int cx = cell_of(g, qx, g->ox, g->nx), cy = cell_of(g, qy, g->oy, g->ny);
    int count = 0;
    for (int y = cy - span; y <= cy + span; ++y) {
      if (y < 0 || y >= g->ny) continue;
      for (int x = cx - span; x <= cx + span; ++x) {
        if (x < 0 || x >= g->nx) continue;
        const int c = y * g->nx + x;
        for (int s = g->start[c]; s < g->start[c + 1]; ++s) {
          const int j = g->order[s];
          const float dx = qx - xy[2 * j], dy = qy - xy[2 * j + 1];
          const float dz = qz - (dim == 3 ? pts[(size_t)j * 3 + 2] : 0.f);
          float d2 = dx * dx; /* FLANN L2_Simple: result += diff*diff, dimension by dimension */
          d2 += dy * dy;
          d2 += dz * dz;
          if (!(r2 < (double)d2)) ++count;
        }
      }
    }
    keep[i] = (uint8_t)(count >= min_points + 1);
    kept += keep[i];
  }
  nn_grid_free(g);
  free(xy);
  return kept;
}

/* ------------------------------------------------------------------ downsample (quadtree medoid) */
typedef struct {
  const float *pts;
  int *out;
  int n_out;
  float max_size;
} ds_ctx;

static void ds_leaf(ds_ctx *c, const int *idx, int n) {
  float best = FLT_MAX;
  int med = 0;
  for (int a = 0; a < n; ++a) {
    const float ax = c->pts[2 * idx[a]], ay = c->pts[2 * idx[a] + 1];
    float acc = 0.f;
    for (int b = 0; b < n; ++b) {
      const float dx = ax - c->pts[2 * idx[b]], dy = ay - c->pts[2 * idx[b] + 1];
      acc += sqrtf(dx * dx + dy * dy);
    }
    if (acc < best) best = acc, med = a;
  }
  c->out[c->n_out++] = idx[med];
}

static void ds_build(ds_ctx *c, int *idx, int n, float cx, float cy, float radius, int *scratch) {
  if (n == 0) return;
  if (((double)radius * 2.0 <= (double)c->max_size) || n <= 1) {
    ds_leaf(c, idx, n);
    return;
  }
  /* stable 4-way partition by child id */
  int cnt[4] = {0, 0, 0, 0};
  for (int i = 0; i < n; ++i) {
    const int id = (c->pts[2 * idx[i]] > cx ? 1 : 0) | (c->pts[2 * idx[i] + 1] > cy ? 2 : 0);
    cnt[id]++;
  }
  int off[5] = {0, cnt[0], cnt[0] + cnt[1], cnt[0] + cnt[1] + cnt[2], n};
  int cur[4] = {off[0], off[1], off[2], off[3]};
  for (int i = 0; i < n; ++i) {
    const int id = (c->pts[2 * idx[i]] > cx ? 1 : 0) | (c->pts[2 * idx[i] + 1] > cy ? 2 : 0);
    scratch[cur[id]++] = idx[i];
  }
  memcpy(idx, scratch, sizeof(int) * (size_t)n);
  const float hr = radius * 0.5f;
  for (int k = 0; k < 4; ++k) {
    const float ncx = cx + ((k & 1) ? hr : -hr), ncy = cy + ((k & 2) ? hr : -hr);
    ds_build(c, idx + off[k], cnt[k], ncx, ncy, hr, scratch + off[k]);
  }
}

/* pts [n][2]; out_idx[n] <- indices of the kept points in output order; returns their number */
int orc_downsample(const float *pts, int n, float resolution, int32_t *out_idx) {
  if (n == 0) return 0;
  float minx = pts[0], maxx = pts[0], miny = pts[1], maxy = pts[1];
  for (int i = 1; i < n; ++i) {
    const float x = pts[2 * i], y = pts[2 * i + 1];
    if (x < minx) minx = x;
    if (x > maxx) maxx = x;
    if (y < miny) miny = y;
    if (y > maxy) maxy = y;
  }
  const float rx = maxx - minx, ry = maxy - miny;
  const float cx = minx + rx * 0.5f, cy = miny + ry * 0.5f;
  float radius = rx;
  if (radius < ry) radius = ry;
  radius *= 0.5f;
  int *idx = (int *)malloc(sizeof(int) * (size_t)n), *scratch = (int *)malloc(sizeof(int) * (size_t)n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  ds_ctx c = {pts, (int *)out_idx, 0, resolution};
  ds_build(&c, idx, n, cx, cy, radius, scratch);
  free(idx);
  free(scratch);
  return c.n_out;
}
