/* TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
 *
 * CPU restatement of the scan matcher behind `bruce_slam.pcl.ICP.compute`
 * (bruce_slam/src/bruce_slam/cpp/pcl.cpp:198-212 -> libpointmatcher PM::ICP::operator()) under
 * the configuration the reference ships, bruce_slam/config/icp.yaml:1-31:
 *     matcher            KDTreeMatcher  knn 1, epsilon 0, maxDist 10.0          (:5-9)
 *     outlierFilters     MaxDistOutlierFilter 3.0  x  TrimmedDistOutlierFilter 0.8 (:11-15)
 *     errorMinimizer     PointToPointErrorMinimizer                              (:17-20)
 *     checkers           Counter(40) + Differential(0.01 rad, 0.1 m, smooth 4)   (:22-28)
 * libpointmatcher (README.md:51-55 pins commit d478ef2e) and libnabo are NOT vendored in
 * /root/reference and not installed here: **parity unpinned**.  This file restates the published
 * algorithm of PointMatcher<float>::ICP (ICP.cpp compute / computeWithTransformedReference), its
 * KDTreeMatcher, MaxDist / TrimmedDist outlier filters, PointToPoint error minimiser and
 * Counter / Differential transformation checkers, all in float32 (pcl.cpp:12 `PointMatcher<float>`):
 *
 *  1. reference (target) is centred on its mean; T_refIn_refMean = [I | mean].
 *  2. reading (source) <- (T_refIn_refMean^-1 * T_init) * reading, once.
 *  3. loop: step = T_iter * reading; nearest reference point of every step point (squared float32
 *     distance, accepted when <= maxDist^2, else "no match": dist = +inf);
 *     weights = [dist <= maxDistOutlier^2] * [dist <= q], q = element floor(float(n_finite)*ratio) of
 *     the ascending finite distances ("no outlier to filter" if there is none);
 *     kept pairs = weight != 0 ("ErrorMnimizer: no point to minimize" if none; their number is the
 *     inlier count); rigid 2-D fit of kept step points onto their matches: weighted means, 2x2
 *     cross-covariance of the centred pairs, R = U V^T (closed form for 2-D: angle atan2(m10-m01,
 *     m00+m11)), t = mean_ref - R mean_read; T_iter <- [R t] * T_iter;
 *     Counter: stop once 40 iterations were made.  Differential: keep rotation (as quaternion) and
 *     translation of every T_iter; once more than `smooth` are stored, average |angular distance| and
 *     |translation step| over the last `smooth` steps; stop when both fall below the limits.
 *  4. result = T_refIn_refMean * T_iter * (T_refIn_refMean^-1 * T_init).
 *
 * X1 -- the minimiser the YAML keeps commented out (icp.yaml:18-19 `PointToPlaneErrorMinimizer force2D 1`),
 * selected by orc_icp_params.minimizer = 1.  It needs a "normals" descriptor on the reference, which upstream
 * comes from a `SurfaceNormalDataPointsFilter` in referenceDataPointsFilters (knn, default 5):
 *   normals  for every reference point: its knn nearest reference points (itself included, ascending squared
 *            float32 distance, ties by lower index); mean = sum / k; C = sum (p-mean)(p-mean)^T / k; normal = unit
 *            eigenvector of the smaller eigenvalue of C (closed form for the symmetric 2x2; upstream calls
 *            Eigen::EigenSolver, whose sign and rounding are build details -- the fit is invariant to the sign);
 *            C == 0 (all neighbours coincide): normal stays (0,0) like upstream's rank test.  Computed in the
 *            centred frame (upstream: before centring; normals do not depend on a translation).
 *   fit      over the kept pairs (step point q, matched reference r, its normal n):  c = q.x n.y - q.y n.x,
 *            F = (c, n.x, n.y),  A = sum F F^T,  b = -sum F ((q-r).n),  A x = b by Cholesky (A.llt()),
 *            T_step = [Rotation2D(x0) | (x1, x2)]  (PointToPlane.cpp compute_in_place, 2-D branch).
 *            A pivot that is not positive leaves that unknown at 0 (upstream: minimum-norm solution).
 *
 * Every sum over points is a sequential float32 sum in point order (upstream sums with Eigen's
 * vectorised reductions, whose order is a build detail); products a*b+c are NOT contracted
 * (-ffp-contract=off) to mirror an SSE2 build.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct nn_grid nn_grid;
nn_grid *nn_grid_build(const float *pts, int n);
void nn_grid_free(nn_grid *g);
int nn_grid_query(const nn_grid *g, float qx, float qy, float max_d2, float *d2_out);

typedef struct {
  float matcher_max_dist;  /* KDTreeMatcher maxDist (10.0) */
  float outlier_max_dist;  /* MaxDistOutlierFilter maxDist (3.0); <= 0 disables the filter */
  float trim_ratio;        /* TrimmedDistOutlierFilter ratio (0.8); < 0 disables the filter */
  int max_iterations;      /* CounterTransformationChecker maxIterationCount (40) */
  float min_diff_rot;      /* DifferentialTransformationChecker minDiffRotErr (0.01) */
  float min_diff_trans;    /* ... minDiffTransErr (0.1) */
  int smooth_length;       /* ... smoothLength (4); 0 disables the differential checker */
  int flags;               /* bit 0: MaxDist filter compares the squared distance with maxDist itself */
  int minimizer;           /* 0 PointToPointErrorMinimizer (shipped), 1 PointToPlaneErrorMinimizer force2D (X1) */
  int normals_knn;         /* SurfaceNormalDataPointsFilter knn on the reference (5) */
} orc_icp_params;

enum { ORC_ICP_OK = 0, ORC_ICP_NO_OUTLIER = 1, ORC_ICP_NO_POINT = 2, ORC_ICP_NAN_ROT = 3, ORC_ICP_NAN_TRANS = 4,
       ORC_ICP_NOT_RIGID = 5, ORC_ICP_EMPTY_REF = 6 };

static void mat3_mul(const float *a, const float *b, float *c) { /* coefficient-wise, k ascending */
  float r[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float acc = a[i * 3 + 0] * b[0 * 3 + j];
      acc = acc + a[i * 3 + 1] * b[1 * 3 + j];
      acc = acc + a[i * 3 + 2] * b[2 * 3 + j];
      r[i * 3 + j] = acc;
    }
  memcpy(c, r, sizeof(r));
}

static int cmp_f(const void *a, const void *b) {
  const float x = *(const float *)a, y = *(const float *)b;
  return (x > y) - (x < y);
}

/* z and w of the unit quaternion of the 2-D rotation block embedded in a 3x3 identity (Eigen's
 * rotation-matrix -> quaternion conversion, float32) */
static void rot_to_quat(const float *T, float *qw, float *qz) {
  const float m00 = T[0], m01 = T[1], m10 = T[3], m11 = T[4];
  float t = m00 + m11 + 1.0f;
  if (t > 0.f) {
    t = sqrtf(t + 1.0f);
    *qw = 0.5f * t;
    t = 0.5f / t;
    *qz = (m10 - m01) * t;
  } else {
    t = sqrtf(1.0f - m00 - m11 + 1.0f);
    *qz = 0.5f * t;
    t = 0.5f / t;
    *qw = (m10 - m01) * t;
  }
}

/* SurfaceNormalDataPointsFilter (knn, keepNormals) on a 2-D cloud; brute-force neighbours */
void orc_surface_normals(const float *pts, int n, int knn, float *normals) {
  enum { KMAX = 32 };
  if (knn > KMAX) knn = KMAX;
  if (knn < 1) knn = 1;
  for (int i = 0; i < n; ++i) {
    float bd[KMAX];
    int bi[KMAX], k = 0;
    const float qx = pts[2 * i], qy = pts[2 * i + 1];
    for (int j = 0; j < n; ++j) {
      const float dx = qx - pts[2 * j], dy = qy - pts[2 * j + 1];
      const float d2 = dx * dx + dy * dy;
      if (k == knn && !(d2 < bd[k - 1])) continue; /* j ascending: an equal distance never displaces a lower index */
      int at = k < knn ? k++ : knn - 1;
      while (at > 0 && d2 < bd[at - 1]) bd[at] = bd[at - 1], bi[at] = bi[at - 1], --at;
      bd[at] = d2, bi[at] = j;
    }
    float sx = 0.f, sy = 0.f;
    for (int j = 0; j < k; ++j) sx += pts[2 * bi[j]], sy += pts[2 * bi[j] + 1];
    const float kf = (float)k, mx = sx / kf, my = sy / kf;
    float a = 0.f, b = 0.f, c = 0.f;
    for (int j = 0; j < k; ++j) {
      const float ux = pts[2 * bi[j]] - mx, uy = pts[2 * bi[j] + 1] - my;
      a += ux * ux, b += ux * uy, c += uy * uy;
    }
    a /= kf, b /= kf, c /= kf;
    float nx = 0.f, ny = 0.f;
    if (a != 0.f || b != 0.f || c != 0.f) {
      /* major axis u of [[a b][b c]]: (r + d, b) if d >= 0 else (b, r - d), d = (a-c)/2, r = sqrt(d^2 + b^2);
       * the normal is u turned by 90 degrees; isotropic (r == 0): Eigen's identity basis, first column */
      const float d = 0.5f * (a - c), r = sqrtf(d * d + b * b);
      if (r == 0.f) {
        nx = 1.f, ny = 0.f;
      } else {
        const float ux = d >= 0.f ? r + d : b, uy = d >= 0.f ? b : r - d;
        const float len = sqrtf(ux * ux + uy * uy);
        nx = -(uy / len), ny = ux / len;
      }
    }
    normals[2 * i] = nx, normals[2 * i + 1] = ny;
  }
}

/* x = A^-1 b for the symmetric 3x3 A (upper triangle a00 a01 a02 a11 a12 a22) by Cholesky, float32 */
static void solve_llt3(const float *A, const float *b, float *x) {
  const float a00 = A[0], a01 = A[1], a02 = A[2], a11 = A[3], a12 = A[4], a22 = A[5];
  float l00 = 0.f, l10 = 0.f, l20 = 0.f, l11 = 0.f, l21 = 0.f, l22 = 0.f;
  const int p0 = a00 > 0.f;
  if (p0) l00 = sqrtf(a00), l10 = a01 / l00, l20 = a02 / l00;
  const float d1 = a11 - l10 * l10;
  const int p1 = d1 > 0.f;
  if (p1) l11 = sqrtf(d1), l21 = (a12 - l20 * l10) / l11;
  const float d2 = (a22 - l20 * l20) - l21 * l21;
  const int p2 = d2 > 0.f;
  if (p2) l22 = sqrtf(d2);
  /* L y = b, L^T x = y; an unknown whose pivot failed stays 0 */
  const float y0 = p0 ? b[0] / l00 : 0.f;
  const float y1 = p1 ? (b[1] - l10 * y0) / l11 : 0.f;
  const float y2 = p2 ? ((b[2] - l20 * y0) - l21 * y1) / l22 : 0.f;
  x[2] = p2 ? y2 / l22 : 0.f;
  x[1] = p1 ? (y1 - l21 * x[2]) / l11 : 0.f;
  x[0] = p0 ? ((y0 - l10 * x[1]) - l20 * x[2]) / l00 : 0.f;
}

int orc_icp(const float *src, int ns, const float *tgt, int nt, const float *guess /* 3x3 row-major */,
            const orc_icp_params *prm, float *T_out, int *iters_out, int *inliers_out) {
  memcpy(T_out, guess, 9 * sizeof(float));
  *iters_out = 0;
  *inliers_out = 0;
  if (nt <= 0) return ORC_ICP_EMPTY_REF;
  /* RigidTransformation::checkParameters on the initial transform */
  {
    const float det = guess[0] * guess[4] - guess[1] * guess[3];
    if (fabsf(1.0f - det) > 0.001f || det != det) return ORC_ICP_NOT_RIGID;
  }
  /* 1. centre the reference */
  float sx = 0.f, sy = 0.f;
  for (int i = 0; i < nt; ++i) sx += tgt[2 * i], sy += tgt[2 * i + 1];
  const float mx = sx / (float)nt, my = sy / (float)nt;
  float *ref = (float *)malloc(sizeof(float) * 2 * (size_t)nt);
  for (int i = 0; i < nt; ++i) ref[2 * i] = tgt[2 * i] - mx, ref[2 * i + 1] = tgt[2 * i + 1] - my;
  nn_grid *grid = nn_grid_build(ref, nt);
  float *normals = NULL;
  if (prm->minimizer == 1) {
    normals = (float *)malloc(sizeof(float) * 2 * (size_t)nt);
    orc_surface_normals(ref, nt, prm->normals_knn, normals);
  }

  /* 2. reading into the centred frame: T_refMean_dataIn = [I | -mean] * T_init */
  const float Tmean[9] = {1, 0, mx, 0, 1, my, 0, 0, 1}, Tmean_inv[9] = {1, 0, -mx, 0, 1, -my, 0, 0, 1};
  float T0[9];
  mat3_mul(Tmean_inv, guess, T0);
  float *reading = (float *)malloc(sizeof(float) * 2 * (size_t)(ns > 0 ? ns : 1));
  for (int i = 0; i < ns; ++i) {
    const float x = src[2 * i], y = src[2 * i + 1];
    reading[2 * i] = (T0[0] * x + T0[1] * y) + T0[2];
    reading[2 * i + 1] = (T0[3] * x + T0[4] * y) + T0[5];
  }
  float *step = (float *)malloc(sizeof(float) * 2 * (size_t)(ns > 0 ? ns : 1));
  float *dist = (float *)malloc(sizeof(float) * (size_t)(ns > 0 ? ns : 1));
  float *sorted = (float *)malloc(sizeof(float) * (size_t)(ns > 0 ? ns : 1));
  int *match = (int *)malloc(sizeof(int) * (size_t)(ns > 0 ? ns : 1));

  float Ti[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  const float max_d2 = prm->matcher_max_dist * prm->matcher_max_dist;
  const float out_d2 = (prm->flags & 1) ? prm->outlier_max_dist : prm->outlier_max_dist * prm->outlier_max_dist;
  /* differential checker history (ring of the last smooth_length + 1 entries is enough) */
  enum { HMAX = 64 };
  float hq_w[HMAX], hq_z[HMAX], ht_x[HMAX], ht_y[HMAX];
  int hn = 0;
  const int smooth = prm->smooth_length > HMAX - 1 ? HMAX - 1 : prm->smooth_length;
  rot_to_quat(Ti, &hq_w[0], &hq_z[0]);
  ht_x[0] = Ti[2], ht_y[0] = Ti[5];
  hn = 1;

  int status = ORC_ICP_OK, iterate = 1, count = 0, inliers = 0;
  while (iterate) {
    for (int i = 0; i < ns; ++i) {
      const float x = reading[2 * i], y = reading[2 * i + 1];
      step[2 * i] = (Ti[0] * x + Ti[1] * y) + Ti[2];
      step[2 * i + 1] = (Ti[3] * x + Ti[4] * y) + Ti[5];
    }
    int n_fin = 0;
    for (int i = 0; i < ns; ++i) {
      match[i] = nn_grid_query(grid, step[2 * i], step[2 * i + 1], max_d2, &dist[i]);
      if (match[i] >= 0) sorted[n_fin++] = dist[i];
    }
    float limit = INFINITY;
    if (prm->trim_ratio >= 0.f) {
      if (n_fin == 0) {
        status = ORC_ICP_NO_OUTLIER;
        break;
      }
      if (prm->trim_ratio == 1.0f) {
        limit = sorted[0];
        for (int i = 1; i < n_fin; ++i)
          if (sorted[i] > limit) limit = sorted[i];
      } else {
        qsort(sorted, (size_t)n_fin, sizeof(float), cmp_f);
        limit = sorted[(size_t)((float)n_fin * prm->trim_ratio)];
      }
    }
    /* kept pairs and their weighted (0/1) means */
    int n_keep = 0;
    float srx = 0.f, sry = 0.f, sfx = 0.f, sfy = 0.f;
    for (int i = 0; i < ns; ++i) {
      int keep = match[i] >= 0;
      if (prm->outlier_max_dist > 0.f) keep = keep && (dist[i] <= out_d2);
      if (prm->trim_ratio >= 0.f) keep = keep && (dist[i] <= limit);
      if (!keep) {
        match[i] = -1; /* weight 0 */
        continue;
      }
      ++n_keep;
      srx += step[2 * i], sry += step[2 * i + 1];
      sfx += ref[2 * match[i]], sfy += ref[2 * match[i] + 1];
    }
    if (n_keep == 0) {
      status = ORC_ICP_NO_POINT;
      break;
    }
    inliers = n_keep;
    float dT[9];
    if (prm->minimizer == 1) {
      float A[6] = {0, 0, 0, 0, 0, 0}, bb[3] = {0, 0, 0}, x[3];
      for (int i = 0; i < ns; ++i) {
        if (match[i] < 0) continue;
        const float qx = step[2 * i], qy = step[2 * i + 1];
        const float nx = normals[2 * match[i]], ny = normals[2 * match[i] + 1];
        const float cr = qx * ny - qy * nx;
        const float dp = (qx - ref[2 * match[i]]) * nx + (qy - ref[2 * match[i] + 1]) * ny;
        A[0] += cr * cr, A[1] += cr * nx, A[2] += cr * ny, A[3] += nx * nx, A[4] += nx * ny, A[5] += ny * ny;
        bb[0] += cr * dp, bb[1] += nx * dp, bb[2] += ny * dp;
      }
      bb[0] = -bb[0], bb[1] = -bb[1], bb[2] = -bb[2];
      solve_llt3(A, bb, x);
      const float c = (float)cos((double)x[0]), s = (float)sin((double)x[0]);
      const float d[9] = {c, -s, x[1], s, c, x[2], 0, 0, 1};
      memcpy(dT, d, sizeof(d));
    } else {
    const float winv = 1.0f / (float)n_keep;
    const float mrx = srx * winv, mry = sry * winv, mfx = sfx * winv, mfy = sfy * winv;
    float m00 = 0.f, m01 = 0.f, m10 = 0.f, m11 = 0.f;
    for (int i = 0; i < ns; ++i) {
      if (match[i] < 0) continue;
      const float px = step[2 * i] - mrx, py = step[2 * i + 1] - mry;
      const float qx = ref[2 * match[i]] - mfx, qy = ref[2 * match[i] + 1] - mfy;
      m00 += qx * px, m01 += qx * py, m10 += qy * px, m11 += qy * py;
    }
    const float a = m00 + m11, b = m10 - m01;
    const float h = sqrtf(a * a + b * b);
    float c = 1.f, s = 0.f;
    if (h > 0.f) c = a / h, s = b / h;
    const float tx = mfx - (c * mrx + (-s) * mry), ty = mfy - (s * mrx + c * mry);
    const float d[9] = {c, -s, tx, s, c, ty, 0, 0, 1};
    memcpy(dT, d, sizeof(d));
    }
    mat3_mul(dT, Ti, Ti);

    /* checkers: Counter first, then Differential */
    ++count;
    int counter_stop = 0;
    if (count >= prm->max_iterations) iterate = 0, counter_stop = 1;
    if (!counter_stop && smooth > 0) {
      if (hn == HMAX) { /* drop the oldest */
        memmove(hq_w, hq_w + 1, sizeof(float) * (HMAX - 1));
        memmove(hq_z, hq_z + 1, sizeof(float) * (HMAX - 1));
        memmove(ht_x, ht_x + 1, sizeof(float) * (HMAX - 1));
        memmove(ht_y, ht_y + 1, sizeof(float) * (HMAX - 1));
        hn = HMAX - 1;
      }
      rot_to_quat(Ti, &hq_w[hn], &hq_z[hn]);
      ht_x[hn] = Ti[2], ht_y[hn] = Ti[5];
      ++hn;
      float vr = 0.f, vt = 0.f;
      if (hn > smooth) {
        for (int i = hn - 1; i >= hn - smooth; --i) {
          /* d = q_i * conj(q_{i-1}); angular distance = 2 atan2(|vec|, |w|) */
          const float dw = hq_w[i] * hq_w[i - 1] + hq_z[i] * hq_z[i - 1];
          const float dz = hq_z[i] * hq_w[i - 1] - hq_w[i] * hq_z[i - 1];
          vr += fabsf(2.0f * atan2f(fabsf(dz), fabsf(dw)));
          const float ex = ht_x[i] - ht_x[i - 1], ey = ht_y[i] - ht_y[i - 1];
          vt += fabsf(sqrtf(ex * ex + ey * ey));
        }
        vr /= (float)smooth;
        vt /= (float)smooth;
        if (vr < prm->min_diff_rot && vt < prm->min_diff_trans) iterate = 0;
      }
      if (vr != vr) {
        status = ORC_ICP_NAN_ROT;
        break;
      }
      if (vt != vt) {
        status = ORC_ICP_NAN_TRANS;
        break;
      }
    }
  }
  *iters_out = count;
  *inliers_out = inliers;
  if (status == ORC_ICP_OK) {
    float tmp[9];
    mat3_mul(Tmean, Ti, tmp);
    mat3_mul(tmp, T0, T_out);
  }
  nn_grid_free(grid);
  free(ref), free(reading), free(step), free(dist), free(sorted), free(match), free(normals);
  return status;
}
