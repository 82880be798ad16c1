/* TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
 *
 * CPU restatement ("port") of the reference's 1-D range-axis CFAR detector,
 * bruce_slam/src/bruce_slam/cpp/cfar.cpp:
 *     ca   :10-28    soca :30-51    goca :53-74    os   :76-96
 *     ca2  :98-118   soca2:120-143  goca2:145-168  os2  :170-192
 * One routine parameterised by `alg`; what is kept identical to the reference:
 *   - per beam (column), per range bin (row) in [T+G, R-T-G): one ascending scan
 *     i = row-T-G .. row+T+G of the column;
 *   - float32 accumulators that start at 0 and take cells in ascending i
 *     (a single accumulator for CA, one each for leading/lagging in SOCA/GOCA);
 *   - the guard test `abs(i-row) > G` (CA, OS) and `(i-row) > G` / `< -G`
 *     (SOCA, GOCA);
 *   - thresholds evaluated in double exactly as written there:
 *         CA    tau * sum / (2.0 * T)      (cfar.cpp:24)
 *         SOCA  tau * min(lead,lag) / T    (cfar.cpp:46-47; T promoted int->double)
 *         GOCA  tau * max(lead,lag) / T    (cfar.cpp:69-70)
 *         OS    tau * kth_smallest(train)  (cfar.cpp:91-92; k is 0-based)
 *     and compared with strict `>` against the float cell promoted to double;
 *   - the "2" variants store that double rounded to float32 (cfar.cpp:114);
 *   - border rows stay 0.
 * Pinned against the unmodified reference source compiled in oracle/_ref (see
 * oracle/Makefile, tests/test_oracle_cfar.py).
 *
 * Layout: `img` is addressed through element strides so the same routine reads a
 * row-major numpy image or a column-major copy; outputs are row-major [R][B].
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ORC_CA = 0, ORC_SOCA = 1, ORC_GOCA = 2, ORC_OS = 3 };

static int cmp_float(const void *a, const void *b) {
  float x = *(const float *)a, y = *(const float *)b;
  return (x > y) - (x < y);
}

/* returns 0, or -1 on bad arguments */
int orc_cfar(int alg, const float *img, long stride_r, long stride_c, int R, int B, int T, int G,
             int k, double tau, uint8_t *mask, float *thr /* may be NULL */) {
  if (alg < 0 || alg > 3 || T < 0 || G < 0) return -1;
  if (alg == ORC_OS && (k < 0 || k >= 2 * T)) return -1;
  memset(mask, 0, (size_t)R * (size_t)B);
  if (thr) memset(thr, 0, (size_t)R * (size_t)B * sizeof(float));
  const int half = T + G;
  float *train = (float *)malloc(sizeof(float) * (size_t)(2 * T > 0 ? 2 * T : 1));
  if (!train) return -1;

  for (int col = 0; col < B; ++col) {
    const float *colp = img + (long)col * stride_c;
    for (int row = half; row < R - half; ++row) {
      const float cut = colp[(long)row * stride_r];
      double t;
      if (alg == ORC_CA) {
        float acc = 0;
        for (int i = row - half; i < row + half + 1; ++i)
          if (abs(i - row) > G) acc += colp[(long)i * stride_r];
        t = tau * acc / (2.0 * T);
      } else if (alg == ORC_OS) {
        int n = 0;
        for (int i = row - half; i < row + half + 1; ++i)
          if (abs(i - row) > G) train[n++] = colp[(long)i * stride_r];
        qsort(train, (size_t)n, sizeof(float), cmp_float); /* value of nth_element(k) */
        t = tau * train[k];
      } else {
        float lead = 0.0f, lag = 0.0f;
        for (int i = row - half; i < row + half + 1; ++i) {
          if ((i - row) > G)
            lag += colp[(long)i * stride_r];
          else if ((i - row) < -G)
            lead += colp[(long)i * stride_r];
        }
        float pick;
        if (alg == ORC_SOCA)
          pick = lag < lead ? lag : lead; /* std::min(lead, lag) */
        else
          pick = lead < lag ? lag : lead; /* std::max(lead, lag) */
        t = tau * pick / T;
      }
      mask[(size_t)row * B + col] = (uint8_t)(cut > t);
      if (thr) thr[(size_t)row * B + col] = (float)t;
    }
  }
  free(train);
  return 0;
}

/* Convenience used by the CPU baseline timing: uint8 image -> float32 (the
 * dtype conversion pybind11 performs at cfar.cpp:10's `const MatrixXf&`), CFAR,
 * then the node's amplitude gate `peaks &= img > threshold`
 * (bruce_slam/src/bruce_slam/feature_extraction.py:223-224). */
int orc_cfar_u8(int alg, const uint8_t *img, int R, int B, int T, int G, int k, double tau,
                int threshold /* <0: no gate */, uint8_t *mask) {
  float *f = (float *)malloc(sizeof(float) * (size_t)R * (size_t)B);
  if (!f) return -1;
  /* column-major copy like the Eigen matrix pybind builds */
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < B; ++c) f[(size_t)c * R + r] = (float)img[(size_t)r * B + c];
  int rc = orc_cfar(alg, f, 1, R, R, B, T, G, k, tau, mask, NULL);
  free(f);
  if (rc) return rc;
  if (threshold >= 0)
    for (size_t i = 0; i < (size_t)R * (size_t)B; ++i) mask[i] &= (uint8_t)(img[i] > threshold);
  return 0;
}
