// Batched per-keyframe front end: polar frames -> CFAR -> Cartesian cloud -> filters -> sequential
// scan matching against the window of previous frames, one C-ABI call per batch.
//
// This is the GPU form of what the reference does one message at a time in
// FeatureExtraction.callback (feature_extraction.py:196-252) followed, per keyframe, by
// SLAM.add_sequential_scan_matching (slam.py:718-837): source = the new frame's cloud, target =
// get_points(previous `window` frames, expressed in the previous frame) (slam.py:632-633,229-292:
// Keyframe.transform_points + pcl.downsample), guess = odometry between the two poses, then
// pcl.ICP.compute.  The pose-graph side (ISAM2) stays on the CPU and consumes the SE(2) results.
// Frames of one batch are independent given their odometry poses, which is what lets a backlog of
// keyframes be matched in parallel (and sharded over GPUs by the caller).
//
// Device layout: every per-frame cloud lives at a fixed stride (`cap_points` rows) with a count
// array, so all stages run without packing or host round trips; the host flavour overlaps the
// host->device copy of frame chunk k+1 with the kernels of chunk k on two streams.
#include <vector>

#include "common.cuh"

namespace sfe {
int cfar_run(sfe_ctx *ctx, const void *img, int dtype, int F, int R, int B, int alg, int T, int G, int k,
             double tau, int gate_on, double gate, uint8_t *mask, float *thr, uint32_t *bits, int force_exact);
int cart_points_run(sfe_ctx *ctx, const sfe_maps *m, const uint8_t *mask, const uint32_t *bits, int F, int cap,
                    int32_t *ij, float *xy, int32_t *count);
int downsample_run(sfe_ctx *ctx, const float *pts, const int *off, const int *cnt, int n_clouds, int dim, int n_max,
                   float resolution, float *out_pts, int32_t *out_idx, int32_t *out_count, int n_split = 0);
int remove_outlier_run(sfe_ctx *ctx, const float *pts, const int *off, const int *cnt, int n_clouds, int dim, int n_max,
                       double radius, int min_points, float *out_pts, int32_t *out_idx, int32_t *out_count,
                       int n_split = 0);
int icp_run(sfe_ctx *ctx, const sfe_icp_params *prm, const float *src_pts, const int *src_off, const int *src_cnt,
            const float *tgt_pts, const int *tgt_off, const int *tgt_cnt, int min_points, const int *src_id,
            const int *tgt_id, int P, int ns_max, int nt_max, const float *guess, float *T_out, int *iters,
            int *inliers, int *status, const int *raw_cnt = nullptr, int raw_cap = 0, int class_mode = 0,
            int class_ns = 0, int class_nt = 0, int force_threads = 0);

// Size class of the first ICP launch (fe_match).  Measured on the config-4 replay (filtered clouds: median 326,
// max 639 points; window submaps <= 1808 raw points): one launch sized for the capacities (1024 / 3072, 4 CTAs per
// SM) 2.40 ms per 4096 frames with 256 threads, 2.59 ms with 128; this class with 128 threads (6 CTAs per SM) 2.11 ms;
// a class that leaves 14 % of the frames to the second launch (512 points) 3.11 ms -- the tail of a nearly empty
// launch costs a full problem latency.
constexpr int FE_ICP_SMALL_SRC = 640, FE_ICP_SMALL_TGT = 1536, FE_ICP_SMALL_THREADS = 128;
// Size classes of the cloud filters (cloud.cu: two launches by cloud size).  Config-4 replay: 3-5 k raw Cartesian
// points per frame, ~580 after the voxel filter, ~400 after the outlier filter, window submaps <= 1.8 k raw points.
// Measured (ms per 4096 frames): outlier filter 0.145 -> 0.113 and submap down-sampling 0.241 -> 0.215 with a
// small-cloud class; the per-frame down-sampling is best left as one launch (0.79; 0.90 / 0.87 with a class
// boundary at 4608 / 5632 points: too many frames land in the second, poorly filled launch).
constexpr int FE_DS_SPLIT = 0, FE_RO_SPLIT = 1024, FE_SUBMAP_SPLIT = 2048;

// T_ab = pose_a^-1 * pose_b as float32 3x3 (gtsam Pose2::between, then matrix().astype(float32)).
// Evaluated on the host in double (libm), so the float32 matrices the kernels see are the ones a
// Python caller would compute.  Rotation as gtsam forms it: Rot2 holds (cos, sin) and between() multiplies
// r_a^-1 * r_b, i.e. c = ca*cb + sa*sb, s = ca*sb - sa*cb (not cos/sin of the angle difference; the two differ by
// ~1e-16, which matters only on a float32 rounding boundary); translation = r_a.unrotate(t_b - t_a).
static void pose_between(const double *a, const double *b, float *T) {
  const double ca = cos(a[2]), sa = sin(a[2]), cb = cos(b[2]), sb = sin(b[2]);
  const double dx = b[0] - a[0], dy = b[1] - a[1];
  const double x = ca * dx + sa * dy, y = -sa * dx + ca * dy;
  const double c = ca * cb + sa * sb, s = ca * sb - sa * cb;
  T[0] = (float)c, T[1] = (float)-s, T[2] = (float)x;
  T[3] = (float)s, T[4] = (float)c, T[5] = (float)y;
  T[6] = 0.f, T[7] = 0.f, T[8] = 1.f;
}

// A6: the keyframe cloud as SLAM holds it.  The feature node publishes xyz = [p0, 0, p1]
// (feature_extraction.py:182); SLAM reads np.c_[x, -z] = (p0, -p1) (slam_ros.py:169-170).  Negation is exact, so
// flipping the float32 cloud after the filters equals what the reference's float32 message round trip gives.
__global__ void flip_lateral_kernel(float *__restrict__ xy, const int32_t *__restrict__ cnt, int cap, int n) {
  const int f = blockIdx.x;
  if (f >= n) return;
  const int k = min(cnt[f], cap);
  float *p = xy + 2 * (size_t)f * cap;
  for (int i = threadIdx.x; i < k; i += blockDim.x) p[2 * i + 1] = -p[2 * i + 1];
}

// offsets helper: off[i] = i * stride
__global__ void fill_offsets_kernel(int32_t *off, int n, int stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n) off[i] = i * stride;
}

// Target assembly (slam.py:229-292 without the final downsample): for frame i gather the clouds of frames
// i-window .. i-1, each moved into frame (i-1)'s coordinates with Keyframe.transform_points
// (slam_objects.py:178-198: float32 points @ R^T + t), concatenated oldest first.  rel[i][w] is the
// float32 transform of window slot w (frame i-window+w), rel[i][window] the ICP guess.  One CTA per frame.
// Frames of the PREVIOUS call kept on the device (sfe_frontend_set_carry): `carry_n` clouds in carry_xy / carry_cnt
// precede frame 0 of this batch, so virtual frame u = carry_n + (batch index).
__global__ void __launch_bounds__(256)
    assemble_targets_kernel(const float *__restrict__ clouds, const int32_t *__restrict__ counts, int cap,
                            const float *__restrict__ rel, int f0, int F, int window, float *__restrict__ tgt,
                            int32_t *__restrict__ tgt_count, int tgt_cap, float *__restrict__ guess,
                            const float *__restrict__ carry_xy, const int32_t *__restrict__ carry_cnt, int carry_n) {
  const int i = f0 + blockIdx.x;  // frames [f0, F) of the batch
  if (i >= F) return;
  const float *myrel = rel + (size_t)i * (window + 1) * 9;
  if (threadIdx.x < 9) guess[9 * (size_t)i + threadIdx.x] = myrel[window * 9 + threadIdx.x];
  float *out = tgt + 2 * (size_t)i * tgt_cap;
  int base = 0;
  for (int w = 0; w < window; ++w) {
    const int u = i + carry_n - window + w;  // virtual index of the window frame
    if (u < 0) continue;
    const float *T = myrel + w * 9;
    const float t0 = T[0], t1 = T[1], t2 = T[2], t3 = T[3], t4 = T[4], t5 = T[5];
    const bool carried = u < carry_n;
    const int n = min(carried ? carry_cnt[u] : counts[u - carry_n], cap);
    const float *src = carried ? carry_xy + 2 * (size_t)u * cap : clouds + 2 * (size_t)(u - carry_n) * cap;
    for (int j = threadIdx.x; j < n && base + j < tgt_cap; j += blockDim.x) {
      const float x = src[2 * j], y = src[2 * j + 1];
      // numpy's float32 `points.dot(R.T) + t` (slam_objects.py:196): fma(y, r01, x * r00) + tx -- the same
      // convention as the costmap kernels (globalinit.cu), pinned against numpy in tests/test_oracle_globalinit.py
      out[2 * (base + j)] = __fadd_rn(__fmaf_rn(y, t1, __fmul_rn(x, t0)), t2);
      out[2 * (base + j) + 1] = __fadd_rn(__fmaf_rn(y, t4, __fmul_rn(x, t3)), t5);
    }
    base = min(base + n, tgt_cap);
  }
  if (threadIdx.x == 0) tgt_count[i] = base;
}

}  // namespace sfe

using namespace sfe;

struct sfe_frontend {
  sfe_ctx *ctx;
  const sfe_maps *maps;
  sfe_frontend_params p;
  int max_frames;
  // device buffers (sized for max_frames)
  uint8_t *frames;      // [max_frames][R][B]      (host flavour only)
  uint32_t *bits;       // [max_frames][R][wpr]
  int32_t *ij;          // [max_frames][cap][2]
  float *xy_a, *xy_b;   // [max_frames][cap][2] ping-pong
  int32_t *idx;         // [max_frames][cap]
  int32_t *cnt_a, *cnt_b, *cnt_c;
  int32_t *off_pts, *off_tgt;
  float *tgt_a, *tgt_b; // [max_frames][window*cap][2]
  int32_t *tgt_idx, *tcnt_a, *tcnt_b;
  float *rel;           // [max_frames][window+1][9] relative transforms (window slots, then the guess)
  float *rel_host[2];   // pinned staging of the same, double-buffered (no stream sync per call)
  cudaEvent_t ev_rel[2];
  int rel_turn;
  float *guess, *T;
  int32_t *iters, *inliers, *status;
  // continuation across calls: the last `window` clouds (+ host poses) of the previous call
  int carry_on, carry_n;
  float *carry_xy;       // [window][cap][2]
  int32_t *carry_cnt;    // [window]
  double *carry_poses;   // host [window][3]
  cudaStream_t copy_stream;
  cudaEvent_t ev_copy[2], ev_done;
  // optional per-stage timing (CUDA events on the launch stream)
  int timing;
  std::vector<cudaEvent_t> *tev;    // pool
  std::vector<int> *tstage;         // stage id of interval [2k, 2k+1]
  size_t tused;
  double stage_ms[SFE_FE_STAGES];
  long long stage_launches[SFE_FE_STAGES];
};

static void fe_tic(sfe_frontend *fe, int stage) {
  if (!fe->timing) return;
  while (fe->tev->size() < fe->tused + 2) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    fe->tev->push_back(e);
  }
  cudaEventRecord((*fe->tev)[fe->tused], fe->ctx->stream);
  fe->tstage->push_back(stage);
}
static void fe_toc(sfe_frontend *fe) {
  if (!fe->timing) return;
  cudaEventRecord((*fe->tev)[fe->tused + 1], fe->ctx->stream);
  fe->tused += 2;
}

extern "C" {

void sfe_frontend_params_default(sfe_frontend_params *p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->R = 512, p->B = 512;
  p->cfar_alg = SFE_CFAR_SOCA, p->train_hs = 20, p->guard_hs = 5, p->rank = 10;
  p->tau = 2.749063720096473;  // CFAR(40, 10, 0.1, 10).threshold_factor_SOCA
  p->gate_enable = 1, p->gate_threshold = 65.0;
  p->resolution = 0.5f;
  p->outlier_radius = 1.0, p->outlier_min_points = 5;
  p->window = 3;
  p->submap_resolution = 0.5f;
  p->min_points = 50;
  sfe_icp_params_default(&p->icp);
  p->cap_points = 8192;
  p->cap_source = 1024, p->cap_target = 3072;
  p->flip_lateral = 1;
}

#define FE_ALLOC(ptr, bytes)                                                        \
  do {                                                                              \
    cudaError_t e_ = cudaMalloc((void **)&(ptr), (bytes));                          \
    if (e_ != cudaSuccess) {                                                        \
      sfe::set_error("sfe_frontend_create: cudaMalloc(%zu) failed: %s", (size_t)(bytes), cudaGetErrorString(e_)); \
      sfe_frontend_destroy(fe);                                                     \
      return SFE_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

int sfe_frontend_create(sfe_ctx *ctx, const sfe_maps *maps, const sfe_frontend_params *params, int max_frames,
                        sfe_frontend **out) {
  SFE_REQUIRE(ctx && maps && params && out, "sfe_frontend_create: null argument");
  *out = nullptr;
  SFE_REQUIRE(max_frames > 0, "sfe_frontend_create: max_frames must be positive");
  SFE_REQUIRE(params->R == maps->R && params->B == maps->B, "sfe_frontend_create: polar shape %dx%d differs from the maps' %dx%d",
              params->R, params->B, maps->R, maps->B);
  SFE_REQUIRE(params->cap_points > 0 && params->window >= 1 && params->window <= 16, "sfe_frontend_create: bad capacity/window");
  SFE_REQUIRE(params->cap_source > 0 && params->cap_target > 0 && params->cap_source <= 65535 && params->cap_target <= 65535,
              "sfe_frontend_create: cap_source / cap_target must be in [1, 65535]");
  SFE_CUDA(cudaSetDevice(ctx->device));
  sfe_frontend *fe = new sfe_frontend();
  memset(fe, 0, sizeof(*fe));
  fe->ctx = ctx, fe->maps = maps, fe->p = *params, fe->max_frames = max_frames;
  fe->tev = new std::vector<cudaEvent_t>();
  fe->tstage = new std::vector<int>();
  const size_t F = max_frames, cap = params->cap_points, tcap = (size_t)params->window * cap;
  const size_t wpr = (params->B + 31) / 32;
  FE_ALLOC(fe->frames, F * params->R * params->B);
  FE_ALLOC(fe->bits, F * params->R * wpr * 4);
  FE_ALLOC(fe->ij, F * cap * 2 * 4);
  FE_ALLOC(fe->xy_a, F * cap * 2 * 4);
  FE_ALLOC(fe->xy_b, F * cap * 2 * 4);
  FE_ALLOC(fe->idx, F * cap * 4);
  FE_ALLOC(fe->cnt_a, F * 4);
  FE_ALLOC(fe->cnt_b, F * 4);
  FE_ALLOC(fe->cnt_c, F * 4);
  FE_ALLOC(fe->off_pts, (F + 1) * 4);
  FE_ALLOC(fe->off_tgt, (F + 1) * 4);
  FE_ALLOC(fe->tgt_a, F * tcap * 2 * 4);
  FE_ALLOC(fe->tgt_b, F * tcap * 2 * 4);
  FE_ALLOC(fe->tgt_idx, F * tcap * 4);
  FE_ALLOC(fe->tcnt_a, F * 4);
  FE_ALLOC(fe->tcnt_b, F * 4);
  FE_ALLOC(fe->rel, F * (params->window + 1) * 9 * 4);
  for (int k = 0; k < 2; ++k)
    if (cudaMallocHost((void **)&fe->rel_host[k], F * (params->window + 1) * 9 * 4) != cudaSuccess ||
        cudaEventCreateWithFlags(&fe->ev_rel[k], cudaEventDisableTiming) != cudaSuccess) {
      sfe::set_error("sfe_frontend_create: pinned staging allocation failed");
      sfe_frontend_destroy(fe);
      return SFE_ERR_CUDA;
    }
  FE_ALLOC(fe->carry_xy, (size_t)params->window * cap * 2 * 4);
  FE_ALLOC(fe->carry_cnt, (size_t)params->window * 4);
  fe->carry_poses = new double[3 * (size_t)params->window]();
  FE_ALLOC(fe->guess, F * 9 * 4);
  FE_ALLOC(fe->T, F * 9 * 4);
  FE_ALLOC(fe->iters, F * 4);
  FE_ALLOC(fe->inliers, F * 4);
  FE_ALLOC(fe->status, F * 4);
  cudaError_t e = cudaStreamCreateWithFlags(&fe->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fe->ev_copy[0], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fe->ev_copy[1], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fe->ev_done, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    sfe::set_error("sfe_frontend_create: %s", cudaGetErrorString(e));
    sfe_frontend_destroy(fe);
    return SFE_ERR_CUDA;
  }
  fill_offsets_kernel<<<(max_frames + 256) / 256, 256, 0, ctx->stream>>>(fe->off_pts, max_frames, (int)cap);
  fill_offsets_kernel<<<(max_frames + 256) / 256, 256, 0, ctx->stream>>>(fe->off_tgt, max_frames, (int)tcap);
  SFE_CUDA(cudaGetLastError());
  ctx->launches += 2;
  *out = fe;
  return SFE_OK;
}

void sfe_frontend_destroy(sfe_frontend *fe) {
  if (!fe) return;
  cudaSetDevice(fe->ctx->device);
  cudaStreamSynchronize(fe->ctx->stream);
  void *bufs[] = {fe->frames, fe->bits, fe->ij, fe->xy_a, fe->xy_b, fe->idx, fe->cnt_a, fe->cnt_b, fe->cnt_c,
                  fe->off_pts, fe->off_tgt, fe->tgt_a, fe->tgt_b, fe->tgt_idx, fe->tcnt_a, fe->tcnt_b, fe->rel,
                  fe->guess, fe->T, fe->iters, fe->inliers, fe->status, fe->carry_xy, fe->carry_cnt};
  for (void *b : bufs)
    if (b) cudaFree(b);
  delete[] fe->carry_poses;
  for (int k = 0; k < 2; ++k) {
    if (fe->rel_host[k]) cudaFreeHost(fe->rel_host[k]);
    if (fe->ev_rel[k]) cudaEventDestroy(fe->ev_rel[k]);
  }
  if (fe->copy_stream) cudaStreamDestroy(fe->copy_stream);
  for (auto &ev : fe->ev_copy)
    if (ev) cudaEventDestroy(ev);
  if (fe->ev_done) cudaEventDestroy(fe->ev_done);
  if (fe->tev) {
    for (auto e : *fe->tev) cudaEventDestroy(e);
    delete fe->tev;
  }
  delete fe->tstage;
  delete fe;
}

// window transforms + guesses from the odometry poses (host, double), uploaded as float32.  With carried frames the
// pose sequence is (carried poses ++ this batch's poses): frame i of the batch is virtual frame v = carry_n + i, its
// reference is virtual frame v - 1 and its window the virtual frames v - window .. v - 1.
static int fe_upload_poses(sfe_frontend *fe, const double *poses, int n) {
  const int W = fe->p.window, C = fe->carry_n;
  const int turn = fe->rel_turn ^= 1;
  SFE_CUDA(cudaEventSynchronize(fe->ev_rel[turn]));  // the copy that last used this staging buffer (two calls ago)
  float *stage = fe->rel_host[turn];
  auto pose_of = [&](int v) { return v < C ? fe->carry_poses + 3 * (size_t)v : poses + 3 * (size_t)(v - C); };
  for (int i = 0; i < n; ++i) {
    float *r = stage + (size_t)i * (W + 1) * 9;
    const int v = C + i;
    for (int w = 0; w < W; ++w) {
      const int u = v - W + w;
      if (u >= 0 && v > 0)
        pose_between(pose_of(v - 1), pose_of(u), r + w * 9);
      else
        for (int q = 0; q < 9; ++q) r[w * 9 + q] = (q % 4 == 0) ? 1.f : 0.f;
    }
    if (v > 0)
      pose_between(pose_of(v - 1), pose_of(v), r + W * 9);
    else
      for (int q = 0; q < 9; ++q) r[W * 9 + q] = (q % 4 == 0) ? 1.f : 0.f;
  }
  SFE_CUDA(cudaMemcpyAsync(fe->rel, stage, sizeof(float) * 9 * (size_t)n * (W + 1), cudaMemcpyHostToDevice,
                           fe->ctx->stream));
  SFE_CUDA(cudaEventRecord(fe->ev_rel[turn], fe->ctx->stream));
  return SFE_OK;
}

static const float *fe_cloud(const sfe_frontend *fe, const int32_t **cnt);

// stages 1-3 on frames [f0, f0+n): CFAR (bit plane) -> Cartesian points -> downsample -> outlier removal.
// Final per-frame clouds end in xy_a / cnt_c (or wherever the last enabled filter wrote; see cloud()).
static int fe_features(sfe_frontend *fe, const uint8_t *frames_dev, int f0, int n) {
  sfe_ctx *ctx = fe->ctx;
  const sfe_frontend_params &p = fe->p;
  const size_t cap = p.cap_points, wpr = (p.B + 31) / 32;
  uint32_t *bits = fe->bits + (size_t)f0 * p.R * wpr;
  fe_tic(fe, SFE_FE_CFAR);
  int rc = cfar_run(ctx, frames_dev, SFE_U8, n, p.R, p.B, p.cfar_alg, p.train_hs, p.guard_hs, p.rank, p.tau,
                    p.gate_enable, p.gate_threshold, nullptr, nullptr, bits, 0);
  fe_toc(fe);
  if (rc != SFE_OK) return rc;
  fe_tic(fe, SFE_FE_CART);
  rc = cart_points_run(ctx, fe->maps, nullptr, bits, n, (int)cap, fe->ij + (size_t)f0 * cap * 2,
                       fe->xy_a + (size_t)f0 * cap * 2, fe->cnt_a + f0);
  fe_toc(fe);
  if (rc != SFE_OK) return rc;
  // feature_extraction.py:241-249
  const float *cur = fe->xy_a;
  const int32_t *cur_cnt = fe->cnt_a;
  if (p.resolution > 0.f) {
    fe_tic(fe, SFE_FE_DOWNSAMPLE);
    static const int ds_split = [] {  // development switch: SFE_FE_DS_SPLIT=<points> (0: one launch)
      const char *e = getenv("SFE_FE_DS_SPLIT");
      return e ? atoi(e) : FE_DS_SPLIT;
    }();
    rc = downsample_run(ctx, cur, fe->off_pts + f0, cur_cnt + f0, n, 2, (int)cap, p.resolution, fe->xy_b, fe->idx,
                        fe->cnt_b + f0, ds_split);
    fe_toc(fe);
    if (rc != SFE_OK) return rc;
    cur = fe->xy_b, cur_cnt = fe->cnt_b;
  }
  if (p.outlier_min_points > 1) {
    float *dst = (cur == fe->xy_a) ? fe->xy_b : fe->xy_a;
    fe_tic(fe, SFE_FE_OUTLIER);
    rc = remove_outlier_run(ctx, cur, fe->off_pts + f0, cur_cnt + f0, n, 2, (int)cap, p.outlier_radius,
                            p.outlier_min_points, dst, fe->idx, fe->cnt_c + f0, FE_RO_SPLIT);
    fe_toc(fe);
    if (rc != SFE_OK) return rc;
  }
  if (p.flip_lateral) {
    const int32_t *cnt;
    float *cloud = const_cast<float *>(fe_cloud(fe, &cnt));
    fe_tic(fe, SFE_FE_OUTLIER);
    flip_lateral_kernel<<<n, 128, 0, ctx->stream>>>(cloud + 2 * (size_t)f0 * cap, cnt + f0, (int)cap, n);
    fe_toc(fe);
    SFE_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  return SFE_OK;
}

static const float *fe_cloud(const sfe_frontend *fe, const int32_t **cnt) {
  const sfe_frontend_params &p = fe->p;
  const bool ds = p.resolution > 0.f, ro = p.outlier_min_points > 1;
  if (ro) {
    *cnt = fe->cnt_c;
    return ds ? fe->xy_a : fe->xy_b;
  }
  if (ds) {
    *cnt = fe->cnt_b;
    return fe->xy_b;
  }
  *cnt = fe->cnt_a;
  return fe->xy_a;
}

// stage 4-5 on frames [f0, f0+n): targets from the window of previous frames (which may precede f0), then ICP
static int fe_match(sfe_frontend *fe, int f0, int n) {
  sfe_ctx *ctx = fe->ctx;
  const sfe_frontend_params &p = fe->p;
  const int cap = p.cap_points, tcap = p.window * cap;
  const int32_t *cnt;
  const float *cloud = fe_cloud(fe, &cnt);
  fe_tic(fe, SFE_FE_SUBMAP);
  assemble_targets_kernel<<<n, 256, 0, ctx->stream>>>(cloud, cnt, cap, fe->rel, f0, f0 + n, p.window, fe->tgt_a,
                                                      fe->tcnt_a, tcap, fe->guess, fe->carry_xy, fe->carry_cnt,
                                                      fe->carry_n);
  SFE_CUDA(cudaGetLastError());
  ctx->launches++;
  const float *tgt = fe->tgt_a;
  const int32_t *tcnt = fe->tcnt_a;
  if (p.submap_resolution > 0.f) {
    int rc = downsample_run(ctx, fe->tgt_a, fe->off_tgt + f0, fe->tcnt_a + f0, n, 2, tcap, p.submap_resolution,
                            fe->tgt_b, fe->tgt_idx, fe->tcnt_b + f0, FE_SUBMAP_SPLIT);
    if (rc != SFE_OK) return rc;
    tgt = fe->tgt_b, tcnt = fe->tcnt_b;
  }
  fe_toc(fe);
  fe_tic(fe, SFE_FE_ICP);
  // Two launches by problem size: a sonar frame's filtered cloud is a few hundred points, far below the
  // capacities the buffers are sized for, and the ICP kernel's occupancy is set by its shared-memory footprint.
  // The first launch is sized for the common case (6 CTAs per SM instead of 4) and leaves bigger problems
  // untouched; the second, sized for the capacities, only takes those.
  const int ns_small = p.cap_source < FE_ICP_SMALL_SRC ? p.cap_source : FE_ICP_SMALL_SRC;
  const int nt_small = p.cap_target < FE_ICP_SMALL_TGT ? p.cap_target : FE_ICP_SMALL_TGT;
  // class_mode 1 leaves oversize problems untouched for the second launch; when the capacities are not larger than
  // the small class there is no second launch and this one must report them itself (SFE_ICP_TOO_LARGE)
  const bool two = ns_small < p.cap_source || nt_small < p.cap_target;
  int rc = icp_run(ctx, &p.icp, cloud, fe->off_pts + f0, cnt + f0, tgt, fe->off_tgt + f0, tcnt + f0, p.min_points,
                   nullptr, nullptr, n, ns_small, nt_small, fe->guess + 9 * (size_t)f0, fe->T + 9 * (size_t)f0,
                   fe->iters + f0, fe->inliers + f0, fe->status + f0, fe->cnt_a + f0, cap, two ? 1 : 0, 0, 0,
                   FE_ICP_SMALL_THREADS);
  if (rc == SFE_OK && two)
    rc = icp_run(ctx, &p.icp, cloud, fe->off_pts + f0, cnt + f0, tgt, fe->off_tgt + f0, tcnt + f0, p.min_points,
                 nullptr, nullptr, n, p.cap_source, p.cap_target, fe->guess + 9 * (size_t)f0,
                 fe->T + 9 * (size_t)f0, fe->iters + f0, fe->inliers + f0, fe->status + f0, fe->cnt_a + f0, cap, 2,
                 ns_small, nt_small);
  fe_toc(fe);
  return rc;
}

// After a batch: the carried window becomes the last `window` frames of (carried ++ batch), stream-ordered after
// the batch's matching (which still reads the old carried clouds).  Slots move towards the front, in ascending
// order, so no frame is overwritten before it has been moved.
static int fe_update_carry(sfe_frontend *fe, const double *poses, int n) {
  if (!fe->carry_on) return SFE_OK;
  const int W = fe->p.window, C = fe->carry_n, total = C + n, keep = total < W ? total : W;
  const size_t cap = fe->p.cap_points, cbytes = cap * 2 * sizeof(float);
  const int32_t *cnt;
  const float *cloud = fe_cloud(fe, &cnt);
  for (int j = 0; j < keep; ++j) {
    const int v = total - keep + j;
    if (v < C) {
      if (v != j) {
        SFE_CUDA(cudaMemcpyAsync(fe->carry_xy + (size_t)j * cap * 2, fe->carry_xy + (size_t)v * cap * 2, cbytes,
                                 cudaMemcpyDeviceToDevice, fe->ctx->stream));
        SFE_CUDA(cudaMemcpyAsync(fe->carry_cnt + j, fe->carry_cnt + v, 4, cudaMemcpyDeviceToDevice, fe->ctx->stream));
        memcpy(fe->carry_poses + 3 * (size_t)j, fe->carry_poses + 3 * (size_t)v, 3 * sizeof(double));
      }
    } else {
      const size_t b = (size_t)(v - C);
      SFE_CUDA(cudaMemcpyAsync(fe->carry_xy + (size_t)j * cap * 2, cloud + b * cap * 2, cbytes, cudaMemcpyDeviceToDevice,
                               fe->ctx->stream));
      SFE_CUDA(cudaMemcpyAsync(fe->carry_cnt + j, cnt + b, 4, cudaMemcpyDeviceToDevice, fe->ctx->stream));
      memcpy(fe->carry_poses + 3 * (size_t)j, poses + 3 * b, 3 * sizeof(double));
    }
  }
  fe->carry_n = keep;
  return SFE_OK;
}

int sfe_frontend_set_carry(sfe_frontend *fe, int enable) {
  SFE_REQUIRE(fe != nullptr, "sfe_frontend_set_carry: null handle");
  fe->carry_on = enable != 0;
  if (!fe->carry_on) fe->carry_n = 0;
  return SFE_OK;
}

int sfe_frontend_set_timing(sfe_frontend *fe, int enable) {
  SFE_REQUIRE(fe != nullptr, "sfe_frontend_set_timing: null handle");
  fe->timing = enable != 0;
  fe->tused = 0;
  fe->tstage->clear();
  for (int i = 0; i < SFE_FE_STAGES; ++i) fe->stage_ms[i] = 0.0, fe->stage_launches[i] = 0;
  return SFE_OK;
}

int sfe_frontend_get_timing(sfe_frontend *fe, double *stage_ms, int64_t *stage_calls) {
  SFE_REQUIRE(fe && stage_ms, "sfe_frontend_get_timing: null argument");
  SFE_CUDA(cudaSetDevice(fe->ctx->device));
  SFE_CUDA(cudaStreamSynchronize(fe->ctx->stream));
  for (size_t k = 0; k + 1 < fe->tused + 1 && 2 * k + 1 < fe->tused + 1 && k < fe->tstage->size(); ++k) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, (*fe->tev)[2 * k], (*fe->tev)[2 * k + 1]) == cudaSuccess) {
      fe->stage_ms[(*fe->tstage)[k]] += ms;
      fe->stage_launches[(*fe->tstage)[k]]++;
    }
  }
  fe->tused = 0;
  fe->tstage->clear();
  for (int i = 0; i < SFE_FE_STAGES; ++i) {
    stage_ms[i] = fe->stage_ms[i];
    if (stage_calls) stage_calls[i] = fe->stage_launches[i];
  }
  return SFE_OK;
}

int sfe_frontend_run_dev(sfe_frontend *fe, const uint8_t *frames_dev, const double *poses_host, int n_frames) {
  SFE_REQUIRE(fe && frames_dev && poses_host, "sfe_frontend_run_dev: null argument");
  SFE_REQUIRE(n_frames >= 0 && n_frames <= fe->max_frames, "sfe_frontend_run_dev: %d frames exceed max_frames %d",
              n_frames, fe->max_frames);
  if (n_frames == 0) return SFE_OK;
  SFE_CUDA(cudaSetDevice(fe->ctx->device));
  int rc = fe_upload_poses(fe, poses_host, n_frames);
  if (rc != SFE_OK) return rc;
  rc = fe_features(fe, frames_dev, 0, n_frames);
  if (rc != SFE_OK) return rc;
  rc = fe_match(fe, 0, n_frames);
  if (rc != SFE_OK) return rc;
  return fe_update_carry(fe, poses_host, n_frames);
}

int sfe_frontend_results_dev(const sfe_frontend *fe, const float **T, const int32_t **iters, const int32_t **inliers,
                             const int32_t **status, const float **cloud_xy, const int32_t **cloud_count,
                             int32_t *cloud_stride) {
  SFE_REQUIRE(fe != nullptr, "sfe_frontend_results_dev: null handle");
  if (T) *T = fe->T;
  if (iters) *iters = fe->iters;
  if (inliers) *inliers = fe->inliers;
  if (status) *status = fe->status;
  const int32_t *cnt;
  const float *cloud = fe_cloud(fe, &cnt);
  if (cloud_xy) *cloud_xy = cloud;
  if (cloud_count) *cloud_count = cnt;
  if (cloud_stride) *cloud_stride = fe->p.cap_points;
  return SFE_OK;
}

int sfe_frontend_run_host(sfe_frontend *fe, const uint8_t *frames_host, const double *poses_host, int n_frames,
                          int chunk_frames, float *T_host, int32_t *iters_host, int32_t *inliers_host,
                          int32_t *status_host, int32_t *npoints_host) {
  SFE_REQUIRE(fe && frames_host && poses_host && T_host && iters_host && inliers_host && status_host,
              "sfe_frontend_run_host: null argument");
  SFE_REQUIRE(n_frames >= 0 && n_frames <= fe->max_frames, "sfe_frontend_run_host: %d frames exceed max_frames %d",
              n_frames, fe->max_frames);
  if (n_frames == 0) return SFE_OK;
  sfe_ctx *ctx = fe->ctx;
  SFE_CUDA(cudaSetDevice(ctx->device));
  const size_t fbytes = (size_t)fe->p.R * fe->p.B;
  if (chunk_frames <= 0) chunk_frames = 256;
  int rc0 = fe_upload_poses(fe, poses_host, n_frames);
  if (rc0 != SFE_OK) return rc0;
  // frame chunks stream in on the copy stream; the feature stages of chunk k run while chunk k+1 copies
  int k = 0;
  for (int f0 = 0; f0 < n_frames; f0 += chunk_frames, ++k) {
    const int n = (n_frames - f0) < chunk_frames ? (n_frames - f0) : chunk_frames;
    SFE_CUDA(cudaMemcpyAsync(fe->frames + (size_t)f0 * fbytes, frames_host + (size_t)f0 * fbytes, (size_t)n * fbytes,
                             cudaMemcpyHostToDevice, fe->copy_stream));
    SFE_CUDA(cudaEventRecord(fe->ev_copy[k & 1], fe->copy_stream));
    SFE_CUDA(cudaStreamWaitEvent(ctx->stream, fe->ev_copy[k & 1], 0));
    int rc = fe_features(fe, fe->frames + (size_t)f0 * fbytes, f0, n);
    if (rc != SFE_OK) return rc;
    rc = fe_match(fe, f0, n);  // its window only reaches back into chunks already processed
    if (rc != SFE_OK) return rc;
  }
  int rc = fe_update_carry(fe, poses_host, n_frames);
  if (rc != SFE_OK) return rc;
  SFE_CUDA(cudaMemcpyAsync(T_host, fe->T, sizeof(float) * 9 * (size_t)n_frames, cudaMemcpyDeviceToHost, ctx->stream));
  SFE_CUDA(cudaMemcpyAsync(iters_host, fe->iters, 4 * (size_t)n_frames, cudaMemcpyDeviceToHost, ctx->stream));
  SFE_CUDA(cudaMemcpyAsync(inliers_host, fe->inliers, 4 * (size_t)n_frames, cudaMemcpyDeviceToHost, ctx->stream));
  SFE_CUDA(cudaMemcpyAsync(status_host, fe->status, 4 * (size_t)n_frames, cudaMemcpyDeviceToHost, ctx->stream));
  if (npoints_host) {
    const int32_t *cnt;
    fe_cloud(fe, &cnt);
    SFE_CUDA(cudaMemcpyAsync(npoints_host, cnt, 4 * (size_t)n_frames, cudaMemcpyDeviceToHost, ctx->stream));
  }
  SFE_CUDA(cudaStreamSynchronize(ctx->stream));
  return SFE_OK;
}

}  // extern "C"
