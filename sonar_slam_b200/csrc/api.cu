// libsonarfe C ABI: context, error reporting, host-buffer conveniences.
// Declarations and reference citations: include/sonarfe.h.
#include <cudaTypedefs.h>

#include <mutex>

#include "common.cuh"

namespace sfe {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ensure(sfe_ctx *ctx, Buffer &b, size_t bytes) {
  if (bytes <= b.cap) return SFE_OK;
  // grow-only; the old block may still be in use by enqueued work -> stream-ordered free
  if (b.ptr) SFE_CUDA(cudaFreeAsync(b.ptr, ctx->stream));
  if (b.ptr == ctx->cfar_lut_buf) ctx->cfar_lut_buf = nullptr;  // the cached table went with the block
  b.ptr = nullptr, b.cap = 0;
  size_t want = bytes + bytes / 4 + 256;
  SFE_CUDA(cudaMallocAsync(&b.ptr, want, ctx->stream));
  b.cap = want;
  return SFE_OK;
}

int ensure_pinned(sfe_ctx *ctx, size_t bytes) {
  if (bytes <= ctx->pinned_cap) return SFE_OK;
  if (ctx->pinned) {
    SFE_CUDA(cudaStreamSynchronize(ctx->stream));
    SFE_CUDA(cudaFreeHost(ctx->pinned));
  }
  ctx->pinned = nullptr, ctx->pinned_cap = 0;
  SFE_CUDA(cudaMallocHost(&ctx->pinned, bytes * 2));
  ctx->pinned_cap = bytes * 2;
  return SFE_OK;
}

int encode_tensor_map_3d(CUtensorMap *map, CUtensorMapDataType dt, size_t es, const void *base, uint64_t d0,
                         uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1, uint32_t b2) {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  });
  if (!encode) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return SFE_ERR_CUDA;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * es, d0 * d1 * es};  // bytes, dims 1..2
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(map, dt, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) for dims (%llu,%llu,%llu) box (%u,%u,%u)", (int)r,
              (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, b0, b1, b2);
    return SFE_ERR_CUDA;
  }
  return SFE_OK;
}

int cfar_run(sfe_ctx *ctx, const void *img, int dtype, int F, int R, int B, int alg, int T, int G, int k,
             double tau, int gate_on, double gate, uint8_t *mask, float *thr, uint32_t *bits, int force_exact);

}  // namespace sfe

using namespace sfe;

extern "C" {

int sfe_version(void) { return SFE_VERSION; }
const char *sfe_last_error(void) { return g_err; }

int sfe_ctx_create(int device, void *cuda_stream, int own_stream, sfe_ctx **out) {
  SFE_REQUIRE(out != nullptr, "sfe_ctx_create: null out pointer");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("sfe_ctx_create: no CUDA device available (%s); libsonarfe has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return SFE_ERR_CUDA;
  }
  SFE_REQUIRE(device >= 0 && device < n, "sfe_ctx_create: device %d out of range [0, %d)", device, n);
  SFE_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  SFE_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("sfe_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
              prop.major, prop.minor);
    return SFE_ERR_UNSUPPORTED;
  }
  sfe_ctx *ctx = new sfe_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  if (!own_stream) {
    ctx->stream = (cudaStream_t)cuda_stream;
  } else {
    cudaError_t e2 = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e2 != cudaSuccess) {
      set_error("cudaStreamCreate: %s", cudaGetErrorString(e2));
      delete ctx;
      return SFE_ERR_CUDA;
    }
    ctx->owns_stream = true;
  }
  *out = ctx;
  return SFE_OK;
}

void sfe_ctx_destroy(sfe_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto &b : ctx->scratch)
    if (b.ptr) cudaFree(b.ptr);
  for (auto &b : ctx->stage_in)
    if (b.ptr) cudaFree(b.ptr);
  for (auto &b : ctx->stage_out)
    if (b.ptr) cudaFree(b.ptr);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int sfe_sync(sfe_ctx *ctx) {
  SFE_REQUIRE(ctx != nullptr, "sfe_sync: null context");
  SFE_CUDA(cudaStreamSynchronize(ctx->stream));
  return SFE_OK;
}

uint64_t sfe_launch_count(const sfe_ctx *ctx) { return ctx ? ctx->launches : 0; }

int sfe_copy_to_host(sfe_ctx *ctx, void *dst_host, const void *src_dev, uint64_t bytes) {
  SFE_REQUIRE(ctx && (bytes == 0 || (dst_host && src_dev)), "sfe_copy_to_host: null argument");
  if (bytes == 0) return SFE_OK;
  SFE_CUDA(cudaSetDevice(ctx->device));
  SFE_CUDA(cudaMemcpyAsync(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
  SFE_CUDA(cudaStreamSynchronize(ctx->stream));
  return SFE_OK;
}

int sfe_cfar_dev(sfe_ctx *ctx, const void *img_dev, int dtype, int n_frames, int R, int B, int alg, int train_hs,
                 int guard_hs, int k, double tau, int gate_enable, double gate_threshold, uint8_t *mask_dev,
                 float *thr_dev, uint32_t *bits_dev) {
  SFE_REQUIRE(ctx != nullptr, "sfe_cfar_dev: null context");
  SFE_CUDA(cudaSetDevice(ctx->device));
  return cfar_run(ctx, img_dev, dtype, n_frames, R, B, alg, train_hs, guard_hs, k, tau, gate_enable, gate_threshold,
                  mask_dev, thr_dev, bits_dev, 0);
}

int sfe_cfar_host(sfe_ctx *ctx, const void *img_host, int dtype, int n_frames, int R, int B, int alg, int train_hs,
                  int guard_hs, int k, double tau, int gate_enable, double gate_threshold, uint8_t *mask_host,
                  float *thr_host) {
  SFE_REQUIRE(ctx != nullptr, "sfe_cfar_host: null context");
  SFE_REQUIRE(dtype == SFE_U8 || dtype == SFE_F32, "sfe_cfar_host: dtype must be SFE_U8 or SFE_F32");
  SFE_REQUIRE(n_frames >= 0 && R >= 0 && B >= 0, "sfe_cfar_host: negative shape");
  SFE_CUDA(cudaSetDevice(ctx->device));
  const size_t cells = (size_t)n_frames * R * B;
  if (cells == 0) return SFE_OK;
  SFE_REQUIRE(img_host != nullptr && mask_host != nullptr, "sfe_cfar_host: null image or mask pointer");
  const size_t es = dtype == SFE_U8 ? 1 : 4;
  int rc;
  if ((rc = ensure(ctx, ctx->stage_in[0], cells * es)) != SFE_OK) return rc;
  if ((rc = ensure(ctx, ctx->stage_out[0], cells)) != SFE_OK) return rc;
  if (thr_host && (rc = ensure(ctx, ctx->stage_out[1], cells * sizeof(float))) != SFE_OK) return rc;
  SFE_CUDA(cudaMemcpyAsync(ctx->stage_in[0].ptr, img_host, cells * es, cudaMemcpyHostToDevice, ctx->stream));
  rc = cfar_run(ctx, ctx->stage_in[0].ptr, dtype, n_frames, R, B, alg, train_hs, guard_hs, k, tau, gate_enable,
                gate_threshold, (uint8_t *)ctx->stage_out[0].ptr, thr_host ? (float *)ctx->stage_out[1].ptr : nullptr,
                nullptr, 0);
  if (rc != SFE_OK) return rc;
  SFE_CUDA(cudaMemcpyAsync(mask_host, ctx->stage_out[0].ptr, cells, cudaMemcpyDeviceToHost, ctx->stream));
  if (thr_host)
    SFE_CUDA(cudaMemcpyAsync(thr_host, ctx->stage_out[1].ptr, cells * sizeof(float), cudaMemcpyDeviceToHost,
                             ctx->stream));
  SFE_CUDA(cudaStreamSynchronize(ctx->stream));
  return SFE_OK;
}

}  // extern "C"

// ============================================================================ clouds / match / ICP
namespace sfe {
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
int match_run(sfe_ctx *ctx, const float *ref_pts, const int *ref_off, const float *in_pts, const int *in_off, int P,
              int nt_max, float max_dist, int32_t *ids, float *dists);

// carve `n` sub-buffers out of one grow-only scratch block (256 B aligned each)
struct Carver {
  char *base = nullptr;
  size_t off = 0;
  void *take(size_t bytes) {
    void *p = base ? base + off : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  }
};
}  // namespace sfe

extern "C" {

int sfe_downsample_dev(sfe_ctx *ctx, const float *pts_dev, const int32_t *off_dev, int n_clouds, int dim, int n_max,
                       float resolution, float *out_pts_dev, int32_t *out_idx_dev, int32_t *out_count_dev) {
  SFE_REQUIRE(ctx != nullptr, "sfe_downsample_dev: null context");
  SFE_CUDA(cudaSetDevice(ctx->device));
  return downsample_run(ctx, pts_dev, off_dev, nullptr, n_clouds, dim, n_max, resolution, out_pts_dev, out_idx_dev,
                        out_count_dev);
}

int sfe_remove_outlier_dev(sfe_ctx *ctx, const float *pts_dev, const int32_t *off_dev, int n_clouds, int dim,
                           int n_max, double radius, int min_points, float *out_pts_dev, int32_t *out_idx_dev,
                           int32_t *out_count_dev) {
  SFE_REQUIRE(ctx != nullptr, "sfe_remove_outlier_dev: null context");
  SFE_CUDA(cudaSetDevice(ctx->device));
  return remove_outlier_run(ctx, pts_dev, off_dev, nullptr, n_clouds, dim, n_max, radius, min_points, out_pts_dev,
                            out_idx_dev, out_count_dev);
}

static int cloud_filter_host(sfe_ctx *ctx, int which, const float *pts_host, int n, int dim, float resolution,
                             double radius, int min_points, float *out_pts_host, int32_t *out_idx_host,
                             int32_t *n_out) {
  SFE_REQUIRE(ctx && n_out, "cloud filter: null context or n_out");
  SFE_REQUIRE(n >= 0, "cloud filter: negative point count");
  SFE_REQUIRE(dim == 2 || dim == 3, "cloud filter: points must have 2 or 3 columns (got %d)", dim);
  *n_out = 0;
  if (n == 0) return SFE_OK;
  SFE_REQUIRE(pts_host && out_pts_host && out_idx_host, "cloud filter: null pointer");
  SFE_CUDA(cudaSetDevice(ctx->device));
  Carver cv;
  for (int pass = 0; pass < 2; ++pass) {
    cv.off = 0;
    float *d_pts = (float *)cv.take(sizeof(float) * (size_t)n * dim);
    int32_t *d_off = (int32_t *)cv.take(2 * sizeof(int32_t));
    float *d_out = (float *)cv.take(sizeof(float) * (size_t)n * dim);
    int32_t *d_idx = (int32_t *)cv.take(sizeof(int32_t) * (size_t)n);
    int32_t *d_cnt = (int32_t *)cv.take(sizeof(int32_t));
    if (pass == 0) {
      int rc = ensure(ctx, ctx->stage_in[1], cv.off);
      if (rc != SFE_OK) return rc;
      cv.base = (char *)ctx->stage_in[1].ptr;
      continue;
    }
    const int32_t off[2] = {0, n};
    SFE_CUDA(cudaMemcpyAsync(d_pts, pts_host, sizeof(float) * (size_t)n * dim, cudaMemcpyHostToDevice, ctx->stream));
    SFE_CUDA(cudaMemcpyAsync(d_off, off, sizeof(off), cudaMemcpyHostToDevice, ctx->stream));
    int rc = which == 0 ? downsample_run(ctx, d_pts, d_off, nullptr, 1, dim, n, resolution, d_out, d_idx, d_cnt)
                        : remove_outlier_run(ctx, d_pts, d_off, nullptr, 1, dim, n, radius, min_points, d_out, d_idx,
                                             d_cnt);
    if (rc != SFE_OK) return rc;
    SFE_CUDA(cudaMemcpyAsync(n_out, d_cnt, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SFE_CUDA(cudaMemcpyAsync(out_pts_host, d_out, sizeof(float) * (size_t)n * dim, cudaMemcpyDeviceToHost,
                             ctx->stream));
    SFE_CUDA(cudaMemcpyAsync(out_idx_host, d_idx, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    SFE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return SFE_OK;
}

int sfe_downsample_host(sfe_ctx *ctx, const float *pts_host, int n, int dim, float resolution, float *out_pts_host,
                        int32_t *out_idx_host, int32_t *n_out) {
  return cloud_filter_host(ctx, 0, pts_host, n, dim, resolution, 0.0, 0, out_pts_host, out_idx_host, n_out);
}

int sfe_remove_outlier_host(sfe_ctx *ctx, const float *pts_host, int n, int dim, double radius, int min_points,
                            float *out_pts_host, int32_t *out_idx_host, int32_t *n_out) {
  return cloud_filter_host(ctx, 1, pts_host, n, dim, 0.f, radius, min_points, out_pts_host, out_idx_host, n_out);
}

int sfe_match_dev(sfe_ctx *ctx, const float *ref_pts_dev, const int32_t *ref_off_dev, const float *in_pts_dev,
                  const int32_t *in_off_dev, int n_pairs, int n_ref_max, float max_dist, int32_t *ids_dev,
                  float *dists_dev) {
  SFE_REQUIRE(ctx != nullptr, "sfe_match_dev: null context");
  SFE_CUDA(cudaSetDevice(ctx->device));
  return match_run(ctx, ref_pts_dev, ref_off_dev, in_pts_dev, in_off_dev, n_pairs, n_ref_max, max_dist, ids_dev,
                   dists_dev);
}

int sfe_match_host(sfe_ctx *ctx, const float *ref_host, int n_ref, const float *in_host, int n_in, float max_dist,
                   int32_t *ids_host, float *dists_host) {
  SFE_REQUIRE(ctx != nullptr, "sfe_match_host: null context");
  SFE_REQUIRE(n_ref >= 0 && n_in >= 0, "sfe_match_host: negative point count");
  if (n_in == 0) return SFE_OK;
  SFE_REQUIRE(in_host && ids_host && dists_host && (ref_host || n_ref == 0), "sfe_match_host: null pointer");
  SFE_CUDA(cudaSetDevice(ctx->device));
  Carver cv;
  for (int pass = 0; pass < 2; ++pass) {
    cv.off = 0;
    float *d_ref = (float *)cv.take(sizeof(float) * 2 * (size_t)(n_ref > 0 ? n_ref : 1));
    float *d_in = (float *)cv.take(sizeof(float) * 2 * (size_t)n_in);
    int32_t *d_off = (int32_t *)cv.take(4 * sizeof(int32_t));
    int32_t *d_ids = (int32_t *)cv.take(sizeof(int32_t) * (size_t)n_in);
    float *d_d = (float *)cv.take(sizeof(float) * (size_t)n_in);
    if (pass == 0) {
      int rc = ensure(ctx, ctx->stage_in[1], cv.off);
      if (rc != SFE_OK) return rc;
      cv.base = (char *)ctx->stage_in[1].ptr;
      continue;
    }
    const int32_t off[4] = {0, n_ref, 0, n_in};
    if (n_ref > 0)
      SFE_CUDA(cudaMemcpyAsync(d_ref, ref_host, sizeof(float) * 2 * (size_t)n_ref, cudaMemcpyHostToDevice, ctx->stream));
    SFE_CUDA(cudaMemcpyAsync(d_in, in_host, sizeof(float) * 2 * (size_t)n_in, cudaMemcpyHostToDevice, ctx->stream));
    SFE_CUDA(cudaMemcpyAsync(d_off, off, sizeof(off), cudaMemcpyHostToDevice, ctx->stream));
    int rc = match_run(ctx, d_ref, d_off, d_in, d_off + 2, 1, n_ref, max_dist, d_ids, d_d);
    if (rc != SFE_OK) return rc;
    SFE_CUDA(cudaMemcpyAsync(ids_host, d_ids, sizeof(int32_t) * (size_t)n_in, cudaMemcpyDeviceToHost, ctx->stream));
    SFE_CUDA(cudaMemcpyAsync(dists_host, d_d, sizeof(float) * (size_t)n_in, cudaMemcpyDeviceToHost, ctx->stream));
    SFE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return SFE_OK;
}

void sfe_icp_params_default(sfe_icp_params *p) {
  if (!p) return;
  p->matcher_max_dist = 10.0f;
  p->outlier_max_dist = 3.0f;
  p->trim_ratio = 0.8f;
  p->max_iterations = 40;
  p->min_diff_rot = 0.01f;
  p->min_diff_trans = 0.1f;
  p->smooth_length = 4;
  p->flags = 0;
  p->minimizer = 0;
  p->normals_knn = 5;
}

const char *sfe_icp_status_message(int status) {
  switch (status) {
    case SFE_ICP_SUCCESS: return "success";
    case SFE_ICP_NO_OUTLIER: return "no outlier to filter";
    case SFE_ICP_NO_POINT: return "ErrorMnimizer: no point to minimize";
    case SFE_ICP_NAN_ROT: return "abs rotation norm not a number";
    case SFE_ICP_NAN_TRANS: return "abs translation norm not a number";
    case SFE_ICP_NOT_RIGID: return "RigidTransformation: Error, rotation matrix is not orthogonal.";
    case SFE_ICP_EMPTY_REF: return "reference cloud is empty";
    case SFE_ICP_SKIPPED: return "not enough points";
    case SFE_ICP_TOO_LARGE: return "cloud exceeds the configured capacity";
    default: return "unknown ICP status";
  }
}

int sfe_icp_dev(sfe_ctx *ctx, const sfe_icp_params *prm, const float *src_pts_dev, const int32_t *src_off_dev,
                const float *tgt_pts_dev, const int32_t *tgt_off_dev, const int32_t *src_id_dev,
                const int32_t *tgt_id_dev, int n_problems, int ns_max, int nt_max, const float *guess_dev, float *T_dev,
                int32_t *iters_dev, int32_t *inliers_dev, int32_t *status_dev) {
  SFE_REQUIRE(ctx != nullptr, "sfe_icp_dev: null context");
  SFE_CUDA(cudaSetDevice(ctx->device));
  return icp_run(ctx, prm, src_pts_dev, src_off_dev, nullptr, tgt_pts_dev, tgt_off_dev, nullptr, 0, src_id_dev,
                 tgt_id_dev, n_problems, ns_max, nt_max, guess_dev, T_dev, iters_dev, inliers_dev, status_dev);
}

int sfe_icp_host(sfe_ctx *ctx, const sfe_icp_params *prm, const float *src_host, int ns, const float *tgt_host, int nt,
                 const float *guess_host, int n_guesses, float *T_host, int32_t *iters_host, int32_t *inliers_host,
                 int32_t *status_host) {
  SFE_REQUIRE(ctx && prm, "sfe_icp_host: null context or parameters");
  SFE_REQUIRE(ns >= 0 && nt >= 0 && n_guesses >= 0, "sfe_icp_host: negative size");
  if (n_guesses == 0) return SFE_OK;
  SFE_REQUIRE(guess_host && T_host && iters_host && inliers_host && status_host, "sfe_icp_host: null pointer");
  SFE_REQUIRE((src_host || ns == 0) && (tgt_host || nt == 0), "sfe_icp_host: null cloud pointer");
  SFE_CUDA(cudaSetDevice(ctx->device));
  Carver cv;
  for (int pass = 0; pass < 2; ++pass) {
    cv.off = 0;
    float *d_src = (float *)cv.take(sizeof(float) * 2 * (size_t)(ns > 0 ? ns : 1));
    float *d_tgt = (float *)cv.take(sizeof(float) * 2 * (size_t)(nt > 0 ? nt : 1));
    int32_t *d_off = (int32_t *)cv.take(4 * sizeof(int32_t));
    int32_t *d_zero = (int32_t *)cv.take(sizeof(int32_t) * (size_t)n_guesses);
    float *d_guess = (float *)cv.take(sizeof(float) * 9 * (size_t)n_guesses);
    float *d_T = (float *)cv.take(sizeof(float) * 9 * (size_t)n_guesses);
    int32_t *d_res = (int32_t *)cv.take(sizeof(int32_t) * 3 * (size_t)n_guesses);
    if (pass == 0) {
      int rc = ensure(ctx, ctx->stage_in[1], cv.off);
      if (rc != SFE_OK) return rc;
      cv.base = (char *)ctx->stage_in[1].ptr;
      continue;
    }
    const int32_t off[4] = {0, ns, 0, nt};
    if (ns > 0)
      SFE_CUDA(cudaMemcpyAsync(d_src, src_host, sizeof(float) * 2 * (size_t)ns, cudaMemcpyHostToDevice, ctx->stream));
    if (nt > 0)
      SFE_CUDA(cudaMemcpyAsync(d_tgt, tgt_host, sizeof(float) * 2 * (size_t)nt, cudaMemcpyHostToDevice, ctx->stream));
    SFE_CUDA(cudaMemcpyAsync(d_off, off, sizeof(off), cudaMemcpyHostToDevice, ctx->stream));
    SFE_CUDA(cudaMemsetAsync(d_zero, 0, sizeof(int32_t) * (size_t)n_guesses, ctx->stream));
    SFE_CUDA(cudaMemcpyAsync(d_guess, guess_host, sizeof(float) * 9 * (size_t)n_guesses, cudaMemcpyHostToDevice,
                             ctx->stream));
    int rc = icp_run(ctx, prm, d_src, d_off, nullptr, d_tgt, d_off + 2, nullptr, 0, d_zero, d_zero, n_guesses, ns, nt,
                     d_guess, d_T, d_res, d_res + n_guesses, d_res + 2 * n_guesses);
    if (rc != SFE_OK) return rc;
    SFE_CUDA(cudaMemcpyAsync(T_host, d_T, sizeof(float) * 9 * (size_t)n_guesses, cudaMemcpyDeviceToHost, ctx->stream));
    SFE_CUDA(cudaMemcpyAsync(iters_host, d_res, sizeof(int32_t) * (size_t)n_guesses, cudaMemcpyDeviceToHost,
                             ctx->stream));
    SFE_CUDA(cudaMemcpyAsync(inliers_host, d_res + n_guesses, sizeof(int32_t) * (size_t)n_guesses,
                             cudaMemcpyDeviceToHost, ctx->stream));
    SFE_CUDA(cudaMemcpyAsync(status_host, d_res + 2 * n_guesses, sizeof(int32_t) * (size_t)n_guesses,
                             cudaMemcpyDeviceToHost, ctx->stream));
    SFE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return SFE_OK;
}

}  // extern "C"
