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
