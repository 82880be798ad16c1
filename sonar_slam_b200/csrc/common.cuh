// Shared device/host helpers for libsonarfe (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/sonarfe.h"

namespace sfe {

// ---------------------------------------------------------------- error plumbing
void set_error(const char *fmt, ...);  // thread-local message behind sfe_last_error()

#define SFE_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      sfe::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      (void)cudaGetLastError(); /* do not leave the error for an unrelated later call */   \
      return SFE_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define SFE_REQUIRE(cond, ...)         \
  do {                                 \
    if (!(cond)) {                     \
      sfe::set_error(__VA_ARGS__);     \
      return SFE_ERR_ARG;              \
    }                                  \
  } while (0)

// ---------------------------------------------------------------- context
struct Buffer {  // grow-only device scratch
  void *ptr = nullptr;
  size_t cap = 0;
};

}  // namespace sfe

struct sfe_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  int sm_count = 148;
  int max_smem_optin = 0;
  sfe::Buffer scratch[8];   // per-subsystem scratch (see SCR_* below)
  sfe::Buffer stage_in[2];  // device staging for the *_host entry points
  sfe::Buffer stage_out[4];
  void *pinned = nullptr;   // small pinned host block for async result reads
  size_t pinned_cap = 0;
  uint64_t launches = 0;    // kernels launched through this context (bench: gpu_launches)
  // per-context caches (they describe buffers / attributes owned by THIS context, so they live here and not in
  // thread-local statics: a new context may get a recycled device address, two contexts may share a thread)
  double cfar_lut_key[6] = {-1, 0, 0, 0, 0, 0};  // parameters of the table held in scratch[SCR_CFAR_LUT]
  const void *cfar_lut_buf = nullptr;            // ... and the buffer it was uploaded to (null = none)
  struct OccEntry { size_t smem; int threads, per_sm, variant; } icp_occ[8] = {};
  int icp_occ_next = 0;
};

struct sfe_maps {  // per-geometry polar->Cartesian sampling table (featx.cu)
  int rows, cols;  // Cartesian image
  int R, B;        // polar image
  double width, height;
  void *table;     // device MapEntry[rows*cols]
  int32_t *inv_off;  // device [R*B + 1]: for every polar cell, the Cartesian pixels it can light ...
  int32_t *inv_idx;  // ... as a CSR list of pixel indices
  float *metres;     // device [cols + rows]: lateral metres of every column, then forward metres of every row
                     // (feature_extraction.py:235-237 evaluated per column / per row in float64 on the host)
  int device;
};

namespace sfe {
enum { SCR_CFAR_FLAGS = 0, SCR_FEATX = 1, SCR_CLOUD = 2, SCR_ICP = 3, SCR_MISC = 4, SCR_CFAR_LUT = 5 };

int ensure(sfe_ctx *ctx, Buffer &b, size_t bytes);  // (re)allocate if too small
int ensure_pinned(sfe_ctx *ctx, size_t bytes);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda needed)
int encode_tensor_map_3d(CUtensorMap *map, CUtensorMapDataType dt, size_t elem_bytes, const void *base,
                         uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1, uint32_t b2);

// ---------------------------------------------------------------- device PTX wrappers
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// try_wait with a suspend-time hint: a waiting thread may sleep in hardware until the phase completes (or the
// hint expires) instead of spinning on the barrier.
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// TMA: global -> shared tile load of a 3-D tensor (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// TMA: shared -> global tile store (SASS: UTMASTG)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#endif  // __CUDACC__

}  // namespace sfe
