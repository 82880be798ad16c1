// CFAR detector kernels (sm_100a).
//
// Replaces bruce_slam/src/bruce_slam/cpp/cfar.cpp:10-192 (ca/soca/goca/os and the
// "2" variants) for batches of polar sonar frames laid out [frame][range][beam].
//
// Two kernels:
//
//  cfar_ring_tma_kernel   the streaming path for the shipped configuration
//      (train_hs = 20, guard_hs = 5; CA / SOCA / GOCA).  One CTA owns a strip of
//      128 beams of one frame and marches down the range axis once.  A producer
//      warp streams [16 range bins x 128 beams] boxes of the frame through a
//      4-stage shared-memory ring with TMA (cp.async.bulk.tensor + mbarrier);
//      each of the 128 consumer threads owns ONE beam, reads every cell of its
//      beam exactly once from shared memory and keeps the last 64 range bins in
//      a register ring, so the leading/lagging window sums are two sliding
//      adds + two sliding subtracts per cell and nothing is ever re-read from
//      shared or global memory.  The 0/1 mask leaves through a double-buffered
//      shared tile and TMA stores.  HBM traffic = the image once + the mask
//      once (no halo: a strip spans the whole range axis).
//
//      Exactness: for integer-valued cells with |x| <= 65535 every float32 sum
//      is exact whatever the order, so sliding sums equal the reference's
//      sequential sums.  The double-precision threshold compare of cfar.cpp
//      (`img > tau * sum / train_hs`) is decided by a float32 estimate when the
//      cell is farther than a 1e-6 relative margin from the threshold and by the
//      identical double expression otherwise.  A strip that sees any other
//      value (fraction, |x| > 65535, inf, nan) raises a flag and is re-done by ...
//
//  cfar_exact_kernel      ... the general path: any train_hs/guard_hs, OS-CFAR, any
//      float input.  It accumulates each window in the reference's order
//      (ascending range, one float32 accumulator per sum) and evaluates the
//      threshold in double exactly as written in cfar.cpp, so it is bit-exact for
//      arbitrary float32 images.  It is a device path, not a CPU fallback.
#include "common.cuh"

namespace sfe {

constexpr int CF_W = 128;     // beams per strip / consumer threads per CTA
constexpr int CF_CH = 16;     // range bins per TMA box
constexpr int CF_NSTAGE = 4;  // input ring depth (stages of CF_CH rows)
constexpr int CF_RING = 32;   // two 32-deep register rings per thread (cells, window sums)
constexpr int CF_T = 20;      // train_hs of the streaming path
constexpr int CF_G = 5;       // guard_hs of the streaming path
constexpr int CF_HALF = CF_T + CF_G;

struct CfarParams {
  int F, R, B, strips;
  int alg, T, G, k;
  double tau;   // threshold factor (exact compare)
  double div;   // 2*T (CA), T (SOCA/GOCA), 1 (OS)
  float c;      // float(tau / div): float32 estimate of the threshold slope
  int gate_on;
  float gate_f;   // largest float g with (x > g) <=> ((double)x > gate) for every float x
  double gate_d;
  uint8_t *mask;  // may be null
  float *thr;     // may be null
  uint32_t *bits; // may be null
  int words_per_row;
  uint8_t *flags; // per (frame, strip): streaming path saw a non-integer cell
  const CUtensorMap *out_map;  // device-side: address of the kernel's __grid_constant__ copy
};

__device__ __forceinline__ float cell_to_float(float v) { return v; }
__device__ __forceinline__ float cell_to_float(uint8_t v) {
  // exact u8 -> f32 on the FMA/ALU pipes (no I2F): 2^23 + v, minus 2^23
  return __uint_as_float(0x4B000000u | (uint32_t)v) - 8388608.0f;
}

// Rare path of the streaming kernel: the float32 estimate is within its own
// uncertainty of the threshold, decide with the reference's double expression.
__device__ __noinline__ bool cfar_exact_compare(double tau, float S, double div, float xc) {
  return (double)xc > tau * (double)S / div;
}

struct CfarStep {  // per-thread streaming state (all in registers; indices are static after unrolling)
  float xr[CF_RING];  // x[rn - a]      at slot (j - a) & 31
  float wr[CF_RING];  // W[rn - a] = sum of the 20 cells ending at rn - a
  float w;            // W[rn - 1]
  float bad;          // > 0 once a non-integer cell was seen
  float mx;           // max |cell|
};

// One range bin.  J = position in the 32-step unrolled body; EDGE = this 16-row block
// touches the image border (row validity must be checked per row).
template <typename InT, int ALG, bool WITH_THR, bool MASK, bool BITS, bool EDGE, int J>
__device__ __forceinline__ void cfar_step(CfarStep &s, const InT (*__restrict__ tile)[CF_W],
                                          uint8_t (*__restrict__ obuf)[CF_CH][CF_W], const CfarParams &p,
                                          const int rn, const bool has_chunk, const int tid, const int f,
                                          const int col0, const float c, const float gate) {
  float xn = 0.f;
  if (!EDGE || has_chunk) xn = cell_to_float(tile[J % CF_CH][tid]);
  if (sizeof(InT) == 4) {
    // integer-valued?  (x + 1.5*2^23) - 1.5*2^23 == x  <=>  x integer and |x| < 2^22
    const float rt = (xn + 12582912.0f) - 12582912.0f;
    s.bad = fmaxf(s.bad, fabsf(rt - xn));
    s.mx = fmaxf(s.mx, fabsf(xn));
  }
  // cell under test r = rn - 25:
  //   lagging window r+6 .. r+25  = the 20 cells ending at rn       -> W[rn]
  //   leading window r-25 .. r-6  = the 20 cells ending at rn - 31  -> W[rn-31]
  const float x20 = s.xr[(J + CF_RING - CF_T) % CF_RING];
  const float xc = s.xr[(J + CF_RING - CF_HALF) % CF_RING];
  const float lead = s.wr[(J + CF_RING - (CF_HALF + CF_G + 1)) % CF_RING];
  const float lag = (s.w + xn) - x20;
  s.w = lag;
  s.wr[J] = lag;
  s.xr[J] = xn;

  const int r = rn - CF_HALF;
  if (EDGE && r < 0) return;
  bool pass = false;
  if (!EDGE || (r >= CF_HALF && r < p.R - CF_HALF)) {
    float S;
    if (ALG == SFE_CFAR_CA)
      S = lead + lag;
    else if (ALG == SFE_CFAR_SOCA)
      S = fminf(lead, lag);
    else
      S = fmaxf(lead, lag);
    if (WITH_THR) {
      const double d = p.tau * (double)S / p.div;
      pass = (double)xc > d;
      if (col0 + tid < p.B) p.thr[((size_t)f * p.R + r) * p.B + col0 + tid] = (float)d;
    } else {
      const float u = fmaf(-S, c, xc);       // xc - S*c
      const float m = fabsf(xc) * 2e-6f;      // >> float32 error of S*c and of the subtraction
      pass = u > m;
      if (fabsf(u) <= m) pass = cfar_exact_compare(p.tau, S, p.div, xc);
    }
    pass = pass && (xc > gate);  // gate = -inf when the amplitude gate is off
  }
  if (BITS) {
    if (!EDGE || r < p.R) {
      const unsigned b = __ballot_sync(0xffffffffu, pass && (col0 + tid < p.B));
      const int w = (col0 >> 5) + (tid >> 5);
      if ((tid & 31) == 0 && w < p.words_per_row) p.bits[((size_t)f * p.R + r) * p.words_per_row + w] = b;
    }
  }
  if (MASK) {
    constexpr int orow = (J + CF_RING - CF_HALF) % CF_CH;  // == r % 16
    // r = 32*body + J - 25  ->  (r / 16) & 1 depends on J only
    constexpr int ob = ((J + 2 * CF_RING - CF_HALF) / CF_CH) & 1;
    obuf[ob][orow][tid] = pass ? 1 : 0;
    if (orow == CF_CH - 1) {
      fence_proxy_async_smem();
      if (tid == 0) tma_wait_read<0>();  // the other buffer's store has finished reading
      named_bar_sync(1, CF_W);
      if (tid == 0) {
        tma_store_3d(p.out_map, &obuf[ob][0][0], col0, r - (CF_CH - 1), f);
        tma_commit();
      }
    }
  }
}

template <typename InT, int ALG, bool WITH_THR, bool MASK, bool BITS, bool EDGE, int Q>
__device__ __forceinline__ void cfar_block16(CfarStep &s, const InT (*__restrict__ tile)[CF_W],
                                             uint8_t (*__restrict__ obuf)[CF_CH][CF_W], const CfarParams &p,
                                             const int rn0, const bool has_chunk, const int tid, const int f,
                                             const int col0, const float c, const float gate) {
#define SFE_STEP(I) \
  cfar_step<InT, ALG, WITH_THR, MASK, BITS, EDGE, Q * CF_CH + I>(s, tile, obuf, p, rn0 + I, has_chunk, tid, f, col0, c, gate);
  SFE_STEP(0) SFE_STEP(1) SFE_STEP(2) SFE_STEP(3) SFE_STEP(4) SFE_STEP(5) SFE_STEP(6) SFE_STEP(7)
  SFE_STEP(8) SFE_STEP(9) SFE_STEP(10) SFE_STEP(11) SFE_STEP(12) SFE_STEP(13) SFE_STEP(14) SFE_STEP(15)
#undef SFE_STEP
}

template <typename InT, int ALG, bool WITH_THR, bool MASK, bool BITS>
__global__ void __launch_bounds__(CF_W + 32)
    cfar_ring_tma_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                         CfarParams p) {
  __shared__ __align__(128) InT tile[CF_NSTAGE][CF_CH][CF_W];
  __shared__ __align__(128) uint8_t obuf[2][CF_CH][CF_W];
  __shared__ __align__(8) uint64_t full_bar[CF_NSTAGE];
  __shared__ __align__(8) uint64_t empty_bar[CF_NSTAGE];

  const int tid = threadIdx.x;
  const int f = blockIdx.x / p.strips;
  const int col0 = (blockIdx.x % p.strips) * CF_W;
  const int R = p.R;
  p.out_map = &out_map;

  if (tid == 0) {
    for (int s = 0; s < CF_NSTAGE; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CF_W / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= CF_W) {
    // ------------------------------------------------------------ producer warp
    if (tid == CF_W) {
      prefetch_tmap(&in_map);
      const int nchunks = (R + CF_CH - 1) / CF_CH;
      for (int c = 0; c < nchunks; ++c) {
        const int st = c % CF_NSTAGE, it = c / CF_NSTAGE;
        if (it > 0) mbar_wait(&empty_bar[st], (it - 1) & 1);
        mbar_arrive_expect_tx(&full_bar[st], CF_CH * CF_W * (int)sizeof(InT));
        tma_load_3d(&tile[st][0][0], &in_map, &full_bar[st], col0, c * CF_CH, f);
      }
    }
    return;
  }

  // -------------------------------------------------------------- consumers (one beam each)
  CfarStep s;
#pragma unroll
  for (int i = 0; i < CF_RING; ++i) s.xr[i] = 0.f, s.wr[i] = 0.f;
  s.w = 0.f, s.bad = 0.f, s.mx = 0.f;
  const float c = p.c;
  const float gate = p.gate_on ? p.gate_f : -INFINITY;

  const int r_end = ((R + CF_CH - 1) / CF_CH) * CF_CH;  // rows emitted (TMA store clips rows >= R)
  const int rn_end = r_end + CF_HALF;                   // newest-row index runs [0, rn_end)
  const int nblk = (rn_end + CF_CH - 1) / CF_CH;        // 16-row blocks to run

  // block b covers newest rows rn0 = 16 b .. 16 b + 15; its TMA box (if any) sits in stage b % 4 and
  // completes phase (b / 4) & 1 of that stage's barrier.
  for (int b2 = 0; b2 * 2 < nblk; ++b2) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int b = b2 * 2 + q;
      if (b < nblk) {
        const int rn0 = b * CF_CH;
        const bool has_chunk = rn0 < R;
        const int st = b % CF_NSTAGE;
        if (has_chunk) mbar_wait(&full_bar[st], (b / CF_NSTAGE) & 1);
        const int r0 = rn0 - CF_HALF;
        const bool interior = (r0 >= CF_HALF) && (r0 + CF_CH - 1 < R - CF_HALF);
        if (interior) {
          if (q == 0)
            cfar_block16<InT, ALG, WITH_THR, MASK, BITS, false, 0>(s, tile[st], obuf, p, rn0, true, tid, f, col0, c, gate);
          else
            cfar_block16<InT, ALG, WITH_THR, MASK, BITS, false, 1>(s, tile[st], obuf, p, rn0, true, tid, f, col0, c, gate);
        } else {
          if (q == 0)
            cfar_block16<InT, ALG, WITH_THR, MASK, BITS, true, 0>(s, tile[st], obuf, p, rn0, has_chunk, tid, f, col0, c, gate);
          else
            cfar_block16<InT, ALG, WITH_THR, MASK, BITS, true, 1>(s, tile[st], obuf, p, rn0, has_chunk, tid, f, col0, c, gate);
        }
        if (has_chunk) {
          __syncwarp();
          if ((tid & 31) == 0) mbar_arrive(&empty_bar[st]);
        }
      }
    }
  }
  if (sizeof(InT) == 4 && (s.bad > 0.f || !(s.mx <= 262144.0f))) p.flags[blockIdx.x] = 1;
  if (MASK && tid == 0) tma_wait_all<0>();
}

// General / exact path.  One CTA per (frame, strip of CF_W beams); thread = beam.
template <typename InT>
__global__ void __launch_bounds__(CF_W) cfar_exact_kernel(const InT *__restrict__ img, const CfarParams p,
                                                          const uint8_t *__restrict__ only_flagged) {
  extern __shared__ float train_smem[];  // OS only: [2*T][CF_W]
  if (only_flagged != nullptr && only_flagged[blockIdx.x] == 0) return;
  const int tid = threadIdx.x;
  const int f = blockIdx.x / p.strips;
  const int col0 = (blockIdx.x % p.strips) * CF_W;
  const int col = col0 + tid;
  const bool live = col < p.B;
  const int R = p.R, B = p.B, T = p.T, G = p.G;
  const int half = T + G;
  const InT *colp = img + (size_t)f * R * B + (live ? col : 0);

  for (int r = 0; r < R; ++r) {
    bool pass = false;
    if (r >= half && r < R - half) {
      double d;
      if (p.alg == SFE_CFAR_CA) {
        float acc = 0.f;
        for (int i = r - half; i <= r + half; ++i)
          if (abs(i - r) > G) acc = __fadd_rn(acc, cell_to_float(colp[(size_t)i * B]));
        d = p.tau * (double)acc / p.div;
      } else if (p.alg == SFE_CFAR_OS) {
        int n = 0;
        for (int i = r - half; i <= r + half; ++i)
          if (abs(i - r) > G) train_smem[(n++) * CF_W + tid] = cell_to_float(colp[(size_t)i * B]);
        float v = __int_as_float(0x7fc00000);
        for (int a = 0; a < n; ++a) {  // k-th smallest by rank counting
          const float va = train_smem[a * CF_W + tid];
          int less = 0, leq = 0;
          for (int b = 0; b < n; ++b) {
            const float vb = train_smem[b * CF_W + tid];
            less += vb < va;
            leq += vb <= va;
          }
          if (less <= p.k && p.k < leq) {
            v = va;
            break;
          }
        }
        d = p.tau * (double)v;
      } else {
        float ld = 0.f, lg = 0.f;
        for (int i = r - half; i < r - G; ++i) ld = __fadd_rn(ld, cell_to_float(colp[(size_t)i * B]));
        for (int i = r + G + 1; i <= r + half; ++i) lg = __fadd_rn(lg, cell_to_float(colp[(size_t)i * B]));
        // std::min(lead, lag) / std::max(lead, lag) as written in cfar.cpp:46,69
        const float S = (p.alg == SFE_CFAR_SOCA) ? (lg < ld ? lg : ld) : (ld < lg ? lg : ld);
        d = p.tau * (double)S / p.div;
      }
      const float xc = cell_to_float(colp[(size_t)r * B]);
      pass = (double)xc > d;
      if (p.gate_on) pass = pass && ((double)xc > p.gate_d);
      if (p.thr != nullptr && live) p.thr[((size_t)f * R + r) * B + col] = (float)d;
    }
    pass = pass && live;
    if (p.mask != nullptr && live) p.mask[((size_t)f * R + r) * B + col] = pass ? 1 : 0;
    if (p.bits != nullptr) {
      const unsigned b = __ballot_sync(0xffffffffu, pass);
      const int w = (col0 >> 5) + (tid >> 5);
      if ((tid & 31) == 0 && w < p.words_per_row) p.bits[((size_t)f * R + r) * p.words_per_row + w] = b;
    }
  }
}

// ---------------------------------------------------------------------------- host side
// float g such that (x > g) <=> ((double)x > t) for every float x
static float gate_as_float(double t) {
  if (t != t) return __builtin_nanf("");  // x > nan is false, like numpy
  if (t >= 3.4028234663852886e38) return 3.4028234663852886e38f;  // only +inf can pass
  if (t < -3.4028234663852886e38) return -INFINITY;
  float g = (float)t;
  if ((double)g > t) g = nextafterf(g, -INFINITY);  // largest float <= t
  return g;
}

template <typename InT, int ALG, bool WITH_THR, bool MASK, bool BITS>
static int launch_ring(sfe_ctx *ctx, const CUtensorMap &in_map, const CUtensorMap &out_map, const CfarParams &p) {
  cfar_ring_tma_kernel<InT, ALG, WITH_THR, MASK, BITS>
      <<<p.F * p.strips, CF_W + 32, 0, ctx->stream>>>(in_map, out_map, p);
  SFE_CUDA(cudaGetLastError());
  ctx->launches++;
  return SFE_OK;
}

template <typename InT, int ALG>
static int launch_ring_out(sfe_ctx *ctx, const CUtensorMap &in_map, const CUtensorMap &out_map, const CfarParams &p) {
  const bool m = p.mask != nullptr, b = p.bits != nullptr;
  if (p.thr != nullptr) {  // the "2" variants: rarely used, one instantiation
    if (m && b) return launch_ring<InT, ALG, true, true, true>(ctx, in_map, out_map, p);
    if (m) return launch_ring<InT, ALG, true, true, false>(ctx, in_map, out_map, p);
    if (b) return launch_ring<InT, ALG, true, false, true>(ctx, in_map, out_map, p);
    return launch_ring<InT, ALG, true, false, false>(ctx, in_map, out_map, p);
  }
  if (m && b) return launch_ring<InT, ALG, false, true, true>(ctx, in_map, out_map, p);
  if (m) return launch_ring<InT, ALG, false, true, false>(ctx, in_map, out_map, p);
  if (b) return launch_ring<InT, ALG, false, false, true>(ctx, in_map, out_map, p);
  return SFE_OK;  // nothing to write
}

template <typename InT>
static int launch_ring_alg(sfe_ctx *ctx, const CUtensorMap &in_map, const CUtensorMap &out_map, const CfarParams &p) {
  switch (p.alg) {
    case SFE_CFAR_CA:
      return launch_ring_out<InT, SFE_CFAR_CA>(ctx, in_map, out_map, p);
    case SFE_CFAR_SOCA:
      return launch_ring_out<InT, SFE_CFAR_SOCA>(ctx, in_map, out_map, p);
    default:
      return launch_ring_out<InT, SFE_CFAR_GOCA>(ctx, in_map, out_map, p);
  }
}

template <typename InT>
static int launch_exact(sfe_ctx *ctx, const InT *img, const CfarParams &p, const uint8_t *only_flagged) {
  size_t smem = 0;
  if (p.alg == SFE_CFAR_OS) {
    smem = (size_t)2 * p.T * CF_W * sizeof(float);
    if (smem > 48 * 1024) {
      if (smem > (size_t)ctx->max_smem_optin) {
        set_error("OS-CFAR: train_hs=%d needs %zu B of shared memory (max %d)", p.T, smem, ctx->max_smem_optin);
        return SFE_ERR_UNSUPPORTED;
      }
      SFE_CUDA(cudaFuncSetAttribute(cfar_exact_kernel<InT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
  }
  cfar_exact_kernel<InT><<<p.F * p.strips, CF_W, smem, ctx->stream>>>(img, p, only_flagged);
  SFE_CUDA(cudaGetLastError());
  ctx->launches++;
  return SFE_OK;
}

int cfar_run(sfe_ctx *ctx, const void *img, int dtype, int F, int R, int B, int alg, int T, int G, int k,
             double tau, int gate_on, double gate, uint8_t *mask, float *thr, uint32_t *bits, int force_exact) {
  SFE_REQUIRE(ctx != nullptr, "cfar: null context");
  SFE_REQUIRE(img != nullptr || F == 0, "cfar: null image pointer");
  SFE_REQUIRE(dtype == SFE_U8 || dtype == SFE_F32, "cfar: dtype must be SFE_U8 or SFE_F32 (got %d)", dtype);
  SFE_REQUIRE(F >= 0 && R >= 0 && B >= 0, "cfar: negative shape (%d, %d, %d)", F, R, B);
  SFE_REQUIRE(alg >= SFE_CFAR_CA && alg <= SFE_CFAR_OS, "cfar: unknown variant %d", alg);
  SFE_REQUIRE(T >= 0 && G >= 0, "cfar: train_hs/guard_hs must be >= 0 (got %d, %d)", T, G);
  SFE_REQUIRE(alg != SFE_CFAR_OS || (k >= 0 && k < 2 * T), "cfar: OS rank k=%d outside [0, %d)", k, 2 * T);
  if (F == 0 || R == 0 || B == 0) return SFE_OK;

  CfarParams p{};
  p.F = F, p.R = R, p.B = B;
  p.strips = (B + CF_W - 1) / CF_W;
  p.alg = alg, p.T = T, p.G = G, p.k = k;
  p.tau = tau;
  p.div = alg == SFE_CFAR_CA ? 2.0 * T : (alg == SFE_CFAR_OS ? 1.0 : (double)T);
  p.c = (float)(tau / p.div);
  p.gate_on = gate_on != 0;
  p.gate_d = gate;
  p.gate_f = gate_as_float(gate);
  p.mask = mask, p.thr = thr, p.bits = bits;
  p.words_per_row = (B + 31) / 32;
  const size_t es = dtype == SFE_U8 ? 1 : 4;

  // the streaming path's preconditions (shape of the register ring; TMA alignment rules)
  bool fast = !force_exact && T == CF_T && G == CF_G && alg != SFE_CFAR_OS && R > 2 * CF_HALF;
  fast = fast && ((uintptr_t)img % 16 == 0) && ((size_t)B * es % 16 == 0);
  if (mask != nullptr) fast = fast && ((uintptr_t)mask % 16 == 0) && (B % 16 == 0);
  fast = fast && isfinite(tau) && isfinite((double)p.c);

  if (thr != nullptr) SFE_CUDA(cudaMemsetAsync(thr, 0, (size_t)F * R * B * sizeof(float), ctx->stream));

  if (fast) {
    CUtensorMap in_map, out_map;
    int rc = encode_tensor_map_3d(&in_map, dtype == SFE_U8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                  es, img, B, R, F, CF_W, CF_CH, 1);
    if (rc != SFE_OK) return rc;
    if (mask != nullptr) {
      rc = encode_tensor_map_3d(&out_map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, mask, B, R, F, CF_W, CF_CH, 1);
      if (rc != SFE_OK) return rc;
    } else {
      out_map = in_map;  // unused
    }
    if (dtype == SFE_F32) {
      rc = ensure(ctx, ctx->scratch[SCR_CFAR_FLAGS], (size_t)F * p.strips);
      if (rc != SFE_OK) return rc;
      p.flags = (uint8_t *)ctx->scratch[SCR_CFAR_FLAGS].ptr;
      SFE_CUDA(cudaMemsetAsync(p.flags, 0, (size_t)F * p.strips, ctx->stream));
      rc = launch_ring_alg<float>(ctx, in_map, out_map, p);
      if (rc != SFE_OK) return rc;
      return launch_exact<float>(ctx, (const float *)img, p, p.flags);  // re-does flagged strips only
    }
    return launch_ring_alg<uint8_t>(ctx, in_map, out_map, p);
  }
  if (dtype == SFE_F32) return launch_exact<float>(ctx, (const float *)img, p, nullptr);
  return launch_exact<uint8_t>(ctx, (const uint8_t *)img, p, nullptr);
}

}  // namespace sfe
