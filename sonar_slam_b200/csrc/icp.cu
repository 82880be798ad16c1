// ICP scan matcher and nearest-neighbour matcher (sm_100a).
//
// Replaces bruce_slam.pcl.ICP.compute and bruce_slam.pcl.match
// (bruce_slam/src/bruce_slam/cpp/pcl.cpp:161-174,185-212), i.e. libpointmatcher's
// PM::ICP::operator() under bruce_slam/config/icp.yaml: KD-tree NN (knn 1, maxDist 10) ->
// MaxDist(3.0) x TrimmedDist(0.8) outlier weights -> point-to-point rigid fit ->
// Counter(40) + Differential(0.01 rad, 0.1 m, smooth 4) checkers, all float32.
// Algorithm statement and conventions: oracle/icp_ref.c (same steps, same float32 expressions).
//
// One CTA solves one (source, target, guess) problem end to end: the target is centred on its
// mean and counting-sorted into a uniform grid held in shared memory (grid.cuh), the source is
// moved into the centred frame once, and every iteration runs entirely on chip:
//   transform + grid NN search per source point (seeded with the previous iteration's match: only the
//   cells overlapping the seed's disc are scanned)  ->  exact k-th smallest distance (4-pass radix
//   select on the float bits)  ->  0/1 weights  ->  warp-shuffle / shared-memory reductions of the
//   weighted means and the 2x2 cross-covariance (float32 products, float64 accumulation, fixed
//   order: deterministic)  ->  closed-form 2-D rotation and translation, T_iter update and the
//   Counter / Differential checkers on one thread.
// CTAs are persistent over the problem list (batch of keyframe pairs / of initial guesses).
#include <mutex>

#include "grid.cuh"

namespace sfe {

#ifndef SFE_SEQ_PIPE_MIN
#define SFE_SEQ_PIPE_MIN 256  // CTA sizes from which the sequential sums are software-pipelined (register budget)
#endif
constexpr int ICP_THREADS = 512;
constexpr int ICP_HIST = 16;  // differential-checker history kept (>= smoothLength + 1)
constexpr int ICP_COARSE = 8;            // fine cells per coarse cell edge
constexpr int ICP_COARSE_WORDS = 96;     // bitmap words: cnx*cny <= 2048 + margin
constexpr int ICP_PLANE_SCRATCH = 2 * 320;  // floats (seq_sum9_warp: two buffers of 9 rows, 36 apart, 32 used)
constexpr float ICP_PRUNED = 3.0e38f;    // "finite, but farther than we needed to know"

enum { ICP_OK = 0, ICP_NO_OUTLIER = 1, ICP_NO_POINT = 2, ICP_NAN_ROT = 3, ICP_NAN_TRANS = 4, ICP_NOT_RIGID = 5,
       ICP_EMPTY_REF = 6, ICP_SKIPPED = 7, ICP_TOO_LARGE = 8 };

struct IcpBatch {
  const float *src_pts;
  const int *src_off;
  const float *tgt_pts;
  const int *tgt_off;
  const int *src_cnt;  // optional counts (else off[c+1]-off[c])
  const int *tgt_cnt;
  int min_points;      // problems whose source or target has fewer points are skipped (ICP_SKIPPED)
  const int *raw_cnt;  // optional: un-clamped size of the source's raw cloud; > raw_cap -> ICP_TOO_LARGE
  int raw_cap;
  // size classes: a batch can be served by two launches with different shared-memory footprints.
  //   1: this launch only takes problems that fit (ns <= ns_max, nt <= nt_max) and leaves the others untouched;
  //   2: this launch skips the problems with ns <= class_ns and nt <= class_nt (the other launch's share)
  int class_mode, class_ns, class_nt;
  const int *src_id;  // may be null (problem p uses source p)
  const int *tgt_id;  // may be null
  const float *guess; // [P][9] row-major 3x3
  float *T_out;       // [P][9]
  int *iters, *inliers, *status;
  int P, ns_max, nt_max, max_cells;
  float cell_scale;  // grid cell = cell_scale * sqrt(area / n)
  sfe_icp_params prm;
  uint16_t *orig_ws;  // [slots][nt_max]
  float2 *nrm_ws;     // [slots][nt_max] point-to-plane only: normal of the target point at every sorted position
  int use_order;      // search passes visit the source points in spatially sorted order (big problems only)
  float margin_mult;  // certificate size in units of the point's last step (one-pass path)
  int small_mult;     // problems with ns <= small_mult * blockDim.x (and nt <= 4096) take the one-pass exact path
  int slot_by_smid;   // workspace slot = %smid (one CTA per SM, one CTA per problem) instead of blockIdx.x
};

constexpr int ICP_SM_SLOTS = 256;  // >= %nsmid on every sm_100 part
__device__ __forceinline__ unsigned icp_smid() {
  unsigned r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}

__device__ __forceinline__ void mat3_mul_rn(const float *a, const float *b, float *c) {
  float r[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float acc = __fmul_rn(a[i * 3 + 0], b[0 * 3 + j]);
      acc = __fadd_rn(acc, __fmul_rn(a[i * 3 + 1], b[1 * 3 + j]));
      acc = __fadd_rn(acc, __fmul_rn(a[i * 3 + 2], b[2 * 3 + j]));
      r[i * 3 + j] = acc;
    }
#pragma unroll
  for (int i = 0; i < 9; ++i) c[i] = r[i];
}

__device__ __forceinline__ float2 apply_T(const float *T, float x, float y) {
  float2 o;
  o.x = __fadd_rn(__fadd_rn(__fmul_rn(T[0], x), __fmul_rn(T[1], y)), T[2]);
  o.y = __fadd_rn(__fadd_rn(__fmul_rn(T[3], x), __fmul_rn(T[4], y)), T[5]);
  return o;
}

// Eigen's rotation-matrix -> quaternion conversion for a 2-D rotation embedded in 3x3 (w and z)
__device__ __forceinline__ void rot_to_quat(const float *T, float &qw, float &qz) {
  const float m00 = T[0], m01 = T[1], m10 = T[3], m11 = T[4];
  float t = __fadd_rn(__fadd_rn(m00, m11), 1.0f);
  if (t > 0.f) {
    t = sqrtf(__fadd_rn(t, 1.0f));
    qw = __fmul_rn(0.5f, t);
    t = __fdiv_rn(0.5f, t);
    qz = __fmul_rn(__fsub_rn(m10, m01), t);
  } else {
    t = sqrtf(__fadd_rn(__fsub_rn(__fsub_rn(1.0f, m00), m11), 1.0f));
    qz = __fmul_rn(0.5f, t);
    t = __fdiv_rn(0.5f, t);
    qw = __fmul_rn(__fsub_rn(m10, m01), t);
  }
}

// Sum K doubles over the CTA with ONE barrier: every warp publishes its partial sums, and after the barrier
// every warp reduces the (<= 16) partials itself with a butterfly, so all threads hold the bit-identical total.
// `scratch` has two halves used alternately (`phase` counts the calls): the half written by call n is last read
// before call n+1's barrier and not rewritten before call n+2.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double *scratch /* [2][5][16] smem */, int &phase,
                                          double (&out)[K]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  static_assert(K <= 5, "scratch holds five partial sums per warp");
  double *half = scratch + (phase & 1) * (5 * 16);
  ++phase;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double x = v[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    if (lane == 0) half[k * 16 + warp] = x;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double x = lane < nwarps ? half[k * 16 + lane] : 0.0;
#pragma unroll
    for (int d = 8; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);  // nwarps <= 16
    out[k] = __shfl_sync(0xffffffffu, x, 0);
  }
}

// same for one int
__device__ __forceinline__ int block_total(int v, int *scratch /* [2][16] smem */, int &phase) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  int *half = scratch + (phase & 1) * 16;
  ++phase;
  v = __reduce_add_sync(0xffffffffu, v);
  if (lane == 0) half[warp] = v;
  __syncthreads();
  return __reduce_add_sync(0xffffffffu, lane < nwarps ? half[lane] : 0);
}

struct IcpShared {  // small fixed-size part of the shared state
  float Ti[9];
  float Tprev[9];  // T_iter of the previous iteration (how far did every point move since?)
  float T0[9];
  float mean[2];
  float bbox[4];
  int iterate, status, count, inliers;
  int sel_bin, sel_k;
  uint32_t sel_val;  // radix select: the element found alone in its bin
  int hist[3][256];  // radix-select histograms, used in rotation (block_select_kth)
  int scan[36];
  int tot[2][16];
  double red[2 * 5 * 16];  // <= 16 warps per CTA
  float wred[4 * 16];      // per-warp partials of the bounding-box / maximum reductions
  float hq_w[ICP_HIST], hq_z[ICP_HIST], ht_x[ICP_HIST], ht_y[ICP_HIST];
  int hn;
  float seq[4];               // results of the sequential sums
  __align__(16) float seq_buf[2 * 4 * 36];  // staging of seq_sum4_warp
  int cnx, cny;               // coarse occupancy grid (ICP_COARSE x ICP_COARSE fine cells per coarse cell)
  uint32_t coarse[ICP_COARSE_WORDS];
};

// k-th smallest (0-based) of the finite entries of vals[0..n): 4-pass radix select on the float bits
// (non-negative floats order like their bit patterns).  All threads call it; kk < number of finite entries.
// One barrier per pass: pass number `pass` (counted over the CTA's lifetime) fills histogram pass % 3 and clears
// histogram (pass + 1) % 3 -- last read two passes ago -- before its barrier; after the barrier every warp scans
// the 256 bins itself (8 per lane), so no second barrier is needed to publish the selected bin.
__device__ __forceinline__ float block_select_kth(const float *vals, int n, int kk, IcpShared &sh, int &pass) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
  uint32_t prefix = 0, mask = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    int *hist = sh.hist[pass % 3], *next = sh.hist[(pass + 1) % 3];
    ++pass;
    for (int h = tid; h < 256; h += nthr) next[h] = 0;
    for (int i = tid; i < n; i += nthr) {
      const float v = vals[i];
      if (!(v < INFINITY)) continue;
      const uint32_t u = __float_as_uint(v);
      if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1);
    }
    __syncthreads();
    int c[8], sum = 0;  // lane l owns bins 8l .. 8l+7
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j] = hist[lane * 8 + j], sum += c[j];
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    // the lane whose 8 bins contain rank kk, then lanes 0..7 look at those 8 bins together
    const unsigned owner = __ballot_sync(0xffffffffu, kk >= incl - sum && kk < incl);
    const int src = owner ? __ffs(owner) - 1 : 31;
    const int base = __shfl_sync(0xffffffffu, incl - sum, src);  // number of entries in the bins before lane src's
    const int cj = hist[src * 8 + (lane & 7)];
    int ij = cj;
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, ij, d);
      if ((lane & 7) >= d) ij += t;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, kk - base < ij) & 0xffu;
    const int j = hit ? __ffs(hit) - 1 : 7;
    const int bin = src * 8 + j;
    const int in_bin = __shfl_sync(0xffffffffu, cj, j);
    kk = kk - base - (__shfl_sync(0xffffffffu, ij, j) - in_bin);
    prefix |= (uint32_t)bin << shift;
    mask |= 255u << shift;
    if (in_bin == 1 && shift > 0) {
      // the wanted element is alone in its bin: no need to resolve its remaining bits digit by digit -- the one
      // thread that holds it publishes it (typically after two of the four passes for a few hundred values)
      for (int i = tid; i < n; i += nthr) {
        const float v = vals[i];
        if (v < INFINITY && (__float_as_uint(v) & mask) == prefix) sh.sel_val = __float_as_uint(v);
      }
      __syncthreads();
      return __uint_as_float(sh.sel_val);  // (next written after at least one more barrier)
    }
  }
  return __uint_as_float(prefix);
}

// Exact SEQUENTIAL float32 sums in index order (the accumulation order of oracle/icp_ref.c; default mode), computed
// by one warp.  The dependent chain is the float add alone (~4 cycles per term); terms beyond n (or of dropped
// pairs) are +0, which leaves a sum unchanged bit for bit.  All 32 lanes of the warp must call these.
// Four sequential sums at once (the components of one float4 term per point), by ONE warp with few instructions:
// lane j evaluates the four terms of point base + j and parks them, transposed, in shared memory ([component][32],
// rows padded to 36 floats so that the four reading lanes hit different banks); lanes 0..3 then add "their"
// component's 32 values in order (eight LDS.128 + 32 dependent adds).  ~55 warp instructions per 32 points for all
// four sums; result: lane c (c < 4) returns the sum of component c.
// `scratch` = [2][4][36] floats (double-buffered: a batch is written while the previous one may still be read).
template <bool PIPELINED, typename F>
__device__ __forceinline__ float seq_sum4_warp(int n, float *scratch, F term4) {
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  if (!PIPELINED) {
    int it = 0;
    for (int base = 0; base < n; base += 32, ++it) {
      float *buf = scratch + (it & 1) * (4 * 36);
      const float4 t = base + lane < n ? term4(base + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      buf[0 * 36 + lane] = t.x, buf[1 * 36 + lane] = t.y, buf[2 * 36 + lane] = t.z, buf[3 * 36 + lane] = t.w;
      __syncwarp();
      if (lane < 4) {
        const float4 *v = reinterpret_cast<const float4 *>(buf + lane * 36);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 a = v[k];
          s = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s, a.x), a.y), a.z), a.w);
        }
      }
      // (the buffer written two batches from now is this one: every lane passes the next batch's __syncwarp first)
    }
    return s;
  }
  // Software-pipelined form (CTA sizes with registers to spare): while lanes 0..3 add batch k from registers, all
  // lanes evaluate and park batch k + 1, so the add chain never waits for shared memory.
  float4 cur[8];
  {
    const float4 t = lane < n ? term4(lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    scratch[0 * 36 + lane] = t.x, scratch[1 * 36 + lane] = t.y, scratch[2 * 36 + lane] = t.z, scratch[3 * 36 + lane] = t.w;
    __syncwarp();
    const float4 *v = reinterpret_cast<const float4 *>(scratch + (lane & 3) * 36);
#pragma unroll
    for (int k = 0; k < 8; ++k) cur[k] = v[k];
  }
  int it = 1;
  for (int base = 0; base < n; base += 32, ++it) {
    const int nb = base + 32;
    float *buf = scratch + (it & 1) * (4 * 36);
    if (nb < n) {  // (warp-uniform) park the next batch
      const float4 t = nb + lane < n ? term4(nb + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      buf[0 * 36 + lane] = t.x, buf[1 * 36 + lane] = t.y, buf[2 * 36 + lane] = t.z, buf[3 * 36 + lane] = t.w;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
      s = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s, cur[k].x), cur[k].y), cur[k].z), cur[k].w);
    __syncwarp();
    if (nb < n) {
      const float4 *v = reinterpret_cast<const float4 *>(buf + (lane & 3) * 36);
#pragma unroll
      for (int k = 0; k < 8; ++k) cur[k] = v[k];
    }
  }
  return s;  // lanes c, c + 4, c + 8, ... all hold the sum of component c
}

// Point-to-plane: the nine sequential sums of the normal equations in one pass.  Same scheme as seq_sum4_warp with
// nine rows -- lanes 0..8 each add one component's 32 values per batch, in point order -- and the per-pair normal
// (a gather from global memory, ~700 cycles) requested for eight batches at once so that its latency is paid once
// per 256 points, not once per batch.  `scratch` = 2 x 320 floats (nine rows 36 floats apart, 32 used).  term9(i, m, n, v): the nine terms of kept
// pair i (match m, normal n).  Lane c (c < 9) returns the sum of component c.
template <typename F>
__device__ __forceinline__ float seq_sum9_warp(int n, float *scratch, const uint16_t *match, const float2 *nrm, F term9) {
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  int it = 0;
  for (int chunk = 0; chunk < n; chunk += 256) {
    int mm[8];
    float2 nn[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = chunk + 32 * k + lane;
      mm[k] = i < n ? match[i] : 0xffff;
      nn[k] = mm[k] != 0xffff ? nrm[mm[k]] : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (chunk + 32 * k >= n) break;  // (warp-uniform)
      float *buf = scratch + (it & 1) * 320;
      ++it;
      float v[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (mm[k] != 0xffff) term9(chunk + 32 * k + lane, mm[k], nn[k], v);
#pragma unroll
      for (int c = 0; c < 9; ++c) buf[c * 36 + lane] = v[c];
      __syncwarp();
      if (lane < 9) {  // (8-byte loads: the scratch may sit on dist[], which is only 8-byte aligned)
        const float2 *r = reinterpret_cast<const float2 *>(buf + lane * 36);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 a = r[j];
          s = __fadd_rn(__fadd_rn(s, a.x), a.y);
        }
      }
      // (this buffer is rewritten two batches from now: every lane passes the next batch's __syncwarp first)
    }
  }
  return s;
}

// One sequential sum over values in global memory, p[0], p[stride], ... (the reference cloud's mean, once per
// problem): lane j holds term base + j and a shuffle hands term j to every lane, so all lanes carry the same
// running sum; eight batches (256 terms) are requested at once and the next eight are in flight while these are
// added, so the L2 latency hides behind the ~1000-cycle add chain
__device__ __forceinline__ float seq_sum_warp_global(const float *p, int n, int stride) {
  const int lane = threadIdx.x & 31;
  float s = 0.f, cur[8], nxt[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = k * 32 + lane;
    cur[k] = i < n ? p[(size_t)i * stride] : 0.f;
  }
  for (int base = 0; base < n; base += 256) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = base + 256 + k * 32 + lane;
      nxt[k] = i < n ? p[(size_t)i * stride] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (base + k * 32 < n) {  // (warp-uniform)
#pragma unroll
        for (int j = 0; j < 32; ++j) s = __fadd_rn(s, __shfl_sync(0xffffffffu, cur[k], j));
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) cur[k] = nxt[k];
  }
  return s;
}

// One instantiation per CTA size so that the register budget follows the launch shape (a single
// __launch_bounds__(512) build capped the 128-thread class at 64 registers and spilled).
// PLANE: errorMinimizer = PointToPlaneErrorMinimizer (sfe_icp_params.minimizer 1); a separate instantiation so that
// the shipped point-to-point kernels keep their register allocation.
template <int THREADS, int MINB, bool PLANE>
__global__ void __launch_bounds__(THREADS, MINB) icp_kernel(const IcpBatch b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [IcpShared][sorted float2 nt_max][cells u32][reading float2 ns_max][dist f32 ns_max][match u16 ns_max]
  //         [prev u16 ns_max][qstate u8 ns_max][slack f32 ns_max][order u16 ns_max]
  IcpShared &sh = *reinterpret_cast<IcpShared *>(smem_raw);
  size_t off = (sizeof(IcpShared) + 15) & ~size_t(15);
  float2 *sorted = reinterpret_cast<float2 *>(smem_raw + off);
  off += sizeof(float2) * (size_t)b.nt_max;
  uint32_t *cells = reinterpret_cast<uint32_t *>(smem_raw + off);
  off += sizeof(uint32_t) * (size_t)((b.max_cells + 2) / 2 + 1);
  off = (off + 7) & ~size_t(7);
  float2 *reading = reinterpret_cast<float2 *>(smem_raw + off);
  off += sizeof(float2) * (size_t)b.ns_max;
  float *dist = reinterpret_cast<float *>(smem_raw + off);
  off += sizeof(float) * (size_t)b.ns_max;
  uint16_t *match = reinterpret_cast<uint16_t *>(smem_raw + off);
  off += sizeof(uint16_t) * (size_t)b.ns_max;
  uint16_t *prev = reinterpret_cast<uint16_t *>(smem_raw + off);  // NN of the previous iteration (search seed)
  off += sizeof(uint16_t) * (size_t)b.ns_max;
  uint8_t *qstate = reinterpret_cast<uint8_t *>(smem_raw + off);
  off += ((size_t)b.ns_max + 3) & ~size_t(3);
  // certified-match margin per source point (metres): every target point other than prev[i] is farther from the
  // point's current position than (distance to prev[i]) as long as that distance stays below slack[i]
  float *slack = reinterpret_cast<float *>(smem_raw + off);
  off += sizeof(float) * (size_t)b.ns_max;
  // processing order of the source points in the search passes: sorted by a coarse cell of the target grid, so that
  // the lanes of a warp look at neighbouring cells (similar amounts of work, the same shared-memory lines); results
  // are stored by point index and the sums run in point order, so the order is invisible in the results
  uint16_t *order = reinterpret_cast<uint16_t *>(smem_raw + off);
  if (b.use_order) off += sizeof(uint16_t) * (size_t)b.ns_max;
  off = (off + 15) & ~size_t(15);
  // PLANE only: 640 floats for seq_sum9_warp.  dist[] is dead between the weights (3c) and the next iteration's
  // search, which rewrites all of it: launches with room for 640 source points (the front end's small class is
  // exactly that, and sized to six CTAs per SM to the byte; 2 000 x 20 000 points leave nothing to spare at all)
  // park the scratch there, smaller ones get their own 2.5 KB.
  float *pl_scratch = b.ns_max >= ICP_PLANE_SCRATCH ? dist : reinterpret_cast<float *>(smem_raw + off);
  (void)pl_scratch;

  const int tid = threadIdx.x, nthr = blockDim.x;
  int red_phase = 0, tot_phase = 0, sel_pass = 0;  // rotation counters of the one-barrier reductions (CTA-uniform)
  for (int h = threadIdx.x; h < 3 * 256; h += blockDim.x) (&sh.hist[0][0])[h] = 0;
  // Problems that allow one CTA per SM (config 3: ~200 KB of shared memory) are launched one CTA per problem, not
  // as persistent CTAs: problems of unequal length (checker mode: 4-40 iterations) balance themselves, and SMs
  // free up every few milliseconds so that concurrent NCCL send/recv kernels (the config-5 scatter) get scheduled.
  // The per-CTA global workspace is then indexed by the SM the CTA runs on.
  const unsigned slot = b.slot_by_smid ? icp_smid() : blockIdx.x;
  if (b.slot_by_smid && slot >= ICP_SM_SLOTS) __trap();
  uint16_t *orig = b.orig_ws + (size_t)slot * b.nt_max;
  const sfe_icp_params prm = b.prm;
  const float max_d2 = __fmul_rn(prm.matcher_max_dist, prm.matcher_max_dist);
  const float out_d2 = (prm.flags & 1) ? prm.outlier_max_dist : __fmul_rn(prm.outlier_max_dist, prm.outlier_max_dist);
  const int smooth = min(prm.smooth_length, ICP_HIST - 1);
  // Default: every sum over points is a sequential float32 sum in point order, as in oracle/icp_ref.c -- results
  // are bit-identical to the oracle.  flags bit 1: float32 terms accumulated in float64 (order-independent,
  // ~1e-4 closer to exact arithmetic, a few per cent faster; deviates from the oracle by up to ~2e-3 m on
  // ill-conditioned scans because the float32-sequential sums themselves are that noisy).
  const bool seq = (prm.flags & 2) == 0;

  for (int p = blockIdx.x; p < b.P; p += gridDim.x) {
    const int si = b.src_id ? b.src_id[p] : p, ti = b.tgt_id ? b.tgt_id[p] : p;
    const float *src = b.src_pts + 2 * (size_t)b.src_off[si];
    const int ns = b.src_cnt ? b.src_cnt[si] : b.src_off[si + 1] - b.src_off[si];
    const float *tgt = b.tgt_pts + 2 * (size_t)b.tgt_off[ti];
    const int nt = b.tgt_cnt ? b.tgt_cnt[ti] : b.tgt_off[ti + 1] - b.tgt_off[ti];
    const float *guess = b.guess + 9 * (size_t)p;
    if (b.class_mode == 1 && (ns > b.ns_max || nt > b.nt_max)) continue;
    if (b.class_mode == 2 && ns <= b.class_ns && nt <= b.class_nt) continue;
    __syncthreads();  // previous problem fully retired before shared state is reused

    // ---- admission checks (thread 0), failure leaves T = guess
    if (tid == 0) {
      int st = ICP_OK;
      if (ns > b.ns_max || nt > b.nt_max || (b.raw_cnt && b.raw_cnt[si] > b.raw_cap)) st = ICP_TOO_LARGE;
      else if (ns < b.min_points || nt < b.min_points) st = ICP_SKIPPED;
      else if (nt <= 0) st = ICP_EMPTY_REF;
      const float det = __fsub_rn(__fmul_rn(guess[0], guess[4]), __fmul_rn(guess[1], guess[3]));
      if (st == ICP_OK && (!(fabsf(__fsub_rn(1.0f, det)) <= 0.001f))) st = ICP_NOT_RIGID;
      sh.status = st;
      sh.count = 0, sh.inliers = 0;
    }
    __syncthreads();
    if (sh.status != ICP_OK) {
      if (tid < 9) b.T_out[9 * (size_t)p + tid] = guess[tid];
      if (tid == 0) b.iters[p] = 0, b.inliers[p] = 0, b.status[p] = sh.status;
      continue;
    }

    // ---- 1. mean of the reference, bounding box of the centred reference
    {
      float mx, my;
      if (seq) {
        if (tid < 64) {  // warps 0 and 1: x and y
          const float sum = seq_sum_warp_global(tgt + (tid >> 5), nt, 2);
          if ((tid & 31) == 0) sh.seq[tid >> 5] = __fdiv_rn(sum, (float)nt);
        }
        __syncthreads();
        mx = sh.seq[0], my = sh.seq[1];
      } else {
        double s[2] = {0.0, 0.0}, tot[2];
        for (int i = tid; i < nt; i += nthr) s[0] += (double)tgt[2 * i], s[1] += (double)tgt[2 * i + 1];
        block_sum<2>(s, sh.red, red_phase, tot);
        mx = (float)(tot[0] / (double)nt), my = (float)(tot[1] / (double)nt);
      }
      float mn_x = INFINITY, mn_y = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
      for (int i = tid; i < nt; i += nthr) {
        const float x = tgt[2 * i] - mx, y = tgt[2 * i + 1] - my;
        mn_x = fminf(mn_x, x), mxx = fmaxf(mxx, x), mn_y = fminf(mn_y, y), mxy = fmaxf(mxy, y);
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        mn_x = fminf(mn_x, __shfl_xor_sync(0xffffffffu, mn_x, d));
        mn_y = fminf(mn_y, __shfl_xor_sync(0xffffffffu, mn_y, d));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, d));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, d));
      }
      float *fr = sh.wred;
      if ((tid & 31) == 0) {
        fr[(tid >> 5) * 4 + 0] = mn_x, fr[(tid >> 5) * 4 + 1] = mn_y;
        fr[(tid >> 5) * 4 + 2] = mxx, fr[(tid >> 5) * 4 + 3] = mxy;
      }
      __syncthreads();
      if (tid == 0) {
        float a = INFINITY, bb = INFINITY, c = -INFINITY, d = -INFINITY;
        for (int w = 0; w < (nthr >> 5); ++w) {
          a = fminf(a, fr[w * 4 + 0]), bb = fminf(bb, fr[w * 4 + 1]);
          c = fmaxf(c, fr[w * 4 + 2]), d = fmaxf(d, fr[w * 4 + 3]);
        }
        sh.bbox[0] = a, sh.bbox[1] = bb, sh.bbox[2] = c, sh.bbox[3] = d;
        sh.mean[0] = mx, sh.mean[1] = my;
      }
      __syncthreads();
    }
    const float mx = sh.mean[0], my = sh.mean[1];
    GridView g;
    grid_geometry(nt, sh.bbox[0], sh.bbox[1], sh.bbox[2], sh.bbox[3], 0.05f, g, b.max_cells, b.cell_scale);
    grid_build(tgt, 2, nt, mx, my, g, sorted, cells, orig, sh.scan);
    // coarse occupancy: which 8x8 blocks of cells hold any point (answers "is anything within maxDist?"
    // for far-away source points without walking the fine grid)
    {
      const int cnx = (g.nx + ICP_COARSE - 1) / ICP_COARSE, cny = (g.ny + ICP_COARSE - 1) / ICP_COARSE;
      if (tid == 0) sh.cnx = cnx, sh.cny = cny;
      for (int w = tid; w < ICP_COARSE_WORDS; w += nthr) sh.coarse[w] = 0;
      __syncthreads();
      if (cnx * cny <= ICP_COARSE_WORDS * 32) {
        for (int cc = tid; cc < cnx * cny; cc += nthr) {
          const int ccy = cc / cnx, ccx = cc - ccy * cnx;
          bool occ = false;
          for (int y = ccy * ICP_COARSE; y < min((ccy + 1) * ICP_COARSE, g.ny) && !occ; ++y) {
            const int xa = ccx * ICP_COARSE, xb = min((ccx + 1) * ICP_COARSE, g.nx);
            occ = g.cstart[y * g.nx + xb] > g.cstart[y * g.nx + xa];
          }
          if (occ) atomicOr(&sh.coarse[cc >> 5], 1u << (cc & 31));
        }
      } else if (tid == 0) {
        sh.cnx = 0;  // too many coarse cells: the test is skipped
      }
      __syncthreads();
    }

    // ---- 1b. point-to-plane: normals of the (centred) target, one thread per sorted position, kept in global
    // memory (8 B per target point; the iterations gather them through L2)
    float2 *nrm = nullptr;
    if constexpr (PLANE) {
      nrm = b.nrm_ws + (size_t)slot * b.nt_max;
      const int knn = min(max(prm.normals_knn, 1), GRID_KNN_MAX);
      // best-knn list per thread: the per-source arrays (reading .. slack) are idle until step 2
      unsigned char *scr = reinterpret_cast<unsigned char *>(reading);
      const size_t scr_bytes = (size_t)(reinterpret_cast<unsigned char *>(slack + b.ns_max) - scr);
      if (scr_bytes >= (size_t)nthr * knn * 6) {
        float *bd = reinterpret_cast<float *>(scr) + tid;
        uint16_t *bi = reinterpret_cast<uint16_t *>(scr + sizeof(float) * (size_t)nthr * knn) + tid;
        for (int j = tid; j < nt; j += nthr) nrm[j] = grid_surface_normal(g, j, knn, bd, bi, nthr);
      } else {
        float bd[GRID_KNN_MAX];
        uint16_t bi[GRID_KNN_MAX];
        for (int j = tid; j < nt; j += nthr) nrm[j] = grid_surface_normal(g, j, knn, bd, bi, 1);
      }
      // (published to the CTA, and the scratch released, by the barriers of step 2)
    }

    // ---- 2. reading into the centred frame; T_iter = I; checker history
    if (tid == 0) {
      const float Tinv[9] = {1.f, 0.f, -mx, 0.f, 1.f, -my, 0.f, 0.f, 1.f};
      float T0[9];
      mat3_mul_rn(Tinv, guess, T0);
#pragma unroll
      for (int i = 0; i < 9; ++i) sh.T0[i] = T0[i], sh.Ti[i] = sh.Tprev[i] = (i % 4 == 0) ? 1.f : 0.f;
      rot_to_quat(sh.Ti, sh.hq_w[0], sh.hq_z[0]);
      sh.ht_x[0] = 0.f, sh.ht_y[0] = 0.f;
      sh.hn = 1;
      sh.iterate = 1;
      sh.sel_val = 0;  // (also the one-pass path's list counter, see 3a)
    }
    __syncthreads();
    for (int i = tid; i < ns; i += nthr) {
      reading[i] = apply_T(sh.T0, src[2 * i], src[2 * i + 1]);
      prev[i] = 0xffff;
      slack[i] = 0.f;
    }
    __syncthreads();
    if (b.use_order) {  // counting sort of the point indices by a 16 x 16 bucket grid over the target's bounding box.  Counters: the
       // radix-select histogram that the next selection pass neither fills nor expects to be clear (pass p fills
       // hist[p % 3] and clears hist[(p + 1) % 3]; hist[(p + 2) % 3] is cleared by pass p + 1 before it is used)
      int *cnt = sh.hist[(sel_pass + 2) % 3];
      for (int h = tid; h < 256; h += nthr) cnt[h] = 0;
      __syncthreads();
      const float bw = fmaxf(sh.bbox[2] - sh.bbox[0], 1e-6f) * (1.f / 16.f), bh = fmaxf(sh.bbox[3] - sh.bbox[1], 1e-6f) * (1.f / 16.f);
      auto bucket = [&](int i) {
        const int bx = min(max((int)((reading[i].x - sh.bbox[0]) / bw), 0), 15);
        const int by = min(max((int)((reading[i].y - sh.bbox[1]) / bh), 0), 15);
        return by * 16 + bx;
      };
      for (int i = tid; i < ns; i += nthr) atomicAdd(&cnt[bucket(i)], 1);
      __syncthreads();
      if (tid < 32) {  // exclusive scan of the 256 counters by one warp (8 per lane)
        int c[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = cnt[tid * 8 + j], sum += c[j];
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, d);
          if (tid >= d) incl += t;
        }
        int run = incl - sum;
#pragma unroll
        for (int j = 0; j < 8; ++j) cnt[tid * 8 + j] = run, run += c[j];
      }
      __syncthreads();
      for (int i = tid; i < ns; i += nthr) order[atomicAdd(&cnt[bucket(i)], 1)] = (uint16_t)i;
      __syncthreads();
    }

    // ---- 3. iterations
    while (true) {
      float Ti[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) Ti[i] = sh.Ti[i];
      // 3a. match.  Exact nearest neighbours are only needed for the pairs that can get a non-zero weight:
      //   pass A  every source point scans its 3x3 block of cells: that either settles its nearest neighbour
      //           or at least yields a candidate, i.e. an UPPER bound on its NN distance;
      //   pass B  points without a candidate inside maxDist learn whether ANY target point lies within
      //           maxDist (they then count as "finite" in the trimmed quantile) -- from the coarse
      //           occupancy grid when that is conclusive, else from a search that stops at the first hit;
      //   pass C  with n_finite known, kk = floor(n_finite * ratio).  The kk-th smallest of the upper bounds
      //           (maxDist^2 for a finite point without candidate) bounds the kk-th smallest true distance
      //           from above: call it U.  An unsettled point whose search proves "farther than U" has weight
      //           0 and needs no exact distance.  Results are identical to an exhaustive search (same
      //           quantile element, same kept pairs); the far outliers of a scan stop costing O(area).
      int n_fin = 0;
      const bool small = ns <= b.small_mult * nthr && nt <= 4096;  // few, cheap searches: one exact pass, no pruning
      // Certified matches.  A search that scans at least the 3x3 block also yields a lower bound L on the distance
      // from the point to every target point OTHER than its match (grid.cuh: nn_query_certified).  The point then
      // moves a little every iteration; by the triangle inequality every other target point stays farther than
      // L - (path length since).  slack[i] holds that margin, shrunk by safety terms far above float32 rounding:
      // while the distance to the old match is below it, the old match is the unique nearest neighbour -- the
      // answer an exhaustive search would give, ties impossible -- and costs one distance instead of a search.
      // Used for the small problems of the front end (window submaps thinned to 0.5 m: runner-ups are decimetres
      // away and a certificate lasts several iterations; ICP stage 2.53 -> 2.36 ms per 4096 frames).
      float Tp[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) Tp[i] = sh.Tprev[i];
      auto certified = [&](int i, const float2 q, int seed, float &d2_out, float &step) -> bool {
        const float2 qo = apply_T(Tp, reading[i].x, reading[i].y);
        step = sqrtf(dist2_rn(q.x - qo.x, q.y - qo.y));
        const float sl = slack[i] - (step * 1.00001f + 2e-5f);
        slack[i] = sl;
        const float2 t = sorted[seed];
        const float d2 = dist2_rn(q.x - t.x, q.y - t.y);
        d2_out = d2;
        return d2 <= max_d2 && sqrtf(d2) * 1.00001f + 1e-4f < sl;
      };
      auto certify = [&](int i, float lb2) {  // margin after a certifying search
        slack[i] = lb2 < 3.0e38f ? sqrtf(lb2) * 0.99999f - 1e-4f : 3.0e38f;
      };
      // certificate size: a few of the point's steps (steps shrink as the scan converges), a fraction of a cell at most
      auto margin_for = [&](float step) { return fminf(fmaxf(b.margin_mult * step, 0.01f * g.cell), 0.35f * g.cell); };
      if (small) {
        if constexpr (THREADS <= 256) {
          // Two halves, so that the lanes of a warp stay on one code path.  First every point tests its certificate
          // (one distance; points without a seed search at once: that is the first iteration, all lanes together);
          // the points whose certificate failed are only LISTED -- in match[], which is rewritten below anyway, with
          // the counter in sel_val (zeroed by the single-thread step of the previous iteration).  Then the list is
          // searched densely.  In one loop the re-search ran with 2.4 of 32 lanes active and took 25 % of this
          // kernel's warp instructions on the front end's problems (tools/ncu_call_sites.py).
          uint16_t *todo = match;
          for (int base = 0; base < ns; base += nthr) {
            const int ii = base + tid;
            bool later = false;
            int i = 0;
            if (ii < ns) {
              i = b.use_order ? order[ii] : ii;
              const float2 q = apply_T(Ti, reading[i].x, reading[i].y);
              const int seed = prev[i];
              float d2s, step, lb2;
              if (seed == 0xffff) {
                const NNResult r = nn_query_certified(g, q.x, q.y, max_d2, &lb2);
                certify(i, lb2);
                dist[i] = r.d2;
                prev[i] = r.pos >= 0 ? (uint16_t)r.pos : (uint16_t)0xffff;
                n_fin += r.pos >= 0;
              } else if (certified(i, q, seed, d2s, step)) {
                dist[i] = d2s;  // prev[i] stays
                n_fin += 1;
              } else {
                later = true;
              }
            }
            const unsigned m = __ballot_sync(0xffffffffu, later);
            if (m) {  // one shared-memory atomic per warp
              const int lane = tid & 31, lead = __ffs(m) - 1;
              int at = 0;
              if (lane == lead) at = (int)atomicAdd(&sh.sel_val, (uint32_t)__popc(m));
              at = __shfl_sync(0xffffffffu, at, lead);
              if (later) todo[at + __popc(m & ((1u << lane) - 1u))] = (uint16_t)i;
            }
          }
          __syncthreads();
          const int n_todo = (int)sh.sel_val;
          for (int k = tid; k < n_todo; k += nthr) {
            const int i = todo[k];
            const float2 q = apply_T(Ti, reading[i].x, reading[i].y);
            const float2 qo = apply_T(Tp, reading[i].x, reading[i].y);
            const float step = sqrtf(dist2_rn(q.x - qo.x, q.y - qo.y));
            float lb2;
            const NNResult r = nn_query_seeded_certified(g, q.x, q.y, max_d2, prev[i], margin_for(step), &lb2);
            certify(i, lb2);
            dist[i] = r.d2;
            prev[i] = r.pos >= 0 ? (uint16_t)r.pos : (uint16_t)0xffff;
            n_fin += r.pos >= 0;
          }
          __syncthreads();
          for (int i = tid; i < ns; i += nthr) match[i] = prev[i];
        } else {
          // (512-thread CTAs rarely see a problem this small: the single loop is kept there, and with it the code of the
          //  config-3 instantiation)
          for (int ii = tid; ii < ns; ii += nthr) {
            const int i = b.use_order ? order[ii] : ii;
            const float2 q = apply_T(Ti, reading[i].x, reading[i].y);
            const int seed = prev[i];
            NNResult r;
            float d2s, step, lb2;
            if (seed == 0xffff) {
              r = nn_query_certified(g, q.x, q.y, max_d2, &lb2);
              certify(i, lb2);
            } else if (certified(i, q, seed, d2s, step)) {
              r.d2 = d2s, r.pos = seed, r.tie = 0;
            } else {
              r = nn_query_seeded_certified(g, q.x, q.y, max_d2, seed, margin_for(step), &lb2);
              certify(i, lb2);
            }
            dist[i] = r.d2;
            match[i] = prev[i] = r.pos >= 0 ? (uint16_t)r.pos : (uint16_t)0xffff;
            n_fin += r.pos >= 0;
          }
        }
      }
      const float stop_a = (0.999f * g.cell) * (0.999f * g.cell);
      for (int ii = tid; ii < ns && !small; ii += nthr) {
        const int i = b.use_order ? order[ii] : ii;
        const float2 q = apply_T(Ti, reading[i].x, reading[i].y);
        NNResult r;
        r.d2 = INFINITY, r.pos = -1, r.tie = 0;
        int exact;
        const int seed = prev[i];
        // (no certificates here: a 20 000-point target puts the runner-up a centimetre from the match, so a
        // certificate rarely outlives an iteration -- measured 2.32 -> 2.59 ms per wave with them)
        const float2 ts = sorted[seed != 0xffff ? seed : 0];
        if (seed != 0xffff && dist2_rn(q.x - ts.x, q.y - ts.y) <= stop_a) {
          r = nn_query_seeded(g, q.x, q.y, max_d2, seed);  // last iteration's match is still close: settle it now
          exact = 1;
        } else {
          exact = nn_search(g, q.x, q.y, max_d2, stop_a, -1, r);  // own cell, then at most the 3x3 block
        }
        if (exact) prev[i] = (r.pos >= 0 && r.d2 <= max_d2) ? (uint16_t)r.pos : (uint16_t)0xffff;
        const bool fin = r.pos >= 0 && r.d2 <= max_d2;
        dist[i] = fin ? r.d2 : INFINITY;
        match[i] = fin ? (uint16_t)r.pos : (uint16_t)0xffff;
        qstate[i] = (exact ? 1 : 0) | (fin ? 2 : 0) | ((fin || exact) ? 4 : 0);  // settled / finite / known
        n_fin += fin;
      }
      // pass B: finiteness of the points that have no candidate yet
      for (int ii = tid; ii < ns && !small; ii += nthr) {
        const int i = b.use_order ? order[ii] : ii;
        if (qstate[i] & 4) continue;
        const float2 q = apply_T(Ti, reading[i].x, reading[i].y);
        int verdict = -1;  // 1 finite, 0 not, -1 unknown
        if (sh.cnx > 0) {
          const float C = g.cell * (float)ICP_COARSE, rad = prm.matcher_max_dist;
          const int xa = max(0, (int)floorf((q.x - rad - g.ox) / C) - 1), xb = min(sh.cnx - 1, (int)floorf((q.x + rad - g.ox) / C) + 1);
          const int ya = max(0, (int)floorf((q.y - rad - g.oy) / C) - 1), yb = min(sh.cny - 1, (int)floorf((q.y + rad - g.oy) / C) + 1);
          if (xa > xb || ya > yb) {
            verdict = 0;
          } else if ((xb - xa + 1) * (yb - ya + 1) <= 256) {
            const float r_in = rad * (1.f - 1e-4f) - g.cell * 1e-3f, r_out = rad * (1.f + 1e-4f) + g.cell * 1e-3f;
            bool any_maybe = false;
            for (int cy2 = ya; cy2 <= yb && verdict != 1; ++cy2)
              for (int cx2 = xa; cx2 <= xb; ++cx2) {
                const int cc = cy2 * sh.cnx + cx2;
                if (!((sh.coarse[cc >> 5] >> (cc & 31)) & 1u)) continue;
                const float x0 = g.ox + (float)cx2 * C, x1 = x0 + C, y0 = g.oy + (float)cy2 * C, y1 = y0 + C;
                const float nx_ = fmaxf(fmaxf(x0 - q.x, q.x - x1), 0.f), ny_ = fmaxf(fmaxf(y0 - q.y, q.y - y1), 0.f);
                const float fx_ = fmaxf(fabsf(x0 - q.x), fabsf(x1 - q.x)), fy_ = fmaxf(fabsf(y0 - q.y), fabsf(y1 - q.y));
                if (fx_ * fx_ + fy_ * fy_ <= r_in * r_in) {
                  verdict = 1;  // an occupied block lies entirely inside the acceptance disc
                  break;
                }
                if (nx_ * nx_ + ny_ * ny_ <= r_out * r_out) any_maybe = true;
              }
            if (verdict != 1 && !any_maybe) verdict = 0;
          }
        }
        if (verdict < 0) {  // inconclusive: grow the block until the first point inside maxDist shows up
          NNResult r;
          r.d2 = INFINITY, r.pos = -1, r.tie = 0;
          const int cx = grid_cell_coord(q.x, g.ox, g.inv_cell, g.nx), cy = grid_cell_coord(q.y, g.oy, g.inv_cell, g.ny);
          const int kmax = max(g.nx, g.ny);
          verdict = 0;
          for (int k = 2;; k = min(2 * k, kmax)) {
            nn_scan_block(g, cx, cy, k, q.x, q.y, r);
            if (r.d2 <= max_d2) {
              verdict = 1;
              break;
            }
            const float b2 = nn_block_bound2(g, q.x, q.y, cx, cy, k);
            if (b2 == INFINITY || b2 > max_d2 || k >= kmax) break;
          }
        }
        qstate[i] |= 4 | (verdict ? 2 : 0) | (verdict ? 0 : 1);  // "nothing within maxDist" is a settled answer
        n_fin += verdict;
        dist[i] = verdict ? max_d2 : INFINITY;  // upper bound of a finite point without candidate
      }
      const int total_fin = block_total(n_fin, &sh.tot[0][0], tot_phase);  // (its barrier also publishes dist[] / qstate[])
      // pass C: settle what still matters
      if (!small) {
        float stop_d2 = INFINITY;
        if (prm.trim_ratio >= 0.f) {
          if (prm.trim_ratio < 1.0f && total_fin > 0) {
            int kk = (int)(size_t)__fmul_rn((float)total_fin, prm.trim_ratio);
            if (kk >= total_fin) kk = total_fin - 1;
            stop_d2 = block_select_kth(dist, ns, kk, sh, sel_pass);  // kk-th smallest upper bound >= kk-th smallest distance
          }
        } else if (prm.outlier_max_dist > 0.f) {
          stop_d2 = out_d2;
        }
        // the unsettled points are few (the scan's outliers) and their searches long: every warp takes the
        // unsettled points of its 32 lanes one after the other and searches for each with all 32 lanes
        for (int i0 = tid & ~31; i0 < ns; i0 += nthr) {
          const int i = i0 + (tid & 31) < ns ? (b.use_order ? (int)order[i0 + (tid & 31)] : i0 + (tid & 31)) : ns;
          const bool mine = i < ns && !(qstate[i] & 1);
          float2 q = make_float2(0.f, 0.f);
          int cand = -1;
          float cand_d2 = INFINITY;
          if (mine) {
            q = apply_T(Ti, reading[i].x, reading[i].y);
            cand = match[i] == 0xffff ? -1 : (int)match[i];
            cand_d2 = cand >= 0 ? dist[i] : INFINITY;
          }
          unsigned todo = __ballot_sync(0xffffffffu, mine);
          while (todo) {
            const int L = __ffs(todo) - 1;
            todo &= todo - 1;
            const float qx = __shfl_sync(0xffffffffu, q.x, L), qy = __shfl_sync(0xffffffffu, q.y, L);
            NNResult r;
            r.pos = __shfl_sync(0xffffffffu, cand, L);
            r.d2 = __shfl_sync(0xffffffffu, cand_d2, L);
            r.tie = 1;  // pass A did not resolve ties; a block scan below replaces this flag with what it finds
            const int exact = nn_search_warp(g, qx, qy, max_d2, stop_d2, 1, r);
            if ((tid & 31) == L) {
              if (exact) {
                const bool fin = r.pos >= 0 && r.d2 <= max_d2;  // == the finiteness found in pass B
                dist[i] = fin ? r.d2 : INFINITY;
                match[i] = prev[i] = fin ? (uint16_t)r.pos : (uint16_t)0xffff;
              } else {
                dist[i] = ICP_PRUNED;  // finite, farther than the quantile bound: weight 0
                match[i] = 0xffff;
              }
            }
          }
        }
      }
      __syncthreads();

      // 3b. trimmed-distance limit = element floor(float(n)*ratio) of the ascending finite distances
      float limit = INFINITY;
      if (prm.trim_ratio >= 0.f) {
        if (total_fin == 0) {
          if (tid == 0) sh.status = ICP_NO_OUTLIER;
          __syncthreads();
          break;
        }
        if (prm.trim_ratio == 1.0f) {
          float m = 0.f;
          for (int i = tid; i < ns; i += nthr)
            if (dist[i] < INFINITY) m = fmaxf(m, dist[i]);
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
          float *fr = sh.wred;
          if ((tid & 31) == 0) fr[tid >> 5] = m;
          __syncthreads();
          m = 0.f;
          for (int w = 0; w < (nthr >> 5); ++w) m = fmaxf(m, fr[w]);
          limit = m;
          __syncthreads();
        } else {
          int kk = (int)(size_t)__fmul_rn((float)total_fin, prm.trim_ratio);
          if (kk >= total_fin) kk = total_fin - 1;
          limit = block_select_kth(dist, ns, kk, sh, sel_pass);
        }
      }

      // 3c. kept pairs: count and sums for the means
      int n_keep;
      float mrx = 0.f, mry = 0.f, mfx = 0.f, mfy = 0.f;
      if (seq) {
        int cnt = 0;
        for (int i = tid; i < ns; i += nthr) {
          bool keep = match[i] != 0xffff;
          if (prm.outlier_max_dist > 0.f) keep = keep && (dist[i] <= out_d2);
          if (prm.trim_ratio >= 0.f) keep = keep && (dist[i] <= limit);
          if (!keep) match[i] = 0xffff;
          cnt += keep;
        }
        n_keep = block_total(cnt, &sh.tot[0][0], tot_phase);  // (its barrier publishes match[])
      } else {
        double s5[5] = {0, 0, 0, 0, 0}, t5[5];
        for (int i = tid; i < ns; i += nthr) {
          bool keep = match[i] != 0xffff;
          if (prm.outlier_max_dist > 0.f) keep = keep && (dist[i] <= out_d2);
          if (prm.trim_ratio >= 0.f) keep = keep && (dist[i] <= limit);
          if (!keep) {
            match[i] = 0xffff;
            continue;
          }
          const float2 q = apply_T(Ti, reading[i].x, reading[i].y);
          const float2 r = sorted[match[i]];
          s5[0] += 1.0, s5[1] += (double)q.x, s5[2] += (double)q.y, s5[3] += (double)r.x, s5[4] += (double)r.y;
        }
        block_sum<5>(s5, sh.red, red_phase, t5);
        n_keep = (int)t5[0];
        mrx = (float)t5[1], mry = (float)t5[2], mfx = (float)t5[3], mfy = (float)t5[4];
      }
      if (n_keep == 0) {
        if (tid == 0) sh.status = ICP_NO_POINT;
        __syncthreads();
        break;
      }
      double t4[4] = {0, 0, 0, 0};
      float pl[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // PLANE: A00 A01 A02 A11 A12 A22, sum F (q-r).n (thread 0)
      if constexpr (PLANE) {
        // 3c'/3d'. point-to-plane normal equations over the kept pairs (oracle/icp_ref.c, minimizer 1):
        //   c = q.x n.y - q.y n.x,  F = (c, n.x, n.y),  A = sum F F^T,  b = -sum F ((q - r).n)
        auto terms = [&](int i, float (&v)[9]) {
          const int m = match[i];
          const float2 q = apply_T(Ti, reading[i].x, reading[i].y);
          const float2 r = sorted[m];
          const float2 n = nrm[m];
          const float cr = __fsub_rn(__fmul_rn(q.x, n.y), __fmul_rn(q.y, n.x));
          const float dp = __fadd_rn(__fmul_rn(__fsub_rn(q.x, r.x), n.x), __fmul_rn(__fsub_rn(q.y, r.y), n.y));
          v[0] = __fmul_rn(cr, cr), v[1] = __fmul_rn(cr, n.x), v[2] = __fmul_rn(cr, n.y), v[3] = __fmul_rn(n.x, n.x);
          v[4] = __fmul_rn(n.x, n.y), v[5] = __fmul_rn(n.y, n.y), v[6] = __fmul_rn(cr, dp), v[7] = __fmul_rn(n.x, dp);
          v[8] = __fmul_rn(n.y, dp);
        };
        if (seq) {
          if (tid < 32) {  // nine sequential sums in one pass; dropped pairs add +0
            const float sum = seq_sum9_warp(ns, pl_scratch, match, nrm, [&](int i, int m, float2 n, float (&v)[9]) {
              const float2 q = apply_T(Ti, reading[i].x, reading[i].y);
              const float2 r = sorted[m];
              const float cr = __fsub_rn(__fmul_rn(q.x, n.y), __fmul_rn(q.y, n.x));
              const float dp = __fadd_rn(__fmul_rn(__fsub_rn(q.x, r.x), n.x), __fmul_rn(__fsub_rn(q.y, r.y), n.y));
              v[0] = __fmul_rn(cr, cr), v[1] = __fmul_rn(cr, n.x), v[2] = __fmul_rn(cr, n.y), v[3] = __fmul_rn(n.x, n.x);
              v[4] = __fmul_rn(n.x, n.y), v[5] = __fmul_rn(n.y, n.y), v[6] = __fmul_rn(cr, dp), v[7] = __fmul_rn(n.x, dp);
              v[8] = __fmul_rn(n.y, dp);
            });
            __syncwarp();
            if (tid < 9) sh.seq_buf[tid] = sum;
          }
          __syncthreads();
#pragma unroll
          for (int k = 0; k < 9; ++k) pl[k] = sh.seq_buf[k];
          __syncthreads();  // seq_buf is scratch again next iteration
        } else {
          double s5[5] = {0, 0, 0, 0, 0}, s4[4] = {0, 0, 0, 0}, o5[5], o4[4];
          for (int i = tid; i < ns; i += nthr) {
            if (match[i] == 0xffff) continue;
            float v[9];
            terms(i, v);
#pragma unroll
            for (int k = 0; k < 5; ++k) s5[k] += (double)v[k];
#pragma unroll
            for (int k = 0; k < 4; ++k) s4[k] += (double)v[5 + k];
          }
          block_sum<5>(s5, sh.red, red_phase, o5);
          block_sum<4>(s4, sh.red, red_phase, o4);
#pragma unroll
          for (int k = 0; k < 5; ++k) pl[k] = (float)o5[k];
#pragma unroll
          for (int k = 0; k < 4; ++k) pl[5 + k] = (float)o4[k];
        }
      } else {
        if (seq) {
          // warp 0 forms the four sums (step x, step y, matched reference x, y); dropped pairs add +0
          if (tid < 32) {
            const float sum = seq_sum4_warp<(THREADS >= SFE_SEQ_PIPE_MIN)>(ns, sh.seq_buf, [&](int i) -> float4 {
              const int m = match[i];
              if (m == 0xffff) return make_float4(0.f, 0.f, 0.f, 0.f);
              const float2 q = apply_T(Ti, reading[i].x, reading[i].y);
              const float2 r = sorted[m];
              return make_float4(q.x, q.y, r.x, r.y);
            });
            if (tid < 4) sh.seq[tid] = sum;
          }
          __syncthreads();
          mrx = sh.seq[0], mry = sh.seq[1], mfx = sh.seq[2], mfy = sh.seq[3];
          __syncthreads();  // sh.seq is rewritten below
        }
        const float winv = __fdiv_rn(1.0f, (float)n_keep);
        mrx = __fmul_rn(mrx, winv), mry = __fmul_rn(mry, winv);
        mfx = __fmul_rn(mfx, winv), mfy = __fmul_rn(mfy, winv);

        // 3d. cross-covariance of the centred pairs
        if (seq) {
          if (tid < 32) {  // m00 = qx*px, m01 = qx*py, m10 = qy*px, m11 = qy*py
            const float sum = seq_sum4_warp<(THREADS >= SFE_SEQ_PIPE_MIN)>(ns, sh.seq_buf, [&](int i) -> float4 {
              const int m = match[i];
              if (m == 0xffff) return make_float4(0.f, 0.f, 0.f, 0.f);
              const float2 q = apply_T(Ti, reading[i].x, reading[i].y);
              const float2 r = sorted[m];
              const float px = __fsub_rn(q.x, mrx), py = __fsub_rn(q.y, mry);
              const float qx = __fsub_rn(r.x, mfx), qy = __fsub_rn(r.y, mfy);
              return make_float4(__fmul_rn(qx, px), __fmul_rn(qx, py), __fmul_rn(qy, px), __fmul_rn(qy, py));
            });
            if (tid < 4) sh.seq[tid] = sum;
          }
          __syncthreads();
          t4[0] = (double)sh.seq[0], t4[1] = (double)sh.seq[1], t4[2] = (double)sh.seq[2], t4[3] = (double)sh.seq[3];
        } else {
          double s4[4] = {0, 0, 0, 0};
          for (int i = tid; i < ns; i += nthr) {
            if (match[i] == 0xffff) continue;
            const float2 q = apply_T(Ti, reading[i].x, reading[i].y);
            const float2 r = sorted[match[i]];
            const float px = __fsub_rn(q.x, mrx), py = __fsub_rn(q.y, mry);
            const float qx = __fsub_rn(r.x, mfx), qy = __fsub_rn(r.y, mfy);
            s4[0] += (double)__fmul_rn(qx, px), s4[1] += (double)__fmul_rn(qx, py);
            s4[2] += (double)__fmul_rn(qy, px), s4[3] += (double)__fmul_rn(qy, py);
          }
          block_sum<4>(s4, sh.red, red_phase, t4);
        }
      }

      // 3e. rigid fit, T_iter update, checkers (one thread)
      if (tid == 0) {
        float c = 1.f, s = 0.f, tx, ty;
        if constexpr (PLANE) {
          // A x = b by Cholesky (A.llt()); an unknown whose pivot is not positive stays 0
          const float a00 = pl[0], a01 = pl[1], a02 = pl[2], a11 = pl[3], a12 = pl[4], a22 = pl[5];
          const float b0 = -pl[6], b1 = -pl[7], b2 = -pl[8];
          float l00 = 0.f, l10 = 0.f, l20 = 0.f, l11 = 0.f, l21 = 0.f, l22 = 0.f;
          const bool p0 = a00 > 0.f;
          if (p0) l00 = __fsqrt_rn(a00), l10 = __fdiv_rn(a01, l00), l20 = __fdiv_rn(a02, l00);
          const float d1 = __fsub_rn(a11, __fmul_rn(l10, l10));
          const bool p1 = d1 > 0.f;
          if (p1) l11 = __fsqrt_rn(d1), l21 = __fdiv_rn(__fsub_rn(a12, __fmul_rn(l20, l10)), l11);
          const float d2 = __fsub_rn(__fsub_rn(a22, __fmul_rn(l20, l20)), __fmul_rn(l21, l21));
          const bool p2 = d2 > 0.f;
          if (p2) l22 = __fsqrt_rn(d2);
          const float y0 = p0 ? __fdiv_rn(b0, l00) : 0.f;
          const float y1 = p1 ? __fdiv_rn(__fsub_rn(b1, __fmul_rn(l10, y0)), l11) : 0.f;
          const float y2 = p2 ? __fdiv_rn(__fsub_rn(__fsub_rn(b2, __fmul_rn(l20, y0)), __fmul_rn(l21, y1)), l22) : 0.f;
          const float x2 = p2 ? __fdiv_rn(y2, l22) : 0.f;
          const float x1 = p1 ? __fdiv_rn(__fsub_rn(y1, __fmul_rn(l21, x2)), l11) : 0.f;
          const float x0 = p0 ? __fdiv_rn(__fsub_rn(__fsub_rn(y0, __fmul_rn(l10, x1)), __fmul_rn(l20, x2)), l00) : 0.f;
          c = (float)cos((double)x0), s = (float)sin((double)x0);  // Rotation2D(x0)
          tx = x1, ty = x2;
        } else {
          const float m00 = (float)t4[0], m01 = (float)t4[1], m10 = (float)t4[2], m11 = (float)t4[3];
          const float a = __fadd_rn(m00, m11), bq = __fsub_rn(m10, m01);
          const float h = sqrtf(__fadd_rn(__fmul_rn(a, a), __fmul_rn(bq, bq)));
          if (h > 0.f) c = __fdiv_rn(a, h), s = __fdiv_rn(bq, h);
          tx = __fsub_rn(mfx, __fadd_rn(__fmul_rn(c, mrx), __fmul_rn(-s, mry)));
          ty = __fsub_rn(mfy, __fadd_rn(__fmul_rn(s, mrx), __fmul_rn(c, mry)));
        }
        const float dT[9] = {c, -s, tx, s, c, ty, 0.f, 0.f, 1.f};
        float Tn[9];
        mat3_mul_rn(dT, Ti, Tn);
#pragma unroll
        for (int i = 0; i < 9; ++i) sh.Tprev[i] = Ti[i], sh.Ti[i] = Tn[i];
        sh.sel_val = 0;  // the next iteration's list counter (the select that also uses it is done)
        sh.inliers = n_keep;
        const int count = ++sh.count;
        bool counter_stop = false;
        if (count >= prm.max_iterations) sh.iterate = 0, counter_stop = true;
        if (!counter_stop && smooth > 0) {
          int hn = sh.hn;
          if (hn == ICP_HIST) {
            for (int i = 0; i + 1 < ICP_HIST; ++i)
              sh.hq_w[i] = sh.hq_w[i + 1], sh.hq_z[i] = sh.hq_z[i + 1], sh.ht_x[i] = sh.ht_x[i + 1],
              sh.ht_y[i] = sh.ht_y[i + 1];
            hn = ICP_HIST - 1;
          }
          rot_to_quat(Tn, sh.hq_w[hn], sh.hq_z[hn]);
          sh.ht_x[hn] = Tn[2], sh.ht_y[hn] = Tn[5];
          ++hn;
          sh.hn = hn;
          float vr = 0.f, vt = 0.f;
          if (hn > smooth) {
            for (int i = hn - 1; i >= hn - smooth; --i) {
              const float dw = __fadd_rn(__fmul_rn(sh.hq_w[i], sh.hq_w[i - 1]), __fmul_rn(sh.hq_z[i], sh.hq_z[i - 1]));
              const float dz = __fsub_rn(__fmul_rn(sh.hq_z[i], sh.hq_w[i - 1]), __fmul_rn(sh.hq_w[i], sh.hq_z[i - 1]));
              vr = __fadd_rn(vr, fabsf(__fmul_rn(2.0f, atan2f(fabsf(dz), fabsf(dw)))));
              const float ex = __fsub_rn(sh.ht_x[i], sh.ht_x[i - 1]), ey = __fsub_rn(sh.ht_y[i], sh.ht_y[i - 1]);
              vt = __fadd_rn(vt, fabsf(sqrtf(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)))));
            }
            vr = __fdiv_rn(vr, (float)smooth);
            vt = __fdiv_rn(vt, (float)smooth);
            if (vr < prm.min_diff_rot && vt < prm.min_diff_trans) sh.iterate = 0;
          }
          if (vr != vr)
            sh.status = ICP_NAN_ROT;
          else if (vt != vt)
            sh.status = ICP_NAN_TRANS;
        }
      }
      __syncthreads();
      if (sh.status != ICP_OK || !sh.iterate) break;
    }

    // ---- 4. back to the caller's frame
    if (tid == 0) {
      if (sh.status == ICP_OK) {
        const float Tm[9] = {1.f, 0.f, mx, 0.f, 1.f, my, 0.f, 0.f, 1.f};
        float tmp[9], out[9], Ti[9], T0[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) Ti[i] = sh.Ti[i], T0[i] = sh.T0[i];
        mat3_mul_rn(Tm, Ti, tmp);
        mat3_mul_rn(tmp, T0, out);
#pragma unroll
        for (int i = 0; i < 9; ++i) b.T_out[9 * (size_t)p + i] = out[i];
      } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) b.T_out[9 * (size_t)p + i] = guess[i];
      }
      b.iters[p] = sh.count;
      b.inliers[p] = sh.inliers;
      b.status[p] = sh.status;
    }
  }
}

// pcl.match: nearest reference point (within max_dist) of every query point; one CTA per (ref, query) pair
struct MatchBatch {
  const float *ref_pts;
  const int *ref_off;
  const float *in_pts;
  const int *in_off;
  int P, nt_max, max_cells;
  float max_dist;
  int32_t *ids;  // [total queries]
  float *dists;
  uint16_t *orig_ws;
};

__global__ void __launch_bounds__(ICP_THREADS) match_kernel(const MatchBatch b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int scan[36];
  __shared__ float bbox_w[4 * 32];
  __shared__ float bbox[4];
  float2 *sorted = reinterpret_cast<float2 *>(smem_raw);
  uint32_t *cells = reinterpret_cast<uint32_t *>(smem_raw + sizeof(float2) * (size_t)b.nt_max);
  const int tid = threadIdx.x, nthr = blockDim.x;
  uint16_t *orig = b.orig_ws + (size_t)blockIdx.x * b.nt_max;
  const float max_d2 = __fmul_rn(b.max_dist, b.max_dist);
  for (int p = blockIdx.x; p < b.P; p += gridDim.x) {
    const float *ref = b.ref_pts + 2 * (size_t)b.ref_off[p];
    const int nt = b.ref_off[p + 1] - b.ref_off[p];
    const float *in = b.in_pts + 2 * (size_t)b.in_off[p];
    const int ns = b.in_off[p + 1] - b.in_off[p];
    int32_t *ids = b.ids + b.in_off[p];
    float *dists = b.dists + b.in_off[p];
    __syncthreads();
    if (nt <= 0) {
      for (int i = tid; i < ns; i += nthr) ids[i] = -1, dists[i] = INFINITY;
      continue;
    }
    float mn_x = INFINITY, mn_y = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (int i = tid; i < nt; i += nthr) {
      const float x = ref[2 * i], y = ref[2 * i + 1];
      mn_x = fminf(mn_x, x), mxx = fmaxf(mxx, x), mn_y = fminf(mn_y, y), mxy = fmaxf(mxy, y);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mn_x = fminf(mn_x, __shfl_xor_sync(0xffffffffu, mn_x, d));
      mn_y = fminf(mn_y, __shfl_xor_sync(0xffffffffu, mn_y, d));
      mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, d));
      mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, d));
    }
    if ((tid & 31) == 0) {
      bbox_w[(tid >> 5) * 4 + 0] = mn_x, bbox_w[(tid >> 5) * 4 + 1] = mn_y;
      bbox_w[(tid >> 5) * 4 + 2] = mxx, bbox_w[(tid >> 5) * 4 + 3] = mxy;
    }
    __syncthreads();
    if (tid == 0) {
      float a = INFINITY, bb = INFINITY, c = -INFINITY, d = -INFINITY;
      for (int w = 0; w < (nthr >> 5); ++w) {
        a = fminf(a, bbox_w[w * 4 + 0]), bb = fminf(bb, bbox_w[w * 4 + 1]);
        c = fmaxf(c, bbox_w[w * 4 + 2]), d = fmaxf(d, bbox_w[w * 4 + 3]);
      }
      bbox[0] = a, bbox[1] = bb, bbox[2] = c, bbox[3] = d;
    }
    __syncthreads();
    GridView g;
    grid_geometry(nt, bbox[0], bbox[1], bbox[2], bbox[3], 0.05f, g, b.max_cells);
    grid_build(ref, 2, nt, 0.f, 0.f, g, sorted, cells, orig, scan);
    for (int i = tid; i < ns; i += nthr) {
      const NNResult r = nn_query(g, in[2 * i], in[2 * i + 1], max_d2);
      ids[i] = r.pos >= 0 ? (int32_t)orig[r.pos] : -1;
      dists[i] = r.d2;
    }
  }
}

// ---------------------------------------------------------------------------- host side
static int pick_max_cells(int nt_max) {
  int c = nt_max + nt_max / 2;  // the grid aims at one point per cell; the slack absorbs elongated bounding boxes
  if (c < 256) c = 256;
  if (c > GRID_MAX_CELLS) c = GRID_MAX_CELLS;
  return c;
}

int icp_run(sfe_ctx *ctx, const sfe_icp_params *prm, const float *src_pts, const int *src_off, const int *src_cnt,
            const float *tgt_pts, const int *tgt_off, const int *tgt_cnt, int min_points, const int *src_id,
            const int *tgt_id, int P, int ns_max, int nt_max, const float *guess, float *T_out, int *iters,
            int *inliers, int *status, const int *raw_cnt, int raw_cap, int class_mode, int class_ns, int class_nt,
            int force_threads) {
  SFE_REQUIRE(ctx && prm, "icp: null context or parameters");
  SFE_REQUIRE(P >= 0 && ns_max >= 0 && nt_max >= 0, "icp: negative sizes");
  if (P == 0) return SFE_OK;
  SFE_REQUIRE(src_pts && src_off && tgt_pts && tgt_off && guess && T_out && iters && inliers && status,
              "icp: null pointer");
  SFE_REQUIRE(nt_max <= 65535 && ns_max <= 65535, "icp: clouds of more than 65535 points are not supported (got %d, %d)",
              ns_max, nt_max);
  SFE_REQUIRE(prm->max_iterations >= 1, "icp: maxIterationCount must be >= 1");
  SFE_REQUIRE(prm->minimizer == 0 || prm->minimizer == 1, "icp: unknown error minimizer %d", prm->minimizer);
  SFE_REQUIRE(prm->minimizer == 0 || (prm->normals_knn >= 3 && prm->normals_knn <= GRID_KNN_MAX),
              "icp: SurfaceNormalDataPointsFilter knn %d is not supported (3..%d)", prm->normals_knn, GRID_KNN_MAX);
  SFE_REQUIRE(prm->smooth_length < ICP_HIST, "icp: smoothLength %d is not supported (at most %d)", prm->smooth_length,
              ICP_HIST - 1);
  IcpBatch b{};
  b.src_pts = src_pts, b.src_off = src_off, b.tgt_pts = tgt_pts, b.tgt_off = tgt_off;
  b.src_cnt = src_cnt, b.tgt_cnt = tgt_cnt, b.min_points = min_points;
  b.raw_cnt = raw_cnt, b.raw_cap = raw_cap;
  b.class_mode = class_mode, b.class_ns = class_ns, b.class_nt = class_nt;
  b.src_id = src_id, b.tgt_id = tgt_id, b.guess = guess;
  b.T_out = T_out, b.iters = iters, b.inliers = inliers, b.status = status;
  b.P = P, b.ns_max = ns_max > 0 ? ns_max : 1, b.nt_max = nt_max > 0 ? nt_max : 1;
  b.max_cells = pick_max_cells(b.nt_max);
  b.use_order = b.ns_max > 1024;  // (the front end's 640-point class is sized to 6 CTAs per SM to the byte)
  b.cell_scale = 1.0f;
  b.prm = *prm;
  {
    // One-pass exact path (per-lane searches with certified matches, no pruning passes) for problems of up to
    // 5 source points per thread: measured on the config-4 replay (sources of ~360 points, 128-thread CTAs), ICP
    // stage per 4096 frames 2.54 ms with a limit of 2 points per thread, 2.23 with 3, 1.82 with 5.
    static const int mult = [] {  // development switch: SFE_ICP_SMALL_MULT
      const char *e = getenv("SFE_ICP_SMALL_MULT");
      return e ? atoi(e) : 5;
    }();
    b.small_mult = mult;
    static const float mm = [] {  // development switch: SFE_ICP_MARGIN_MULT
      const char *e = getenv("SFE_ICP_MARGIN_MULT");
      return e ? (float)atof(e) : 6.f;
    }();
    b.margin_mult = mm;
  }
  auto smem_for = [&](int max_cells) {
    return ((sizeof(IcpShared) + 15) & ~size_t(15)) + sizeof(float2) * (size_t)b.nt_max +
           sizeof(uint32_t) * (size_t)((max_cells + 2) / 2 + 1) + 8 + sizeof(float2) * (size_t)b.ns_max +
           sizeof(float) * (size_t)b.ns_max + 2 * sizeof(uint16_t) * (size_t)b.ns_max + (size_t)b.ns_max + 4 +
           sizeof(float) * (size_t)b.ns_max + (b.use_order ? sizeof(uint16_t) * (size_t)b.ns_max : 0) + 16 +
           (prm->minimizer == 1 && b.ns_max < ICP_PLANE_SCRATCH ? 16 + ICP_PLANE_SCRATCH * sizeof(float) : 0);
  };
  // big problems: trade cell-table entries (coarser cells) for room before giving up
  while (smem_for(b.max_cells) > (size_t)ctx->max_smem_optin && b.max_cells > b.nt_max / 4 + 256)
    b.max_cells -= b.max_cells / 8;
  const size_t smem = smem_for(b.max_cells);
  if (smem > (size_t)ctx->max_smem_optin) {
    set_error("icp: source %d + target %d points need %zu B of shared memory per CTA (limit %d)", ns_max, nt_max, smem,
              ctx->max_smem_optin);
    return SFE_ERR_UNSUPPORTED;
  }
  // CTA size follows the source size (one NN query per thread and iteration is the sweet spot)
  // (measured on the config-4 replay: 256 threads beat 128 and 512 for ~360-point sources)
  const int threads = force_threads > 0 ? force_threads : (b.ns_max <= 640 ? 128 : (b.ns_max <= 1536 ? 256 : ICP_THREADS));
  SFE_REQUIRE(threads == 128 || threads == 256 || threads == 512, "icp: CTA size must be 128, 256 or 512 (got %d)", threads);
  // instantiation: (threads, register budget).  512-thread CTAs whose shared memory allows only one per SM get the
  // full 128 registers.
  const int variant = threads == 128 ? 0 : (threads == 256 ? 1 : (smem > 110 * 1024 ? 3 : 2));
  const bool plane = prm->minimizer == 1;
  const void *fn = plane ? (variant == 0   ? (const void *)icp_kernel<128, 6, true>
                            : variant == 1 ? (const void *)icp_kernel<256, 4, true>
                            : variant == 2 ? (const void *)icp_kernel<512, 2, true>
                                           : (const void *)icp_kernel<512, 1, true>)
                         : (variant == 0   ? (const void *)icp_kernel<128, 6, false>
                            : variant == 1 ? (const void *)icp_kernel<256, 4, false>
                            : variant == 2 ? (const void *)icp_kernel<512, 2, false>
                                           : (const void *)icp_kernel<512, 1, false>);
  // the attribute / occupancy queries are cached in the context per (variant, smem): the front end calls this per
  // copy chunk
  {  // the attribute belongs to (function, device), not to a context: only ever raise it, process-wide
    static std::mutex mu;
    static size_t attr[64][8] = {};
    std::lock_guard<std::mutex> lock(mu);
    size_t &cur = attr[ctx->device & 63][variant + (plane ? 4 : 0)];
    if (smem > cur) {
      SFE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cur = smem;
    }
  }
  int c_per_sm = 0;
  for (const auto &e : ctx->icp_occ)
    if (e.per_sm > 0 && e.smem == smem && e.threads == threads && e.variant == variant + (plane ? 4 : 0)) c_per_sm = e.per_sm;
  if (c_per_sm == 0) {
    int q = 1;
    SFE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, fn, threads, smem));
    c_per_sm = q < 1 ? 1 : q;
    ctx->icp_occ[ctx->icp_occ_next] = {smem, threads, c_per_sm, variant + (plane ? 4 : 0)};
    ctx->icp_occ_next = (ctx->icp_occ_next + 1) % 8;
  }
  const int per_sm = c_per_sm;
  int grid = ctx->sm_count * per_sm;
  if (grid > P) grid = P;
  int slots = grid;
  if (per_sm == 1 && P > grid) b.slot_by_smid = 1, slots = ICP_SM_SLOTS, grid = P;  // one CTA per problem
  const size_t orig_bytes = ((size_t)slots * b.nt_max * sizeof(uint16_t) + 15) & ~size_t(15);
  const size_t nrm_bytes = plane ? (size_t)slots * b.nt_max * sizeof(float2) : 0;
  int rc = ensure(ctx, ctx->scratch[SCR_ICP], orig_bytes + nrm_bytes);
  if (rc != SFE_OK) return rc;
  b.orig_ws = (uint16_t *)ctx->scratch[SCR_ICP].ptr;
  b.nrm_ws = plane ? (float2 *)((char *)ctx->scratch[SCR_ICP].ptr + orig_bytes) : nullptr;
  void *args[] = {(void *)&b};
  SFE_CUDA(cudaLaunchKernel(fn, dim3(grid), dim3(threads), args, smem, ctx->stream));
  ctx->launches++;
  return SFE_OK;
}

int match_run(sfe_ctx *ctx, const float *ref_pts, const int *ref_off, const float *in_pts, const int *in_off, int P,
              int nt_max, float max_dist, int32_t *ids, float *dists) {
  SFE_REQUIRE(ctx, "match: null context");
  SFE_REQUIRE(P >= 0 && nt_max >= 0, "match: negative sizes");
  if (P == 0) return SFE_OK;
  SFE_REQUIRE(ref_pts && ref_off && in_pts && in_off && ids && dists, "match: null pointer");
  SFE_REQUIRE(nt_max <= 65535, "match: reference clouds of more than 65535 points are not supported (got %d)", nt_max);
  MatchBatch b{};
  b.ref_pts = ref_pts, b.ref_off = ref_off, b.in_pts = in_pts, b.in_off = in_off;
  b.P = P, b.nt_max = nt_max > 0 ? nt_max : 1, b.max_cells = pick_max_cells(b.nt_max), b.max_dist = max_dist;
  b.ids = ids, b.dists = dists;
  const size_t smem = sizeof(float2) * (size_t)b.nt_max + sizeof(uint32_t) * (size_t)((b.max_cells + 2) / 2 + 1) + 16;
  if (smem > (size_t)ctx->max_smem_optin) {
    set_error("match: a %d-point reference needs %zu B of shared memory per CTA (limit %d)", nt_max, smem,
              ctx->max_smem_optin);
    return SFE_ERR_UNSUPPORTED;
  }
  SFE_CUDA(cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  SFE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, match_kernel, ICP_THREADS, smem));
  if (per_sm < 1) per_sm = 1;
  int grid = ctx->sm_count * per_sm;
  if (grid > P) grid = P;
  int rc = ensure(ctx, ctx->scratch[SCR_ICP], (size_t)grid * b.nt_max * sizeof(uint16_t));
  if (rc != SFE_OK) return rc;
  b.orig_ws = (uint16_t *)ctx->scratch[SCR_ICP].ptr;
  match_kernel<<<grid, ICP_THREADS, smem, ctx->stream>>>(b);
  SFE_CUDA(cudaGetLastError());
  ctx->launches++;
  return SFE_OK;
}

}  // namespace sfe
