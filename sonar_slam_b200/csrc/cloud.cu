// Point-cloud filters (sm_100a): voxel-medoid down-sampling and radius outlier removal.
//
// Replace bruce_slam.pcl.downsample and bruce_slam.pcl.remove_outlier
// (bruce_slam/src/bruce_slam/cpp/pcl.cpp:128-159 and :54-74), i.e. libpointmatcher's
// OctreeGridDataPointsFilter(maxSizeByNode = resolution, samplingMethod = 3 / medoid) and PCL's
// RadiusOutlierRemoval, as restated in oracle/cloud_ref.c.  One CTA per cloud, clouds packed
// [total][dim] with CSR offsets; results are written compacted at the cloud's own offset together
// with the indices of the surviving points (so descriptor columns can be gathered by the caller,
// which is how the two-argument downsample overload of pcl.cpp:143 is served).
//
// downsample: the quadtree is never materialised.  Splitting always halves the bounding square,
//   so the leaf a point falls into at the size-limited depth D is found by D comparisons against
//   centres computed with the same float32 additions the tree would use; the 2-bit child ids along
//   the path form a key whose ascending order is the tree's depth-first visiting order.  Grouping the
//   points by key with members in index order therefore reproduces the tree's leaves in output order
//   (nodes that the reference stops splitting early because they hold a single point produce the same
//   groups).  The grouping is a counting sort by the first levels of the key in shared memory followed by
//   a parallel rank inside each cell (taken by every cloud whose layout fits the launch's shared memory),
//   or an in-CTA bitonic sort of (key, index) pairs for very large clouds.  Each group then picks its
//   medoid (float32 sequential sum of Euclidean distances, first minimum wins).
// remove_outlier: counts neighbours within the radius on the shared-memory grid of grid.cuh.
#include "grid.cuh"

namespace sfe {

constexpr int CLOUD_THREADS = 512;
constexpr int DS_MAX_DEPTH = 15;
constexpr int DS_FAST_DEPTH = 7;                 // 4^7 = 16384 leaves: counting-sort path
constexpr int DS_FAST_CELLS = 1 << (2 * DS_FAST_DEPTH);
constexpr int DS_WIDE_PREFIX = 6;                // deeper trees: 4^6 cells of the first six levels, 32-bit keys
constexpr int DS_WIDE_CELLS = 1 << (2 * DS_WIDE_PREFIX);

struct CloudBatch {
  const float *pts;  // [total][dim]
  const int *off;    // [n_clouds + 1] first row of every cloud
  const int *cnt;    // optional [n_clouds]: rows used (else off[c+1]-off[c]); clamped to n_max
  int n_clouds, dim, n_max, n_pad, max_cells;
  int n_lo, n_hi;    // size class of this launch: only clouds with n_lo < n <= n_hi are processed (others untouched)
  int n_lay;         // the launch's shared-memory layout / workspace stride is sized for clouds of n_lay points
  float resolution;  // downsample
  double radius;     // remove_outlier
  int min_points;
  float *out_pts;    // [total][dim]
  int32_t *out_idx;  // [total] index (within its cloud) of every surviving point
  int32_t *out_count;  // [n_clouds]
  unsigned long long *sort_ws;  // global sort buffer when n_pad does not fit shared memory
  uint16_t *orig_ws;
  int sort_in_smem;
  int smem_bytes;  // dynamic shared memory of the launch (the counting-sort path is taken by every cloud that fits)
};

// sqrtf of a squared distance.  Every member meets itself once in its leaf's sums: sqrtf(0) leaves the inline fast
// path for the special-operand subroutine (call, test, return with one or two lanes active -- 8 % of this kernel's
// warp instructions and 13 % of its stall samples went there); zero is answered here instead.
__device__ __forceinline__ float sqrt_dist(float d2) {
  if (d2 == 0.f) return 0.f;
  return sqrtf(d2);
}

__global__ void __launch_bounds__(CLOUD_THREADS) downsample_kernel(const CloudBatch b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float red[4 * 32];
  __shared__ float geo[3];  // cx, cy, radius
  __shared__ int depth_s;
  __shared__ int scan[36];
  const int tid = threadIdx.x, nthr = blockDim.x;
  unsigned long long *keys = b.sort_in_smem ? reinterpret_cast<unsigned long long *>(smem_raw)
                                            : b.sort_ws + (size_t)blockIdx.x * b.n_pad;
  float *acc = reinterpret_cast<float *>(smem_raw + (b.sort_in_smem ? sizeof(unsigned long long) * (size_t)b.n_pad : 0));

  for (int cl = blockIdx.x; cl < b.n_clouds; cl += gridDim.x) {
    const int o = b.off[cl], n = min(b.cnt ? b.cnt[cl] : b.off[cl + 1] - o, b.n_max);
    if (!(n > b.n_lo && n <= b.n_hi)) continue;  // the other launch's size class (CTA-uniform)
    const float *pts = b.pts + (size_t)o * b.dim;
    __syncthreads();
    if (n == 0) {
      if (tid == 0) b.out_count[cl] = 0;
      continue;
    }
    // ---- bounding square (Octree::build): centre = min + (max-min)*0.5, radius = max extent * 0.5
    float mn_x = INFINITY, mn_y = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (int i = tid; i < n; i += nthr) {
      const float x = pts[(size_t)i * b.dim], y = pts[(size_t)i * b.dim + 1];
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
      red[(tid >> 5) * 4 + 0] = mn_x, red[(tid >> 5) * 4 + 1] = mn_y;
      red[(tid >> 5) * 4 + 2] = mxx, red[(tid >> 5) * 4 + 3] = mxy;
    }
    __syncthreads();
    if (tid == 0) {
      float a = INFINITY, bb = INFINITY, c = -INFINITY, d = -INFINITY;
      for (int w = 0; w < (nthr >> 5); ++w) {
        a = fminf(a, red[w * 4 + 0]), bb = fminf(bb, red[w * 4 + 1]);
        c = fmaxf(c, red[w * 4 + 2]), d = fmaxf(d, red[w * 4 + 3]);
      }
      const float rx = __fsub_rn(c, a), ry = __fsub_rn(d, bb);
      geo[0] = __fadd_rn(a, __fmul_rn(rx, 0.5f));
      geo[1] = __fadd_rn(bb, __fmul_rn(ry, 0.5f));
      float radius = rx;
      if (radius < ry) radius = ry;
      radius = __fmul_rn(radius, 0.5f);
      geo[2] = radius;
      int D = 0;
      float r = radius;
      while (!((double)r * 2.0 <= (double)b.resolution) && D < DS_MAX_DEPTH) r = __fmul_rn(r, 0.5f), ++D;
      depth_s = D;
    }
    __syncthreads();
    const int D = depth_s;
    const bool wide = D > DS_FAST_DEPTH;
    const size_t a16 = (sizeof(uint16_t) * (size_t)n + 15) & ~size_t(15);  // layout sized by THIS cloud
    const size_t need = (wide ? 2 * a16 : a16) + sizeof(uint32_t) * ((wide ? DS_WIDE_CELLS : DS_FAST_CELLS) / 2 + 4) +
                        a16 + sizeof(float) * (size_t)n + 16;
    if (n <= 65535 && need <= (size_t)b.smem_bytes) {
      // ---- counting-sort path.  The first Dp levels of the path key index a table of cells (Dp = D up to 7
      //      levels = 16384 leaves; deeper trees use 6 levels and finish inside the cell): counting sort by
      //      cell with atomics (members land in arbitrary order), then every member finds its rank inside its
      //      cell by (key, index) -- all members work in parallel -- which restores the reference's member order
      //      (a stable partition of the index list at every level) and groups the leaves of a cell.
      const int Dp = wide ? DS_WIDE_PREFIX : D;
      const int sub_bits = 2 * (D - Dp);
      uint16_t *key16 = reinterpret_cast<uint16_t *>(smem_raw);                        // [n_max] (D <= 7)
      uint32_t *key32 = reinterpret_cast<uint32_t *>(smem_raw);                        // [n_max] (D > 7)
      uint32_t *cellw = reinterpret_cast<uint32_t *>(smem_raw + (wide ? 2 * a16 : a16));
      uint16_t *sidx = reinterpret_cast<uint16_t *>(cellw + (wide ? DS_WIDE_CELLS : DS_FAST_CELLS) / 2 + 4);  // [n_max]
      float *facc = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(sidx) + a16);
      uint16_t *tmp = reinterpret_cast<uint16_t *>(facc);  // sorted member list before it replaces sidx
      const int ncells = 1 << (2 * Dp), nwords = (ncells + 2) / 2;
      for (int w = tid; w < nwords; w += nthr) cellw[w] = 0;
      __syncthreads();
      for (int i = tid; i < n; i += nthr) {
        const float x = pts[(size_t)i * b.dim], y = pts[(size_t)i * b.dim + 1];
        float cx = geo[0], cy = geo[1], r = geo[2];
        unsigned key = 0;
        for (int d = 0; d < D; ++d) {
          const unsigned id = (x > cx ? 1u : 0u) | (y > cy ? 2u : 0u);
          key = (key << 2) | id;
          r = __fmul_rn(r, 0.5f);
          cx = __fadd_rn(cx, (id & 1u) ? r : -r);
          cy = __fadd_rn(cy, (id & 2u) ? r : -r);
        }
        if (wide) key32[i] = key; else key16[i] = (uint16_t)key;
        const unsigned cell = key >> sub_bits;
        atomicAdd(&cellw[cell >> 1], (cell & 1u) ? 0x10000u : 1u);
      }
      __syncthreads();
      {  // inclusive ends per cell (two 16-bit counters per word)
        const int per = (nwords + nthr - 1) / nthr;
        const int w0 = min(tid * per, nwords), w1 = min(w0 + per, nwords);
        int local = 0;
        for (int w = w0; w < w1; ++w) local += (int)(cellw[w] & 0xffffu) + (int)(cellw[w] >> 16);
        int total;
        int run = block_exclusive_scan(local, scan, total);
        for (int w = w0; w < w1; ++w) {
          const uint32_t v = cellw[w];
          const int lo = run + (int)(v & 0xffffu), hi = lo + (int)(v >> 16);
          cellw[w] = (uint32_t)lo | ((uint32_t)hi << 16);
          run = hi;
        }
      }
      __syncthreads();
      for (int i = tid; i < n; i += nthr) {
        const unsigned cell = (wide ? key32[i] : (unsigned)key16[i]) >> sub_bits;
        const uint32_t old = atomicSub(&cellw[cell >> 1], (cell & 1u) ? 0x10000u : 1u);
        sidx[(int)((cell & 1u) ? (old >> 16) : (old & 0xffffu)) - 1] = (uint16_t)i;
      }
      __syncthreads();
      const uint16_t *cstart = reinterpret_cast<const uint16_t *>(cellw);  // [ncells + 1], entry ncells == n
      // rank of every member inside its cell by (key, index)
      for (int a = tid; a < n; a += nthr) {
        const int ia = sidx[a];
        const unsigned ka = wide ? key32[ia] : (unsigned)key16[ia];
        const unsigned cell = ka >> sub_bits;
        const int s0 = cstart[cell], e0 = cstart[cell + 1];
        int rank = 0;
        if (wide) {
          for (int q = s0; q < e0; ++q) {
            const int iq = sidx[q];
            const unsigned kq = key32[iq];
            rank += (kq < ka) || (kq == ka && iq < ia);
          }
        } else {
          for (int q = s0; q < e0; ++q) rank += (int)sidx[q] < ia;
        }
        tmp[s0 + rank] = (uint16_t)ia;
      }
      __syncthreads();
      for (int a = tid; a < n; a += nthr) sidx[a] = tmp[a];
      __syncthreads();
      // per member: float32 sum of distances to the members of its leaf, in member order
      for (int a = tid; a < n; a += nthr) {
        const int ia = sidx[a];
        const unsigned ka = wide ? key32[ia] : (unsigned)key16[ia];
        const unsigned cell = ka >> sub_bits;
        int s0 = cstart[cell], e0 = cstart[cell + 1];
        if (wide) {  // the leaf is the run of equal keys around a
          int s = a, e = a + 1;
          while (s > s0 && key32[sidx[s - 1]] == ka) --s;
          while (e < e0 && key32[sidx[e]] == ka) ++e;
          s0 = s, e0 = e;
        }
        // (dim == 2 on this path -- checked by the host: one 8-byte load per point)
        const float2 *p2 = reinterpret_cast<const float2 *>(pts);
        const float2 pa = p2[ia];
        const float ax = pa.x, ay = pa.y;
        float sum = 0.f;
        int q = s0;
        for (; q + 4 <= e0; q += 4) {  // four members per trip: loads and square roots overlap, the sum stays in order
          const float2 t0 = p2[sidx[q]], t1 = p2[sidx[q + 1]], t2 = p2[sidx[q + 2]], t3 = p2[sidx[q + 3]];
          const float d0 = sqrt_dist(dist2_rn(ax - t0.x, ay - t0.y));
          const float d1 = sqrt_dist(dist2_rn(ax - t1.x, ay - t1.y));
          const float d2 = sqrt_dist(dist2_rn(ax - t2.x, ay - t2.y));
          const float d3 = sqrt_dist(dist2_rn(ax - t3.x, ay - t3.y));
          sum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sum, d0), d1), d2), d3);
        }
        for (; q < e0; ++q) {
          const float2 t = p2[sidx[q]];
          sum = __fadd_rn(sum, sqrt_dist(dist2_rn(ax - t.x, ay - t.y)));
        }
        facc[a] = sum;
      }
      __syncthreads();
      // leaves in key order (= depth-first order).  Every thread owns a contiguous run of MEMBERS of the sorted list
      // (balanced: occupied leaves cluster in a small part of the 4^D cell table, so a split by cell range left a
      // handful of threads with all the work -- 25 % of this kernel's instructions at 3 active lanes); a member is
      // a leaf head when it is the first of its cell (D <= 7: one leaf per cell) or of its run of equal keys.
      {
        const int per = (n + nthr - 1) / nthr;
        const int a0 = min(tid * per, n), a1 = min(a0 + per, n);
        auto is_head = [&](int a) -> bool {
          if (wide) return a == 0 || key32[sidx[a]] != key32[sidx[a - 1]];
          return a == (int)cstart[key16[sidx[a]]];
        };
        int mine = 0;
        for (int a = a0; a < a1; ++a) mine += is_head(a) ? 1 : 0;
        int total;
        int rank = block_exclusive_scan(mine, scan, total);
        // the heads park their position in the output slot of their leaf ...
        for (int a = a0; a < a1; ++a)
          if (is_head(a)) b.out_idx[(size_t)o + rank++] = a;
        __syncthreads();
        // ... and the leaves are then finished densely, one per thread (a head per lane left 2-3 lanes of 32 busy in
        // the medoid search: 16 % of this kernel's instructions); slot r is read and rewritten by the same thread
        for (int r = tid; r < total; r += nthr) {
          const int a = b.out_idx[(size_t)o + r];
          int e0;
          if (wide) {
            const unsigned ka = key32[sidx[a]];
            e0 = a + 1;
            while (e0 < n && key32[sidx[e0]] == ka) ++e0;
          } else {
            e0 = cstart[(int)key16[sidx[a]] + 1];
          }
          float best = 3.402823466e+38f;
          int med = a;
          for (int m = a; m < e0; ++m)
            if (facc[m] < best) best = facc[m], med = m;
          const int idx = sidx[med];
          const size_t dst = (size_t)(o + r);
          for (int d = 0; d < b.dim; ++d) b.out_pts[dst * b.dim + d] = pts[(size_t)idx * b.dim + d];
          b.out_idx[dst] = idx;
        }
        if (tid == 0) b.out_count[cl] = total;
      }
      continue;
    }
    int n_pad = 2;  // bitonic sort size for THIS cloud
    while (n_pad < n) n_pad <<= 1;
    // ---- path key of every point, then sort (key, index)
    for (int i = tid; i < n_pad; i += nthr) {
      unsigned long long kv = ~0ull;
      if (i < n) {
        const float x = pts[(size_t)i * b.dim], y = pts[(size_t)i * b.dim + 1];
        float cx = geo[0], cy = geo[1], r = geo[2];
        unsigned key = 0;
        for (int d = 0; d < D; ++d) {
          const unsigned id = (x > cx ? 1u : 0u) | (y > cy ? 2u : 0u);
          key = (key << 2) | id;
          r = __fmul_rn(r, 0.5f);
          cx = __fadd_rn(cx, (id & 1u) ? r : -r);
          cy = __fadd_rn(cy, (id & 2u) ? r : -r);
        }
        kv = ((unsigned long long)key << 32) | (unsigned)i;
      }
      keys[i] = kv;
    }
    __syncthreads();
    for (int k = 2; k <= n_pad; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < n_pad; i += nthr) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const unsigned long long a = keys[i], c = keys[ixj];
            const bool up = (i & k) == 0;
            if ((a > c) == up) keys[i] = c, keys[ixj] = a;
          }
        }
        __syncthreads();
      }
    }
    // ---- per member: float32 sum of distances to the members of its leaf, in member order
    for (int a = tid; a < n; a += nthr) {
      const unsigned key = (unsigned)(keys[a] >> 32);
      int s = a, e = a + 1;
      while (s > 0 && (unsigned)(keys[s - 1] >> 32) == key) --s;
      while (e < n && (unsigned)(keys[e] >> 32) == key) ++e;
      const int ia = (int)(unsigned)keys[a];
      const float ax = pts[(size_t)ia * b.dim], ay = pts[(size_t)ia * b.dim + 1];
      float sum = 0.f;
      for (int q = s; q < e; ++q) {
        const int iq = (int)(unsigned)keys[q];
        sum = __fadd_rn(sum, sqrt_dist(dist2_rn(ax - pts[(size_t)iq * b.dim], ay - pts[(size_t)iq * b.dim + 1])));
      }
      acc[a] = sum;
    }
    __syncthreads();
    // ---- leaf heads pick the medoid; ordered compaction of the leaves
    int base = 0;
    for (int c0 = 0; c0 < n; c0 += nthr) {
      const int a = c0 + tid;
      bool head = false;
      int med = 0;
      if (a < n) {
        const unsigned key = (unsigned)(keys[a] >> 32);
        head = (a == 0) || ((unsigned)(keys[a - 1] >> 32) != key);
        if (head) {
          float best = 3.402823466e+38f;
          med = a;
          for (int q = a; q < n && (unsigned)(keys[q] >> 32) == key; ++q)
            if (acc[q] < best) best = acc[q], med = q;
        }
      }
      int total;
      const int rank = block_exclusive_scan(head ? 1 : 0, scan, total);
      if (head) {
        const int idx = (int)(unsigned)keys[med];
        const size_t dst = (size_t)(o + base + rank);
        for (int d = 0; d < b.dim; ++d) b.out_pts[dst * b.dim + d] = pts[(size_t)idx * b.dim + d];
        b.out_idx[dst] = idx;
      }
      base += total;
    }
    if (tid == 0) b.out_count[cl] = base;
  }
}

__global__ void __launch_bounds__(CLOUD_THREADS) remove_outlier_kernel(const CloudBatch b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float red[4 * 32];
  __shared__ float bbox[4];
  __shared__ int scan[36];
  float2 *sorted = reinterpret_cast<float2 *>(smem_raw);
  uint32_t *cells = reinterpret_cast<uint32_t *>(smem_raw + sizeof(float2) * (size_t)b.n_lay);
  const int tid = threadIdx.x, nthr = blockDim.x;
  uint16_t *orig = b.orig_ws + (size_t)blockIdx.x * b.n_lay;
  const double r2 = b.radius * b.radius;
  const float rw = (float)(b.radius * (1.0 + 1e-5) + 1e-6);  // search window, slightly widened

  for (int cl = blockIdx.x; cl < b.n_clouds; cl += gridDim.x) {
    const int o = b.off[cl], n = min(b.cnt ? b.cnt[cl] : b.off[cl + 1] - o, b.n_max);
    if (!(n > b.n_lo && n <= b.n_hi)) continue;  // the other launch's size class (CTA-uniform)
    const float *pts = b.pts + (size_t)o * b.dim;
    __syncthreads();
    if (n == 0) {
      if (tid == 0) b.out_count[cl] = 0;
      continue;
    }
    float mn_x = INFINITY, mn_y = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (int i = tid; i < n; i += nthr) {
      const float x = pts[(size_t)i * b.dim], y = pts[(size_t)i * b.dim + 1];
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
      red[(tid >> 5) * 4 + 0] = mn_x, red[(tid >> 5) * 4 + 1] = mn_y;
      red[(tid >> 5) * 4 + 2] = mxx, red[(tid >> 5) * 4 + 3] = mxy;
    }
    __syncthreads();
    if (tid == 0) {
      float a = INFINITY, bb = INFINITY, c = -INFINITY, d = -INFINITY;
      for (int w = 0; w < (nthr >> 5); ++w) {
        a = fminf(a, red[w * 4 + 0]), bb = fminf(bb, red[w * 4 + 1]);
        c = fmaxf(c, red[w * 4 + 2]), d = fmaxf(d, red[w * 4 + 3]);
      }
      bbox[0] = a, bbox[1] = bb, bbox[2] = c, bbox[3] = d;
    }
    __syncthreads();
    GridView g;
    grid_geometry(n, bbox[0], bbox[1], bbox[2], bbox[3], (float)(b.radius * 0.5), g, b.max_cells);
    grid_build(pts, b.dim, n, 0.f, 0.f, g, sorted, cells, orig, scan);

    int base = 0;
    for (int c0 = 0; c0 < n; c0 += nthr) {
      const int i = c0 + tid;
      bool keep = false;
      if (i < n) {
        const float qx = pts[(size_t)i * b.dim], qy = pts[(size_t)i * b.dim + 1];
        const float qz = b.dim == 3 ? pts[(size_t)i * 3 + 2] : 0.f;
        const int xa = grid_cell_coord(qx - rw, g.ox, g.inv_cell, g.nx), xb = grid_cell_coord(qx + rw, g.ox, g.inv_cell, g.nx);
        const int ya = grid_cell_coord(qy - rw, g.oy, g.inv_cell, g.ny), yb = grid_cell_coord(qy + rw, g.oy, g.inv_cell, g.ny);
        int count = 0;
        for (int y = ya; y <= yb; ++y) {
          const int s = g.cstart[y * g.nx + xa], e = g.cstart[y * g.nx + xb + 1];
          for (int q = s; q < e; ++q) {
            const float2 t = sorted[q];
            float d2 = __fmul_rn(qx - t.x, qx - t.x);
            d2 = __fadd_rn(d2, __fmul_rn(qy - t.y, qy - t.y));
            if (b.dim == 3) {
              const float dz = qz - pts[(size_t)orig[q] * 3 + 2];
              d2 = __fadd_rn(d2, __fmul_rn(dz, dz));
            }
            count += !(r2 < (double)d2);
          }
        }
        keep = count >= b.min_points + 1;
      }
      int total;
      const int rank = block_exclusive_scan(keep ? 1 : 0, scan, total);
      if (keep) {
        const size_t dst = (size_t)(o + base + rank);
        for (int d = 0; d < b.dim; ++d) b.out_pts[dst * b.dim + d] = pts[(size_t)i * b.dim + d];
        b.out_idx[dst] = i;
      }
      base += total;
    }
    if (tid == 0) b.out_count[cl] = base;
  }
}

// ---------------------------------------------------------------------------- host side
static int next_pow2(int v) {
  int p = 2;
  while (p < v) p <<= 1;
  return p;
}

// One launch of a size class: layout for clouds of n_lay points, clouds with n_lo < n <= n_hi.
static int downsample_launch(sfe_ctx *ctx, CloudBatch b, int n_lay, int n_lo, int n_hi) {
  b.n_lay = n_lay, b.n_lo = n_lo, b.n_hi = n_hi;
  b.n_pad = next_pow2(n_lay);
  size_t smem_sort = sizeof(unsigned long long) * (size_t)b.n_pad, smem_acc = sizeof(float) * (size_t)n_lay + 16;
  b.sort_in_smem = smem_sort + smem_acc <= (size_t)ctx->max_smem_optin - 4096;
  size_t smem = (b.sort_in_smem ? smem_sort : 0) + smem_acc;
  {
    const size_t a16 = (sizeof(uint16_t) * (size_t)n_lay + 15) & ~size_t(15);
    const size_t lay_a = a16 + sizeof(uint32_t) * (DS_FAST_CELLS / 2 + 4) + a16 + sizeof(float) * (size_t)n_lay + 16;
    const size_t lay_b = 2 * a16 + sizeof(uint32_t) * (DS_WIDE_CELLS / 2 + 4) + a16 + sizeof(float) * (size_t)n_lay + 16;
    const size_t fast = lay_a > lay_b ? lay_a : lay_b;
    if (n_lay <= 65535 && fast <= (size_t)ctx->max_smem_optin - 4096 && fast > smem) smem = fast;
  }
  if (smem > (size_t)ctx->max_smem_optin - 4096) {
    set_error("downsample: clouds of %d points are not supported (shared memory)", n_lay);
    return SFE_ERR_UNSUPPORTED;
  }
  b.smem_bytes = (int)smem;
  SFE_CUDA(cudaFuncSetAttribute(downsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  SFE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, downsample_kernel, CLOUD_THREADS, smem));
  if (per_sm < 1) per_sm = 1;
  int grid = ctx->sm_count * per_sm;
  if (grid > b.n_clouds) grid = b.n_clouds;
  if (!b.sort_in_smem) {
    int rc = ensure(ctx, ctx->scratch[SCR_CLOUD], sizeof(unsigned long long) * (size_t)grid * b.n_pad);
    if (rc != SFE_OK) return rc;
    b.sort_ws = (unsigned long long *)ctx->scratch[SCR_CLOUD].ptr;
  }
  downsample_kernel<<<grid, CLOUD_THREADS, smem, ctx->stream>>>(b);
  SFE_CUDA(cudaGetLastError());
  ctx->launches++;
  return SFE_OK;
}

// n_split > 0 (front end): two launches by cloud size.  The buffers are sized for n_max points per cloud, the
// clouds of a sonar frame are far smaller, and the kernel's occupancy is set by its shared-memory layout: the first
// launch is laid out for clouds of up to n_split points (3-4 CTAs per SM instead of 2) and skips larger ones, the
// second, laid out for n_max, takes only those (normally none: its CTAs read the counts and leave).
int downsample_run(sfe_ctx *ctx, const float *pts, const int *off, const int *cnt, int n_clouds, int dim, int n_max,
                   float resolution, float *out_pts, int32_t *out_idx, int32_t *out_count, int n_split) {
  SFE_REQUIRE(ctx, "downsample: null context");
  SFE_REQUIRE(n_clouds >= 0 && n_max >= 0, "downsample: negative sizes");
  SFE_REQUIRE(dim == 2 || dim == 3, "downsample: points must have 2 or 3 columns (got %d)", dim);
  if (n_clouds == 0) return SFE_OK;
  SFE_REQUIRE(pts && off && out_pts && out_idx && out_count, "downsample: null pointer");
  SFE_REQUIRE(dim == 2, "downsample: only 2-column clouds are supported (the reference only passes [x, y])");
  CloudBatch b{};
  b.pts = pts, b.off = off, b.cnt = cnt, b.n_clouds = n_clouds, b.dim = dim, b.n_max = n_max > 0 ? n_max : 1;
  b.resolution = resolution;
  b.out_pts = out_pts, b.out_idx = out_idx, b.out_count = out_count;
  if (n_split > 0 && n_split < b.n_max) {
    int rc = downsample_launch(ctx, b, n_split, -1, n_split);
    if (rc != SFE_OK) return rc;
    return downsample_launch(ctx, b, b.n_max, n_split, b.n_max);
  }
  return downsample_launch(ctx, b, b.n_max, -1, b.n_max);
}

static int remove_outlier_launch(sfe_ctx *ctx, CloudBatch b, int n_lay, int n_lo, int n_hi) {
  b.n_lay = n_lay, b.n_lo = n_lo, b.n_hi = n_hi;
  b.max_cells = 2 * n_lay < 256 ? 256 : (2 * n_lay > GRID_MAX_CELLS ? GRID_MAX_CELLS : 2 * n_lay);
  const size_t smem = sizeof(float2) * (size_t)n_lay + sizeof(uint32_t) * (size_t)((b.max_cells + 2) / 2 + 1) + 16;
  if (smem > (size_t)ctx->max_smem_optin - 4096) {
    set_error("remove_outlier: clouds of %d points are not supported (shared memory)", n_lay);
    return SFE_ERR_UNSUPPORTED;
  }
  // small clouds: smaller CTAs, more of them per SM (the stages are latency-bound chains; concurrency hides them)
  const int threads = n_lay <= 1024 ? 256 : CLOUD_THREADS;
  SFE_CUDA(cudaFuncSetAttribute(remove_outlier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  SFE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, remove_outlier_kernel, threads, smem));
  if (per_sm < 1) per_sm = 1;
  int grid = ctx->sm_count * per_sm;
  if (grid > b.n_clouds) grid = b.n_clouds;
  // (each class keeps its own workspace: the launches of one call may overlap on the stream's timeline only in
  // order, but the strides differ)
  int rc = ensure(ctx, ctx->scratch[SCR_MISC], (size_t)grid * n_lay * sizeof(uint16_t));
  if (rc != SFE_OK) return rc;
  b.orig_ws = (uint16_t *)ctx->scratch[SCR_MISC].ptr;
  remove_outlier_kernel<<<grid, threads, smem, ctx->stream>>>(b);
  SFE_CUDA(cudaGetLastError());
  ctx->launches++;
  return SFE_OK;
}

int remove_outlier_run(sfe_ctx *ctx, const float *pts, const int *off, const int *cnt, int n_clouds, int dim, int n_max,
                       double radius, int min_points, float *out_pts, int32_t *out_idx, int32_t *out_count,
                       int n_split) {
  SFE_REQUIRE(ctx, "remove_outlier: null context");
  SFE_REQUIRE(n_clouds >= 0 && n_max >= 0, "remove_outlier: negative sizes");
  SFE_REQUIRE(dim == 2 || dim == 3, "remove_outlier: points must have 2 or 3 columns (got %d)", dim);
  SFE_REQUIRE(radius > 0.0, "remove_outlier: radius must be positive");
  if (n_clouds == 0) return SFE_OK;
  SFE_REQUIRE(pts && off && out_pts && out_idx && out_count, "remove_outlier: null pointer");
  SFE_REQUIRE(n_max <= 65535, "remove_outlier: clouds of more than 65535 points are not supported (got %d)", n_max);
  CloudBatch b{};
  b.pts = pts, b.off = off, b.cnt = cnt, b.n_clouds = n_clouds, b.dim = dim, b.n_max = n_max > 0 ? n_max : 1;
  b.radius = radius, b.min_points = min_points;
  b.out_pts = out_pts, b.out_idx = out_idx, b.out_count = out_count;
  if (n_split > 0 && n_split < b.n_max) {
    int rc = remove_outlier_launch(ctx, b, n_split, -1, n_split);
    if (rc != SFE_OK) return rc;
    return remove_outlier_launch(ctx, b, b.n_max, n_split, b.n_max);
  }
  return remove_outlier_launch(ctx, b, b.n_max, -1, b.n_max);
}

}  // namespace sfe
