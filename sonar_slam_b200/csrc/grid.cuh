// Uniform-grid nearest-neighbour machinery shared by the ICP, match and outlier kernels.
//
// This is the GPU counterpart of the KD-tree the reference reaches through libpointmatcher /
// libnabo (KDTreeMatcher, pcl.cpp:161-174 and icp.yaml:5-9) and PCL (RadiusOutlierRemoval,
// pcl.cpp:54-74): an exact search structure, rebuilt per call like the reference does, but laid
// out for a CTA: the reference cloud is counting-sorted by cell (row-major cells, so a row span of
// cells is one contiguous point range) and lives in shared memory together with 16-bit cell
// offsets; a query walks square rings of cells around its own cell until the ring boundary is
// farther than the best match.
//
// Conventions (same as oracle/cloud_ref.c): float32 squared distance dx*dx + dy*dy with each
// product and the sum rounded separately; ties go to the lowest ORIGINAL index.
#pragma once
#include "common.cuh"

namespace sfe {

constexpr int GRID_MAX_CELLS = 16384;  // 16-bit offsets, two per 32-bit word

struct GridView {
  const float2 *pts;        // points sorted by cell
  const uint16_t *cstart;   // [ncells + 1] first sorted position of every cell
  const uint16_t *orig;     // sorted position -> index in the caller's cloud
  int n, nx, ny;
  float ox, oy, cell, inv_cell;
};

__device__ __forceinline__ float dist2_rn(float dx, float dy) {
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

__device__ __forceinline__ int grid_cell_coord(float v, float o, float inv_cell, int n) {
  const int c = (int)floorf((v - o) * inv_cell);
  return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

// block-wide exclusive scan helper: returns the exclusive prefix of `v` over the CTA and the total
__device__ __forceinline__ int block_exclusive_scan(int v, int *warp_sums /* [33] smem */, int &total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nwarps ? warp_sums[lane] : 0;
    int wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += t;
    }
    warp_sums[lane] = wi - w;
    if (lane == 31) warp_sums[32] = wi;
  }
  __syncthreads();
  const int excl = warp_sums[warp] + incl - v;
  total = warp_sums[32];
  __syncthreads();
  return excl;
}

// Choose grid geometry for n points with bounding box [minx,maxx] x [miny,maxy].
__device__ __forceinline__ void grid_geometry(int n, float minx, float miny, float maxx, float maxy, float min_cell,
                                              GridView &g, int max_cells) {
  const float w = maxx - minx, h = maxy - miny;
  float cell = sqrtf((w * h) / (float)(n > 0 ? n : 1));
  if (!(cell > min_cell)) cell = min_cell;
  if (!(cell > 1e-6f)) cell = 1.f;
  // keep the cell table within max_cells (<= GRID_MAX_CELLS)
  for (int it = 0; it < 32; ++it) {
    const float inv = 1.f / cell;
    const long long nx = (long long)(w * inv) + 1, ny = (long long)(h * inv) + 1;
    if (nx * ny <= max_cells && nx < 32768 && ny < 32768) break;
    cell *= 1.5f;
  }
  g.cell = cell;
  g.inv_cell = 1.f / cell;
  g.ox = minx;
  g.oy = miny;
  g.nx = (int)(w * g.inv_cell) + 1;
  g.ny = (int)(h * g.inv_cell) + 1;
  g.n = n;
}

// Counting sort of `n` points (src[i] - (mx,my)) into cells.  All threads of the CTA call this.
//   sorted  [n]            float2, shared or global
//   cells   [ncells/2+1]   uint32 words holding two uint16 offsets each (shared), zeroed here
//   orig    [n]            uint16 original indices (global or shared), may be null
// g.nx/ny/ox/oy/cell must be set (grid_geometry).  n <= 65535.
__device__ __forceinline__ void grid_build(const float *__restrict__ src, int stride, int n, float mx, float my,
                                           GridView &g, float2 *sorted, uint32_t *cells, uint16_t *orig,
                                           int *scan_scratch /* [33] smem */) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int ncells = g.nx * g.ny;
  const int nwords = (ncells + 2) / 2;  // covers entries 0 .. ncells
  for (int w = tid; w < nwords; w += nt) cells[w] = 0;
  __syncthreads();
  // histogram
  for (int i = tid; i < n; i += nt) {
    const float x = src[(size_t)i * stride] - mx, y = src[(size_t)i * stride + 1] - my;
    const int c = grid_cell_coord(y, g.oy, g.inv_cell, g.ny) * g.nx + grid_cell_coord(x, g.ox, g.inv_cell, g.nx);
    atomicAdd(&cells[c >> 1], (c & 1) ? 0x10000u : 1u);
  }
  __syncthreads();
  // inclusive scan over cells: each thread owns a contiguous run of words
  const int per = (nwords + nt - 1) / nt;
  const int w0 = min(tid * per, nwords), w1 = min(w0 + per, nwords);
  int local = 0;
  for (int w = w0; w < w1; ++w) local += (int)(cells[w] & 0xffffu) + (int)(cells[w] >> 16);
  int total;
  int run = block_exclusive_scan(local, scan_scratch, total);
  for (int w = w0; w < w1; ++w) {
    const uint32_t v = cells[w];
    const int lo = run + (int)(v & 0xffffu), hi = lo + (int)(v >> 16);
    cells[w] = (uint32_t)lo | ((uint32_t)hi << 16);  // inclusive ends
    run = hi;
  }
  __syncthreads();
  // scatter from the back of every cell: the inclusive end counts down to the cell start
  for (int i = tid; i < n; i += nt) {
    const float x = src[(size_t)i * stride] - mx, y = src[(size_t)i * stride + 1] - my;
    const int c = grid_cell_coord(y, g.oy, g.inv_cell, g.ny) * g.nx + grid_cell_coord(x, g.ox, g.inv_cell, g.nx);
    const uint32_t old = atomicSub(&cells[c >> 1], (c & 1) ? 0x10000u : 1u);
    const int pos = (int)((c & 1) ? (old >> 16) : (old & 0xffffu)) - 1;
    sorted[pos] = make_float2(x, y);
    if (orig) orig[pos] = (uint16_t)i;
  }
  __syncthreads();
  if (tid == 0) {  // sentinel entry ncells = n
    uint16_t *cs = reinterpret_cast<uint16_t *>(cells);
    cs[ncells] = (uint16_t)n;
  }
  __syncthreads();
  g.pts = sorted;
  g.cstart = reinterpret_cast<const uint16_t *>(cells);
  g.orig = orig;
}

struct NNResult {
  float d2;
  int pos;  // sorted position, -1 if none
};

// scan a contiguous range of sorted points
__device__ __forceinline__ void nn_scan(const GridView &g, int s, int e, float qx, float qy, NNResult &r) {
  for (int p = s; p < e; ++p) {
    const float2 t = g.pts[p];
    const float d2 = dist2_rn(qx - t.x, qy - t.y);
    if (d2 < r.d2) {
      r.d2 = d2;
      r.pos = p;
    } else if (d2 == r.d2 && r.pos >= 0 && g.orig != nullptr) {
      if (g.orig[p] < g.orig[r.pos]) r.pos = p;  // tie: lowest original index
    }
  }
}

// exact nearest neighbour of (qx,qy); accepted only when d2 <= max_d2
__device__ __forceinline__ NNResult nn_query(const GridView &g, float qx, float qy, float max_d2) {
  NNResult r;
  r.d2 = INFINITY;
  r.pos = -1;
  if (g.n <= 0) return r;
  const int cx = grid_cell_coord(qx, g.ox, g.inv_cell, g.nx), cy = grid_cell_coord(qy, g.oy, g.inv_cell, g.ny);
  const int kmax = max(g.nx, g.ny);
  for (int k = 1; k <= kmax; ++k) {
    const int x0 = cx - k, x1 = cx + k, y0 = cy - k, y1 = cy + k;
    const int xa = max(x0, 0), xb = min(x1, g.nx - 1);
    if (k == 1) {  // the whole 3x3 block: three row spans
      for (int y = max(y0, 0); y <= min(y1, g.ny - 1); ++y)
        nn_scan(g, g.cstart[y * g.nx + xa], g.cstart[y * g.nx + xb + 1], qx, qy, r);
    } else {       // ring k: two full rows + two single cells per inner row
      if (y0 >= 0) nn_scan(g, g.cstart[y0 * g.nx + xa], g.cstart[y0 * g.nx + xb + 1], qx, qy, r);
      if (y1 < g.ny) nn_scan(g, g.cstart[y1 * g.nx + xa], g.cstart[y1 * g.nx + xb + 1], qx, qy, r);
      for (int y = max(y0 + 1, 0); y <= min(y1 - 1, g.ny - 1); ++y) {
        if (x0 >= 0) nn_scan(g, g.cstart[y * g.nx + x0], g.cstart[y * g.nx + x0 + 1], qx, qy, r);
        if (x1 < g.nx) nn_scan(g, g.cstart[y * g.nx + x1], g.cstart[y * g.nx + x1 + 1], qx, qy, r);
      }
    }
    // every unvisited point lies outside the (2k+1)^2 block of cells around the query's cell
    float bound = INFINITY;
    if (x0 > 0) bound = fminf(bound, qx - (g.ox + (float)x0 * g.cell));
    if (x1 < g.nx - 1) bound = fminf(bound, (g.ox + (float)(x1 + 1) * g.cell) - qx);
    if (y0 > 0) bound = fminf(bound, qy - (g.oy + (float)y0 * g.cell));
    if (y1 < g.ny - 1) bound = fminf(bound, (g.oy + (float)(y1 + 1) * g.cell) - qy);
    if (bound == INFINITY) break;
    if (bound > 0.f) {
      const float b = bound * (1.0f - 1e-5f) - 1e-6f;  // conservative against float rounding
      if (b > 0.f) {
        const float b2 = b * b * (1.0f - 1e-6f);
        if (b2 > r.d2 || b2 > max_d2) break;
      }
    }
  }
  if (r.pos >= 0 && !(r.d2 <= max_d2)) {
    r.pos = -1;
    r.d2 = INFINITY;
  }
  return r;
}

}  // namespace sfe
