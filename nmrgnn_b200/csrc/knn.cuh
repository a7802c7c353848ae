// k-nearest-neighbour graph builder on the device: the input producer of the forward
// path (what nmrdata.parse_universe + the inv_degree line do on the host in the
// reference: nmrgnn/library.py:111-116, nmrgnn/main.py:239-242).
// Exact brute force within each graph (graphs are a few thousand atoms): one thread per
// query atom, candidates streamed through shared memory in tiles, running top-k kept
// sorted by (distance^2, index).  Padded slots: local index 0 / distance 0.
// The candidate tiles are visited outward from the block's own index range (own tile, +1, -1, +2, ...):
// molecules are stored in chain order, so the k-th best distance becomes tight after one or two tiles and
// almost every later candidate is rejected by a single compare instead of shifting the sorted list.
#pragma once
#include "common.cuh"

namespace nmr {

constexpr int KNN_THREADS = 128;
constexpr int KNN_PARTS = 8;                        // threads per query atom: each scans every 8th candidate
constexpr int KNN_QPB = KNN_THREADS / KNN_PARTS;    // query atoms per block
constexpr int KNN_TILE = 512;
constexpr int KNN_KMAX = 32;

struct KnnArgs {
  const float* pos;          // [n_atoms, 3] nm
  const int64_t* offsets;    // device [n_graphs + 1]
  int32_t* nlist;            // [n_atoms, k]  (global = local + graph offset)
  float* edges;              // [n_atoms, k]
  float* inv_degree;         // [n_atoms]
  int k;
  float cutoff2;             // <= 0: none
};

// (d2, index) lexicographic order: deterministic neighbour order, ties broken by the lower index
__device__ __forceinline__ bool knn_before(float da, int ia, float db, int ib) { return da < db || (da == db && ia < ib); }

// One graph per blockIdx.x, KNN_QPB query atoms per block, KNN_PARTS threads per query: a single thread per query
// runs a long dependent compare chain with one warp per scheduler (measured 0.74 ms for a 2 482-atom protein on 20
// blocks); splitting the candidates 8 ways fills the GPU and an 8-way merge in shared memory restores the exact order.
__global__ void __launch_bounds__(KNN_THREADS) knn_graph_kernel(const KnnArgs p) {
  __shared__ float4 tile[KNN_TILE];
  // running top-k of every thread, slot-major (conflict-free when the lanes touch the same slot).  In shared memory, not
  // in a per-thread array: a dynamically indexed array lives in local memory, and 256 B x 768 threads per SM thrash the
  // L1 (ncu: 78 % of the local sectors missed, the kernel ran 10x longer than its arithmetic).
  __shared__ float md[KNN_KMAX][KNN_THREADS];
  __shared__ int mi[KNN_KMAX][KNN_THREADS];
  __shared__ int mc[KNN_QPB][KNN_PARTS];
  const int g = blockIdx.x;
  const int64_t a0 = p.offsets[g], a1 = p.offsets[g + 1];
  const int64_t n = a1 - a0;
  if ((int64_t)blockIdx.y * KNN_QPB >= n) return;  // whole block out of range
  const int ql = threadIdx.x / KNN_PARTS, part = threadIdx.x % KNN_PARTS;
  const int64_t q_local = (int64_t)blockIdx.y * KNN_QPB + ql;
  const bool active = q_local < n;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (active) {
    const float* q = p.pos + (a0 + q_local) * 3;
    qx = q[0];
    qy = q[1];
    qz = q[2];
  }
  const int tx = threadIdx.x;
  const int k = p.k;
  int count = 0;
  float worst = 3.4e38f;
  int worst_i = 0x7fffffff;
  const int n_tiles = (int)((n + KNN_TILE - 1) / KNN_TILE);
  // tiles are visited outward from the block's own index range (chain-ordered molecules: the k-th best distance
  // becomes tight after one or two tiles and later candidates are rejected by one compare)
  const int tb = (int)(((int64_t)blockIdx.y * KNN_QPB + KNN_QPB / 2) / KNN_TILE);
  for (int sidx = 0, visited = 0; visited < n_tiles; ++sidx) {
    const int dt = (sidx + 1) >> 1;
    const int tix = (sidx & 1) ? tb + dt : tb - dt;
    if (tix < 0 || tix >= n_tiles) continue;       // uniform over the block
    ++visited;
    const int64_t t0 = (int64_t)tix * KNN_TILE;
    const int tn = (int)min((int64_t)KNN_TILE, n - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < tn; i += KNN_THREADS) {
      const float* c = p.pos + (a0 + t0 + i) * 3;
      tile[i] = make_float4(c[0], c[1], c[2], 0.f);
    }
    __syncthreads();
    if (!active) continue;
    for (int i = part; i < tn; i += KNN_PARTS) {
      const float4 c = tile[i];
      const float dx = __fsub_rn(c.x, qx), dy = __fsub_rn(c.y, qy), dz = __fsub_rn(c.z, qz);
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      const int j = (int)(t0 + i);
      if (j == q_local) continue;
      if (p.cutoff2 > 0.f && d2 > p.cutoff2) continue;
      if (count == k && !knn_before(d2, j, worst, worst_i)) continue;
      int pos = count < k ? count : k - 1;
      while (pos > 0 && knn_before(d2, j, md[pos - 1][tx], mi[pos - 1][tx])) {
        md[pos][tx] = md[pos - 1][tx];
        mi[pos][tx] = mi[pos - 1][tx];
        --pos;
      }
      md[pos][tx] = d2;
      mi[pos][tx] = j;
      if (count < k) ++count;
      if (count == k) {
        worst = md[k - 1][tx];
        worst_i = mi[k - 1][tx];
      }
    }
  }
  // 8-way merge of the partial lists (each sorted by (d2, index)) by the query's first thread
  mc[ql][part] = count;
  __syncthreads();
  if (!active || part != 0) return;
  int head[KNN_PARTS];
#pragma unroll
  for (int t = 0; t < KNN_PARTS; ++t) head[t] = 0;
  const int64_t row = (a0 + q_local) * k;
  int deg = 0;
  for (int s = 0; s < k; ++s) {
    int best = -1;
    float bdv = 0.f;
    int biv = 0;
#pragma unroll
    for (int t = 0; t < KNN_PARTS; ++t) {
      if (head[t] < mc[ql][t]) {
        const float dv = md[head[t]][ql * KNN_PARTS + t];
        const int iv = mi[head[t]][ql * KNN_PARTS + t];
        if (best < 0 || knn_before(dv, iv, bdv, biv)) {
          best = t;
          bdv = dv;
          biv = iv;
        }
      }
    }
    int j = 0;
    float d = 0.f;
    if (best >= 0) {
#pragma unroll
      for (int t = 0; t < KNN_PARTS; ++t)
        if (t == best) ++head[t];
      j = biv;
      d = sqrtf(bdv);
    }
    deg += (j > 0);  // library.py:115-116 counts nlist > 0 (a real neighbour with index 0 is not counted)
    p.nlist[row + s] = (int32_t)(a0 + j);
    p.edges[row + s] = d;
  }
  p.inv_degree[a0 + q_local] = deg > 0 ? __fdiv_rn(1.0f, (float)deg) : 0.0f;
}

// ----------------------------------------------------------------------------------------------------------------
// Cell-list form: the brute-force kernel above evaluates n distances per query atom; a protein frame
// needs ~300.  Per graph, one block bins the atoms into a uniform grid (counting sort: histogram, scan, scatter into
// `sorted` = {x, y, z, bits(local index)} in cell order); the query kernel scans the (2r+1)^3 block of cells around
// the query's cell, starting with r = 1.  Everything outside that block is farther than r * cell from the query, so the
// k best are final as soon as the k-th distance is strictly below r * cell (or the block covers the grid); otherwise
// r grows and the block is rescanned.  Same distance expression, same (distance^2, index) order: the output is
// bit-identical to the brute-force kernel (tests/test_gpu_parity.py::test_knn_cell_list_equals_brute_force).
// ----------------------------------------------------------------------------------------------------------------
constexpr int KNN_MAX_CELLS = 8192;
constexpr int KNN_BUILD_THREADS = 1024;
// Atoms per cell OF THE BOUNDING BOX.  The 3 x 3 x 3 block around a query is final once the k-th distance is below the cell
// size c, i.e. once a sphere of radius c holds k atoms: 4.19 c^3 rho >= k, about k / 4 atoms per OCCUPIED cell; a protein
// fills a quarter to a half of its bounding box, hence k / 12.  (Round 2 started with 6 for every k: cells twice too
// wide, 43 x more candidates in the block than in the sphere that matters.)  Any value is exact -- the block grows.
__host__ __device__ __forceinline__ float knn_atoms_per_cell(int k) { return fmaxf(0.5f, (float)k * (1.0f / 12.0f)); }

struct KnnGrid {
  float ox, oy, oz, inv_c;
  float c;
  int nx, ny, nz;
};

struct KnnCellArgs {
  const float* pos;          // [n_atoms, 3] nm
  const int64_t* offsets;    // device [n_graphs + 1]
  float4* sorted;            // [n_atoms] atoms of every graph in cell order
  uint32_t* cell_start;      // [n_graphs][KNN_MAX_CELLS + 1]
  KnnGrid* grid;             // [n_graphs]
  int k;
};

__device__ __forceinline__ int knn_cell_coord(float x, float o, float inv_c, int n) {
  const int c = (int)floorf(__fmul_rn(__fsub_rn(x, o), inv_c));
  return min(max(c, 0), n - 1);
}

__global__ void __launch_bounds__(KNN_BUILD_THREADS) knn_build_cells_kernel(const KnnCellArgs p) {
  __shared__ uint32_t cnt[KNN_MAX_CELLS];
  __shared__ float red[6][32];
  __shared__ uint32_t wsum[32];
  __shared__ KnnGrid gs;
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t a0 = p.offsets[g];
  const int n = (int)(p.offsets[g + 1] - a0);
  if (n <= 0) return;
  // ---- bounding box
  float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int i = tid; i < n; i += KNN_BUILD_THREADS) {
    const float* q = p.pos + (a0 + i) * 3;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      lo[d] = fminf(lo[d], q[d]);
      hi[d] = fmaxf(hi[d], q[d]);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
    if (lane == 0) {
      red[d][warp] = lo[d];
      red[3 + d][warp] = hi[d];
    }
  }
  __syncthreads();
  if (tid == 0) {
    float l[3], e[3];
    for (int d = 0; d < 3; ++d) {
      float a = red[d][0], b = red[3 + d][0];
      for (int w = 1; w < KNN_BUILD_THREADS / 32; ++w) {
        a = fminf(a, red[d][w]);
        b = fmaxf(b, red[3 + d][w]);
      }
      l[d] = a;
      e[d] = fmaxf(b - a, 1e-3f);
    }
    float c = cbrtf(knn_atoms_per_cell(p.k) * e[0] * e[1] * e[2] / (float)n);
    c = fmaxf(c, 0.05f);
    int nx, ny, nz;
    for (;;) {
      nx = (int)(e[0] / c) + 1;
      ny = (int)(e[1] / c) + 1;
      nz = (int)(e[2] / c) + 1;
      if ((int64_t)nx * ny * nz <= KNN_MAX_CELLS) break;
      c *= 1.26f;
    }
    gs.ox = l[0];
    gs.oy = l[1];
    gs.oz = l[2];
    gs.c = c;
    gs.inv_c = 1.0f / c;
    gs.nx = nx;
    gs.ny = ny;
    gs.nz = nz;
    p.grid[g] = gs;
  }
  __syncthreads();
  const KnnGrid G = gs;
  const int cells = G.nx * G.ny * G.nz;
  for (int i = tid; i < cells; i += KNN_BUILD_THREADS) cnt[i] = 0u;
  __syncthreads();
  // ---- histogram
  for (int i = tid; i < n; i += KNN_BUILD_THREADS) {
    const float* q = p.pos + (a0 + i) * 3;
    const int cx = knn_cell_coord(q[0], G.ox, G.inv_c, G.nx), cy = knn_cell_coord(q[1], G.oy, G.inv_c, G.ny),
              cz = knn_cell_coord(q[2], G.oz, G.inv_c, G.nz);
    atomicAdd(&cnt[(cz * G.ny + cy) * G.nx + cx], 1u);
  }
  __syncthreads();
  // ---- exclusive scan over the cells: 8 consecutive cells per thread, block scan of the partial sums
  constexpr int PER = KNN_MAX_CELLS / KNN_BUILD_THREADS;
  uint32_t v[PER], sum = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int c = tid * PER + j;
    v[j] = c < cells ? cnt[c] : 0u;
    sum += v[j];
  }
  uint32_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = wsum[lane], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    wsum[lane] = winc - w;
  }
  __syncthreads();
  uint32_t run = wsum[warp] + inc - sum;
  uint32_t* cs = p.cell_start + (size_t)g * (KNN_MAX_CELLS + 1);
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int c = tid * PER + j;
    if (c < cells) {
      cnt[c] = run;          // becomes the scatter cursor
      cs[c] = run;
    }
    run += v[j];
  }
  if (tid == 0) cs[cells] = (uint32_t)n;
  __syncthreads();
  // ---- scatter (order inside a cell is arbitrary; the query kernel orders candidates by (distance^2, index))
  for (int i = tid; i < n; i += KNN_BUILD_THREADS) {
    const float* q = p.pos + (a0 + i) * 3;
    const float x = q[0], y = q[1], z = q[2];
    const int cx = knn_cell_coord(x, G.ox, G.inv_c, G.nx), cy = knn_cell_coord(y, G.oy, G.inv_c, G.ny),
              cz = knn_cell_coord(z, G.oz, G.inv_c, G.nz);
    const uint32_t slot = atomicAdd(&cnt[(cz * G.ny + cy) * G.nx + cx], 1u);
    p.sorted[a0 + slot] = make_float4(x, y, z, __int_as_float(i));
  }
}

struct KnnQueryArgs {
  KnnArgs out;               // pos / offsets / outputs / k / cutoff2 as for the brute-force kernel
  const float4* sorted;
  const uint32_t* cell_start;
  const KnnGrid* grid;
};

__global__ void __launch_bounds__(KNN_THREADS) knn_query_cells_kernel(const KnnQueryArgs a) {
  // running top-k of every thread, slot-major (conflict-free when the lanes touch the same slot).  In shared memory, not
  // in a per-thread array: a dynamically indexed array lives in local memory, and 256 B x 768 threads per SM thrash the
  // L1 (ncu: 78 % of the local sectors missed, the kernel ran 10x longer than its arithmetic).
  __shared__ float md[KNN_KMAX][KNN_THREADS];
  __shared__ int mi[KNN_KMAX][KNN_THREADS];
  __shared__ int mc[KNN_QPB][KNN_PARTS];
  const KnnArgs& p = a.out;
  const int g = blockIdx.x;
  const int64_t a0 = p.offsets[g];
  const int n = (int)(p.offsets[g + 1] - a0);
  if ((int64_t)blockIdx.y * KNN_QPB >= n) return;
  const int ql = threadIdx.x / KNN_PARTS, part = threadIdx.x % KNN_PARTS, lane = threadIdx.x & 31;
  const unsigned gmask = 0xffu << (lane & 24);                 // the 8 lanes of this query
  const int q_local = blockIdx.y * KNN_QPB + ql;
  if (q_local >= n) return;                                     // whole 8-lane groups leave together
  const KnnGrid G = a.grid[g];
  const uint32_t* cs = a.cell_start + (size_t)g * (KNN_MAX_CELLS + 1);
  const float4* sp = a.sorted + a0;
  const float* q = p.pos + (a0 + q_local) * 3;
  const float qx = q[0], qy = q[1], qz = q[2];
  const int cx = knn_cell_coord(qx, G.ox, G.inv_c, G.nx), cy = knn_cell_coord(qy, G.oy, G.inv_c, G.ny),
            cz = knn_cell_coord(qz, G.oz, G.inv_c, G.nz);
  const int k = p.k;
  const int tx = threadIdx.x;
  const int64_t row = (a0 + q_local) * k;
  for (int r = 1;; ++r) {
    int count = 0;
    float worst = 3.4e38f;
    int worst_i = 0x7fffffff;
    const int x0 = max(cx - r, 0), x1 = min(cx + r, G.nx - 1), y0 = max(cy - r, 0), y1 = min(cy + r, G.ny - 1),
              z0 = max(cz - r, 0), z1 = min(cz + r, G.nz - 1);
    for (int z = z0; z <= z1; ++z)
      for (int y = y0; y <= y1; ++y) {
        // cells x0 .. x1 of a row are contiguous in the sorted order
        const int rb = (z * G.ny + y) * G.nx;
        const int s0 = (int)cs[rb + x0], s1 = (int)cs[rb + x1 + 1];
        for (int i = s0 + part; i < s1; i += KNN_PARTS) {
          const float4 c = sp[i];
          const float dx = __fsub_rn(c.x, qx), dy = __fsub_rn(c.y, qy), dz = __fsub_rn(c.z, qz);
          const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
          const int j = __float_as_int(c.w);
          if (j == q_local) continue;
          if (p.cutoff2 > 0.f && d2 > p.cutoff2) continue;
          if (count == k && !knn_before(d2, j, worst, worst_i)) continue;
          int pos = count < k ? count : k - 1;
          while (pos > 0 && knn_before(d2, j, md[pos - 1][tx], mi[pos - 1][tx])) {
            md[pos][tx] = md[pos - 1][tx];
            mi[pos][tx] = mi[pos - 1][tx];
            --pos;
          }
          md[pos][tx] = d2;
          mi[pos][tx] = j;
          if (count < k) ++count;
          if (count == k) {
            worst = md[k - 1][tx];
            worst_i = mi[k - 1][tx];
          }
        }
      }
    // 8-way merge of the partial lists by the query's first thread
    mc[ql][part] = count;
    __syncwarp(gmask);
    const bool whole = x0 == 0 && y0 == 0 && z0 == 0 && x1 == G.nx - 1 && y1 == G.ny - 1 && z1 == G.nz - 1;
    int done = 1;
    if (part == 0) {
      int head[KNN_PARTS];
#pragma unroll
      for (int t = 0; t < KNN_PARTS; ++t) head[t] = 0;
      // first pass: is the k-th merged distance safely inside the scanned block?
      float dk = 3.4e38f;
      int found = 0;
      for (int s = 0; s < k; ++s) {
        int best = -1;
        float bdv = 0.f;
        int biv = 0;
#pragma unroll
        for (int t = 0; t < KNN_PARTS; ++t)
          if (head[t] < mc[ql][t]) {
            const float dv = md[head[t]][ql * KNN_PARTS + t];
            const int iv = mi[head[t]][ql * KNN_PARTS + t];
            if (best < 0 || knn_before(dv, iv, bdv, biv)) {
              best = t;
              bdv = dv;
              biv = iv;
            }
          }
        if (best < 0) break;
#pragma unroll
        for (int t = 0; t < KNN_PARTS; ++t)
          if (t == best) ++head[t];
        dk = bdv;
        ++found;
      }
      // (0.9999: the cell of a point is floor((x - o) / c) in float32, so a cell boundary is only exact to a few ulp)
      const float reach = __fmul_rn(__fmul_rn((float)r, G.c), 0.9999f);
      const float lim2 = __fmul_rn(reach, reach);
      // with a cutoff nothing beyond it matters: the block is sufficient once it reaches the cutoff
      const bool cut_ok = p.cutoff2 > 0.f && lim2 > p.cutoff2;
      done = whole || cut_ok || (found == k && dk < lim2);
      if (done) {
#pragma unroll
        for (int t = 0; t < KNN_PARTS; ++t) head[t] = 0;
        int deg = 0;
        for (int s = 0; s < k; ++s) {
          int best = -1;
          float bdv = 0.f;
          int biv = 0;
#pragma unroll
          for (int t = 0; t < KNN_PARTS; ++t)
            if (head[t] < mc[ql][t]) {
              const float dv = md[head[t]][ql * KNN_PARTS + t];
              const int iv = mi[head[t]][ql * KNN_PARTS + t];
              if (best < 0 || knn_before(dv, iv, bdv, biv)) {
                best = t;
                bdv = dv;
                biv = iv;
              }
            }
          int j = 0;
          float d = 0.f;
          if (best >= 0) {
#pragma unroll
            for (int t = 0; t < KNN_PARTS; ++t)
              if (t == best) ++head[t];
            j = biv;
            d = sqrtf(bdv);
          }
          deg += (j > 0);  // library.py:115-116 counts nlist > 0 (a real neighbour with index 0 is not counted)
          p.nlist[row + s] = (int32_t)(a0 + j);
          p.edges[row + s] = d;
        }
        p.inv_degree[a0 + q_local] = deg > 0 ? __fdiv_rn(1.0f, (float)deg) : 0.0f;
      }
    }
    done = __shfl_sync(gmask, done, lane & 24);
    if (done) return;
    __syncwarp(gmask);     // the merge buffers are reused by the next, larger block
  }
}

// ----------------------------------------------------------------------------------------------------------------
// Cell-list query, one WARP per query atom (round 2, the default): the eight-threads-per-query kernel above executes
// 36 k warp instructions per warp of four queries -- their row scans and list insertions diverge, every lane keeps its
// own sorted list in shared memory and one lane in eight merges them, twice.  Here the 32 lanes of a warp read 32
// consecutive candidates of a cell row (coalesced float4), and the ONE sorted list of the query lives in registers, entry
// l in lane l (k <= 32): a candidate that beats the current k-th entry is broadcast, every lane compares it with its
// entry, the insert position is a popcount of a ballot and the tail moves up by one shuffle.  No shared memory, no
// divergence between queries, no merge.  Same block walk, same distance expression, same (distance^2, index) order
// as the kernels above: the output is the same bits.
// ----------------------------------------------------------------------------------------------------------------
constexpr int KNN_WARP_QPB = KNN_THREADS / 32;      // query atoms per block

__global__ void __launch_bounds__(KNN_THREADS) knn_query_warp_kernel(const KnnQueryArgs a) {
  const KnnArgs& p = a.out;
  const int g = blockIdx.x;
  const int64_t a0 = p.offsets[g];
  const int n = (int)(p.offsets[g + 1] - a0);
  const int lane = threadIdx.x & 31;
  const int q_local = blockIdx.y * KNN_WARP_QPB + (threadIdx.x >> 5);
  if (q_local >= n) return;                                     // (whole warps)
  const KnnGrid G = a.grid[g];
  const uint32_t* cs = a.cell_start + (size_t)g * (KNN_MAX_CELLS + 1);
  const float4* sp = a.sorted + a0;
  const float* q = p.pos + (a0 + q_local) * 3;
  const float qx = q[0], qy = q[1], qz = q[2];
  const int cx = knn_cell_coord(qx, G.ox, G.inv_c, G.nx), cy = knn_cell_coord(qy, G.oy, G.inv_c, G.ny),
            cz = knn_cell_coord(qz, G.oz, G.inv_c, G.nz);
  const int k = p.k;
  const int64_t row = (a0 + q_local) * k;
  float ed = 3.4e38f;          // this lane's entry of the sorted list (lanes >= count: unused)
  int ei = 0x7fffffff;
  int count = 0;               // warp-uniform
  int px0 = 1, px1 = 0, py0 = 1, py1 = 0, pz0 = 1, pz1 = 0;     // the block already scanned (empty)
  // 32 consecutive candidates of the sorted order: distance, prefilter, insertions
  auto scan = [&](int s0, int s1) {
    for (int base = s0; base < s1; base += 32) {
      const int i = base + lane;
      const bool valid = i < s1;
      const float4 c = sp[valid ? i : s0];
      const float dx = __fsub_rn(c.x, qx), dy = __fsub_rn(c.y, qy), dz = __fsub_rn(c.z, qz);
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      const int j = __float_as_int(c.w);
      bool cand = valid && j != q_local && !(p.cutoff2 > 0.f && d2 > p.cutoff2);
      if (count == k) {        // (uniform) prefilter against the current k-th entry
        const float wd = __shfl_sync(0xffffffffu, ed, k - 1);
        const int wi = __shfl_sync(0xffffffffu, ei, k - 1);
        cand = cand && knn_before(d2, j, wd, wi);
      }
      unsigned mask = __ballot_sync(0xffffffffu, cand);
      while (mask) {           // (uniform) one insertion per surviving candidate
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const float cd = __shfl_sync(0xffffffffu, d2, src);
        const int cj = __shfl_sync(0xffffffffu, j, src);
        if (count == k) {      // the list may have tightened since the prefilter
          const float wd = __shfl_sync(0xffffffffu, ed, k - 1);
          const int wi = __shfl_sync(0xffffffffu, ei, k - 1);
          if (!knn_before(cd, cj, wd, wi)) continue;
        }
        const int pos = __popc(__ballot_sync(0xffffffffu, lane < count && knn_before(ed, ei, cd, cj)));
        const float ud = __shfl_up_sync(0xffffffffu, ed, 1);
        const int ui = __shfl_up_sync(0xffffffffu, ei, 1);
        if (lane > pos && lane < k) {
          ed = ud;
          ei = ui;
        }
        if (lane == pos) {
          ed = cd;
          ei = cj;
        }
        if (count < k) ++count;
      }
    }
  };
  for (int r = 1;; ++r) {
    const int x0 = max(cx - r, 0), x1 = min(cx + r, G.nx - 1), y0 = max(cy - r, 0), y1 = min(cy + r, G.ny - 1),
              z0 = max(cz - r, 0), z1 = min(cz + r, G.nz - 1);
    // only the shell that the previous, smaller block did not cover is scanned; the list carries over
    for (int z = z0; z <= z1; ++z)
      for (int y = y0; y <= y1; ++y) {
        // cells x0 .. x1 of a row are contiguous in the sorted order
        const int rb = (z * G.ny + y) * G.nx;
        if (z >= pz0 && z <= pz1 && y >= py0 && y <= py1) {
          if (x0 < px0) scan((int)cs[rb + x0], (int)cs[rb + px0]);
          if (x1 > px1) scan((int)cs[rb + px1 + 1], (int)cs[rb + x1 + 1]);
        } else {
          scan((int)cs[rb + x0], (int)cs[rb + x1 + 1]);
        }
      }
    px0 = x0; px1 = x1; py0 = y0; py1 = y1; pz0 = z0; pz1 = z1;
    const bool whole = x0 == 0 && y0 == 0 && z0 == 0 && x1 == G.nx - 1 && y1 == G.ny - 1 && z1 == G.nz - 1;
    const float dk = __shfl_sync(0xffffffffu, ed, k - 1);
    // (0.9999: the cell of a point is floor((x - o) / c) in float32, so a cell boundary is only exact to a few ulp)
    const float reach = __fmul_rn(__fmul_rn((float)r, G.c), 0.9999f);
    const float lim2 = __fmul_rn(reach, reach);
    // with a cutoff nothing beyond it matters: the block is sufficient once it reaches the cutoff
    const bool cut_ok = p.cutoff2 > 0.f && lim2 > p.cutoff2;
    if (whole || cut_ok || (count == k && dk < lim2)) {
      const bool have = lane < count;
      const int j = have ? ei : 0;
      // library.py:115-116 counts nlist > 0 (a real neighbour with index 0 is not counted)
      const int deg = __popc(__ballot_sync(0xffffffffu, have && j > 0));
      if (lane < k) {
        p.nlist[row + lane] = (int32_t)(a0 + j);
        p.edges[row + lane] = have ? sqrtf(ed) : 0.0f;
      }
      if (lane == 0) p.inv_degree[a0 + q_local] = deg > 0 ? __fdiv_rn(1.0f, (float)deg) : 0.0f;
      return;
    }
  }
}

}  // namespace nmr
