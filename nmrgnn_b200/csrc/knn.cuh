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
  __shared__ float md[KNN_QPB][KNN_PARTS][KNN_KMAX];
  __shared__ int mi[KNN_QPB][KNN_PARTS][KNN_KMAX];
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
  float bd[KNN_KMAX];
  int bi[KNN_KMAX];
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
      while (pos > 0 && knn_before(d2, j, bd[pos - 1], bi[pos - 1])) {
        bd[pos] = bd[pos - 1];
        bi[pos] = bi[pos - 1];
        --pos;
      }
      bd[pos] = d2;
      bi[pos] = j;
      if (count < k) ++count;
      if (count == k) {
        worst = bd[k - 1];
        worst_i = bi[k - 1];
      }
    }
  }
  // 8-way merge of the partial lists (each sorted by (d2, index)) by the query's first thread
  for (int s = 0; s < count; ++s) {
    md[ql][part][s] = bd[s];
    mi[ql][part][s] = bi[s];
  }
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
        const float dv = md[ql][t][head[t]];
        const int iv = mi[ql][t][head[t]];
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

}  // namespace nmr
