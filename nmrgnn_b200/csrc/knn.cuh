// k-nearest-neighbour graph builder on the device: the input producer of the forward
// path (what nmrdata.parse_universe + the inv_degree line do on the host in the
// reference: nmrgnn/library.py:111-116, nmrgnn/main.py:239-242).
// Exact brute force within each graph (graphs are a few thousand atoms): one thread per
// query atom, candidates streamed through shared memory in tiles, running top-k kept
// sorted by (distance^2, index).  Padded slots: local index 0 / distance 0.
#pragma once
#include "common.cuh"

namespace nmr {

constexpr int KNN_THREADS = 128;
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

__global__ void __launch_bounds__(KNN_THREADS) knn_graph_kernel(const KnnArgs p) {
  __shared__ float4 tile[KNN_TILE];
  const int g = blockIdx.x;
  const int64_t a0 = p.offsets[g], a1 = p.offsets[g + 1];
  const int64_t n = a1 - a0;
  const int64_t q_local = (int64_t)blockIdx.y * KNN_THREADS + threadIdx.x;
  if ((int64_t)blockIdx.y * KNN_THREADS >= n) return;  // whole block out of range
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
  for (int64_t t0 = 0; t0 < n; t0 += KNN_TILE) {
    const int tn = (int)min((int64_t)KNN_TILE, n - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < tn; i += KNN_THREADS) {
      const float* c = p.pos + (a0 + t0 + i) * 3;
      tile[i] = make_float4(c[0], c[1], c[2], 0.f);
    }
    __syncthreads();
    if (!active) continue;
    for (int i = 0; i < tn; ++i) {
      const float4 c = tile[i];
      const float dx = __fsub_rn(c.x, qx), dy = __fsub_rn(c.y, qy), dz = __fsub_rn(c.z, qz);
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      const int j = (int)(t0 + i);
      if (j == q_local) continue;
      if (p.cutoff2 > 0.f && d2 > p.cutoff2) continue;
      if (count == k && !(d2 < worst)) continue;
      // insert keeping (d2, index) order; candidates arrive in index order so ties stay stable
      int pos = count < k ? count : k - 1;
      while (pos > 0 && bd[pos - 1] > d2) {
        bd[pos] = bd[pos - 1];
        bi[pos] = bi[pos - 1];
        --pos;
      }
      bd[pos] = d2;
      bi[pos] = j;
      if (count < k) ++count;
      if (count == k) worst = bd[k - 1];
    }
  }
  if (!active) return;
  const int64_t row = (a0 + q_local) * k;
  int deg = 0;
  for (int s = 0; s < k; ++s) {
    int j = 0;
    float d = 0.f;
    if (s < count) {
      j = bi[s];
      d = sqrtf(bd[s]);
    }
    deg += (j > 0);  // library.py:115-116 counts nlist > 0 (a real neighbour with index 0 is not counted)
    p.nlist[row + s] = (int32_t)(a0 + j);
    p.edges[row + s] = d;
  }
  p.inv_degree[a0 + q_local] = deg > 0 ? __fdiv_rn(1.0f, (float)deg) : 0.0f;
}

}  // namespace nmr
