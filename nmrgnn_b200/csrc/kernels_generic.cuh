// Generic-geometry FP32 kernels: any atom_feature_size / edge_hidden_size / edge_feature_size / layer
// counts of the reference's hyper-parameter space (nmrgnn/model.py:22-36) that the tiled kernels do not
// cover (they need F = 256, H = 128).  Straightforward one-thread-per-output kernels in the reference's
// op order: correctness and coverage, not speed.
#pragma once
#include "common.cuh"

namespace nmr {

// X[e, r] = exp(-(d_e - mu_r)^2 / gap) * (d_e > 0)        (layers.py:137-140, model.py:251-257)
__global__ void __launch_bounds__(256) gen_rbf_kernel(const float* __restrict__ edges, const float* __restrict__ centers,
                                                      float gap, float* __restrict__ X, int64_t n_edges, int H) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n_edges * H) return;
  const int64_t e = i / H;
  const int r = (int)(i % H);
  const float d = edges[e];
  const float diff = d - centers[r];
  X[i] = d > 0.0f ? expf(__fdiv_rn(-__fmul_rn(diff, diff), gap)) : 0.0f;
}

// Y[i, o] = act(sum_k X[i, k] W[k, o] + b[o]) (+ R[i, o]) (* mask_i);  b, R, mask_src optional.
// One warp computes 32 consecutive outputs of one row: W reads coalesce, the X row is broadcast.
__global__ void __launch_bounds__(256) gen_dense_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                        const float* __restrict__ b, const float* __restrict__ R,
                                                        const float* __restrict__ mask_src, float* __restrict__ Y,
                                                        int64_t n_rows, int K, int N, int act) {
  const int64_t i = blockIdx.x;
  const float* x = X + i * K;
  const float m = (mask_src == nullptr || mask_src[i] > 0.0f) ? 1.0f : 0.0f;
  for (int o = threadIdx.x; o < N; o += 256) {
    float acc = 0.0f;
    for (int k = 0; k < K; ++k) acc = fmaf(x[k], W[(size_t)k * N + o], acc);
    if (b != nullptr) acc += b[o];
    acc = apply_act(acc, act);
    if (R != nullptr) acc += R[i * N + o];
    Y[i * N + o] = acc * m;
  }
}

// h_out[i, m] = act(inv_degree[i] * sum_j sum_n e[i,j,n] sum_l h[nl[i,j], l] w[l, m, n]) + h[i, m]
// (layers.py:26-46 + model.py:167).  Block = one atom: T[l, n] = sum_j e[i,j,n] h[nl[i,j], l] is built in
// shared memory first, then every thread contracts it with w for its output features.
__global__ void __launch_bounds__(256) gen_mp_kernel(const float* __restrict__ h_in, const int32_t* __restrict__ nlist,
                                                     const float* __restrict__ efeat, const float* __restrict__ inv_degree,
                                                     const float* __restrict__ w, float* __restrict__ h_out,
                                                     int64_t n_atoms, int K, int F, int E, int act) {
  extern __shared__ float T[];   // [F][E]
  const int64_t i = blockIdx.x;
  for (int t = threadIdx.x; t < F * E; t += 256) {
    const int l = t / E, n = t % E;
    float acc = 0.0f;
    for (int j = 0; j < K; ++j) {
      const float e = efeat[(i * K + j) * E + n];
      if (e != 0.0f) {
        int32_t idx = nlist[i * K + j];
        idx = min(max(idx, 0), (int32_t)(n_atoms - 1));   // validity is flagged by the index check kernel
        acc = fmaf(e, h_in[(size_t)idx * F + l], acc);
      }
    }
    T[t] = acc;
  }
  __syncthreads();
  const float s = inv_degree[i];
  for (int m = threadIdx.x; m < F; m += 256) {
    float acc = 0.0f;
    for (int l = 0; l < F; ++l)
      for (int n = 0; n < E; ++n) acc = fmaf(T[l * E + n], w[((size_t)l * F + m) * E + n], acc);
    h_out[i * F + m] = apply_act(acc * s, act) + h_in[i * F + m];
  }
}

// peaks[i] = sum_c (z_i . Wo[:, c] + bo[c]) * a[i,c] * std[c] + a[i,c] * avg[c]    (model.py:268-273)
__global__ void __launch_bounds__(256) gen_readout_kernel(const float* __restrict__ Z, const float* __restrict__ atoms,
                                                          const float* __restrict__ Wo, const float* __restrict__ bo,
                                                          const float* __restrict__ peak_std, const float* __restrict__ peak_avg,
                                                          float* __restrict__ peaks, int64_t n_atoms, int F2, int C) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n_atoms) return;
  float peak = 0.0f;
  for (int c = 0; c < C; ++c) {
    const float a = atoms[i * C + c];
    if (a != 0.0f) {
      float dot = 0.0f;
      for (int k = 0; k < F2; ++k) dot = fmaf(Z[i * F2 + k], Wo[(size_t)k * C + c], dot);
      peak += (dot + bo[c]) * a * peak_std[c] + a * peak_avg[c];
    }
  }
  peaks[i] = peak;
}

// flags nlist entries outside [0, n_atoms)
__global__ void __launch_bounds__(256) gen_index_check_kernel(const int32_t* __restrict__ nlist, int64_t n, int64_t n_atoms,
                                                              int* err_flag) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n && (nlist[i] < 0 || nlist[i] >= n_atoms)) atomicOr(err_flag, 1);
}

}  // namespace nmr
