// FP32 (FFMA) kernels of the GNN forward for sm_100a.  These are the exact-fp32
// path: used for every geometry the tensor-core path does not cover and as the
// on-device cross-check of the tcgen05 kernels.  One kernel per reference block:
//   edge_mlp_ffma_kernel    RBF -> EdgeFCBlock -> mask      (model.py:251-261)
//   embed_kernel            atoms @ W_e                      (model.py:262)
//   mp_layer_ffma_kernel    gather + bilinear einsum + act + residual (layers.py:26-46, model.py:167)
//   fc_readout_ffma_kernel  FCBlock + out_layer + peak standardisation (model.py:191-196,268-273)
#pragma once
#include "common.cuh"

namespace nmr {

constexpr int MAX_DENSE = 8;  // max Dense layers per block (reference range: 2..6)

// ----------------------------------------------------------------------------------
// Edge MLP.  One CTA = 128 edges.  Activations stay in shared memory between the
// chained [128x128] GEMMs; nothing but the E output features per edge is written.
// ----------------------------------------------------------------------------------
struct EdgeArgs {
  const float* edges;        // [n_edges]
  float* out;                // [n_edges, E]
  int64_t n_edges;
  const float* centers;      // [H] RBF grid
  float gap;
  const float* W[MAX_DENSE]; // hidden: [H,H]; last: [H,E]
  const float* b[MAX_DENSE];
  int n_layers;
  int act;
  const int32_t* nlist;      // optional [n_edges]: validated against n_atoms
  int64_t n_atoms;
  int* err_flag;
};

constexpr int EDGE_H = 128;
constexpr int EDGE_LDX = EDGE_H + 2;
constexpr int EDGE_THREADS = 256;

template <int E>
__host__ __device__ constexpr size_t edge_smem_bytes() {
  return sizeof(float) * (128 * EDGE_LDX + TileGemm<128, EDGE_THREADS>::SMEM_FLOATS + EDGE_H + 128 + EDGE_H * E + E);
}

template <int E>
__global__ void __launch_bounds__(EDGE_THREADS, 2) edge_mlp_ffma_kernel(const EdgeArgs p) {
  using G = TileGemm<128, EDGE_THREADS>;
  extern __shared__ __align__(16) float smem[];
  float* X = smem;                              // [128][EDGE_LDX]
  float* Bs = X + 128 * EDGE_LDX;               // W stream ring
  float* cen = Bs + G::SMEM_FLOATS;             // [H]
  float* dist = cen + EDGE_H;                   // [128]
  float* Wl = dist + 128;                       // [H][E] last layer
  float* bl = Wl + EDGE_H * E;                  // [E]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < EDGE_H; i += EDGE_THREADS) cen[i] = p.centers[i];
  for (int i = tid; i < EDGE_H * E; i += EDGE_THREADS) Wl[i] = p.W[p.n_layers - 1][i];
  if (tid < E) bl[tid] = p.b[p.n_layers - 1][tid];

  int r0, c0;
  G::thread_origin(tid, r0, c0);
  const int64_t n_tiles = (p.n_edges + 127) / 128;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t e0 = tile * 128;
    __syncthreads();  // previous tile's readers of X/dist are done
    if (tid < 128) {
      const int64_t e = e0 + tid;
      float d = 0.0f;
      if (e < p.n_edges) {
        d = p.edges[e];
        if (p.nlist != nullptr) {
          const int32_t idx = p.nlist[e];
          if (idx < 0 || idx >= p.n_atoms) atomicOr(p.err_flag, 1);
        }
      }
      dist[tid] = d;
    }
    __syncthreads();
    // RBF expansion * mask: exp(-(d - mu)^2 / gap), 0 for padded slots (d <= 0)
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const int row = warp * 16 + i;
      const float d = dist[row];
      const bool m = d > 0.0f;
#pragma unroll
      for (int r = lane; r < EDGE_H; r += 32) {
        const float diff = d - cen[r];
        const float v = expf(__fdiv_rn(-__fmul_rn(diff, diff), p.gap));
        X[row * EDGE_LDX + r] = m ? v : 0.0f;
      }
    }
    // hidden layers: X <- act(X @ W + b), in place
    for (int l = 0; l + 1 < p.n_layers; ++l) {
      float acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
      G::run(acc, X, EDGE_LDX, p.W[l], EDGE_H, EDGE_H, Bs, true);
      const float4 bA = *reinterpret_cast<const float4*>(p.b[l] + c0);
      const float4 bB = *reinterpret_cast<const float4*>(p.b[l] + c0 + 64);
      const float bias[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float* xr = X + (r0 + i) * EDGE_LDX + c0;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = apply_act(acc[i][j] + bias[j], p.act);
        *reinterpret_cast<float2*>(xr) = make_float2(v[0], v[1]);
        *reinterpret_cast<float2*>(xr + 2) = make_float2(v[2], v[3]);
        *reinterpret_cast<float2*>(xr + 64) = make_float2(v[4], v[5]);
        *reinterpret_cast<float2*>(xr + 66) = make_float2(v[6], v[7]);
      }
      __syncthreads();
    }
    // last layer (linear, H -> E) * mask; two threads per edge split the k range
    {
      const int row = tid >> 1, half = tid & 1;
      float s[E];
#pragma unroll
      for (int n = 0; n < E; ++n) s[n] = 0.0f;
      const float* xr = X + row * EDGE_LDX + half * (EDGE_H / 2);
      const float* wr = Wl + half * (EDGE_H / 2) * E;
#pragma unroll 8
      for (int k = 0; k < EDGE_H / 2; ++k) {
        const float x = xr[k];
#pragma unroll
        for (int n = 0; n < E; ++n) s[n] = fmaf(x, wr[k * E + n], s[n]);
      }
#pragma unroll
      for (int n = 0; n < E; ++n) s[n] += __shfl_xor_sync(0xffffffffu, s[n], 1);
      const int64_t e = e0 + row;
      if (half == 0 && e < p.n_edges) {
        const bool m = dist[row] > 0.0f;
#pragma unroll
        for (int n = 0; n < E; ++n) p.out[e * E + n] = m ? s[n] + bl[n] : 0.0f;
      }
    }
  }
}

// ----------------------------------------------------------------------------------
// Embedding: nodes[i, :] = atoms[i, :] @ W_e   (Dense without bias)
// ----------------------------------------------------------------------------------
constexpr int EMBED_ATOMS = 16;
__global__ void __launch_bounds__(256) embed_kernel(const float* __restrict__ atoms, const float* __restrict__ We,
                                                    float* __restrict__ nodes, int64_t n_atoms, int C, int F) {
  extern __shared__ __align__(16) float a_s[];  // [EMBED_ATOMS][C]
  const int64_t i0 = (int64_t)blockIdx.x * EMBED_ATOMS;
  const int n_here = (int)min((int64_t)EMBED_ATOMS, n_atoms - i0);
  for (int i = threadIdx.x; i < n_here * C; i += blockDim.x) a_s[i] = atoms[i0 * C + i];
  __syncthreads();
  for (int m = threadIdx.x; m < F; m += blockDim.x) {
    float acc[EMBED_ATOMS];
#pragma unroll
    for (int i = 0; i < EMBED_ATOMS; ++i) acc[i] = 0.0f;
    for (int c = 0; c < C; ++c) {
      const float w = We[c * F + m];
#pragma unroll
      for (int i = 0; i < EMBED_ATOMS; ++i) acc[i] = fmaf(i < n_here ? a_s[i * C + c] : 0.0f, w, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < EMBED_ATOMS; ++i)
      if (i < n_here) nodes[(i0 + i) * F + m] = acc[i];
  }
}

// F = 256 variant used by the forward: 32 atoms per block, 16-byte stores (thread = 4 features of 8 atoms),
// and the per-atom max |h| that the tensor-core MP layer needs for its fp16 range scaling, written in the
// same pass (hmax may be null).
constexpr int EMBED256_ATOMS = 32;
__global__ void __launch_bounds__(256) embed256_kernel(const float* __restrict__ atoms, const float* __restrict__ We,
                                                       float* __restrict__ nodes, float* __restrict__ hmax,
                                                       int64_t n_atoms, int C) {
  extern __shared__ __align__(16) float a_s[];  // [EMBED256_ATOMS][C]
  __shared__ float m_s[EMBED256_ATOMS][2];
  const int64_t i0 = (int64_t)blockIdx.x * EMBED256_ATOMS;
  const int n_here = (int)min((int64_t)EMBED256_ATOMS, n_atoms - i0);
  for (int i = threadIdx.x; i < EMBED256_ATOMS * C; i += 256) a_s[i] = i < n_here * C ? atoms[i0 * C + i] : 0.0f;
  __syncthreads();
  const int cg = threadIdx.x & 63, sub = threadIdx.x >> 6;     // 4 features cg*4.., atoms sub, sub+4, ...
  float4 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < C; ++c) {
    const float4 w = *reinterpret_cast<const float4*>(We + c * 256 + cg * 4);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a = a_s[(sub + 4 * i) * C + c];
      acc[i].x = fmaf(a, w.x, acc[i].x);
      acc[i].y = fmaf(a, w.y, acc[i].y);
      acc[i].z = fmaf(a, w.z, acc[i].z);
      acc[i].w = fmaf(a, w.w, acc[i].w);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = sub + 4 * i;
    if (r < n_here) *reinterpret_cast<float4*>(nodes + (i0 + r) * 256 + cg * 4) = acc[i];
    float m = fmaxf(fmaxf(fabsf(acc[i].x), fabsf(acc[i].y)), fmaxf(fabsf(acc[i].z), fabsf(acc[i].w)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) m_s[r][(threadIdx.x >> 5) & 1] = m;   // the row's 64 threads are 2 warps
  }
  __syncthreads();
  if (hmax != nullptr && threadIdx.x < n_here) hmax[i0 + threadIdx.x] = fmaxf(m_s[threadIdx.x][0], m_s[threadIdx.x][1]);
}

// ----------------------------------------------------------------------------------
// MP layer, F = 256.  One CTA = 128 atoms.  For each 32-feature slice of the input
// nodes the warps gather-aggregate T[i,(n,l)] = sum_j e[i,j,n] h[nl[i,j],l] into
// shared memory (rows come from L1/L2: neighbours of consecutive atoms overlap),
// then the block multiplies that [128 x 32E] slice with the matching rows of the
// re-packed weight W'[(slice,n,l), m] = w[l,m,n].  Epilogue: * inv_degree,
// activation, + residual.  T never reaches HBM.
// ----------------------------------------------------------------------------------
struct MpArgs {
  const float* h_in;       // [n_atoms, 256]
  float* h_out;            // [n_atoms, 256]
  const int32_t* nlist;    // [n_atoms, K]
  const float* efeat;      // [n_atoms, K, E]
  const float* inv_degree; // [n_atoms]
  const float* Wp;         // [(F/32) * 32E, 256] packed, see pack_mp_weight_ffma()
  int64_t n_atoms;
  int K;
  int act;
  int raw;                 // 1: h_out = inv_degree * D (no activation, no residual) -- calibration tap
};

constexpr int MP_F = 256;
constexpr int MP_THREADS = 512;

template <int E>
__host__ __device__ constexpr int mp_ldt() { return 32 * E + 2; }
template <int E>
__host__ inline size_t mp_smem_bytes(int K) {
  return sizeof(float) * (size_t)(128 * mp_ldt<E>() + TileGemm<256, MP_THREADS>::SMEM_FLOATS) +
         (size_t)128 * K * (sizeof(int32_t) + 4 * sizeof(float));
}

template <int E>
__global__ void __launch_bounds__(MP_THREADS, 1) mp_layer_ffma_kernel(const MpArgs p) {
  static_assert(E >= 1 && E <= 4, "edge features are staged as float4");
  using G = TileGemm<256, MP_THREADS>;
  constexpr int LDT = mp_ldt<E>();
  constexpr int KCH = 32 * E;
  extern __shared__ __align__(16) float smem[];
  float* T = smem;                                             // [128][LDT]
  float* Bs = T + 128 * LDT;
  float4* e_s = reinterpret_cast<float4*>(Bs + G::SMEM_FLOATS);  // [128*K]
  int32_t* nl_s = reinterpret_cast<int32_t*>(e_s + 128 * p.K);   // [128*K]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.K;
  int r0, c0;
  G::thread_origin(tid, r0, c0);
  const int64_t n_tiles = (p.n_atoms + 127) / 128;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t a0 = tile * 128;
    const int rows = (int)min((int64_t)128, p.n_atoms - a0);
    __syncthreads();
    for (int i = tid; i < 128 * K; i += MP_THREADS) {
      int32_t idx = 0;
      float4 ev = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < rows * K) {
        idx = p.nlist[a0 * K + i];
        idx = min(max(idx, 0), (int32_t)(p.n_atoms - 1));  // never fault; validity is flagged by the edge kernel
        const float* ep = p.efeat + (a0 * K + i) * E;
        ev.x = ep[0];
        if (E > 1) ev.y = ep[1];
        if (E > 2) ev.z = ep[2];
        if (E > 3) ev.w = ep[3];
      }
      nl_s[i] = idx;
      e_s[i] = ev;
    }
    __syncthreads();

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    for (int c = 0; c < MP_F / 32; ++c) {
      const float* hcol = p.h_in + c * 32 + lane;
      // gather-aggregate this 32-feature slice; warp w owns rows 8w..8w+7
#pragma unroll 1
      for (int i = 0; i < 8; ++i) {
        const int row = warp * 8 + i;
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        const int32_t* nl = nl_s + row * K;
        const float4* ee = e_s + row * K;
#pragma unroll 8
        for (int j = 0; j < K; ++j) {
          const float4 ev = ee[j];
          if (ev.x != 0.f || ev.y != 0.f || ev.z != 0.f || ev.w != 0.f) {
            const float hv = __ldg(hcol + (size_t)nl[j] * MP_F);
            t[0] = fmaf(ev.x, hv, t[0]);
            if (E > 1) t[1] = fmaf(ev.y, hv, t[1]);
            if (E > 2) t[2] = fmaf(ev.z, hv, t[2]);
            if (E > 3) t[3] = fmaf(ev.w, hv, t[3]);
          }
        }
#pragma unroll
        for (int n = 0; n < E; ++n) T[row * LDT + n * 32 + lane] = t[n];
      }
      // acc += T[128 x 32E] @ W'[c*32E .. , :]   (run() barriers before reading T)
      G::run(acc, T, LDT, p.Wp + (size_t)c * KCH * MP_F, MP_F, KCH, Bs, true);
    }

    // epilogue: h_out = act(inv_degree * acc) + h_in
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = r0 + i;
      if (row < rows) {
        const int64_t atom = a0 + row;
        const float s = p.inv_degree[atom];
        const float4 hA = *reinterpret_cast<const float4*>(p.h_in + atom * MP_F + c0);
        const float4 hB = *reinterpret_cast<const float4*>(p.h_in + atom * MP_F + c0 + 64);
        float4 oA, oB;
        if (p.raw) {
          oA = make_float4(acc[i][0] * s, acc[i][1] * s, acc[i][2] * s, acc[i][3] * s);
          oB = make_float4(acc[i][4] * s, acc[i][5] * s, acc[i][6] * s, acc[i][7] * s);
        } else {
          oA.x = apply_act(acc[i][0] * s, p.act) + hA.x;
          oA.y = apply_act(acc[i][1] * s, p.act) + hA.y;
          oA.z = apply_act(acc[i][2] * s, p.act) + hA.z;
          oA.w = apply_act(acc[i][3] * s, p.act) + hA.w;
          oB.x = apply_act(acc[i][4] * s, p.act) + hB.x;
          oB.y = apply_act(acc[i][5] * s, p.act) + hB.y;
          oB.z = apply_act(acc[i][6] * s, p.act) + hB.z;
          oB.w = apply_act(acc[i][7] * s, p.act) + hB.w;
        }
        *reinterpret_cast<float4*>(p.h_out + atom * MP_F + c0) = oA;
        *reinterpret_cast<float4*>(p.h_out + atom * MP_F + c0 + 64) = oB;
      }
    }
  }
}

// ----------------------------------------------------------------------------------
// Node MLP + readout, F = 256.  One CTA = 128 atoms; the node tile stays in shared
// memory across the residual Dense layers; only peaks (and optionally the F/2-wide
// FCBlock output) are written.
// ----------------------------------------------------------------------------------
struct FcArgs {
  const float* nodes;        // [n_atoms, 256]
  const float* atoms;        // [n_atoms, C]
  float* peaks;              // [n_atoms]
  float* fc_nodes;           // optional [n_atoms, 128]
  int64_t n_atoms;
  int C;
  const float* W[MAX_DENSE]; // residual layers [256,256]; last [256,128]
  const float* b[MAX_DENSE];
  int n_layers;
  int act;
  const float* Wo;           // [128, C]
  const float* bo;           // [C]
  const float* peak_std;     // [C]
  const float* peak_avg;     // [C]
};

constexpr int FC_F = 256;
constexpr int FC_F2 = 128;
constexpr int FC_LDX = FC_F + 2;
constexpr int FC_LDZ = FC_F2 + 2;
constexpr int FC_THREADS = 512;
__host__ __device__ constexpr size_t fc_smem_bytes() {
  return sizeof(float) * (128 * FC_LDX + TileGemm<256, FC_THREADS>::SMEM_FLOATS);
}

__global__ void __launch_bounds__(FC_THREADS, 1) fc_readout_ffma_kernel(const FcArgs p) {
  using G = TileGemm<256, FC_THREADS>;
  using G2 = TileGemm<128, FC_THREADS>;
  extern __shared__ __align__(16) float smem[];
  float* X = smem;                  // [128][FC_LDX]; later Z [128][FC_LDZ]
  float* Bs = X + 128 * FC_LDX;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int r0, c0;
  G::thread_origin(tid, r0, c0);
  const int64_t n_tiles = (p.n_atoms + 127) / 128;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t a0 = tile * 128;
    const int rows = (int)min((int64_t)128, p.n_atoms - a0);
    __syncthreads();
    for (int i = tid; i < 128 * (FC_F / 4); i += FC_THREADS) {
      const int row = i / (FC_F / 4), c4 = i % (FC_F / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < rows) v = *reinterpret_cast<const float4*>(p.nodes + (a0 + row) * FC_F + c4 * 4);
      float* xr = X + row * FC_LDX + c4 * 4;
      *reinterpret_cast<float2*>(xr) = make_float2(v.x, v.y);
      *reinterpret_cast<float2*>(xr + 2) = make_float2(v.z, v.w);
    }
    // residual layers: X <- act(X @ W + b) + X
    for (int l = 0; l + 1 < p.n_layers; ++l) {
      float acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
      G::run(acc, X, FC_LDX, p.W[l], FC_F, FC_F, Bs, true);
      const float4 bA = *reinterpret_cast<const float4*>(p.b[l] + c0);
      const float4 bB = *reinterpret_cast<const float4*>(p.b[l] + c0 + 64);
      const float bias[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float* xr = X + (r0 + i) * FC_LDX + c0;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float old = xr[(j & 3) + (j >> 2) * 64];
          v[j] = apply_act(acc[i][j] + bias[j], p.act) + old;
        }
        *reinterpret_cast<float2*>(xr) = make_float2(v[0], v[1]);
        *reinterpret_cast<float2*>(xr + 2) = make_float2(v[2], v[3]);
        *reinterpret_cast<float2*>(xr + 64) = make_float2(v[4], v[5]);
        *reinterpret_cast<float2*>(xr + 66) = make_float2(v[6], v[7]);
      }
      // (the next run() barriers before any thread reads X again)
    }
    // last layer: Z = act(X @ W[256x128] + b); only the 8 warps covering columns 0..127 compute
    {
      float acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
      const int l = p.n_layers - 1;
      const bool active = (warp >> 3) == 0;
      G2::run(acc, X, FC_LDX, p.W[l], FC_F2, FC_F, Bs, active);
      // run() ended with a barrier: X is free, reuse it as Z [128][FC_LDZ]
      if (active) {
        const float4 bA = *reinterpret_cast<const float4*>(p.b[l] + c0);
        const float4 bB = *reinterpret_cast<const float4*>(p.b[l] + c0 + 64);
        const float bias[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float* zr = X + (r0 + i) * FC_LDZ + c0;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = apply_act(acc[i][j] + bias[j], p.act);
          *reinterpret_cast<float2*>(zr) = make_float2(v[0], v[1]);
          *reinterpret_cast<float2*>(zr + 2) = make_float2(v[2], v[3]);
          *reinterpret_cast<float2*>(zr + 64) = make_float2(v[4], v[5]);
          *reinterpret_cast<float2*>(zr + 66) = make_float2(v[6], v[7]);
        }
      }
      __syncthreads();
    }
    if (p.fc_nodes != nullptr) {
      for (int i = tid; i < rows * FC_F2; i += FC_THREADS) {
        const int row = i / FC_F2, k = i % FC_F2;
        p.fc_nodes[(a0 + row) * FC_F2 + k] = X[row * FC_LDZ + k];
      }
    }
    // readout: peaks = sum_c (z @ Wo + bo)[c] * a[c] * std[c] + a[c] * avg[c]; warp per atom,
    // skipping classes with a[c] == 0 (exact for finite activations)
    for (int i = 0; i < 8; ++i) {
      const int row = warp * 8 + i;
      if (row >= rows) break;
      const int64_t atom = a0 + row;
      const float* zr = X + row * FC_LDZ;
      float peak = 0.0f;
      for (int c = 0; c < p.C; ++c) {
        const float a = p.atoms[atom * p.C + c];
        if (a != 0.0f) {
          float dot = 0.0f;
#pragma unroll
          for (int k = lane; k < FC_F2; k += 32) dot = fmaf(zr[k], __ldg(p.Wo + k * p.C + c), dot);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
          const float full = dot + p.bo[c];
          peak += full * a * p.peak_std[c] + a * p.peak_avg[c];
        }
      }
      if (lane == 0) p.peaks[atom] = peak;
    }
  }
}

}  // namespace nmr
