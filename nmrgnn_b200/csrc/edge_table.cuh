// Edge block as a create-time table (option "edge_table", default on for smooth activations).
//
// RBFExpansion -> mask -> EdgeFCBlock -> mask (nmrgnn/model.py:251-261, nmrgnn/layers.py:126-140,
// nmrgnn/model.py:118-138) is a function of ONE scalar per edge: f(d) in R^E for d > 0, exactly 0 for d <= 0.
// The weights are fixed at nmrgnn_create time, so f is tabulated once per model:
//   * nodes d_i = i * h, h = 2^-13 nm, i = -1 .. n+1, evaluated in FP64 on the device (edge_table_nodes_kernel:
//     RBF with the reference's float32 grid, Dense stack, exact activation);
//   * per interval [d_i, d_i+1) and channel the cubic through nodes i-1 .. i+2 in the local coordinate
//     u = (d - d_i) / h, stored as fp32 monomial coefficients {a0, a1, a2, a3} (a0 = f(d_i));
//   * beyond the last interval (d >= n h, chosen where every RBF has underflowed in float32) f is the constant
//     f_inf = EdgeFC(0).
// Because h is a power of two, i = floor(d / h) and u are computed EXACTLY in fp32 (i h is exact and d - i h is a
// Sterbenz difference), so the table adds no error in the argument; the cubic's interpolation error is measured at
// create time against the FP64 evaluation at every interval midpoint (nmrgnn_edge_table_info) and the table is
// rejected (the MLP kernels run instead) if it exceeds 2^-27 of the feature scale.  The result is closer to the
// exact function than an fp32 evaluation of the MLP (error ~1 ulp of the value instead of the rounding of
// three 128-long fp32 dot products).  The exact tcgen05 / FFMA edge kernels stay selectable ("edge_table" = 0).
#pragma once
#include "common.cuh"
#include "kernels_ffma.cuh"
#include "kernels_tc.cuh"   // rec_slot

namespace nmr {

constexpr float EDGE_TAB_H = 1.0f / 8192.0f;      // 2^-13 nm
constexpr float EDGE_TAB_INV_H = 8192.0f;
constexpr int EDGE_TAB_MAX_E = 4;

struct EdgeTabBuildArgs {
  const float* W[MAX_DENSE];   // hidden: [H,H]; last: [H,E]   (fp32 originals)
  const float* b[MAX_DENSE];
  const float* centers;        // [H] float32 RBF grid
  float gap;
  int n_layers;
  int H;
  int E;
  int act;
  int n_nodes;                 // nodes -1 .. n_nodes-2  ->  d = (node - 1) * h ; last node = "infinity" (all RBF = 0)
  double* nodes;               // [n_nodes][E]
  double* mids;                // optional [n_nodes][E]: f at d + h/2 (interpolation check)
};

__device__ __forceinline__ double act_f64(double x, int act) {
  switch (act) {
    case ACT_SOFTPLUS: return fmax(x, 0.0) + log1p(exp(-fabs(x)));
    case ACT_RELU: return fmax(x, 0.0);
    case ACT_TANH: return tanh(x);
    default: return x;
  }
}

// one block per node, H threads (H <= 256): x lives in shared memory, thread o owns output feature o
__global__ void __launch_bounds__(256) edge_table_nodes_kernel(const EdgeTabBuildArgs p) {
  __shared__ double xa[256], xb[256];
  const int node = blockIdx.x >> 1, mid = blockIdx.x & 1;
  if (mid && p.mids == nullptr) return;
  const int o = threadIdx.x, H = p.H;
  const bool inf = node == p.n_nodes - 1;
  const double d = ((double)(node - 1) + (mid ? 0.5 : 0.0)) * (double)EDGE_TAB_H;
  if (o < H) {
    // the reference evaluates the grid and the gap in float32 (layers.py:126-129); d itself is exact here
    const double diff = d - (double)p.centers[o];
    xa[o] = inf ? 0.0 : exp(-(diff * diff) / (double)p.gap);
  }
  __syncthreads();
  double* x = xa;
  double* y = xb;
  for (int l = 0; l + 1 < p.n_layers; ++l) {
    if (o < H) {
      double s = (double)p.b[l][o];
      for (int k = 0; k < H; ++k) s = fma(x[k], (double)p.W[l][(size_t)k * H + o], s);
      y[o] = act_f64(s, p.act);
    }
    __syncthreads();
    double* t = x;
    x = y;
    y = t;
  }
  if (o < p.E) {
    const int l = p.n_layers - 1;
    double s = (double)p.b[l][o];
    for (int k = 0; k < H; ++k) s = fma(x[k], (double)p.W[l][(size_t)k * p.E + o], s);
    (mid ? p.mids : p.nodes)[(size_t)node * p.E + o] = s;
  }
}

// interval i (d in [i h, (i+1) h)) uses nodes i-1 .. i+2  ->  array entries i .. i+3
// cubic through (-1, y0), (0, y1), (1, y2), (2, y3):  a0 = y1,
//   a1 = -y0/3 - y1/2 + y2 - y3/6,  a2 = (y0 + y2)/2 - y1,  a3 = (y3 - y0)/6 + (y1 - y2)/2
__global__ void edge_table_coef_kernel(const double* __restrict__ nodes, const double* __restrict__ mids, int n_intervals,
                                       int E, float4* __restrict__ tab, double* __restrict__ max_err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_intervals) return;
  for (int c = 0; c < E; ++c) {
    const double y0 = nodes[(size_t)i * E + c], y1 = nodes[(size_t)(i + 1) * E + c], y2 = nodes[(size_t)(i + 2) * E + c],
                 y3 = nodes[(size_t)(i + 3) * E + c];
    const double a1 = -y0 / 3.0 - y1 / 2.0 + y2 - y3 / 6.0, a2 = (y0 + y2) / 2.0 - y1, a3 = (y3 - y0) / 6.0 + (y1 - y2) / 2.0;
    const float4 q = make_float4((float)y1, (float)a1, (float)a2, (float)a3);
    tab[(size_t)i * E + c] = q;
    if (mids != nullptr) {
      // interpolation error proper: the cubic (FP64 coefficients) at u = 1/2 against the FP64 function value there
      // (rounding a0 to fp32 costs another <= 1/2 ulp of the value, like any fp32 result)
      const double v = y1 + 0.5 * (a1 + 0.5 * (a2 + 0.5 * a3));
      const double e = fabs(v - mids[(size_t)(i + 1) * E + c]);
      // (positive doubles order like their bit patterns)
      atomicMax(reinterpret_cast<unsigned long long*>(max_err + c), (unsigned long long)__double_as_longlong(e));
      atomicMax(reinterpret_cast<unsigned long long*>(max_err + EDGE_TAB_MAX_E + c),
                (unsigned long long)__double_as_longlong(fabs(y1)));
    }
  }
}

struct EdgeTabArgs {
  const float* edges;        // [n_edges] distances (nm)
  const int32_t* nlist;      // optional [n_edges]: validated against n_atoms (and packed into rec)
  float* out;                // optional [n_edges, E]
  float4* rec;               // optional [n_edges] {e0, e1, e2, bits(idx)} records for the tensor-core MP kernel (E <= 3)
  int64_t n_edges;
  int64_t n_atoms;
  const float4* tab;         // [n_intervals][E] {a0, a1, a2, a3}
  int n_intervals;
  float f_inf[EDGE_TAB_MAX_E];
  int E;
  int* err_flag;
  int rec_k;                 // K (8 or 16) if the records are slot-swizzled (rec_slot), else 0
  int64_t rec_e0;            // edge index of this launch's first edge within the whole batch (chunked launches)
};

// One thread per edge: 8 B in, 16 B (record) and/or 4 E B out; the table (n_intervals * E * 16 B ~ 250 KB) lives in L2 / L1.
__global__ void __launch_bounds__(256) edge_table_kernel(const EdgeTabArgs p) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= p.n_edges) return;
  const float d = p.edges[e];
  float f[EDGE_TAB_MAX_E] = {0.f, 0.f, 0.f, 0.f};
  if (d > 0.0f) {                                   // model.py:251: padded slots (d <= 0, also NaN) contribute exactly 0
    const float t = d * EDGE_TAB_INV_H;             // exact (power of two)
    if (t < (float)p.n_intervals) {
      const int i = (int)t;                         // floor, t >= 0
      const float u = t - (float)i;                 // exact: both are multiples of ulp(t) and |t - i| < 1
      const float4* q = p.tab + (size_t)i * p.E;
#pragma unroll
      for (int c = 0; c < EDGE_TAB_MAX_E; ++c)
        if (c < p.E) {
          const float4 a = __ldg(q + c);
          f[c] = fmaf(fmaf(fmaf(a.w, u, a.z), u, a.y), u, a.x);
        }
    } else {
#pragma unroll
      for (int c = 0; c < EDGE_TAB_MAX_E; ++c) f[c] = p.f_inf[c];
    }
  }
  int32_t idx = 0;
  if (p.nlist != nullptr) {
    idx = p.nlist[e];
    if (idx < 0 || idx >= p.n_atoms) {
      atomicOr(p.err_flag, 1);
      idx = 0;
    }
  }
  if (p.out != nullptr)
    for (int c = 0; c < p.E; ++c) p.out[e * p.E + c] = f[c];
  if (p.rec != nullptr) {
    const int64_t ge = p.rec_e0 + e;
    p.rec[(p.rec_k ? rec_slot(ge, p.rec_k) : ge) - p.rec_e0] = make_float4(f[0], f[1], f[2], __int_as_float(idx));
  }
}

}  // namespace nmr
