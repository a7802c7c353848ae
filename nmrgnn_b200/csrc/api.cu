// C ABI of libnmrgnn_b200.so (see include/nmrgnn_b200.h).  Host-side orchestration:
// weight upload/re-packing, workspace, kernel launches, error reporting.
#include "../../include/nmrgnn_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "kernels_ffma.cuh"
#include "kernels_generic.cuh"
#include "kernels_tc.cuh"
#include "kernels_mp_nsplit.cuh"
#include "kernels_fc_pipe.cuh"
#include "knn.cuh"
#include "edge_table.cuh"
#include "peer_gather.cuh"

using namespace nmr;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

}  // namespace

struct nmrgnn_handle {
  nmrgnn_dims d{};
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // host-buffer calls: H2D of later chunks overlaps the edge kernel of earlier ones
  cudaEvent_t ev_copy[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // The workspaces below belong to the handle: calls must be stream-ordered.  An asynchronous call leaves an event
  // behind; a call on another stream waits for it before touching the workspaces (begin_call / end_call).
  cudaEvent_t ev_last = nullptr;
  cudaStream_t last_stream = nullptr;
  bool last_async = false;
  std::string err;
  std::string path = "ffma";
  int64_t launches = 0;

  // device weights (owned)
  std::vector<float*> owned;
  std::vector<const float*> edge_W, edge_b, fc_W, fc_b;
  const float* embed = nullptr;
  std::vector<const float*> mp_Wp;  // packed for the FFMA kernel
  std::vector<const float*> mp_W;   // original [F,F,E] layout (generic-geometry kernels)
  DevBuf genA, genB;                // generic-geometry activations
  const float *out_W = nullptr, *out_b = nullptr, *peak_std = nullptr, *peak_avg = nullptr;
  const float* centers = nullptr;
  float gap = 0.f;

  // workspace (grow-only)
  DevBuf atoms, nlist, edges, invdeg, efeat, hA, hB, peaks, tmp_in, tmp_out;
  int* err_flag = nullptr;        // device
  int* err_flag_host = nullptr;   // pinned
  bool fast_path = false;         // F=256, H=128, E<=4: tiled kernels available
  bool force_ffma = false;
  DevBuf pos, offs, knn_sorted, knn_cells, knn_grid;   // kNN builder: positions, offsets, cell-list workspaces
  bool knn_cells_on = true;             // option "knn_cells": cell-list search (default) / brute force
  bool knn_warp = true;                 // option "knn_warp": cell-list query by one warp per atom (default) / 8 threads
  std::vector<int64_t> offs_cached;     // graph_offsets currently resident in `offs` (skips the upload when unchanged:
                                        // a frame stream repeats the same offsets, and the call stays graph-capturable)
  // tensor-core path (F=256, H=128, E<=8): pre-split, pre-swizzled operand images
  bool tc_ok = false;
  const uint8_t* edge_img = nullptr;    // [n_hidden][4][hi 8192 | lo 8192]   fp16 split
  const uint8_t* edge_f_img = nullptr;  // [4][hi 1024 | lo 1024]
  const float* edge_bias = nullptr;     // [n_hidden][128]
  float edge_in_scale[MAX_DENSE + 1];   // a-priori power-of-two range scaling per edge layer
  float edge_out_scale[MAX_DENSE + 1];
  bool profile = false;                 // record CUDA events around the stages of nmrgnn_forward
  std::vector<cudaEvent_t> ev;          // [n_mp + 4] stage boundaries of the last profiled forward
  bool ev_valid = false;
  bool mp_tc_ok = false;                // MP layer on tensor cores (F=256, E<=3; K<=16 checked per call)
  std::vector<const uint8_t*> mp_img;   // per layer: [8 passes][E][hi 16384 | lo 16384]
  std::vector<std::vector<float>> mp_w_host;   // originals, kept to re-pack the images when a compensation option changes
  std::vector<std::vector<float>> edge_w_host; // the edge MLP's hidden-layer weights, likewise
  float edge_pos_c = 0.3f;                // measured optimum for the edge MLP (profiles/r02_parity.md)
  std::vector<std::vector<float>> fc_w_host;   // the node MLP's weights, for the same reason
  float fc_pos_c = 0.5f;                // the same for the node MLP's 16-instruction chains (pack_fc_images)
  // single-accumulator form of the MP layers (kernels_tc.cuh mp_layer_tc1_kernel): its own images, scales and constants
  bool mp_one = false;                  // option "mp_single_acc"
  std::vector<const uint8_t*> mp_img1;  // per layer: W' x 2^s pre-scaled, lo image NOT scaled by 2^11
  std::vector<float> mp_wscale_inv;     // per layer 2^-s
  std::vector<float> mp_corr1;          // residual constants of this form (calibrate_mp)
  float mp_pos_c1 = 0.55f;              // slope of its position-dependent compensation (144-instruction chains)
  float mp_pos_c = 0.5f;                // position-dependent compensation slope c' (x 2^-24), see pack_mp_images
  DevBuf rec, hmaxA, hmaxB;
  // round-toward-zero compensation of the tcgen05 accumulation (DESIGN.md "Accumulation model"):
  // multiplicative factors 1 + c applied in the epilogues
  std::vector<float> mp_corr;           // per MP layer, calibrated against the FFMA kernels at create time
  float edge_rz = 1.0f;                 // edge MLP layers (8-instruction chains): analytic constant
  bool fc_tc_ok = false;                // node MLP + readout on tensor cores (F = 256)
  const uint8_t* fc_img = nullptr;      // per layer [8 chunks][2 halves][hi 8192 | lo 8192] (last layer: one half)
  const float* fc_bias = nullptr;       // [n_fc][256]
  float fc_gain[MAX_DENSE], fc_offs[MAX_DENSE];
  float fc_rz = 1.0f;
  // layer-pipelined single-accumulator form of the node MLP (kernels_fc_pipe.cuh): its own images, scales and bounds
  bool fc_pipe = true;                  // option "fc_pipe"
  bool fc_pair = false;                 // option "fc_pair": the pipelined kernel on CTA pairs (cta_group::2)
  const uint8_t* fc_img1p = nullptr;    // its images: per layer [2 ranks][8 chunks][hi 8192 | lo 8192] (last: [hi 4096 | lo 4096])
  const uint8_t* fc_img1 = nullptr;     // per layer [8 chunks][hi 16384 | lo 16384] (last: [hi 8192 | lo 8192]), W x 2^s, lo unscaled
  float fc_wsinv[MAX_DENSE];            // per layer 2^-s
  float fc_g1[MAX_DENSE], fc_o1[MAX_DENSE];   // one-layer growth bound: max|x_{l+1}| <= g max|x_l| + o
  float fc_pos_c1 = 0.5f;               // slope of its position-dependent compensation (48-instruction chains)
  long long* fc_dbg = nullptr;          // diagnostics: per-CTA role cycle counters of the last pipelined node-MLP launch
  bool compensate = true;
  long long* mp_dbg = nullptr;          // diagnostics: per-CTA role cycle counters of the last MP launch
  int64_t tc_min_atoms = 1024;          // calls smaller than this run on the exact-FP32 kernels
  bool mp_l1_prefetch = true;           // option "mp_l1_prefetch"
  bool mp_small_tiles = true;           // option "mp_small_tiles": calls of less than one wave run on 32 / 64 / 96-atom tiles
  bool mp_nsplit = false;               // option "mp_nsplit": MP layers by column-split CTA pairs (kernels_mp_nsplit.cuh)
  int mp_nseg = 1;                      // option "mp_chain_segments": accumulation chains per MP tile (kernels_tc.cuh)
  // edge block as a create-time FP64 table of the scalar function d -> EdgeFC(RBF(d)) (edge_table.cuh)
  bool edge_table = true;               // option "edge_table"
  bool edge_tab_ok = false;             // table built and its interpolation error accepted
  const float4* edge_tab = nullptr;     // [edge_tab_n][E] cubic coefficients
  int edge_tab_n = 0;
  float edge_tab_finf[EDGE_TAB_MAX_E] = {0.f, 0.f, 0.f, 0.f};
  double edge_tab_err = 0.0;            // max interpolation error at interval midpoints / feature scale
  // multi-GPU reassembly over peer memory (peer_gather.cuh)
  int comm_rank = -1, comm_world = 0;
  int64_t comm_capacity = 0;            // floats per (parity, rank) slot, multiple of 4
  float* comm_gbuf = nullptr;           // local gather buffer [2][world][capacity]
  uint32_t* comm_flags = nullptr;       // local flag words [world]
  unsigned int* comm_done = nullptr;    // device counter of the scatter kernel
  float* comm_peer_gbuf[PEER_MAX_WORLD] = {};
  uint32_t* comm_peer_flags[PEER_MAX_WORLD] = {};
  uint32_t comm_epoch = 0;
  bool comm_ready = false;
};

namespace {

int fail(nmrgnn_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  else g_create_error = buf;
  return code;
}

#define CUDA_TRY(h, expr)                                                                      \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return fail((h), _e == cudaErrorMemoryAllocation ? NMRGNN_ERR_OOM : NMRGNN_ERR_CUDA,     \
                  "%s failed: %s", #expr, cudaGetErrorString(_e));                             \
  } while (0)

int ensure(nmrgnn_handle* h, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return NMRGNN_OK;
  if (b.p) CUDA_TRY(h, cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  CUDA_TRY(h, cudaMalloc(&b.p, want));
  b.cap = want;
  return NMRGNN_OK;
}

int upload(nmrgnn_handle* h, const float* host, size_t n, const float** out) {
  float* d = nullptr;
  CUDA_TRY(h, cudaMalloc(&d, n * sizeof(float) + 16));
  h->owned.push_back(d);
  CUDA_TRY(h, cudaMemcpy(d, host, n * sizeof(float), cudaMemcpyHostToDevice));
  *out = d;
  return NMRGNN_OK;
}

int upload_bytes(nmrgnn_handle* h, const void* host, size_t bytes, const uint8_t** out) {
  void* d = nullptr;
  CUDA_TRY(h, cudaMalloc(&d, bytes + 16));
  h->owned.push_back((float*)d);
  CUDA_TRY(h, cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
  *out = (const uint8_t*)d;
  return NMRGNN_OK;
}

// fp32 -> tf32 (round to nearest, ties away, like cvt.rna.tf32.f32) kept in fp32 layout
float tf32_hi(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) != 0x7F800000u) u += 0x1000u;
  u &= 0xFFFFE000u;
  float r;
  std::memcpy(&r, &u, 4);
  return r;
}

// B-operand images for the tcgen05 kernels.  W is [K][ldw] row-major (in -> out); the operand
// tile has one 64-byte swizzled row per output feature n0..n0+rows_tile-1 (rows >= rows_valid are
// zero) and 16 consecutive k per chunk.  Layout: [chunk][hi tile | lo tile], hi = tf32(w), lo = w - hi.
void pack_sw64(const float* W, int K, int ldw, int n0, int rows_valid, int rows_tile, std::vector<uint8_t>& out) {
  const int chunks = K / tc::BK;
  const size_t tile = (size_t)rows_tile * 64;
  out.assign((size_t)chunks * 2 * tile, 0);
  for (int c = 0; c < chunks; ++c)
    for (int n = 0; n < rows_valid; ++n)
      for (int kk = 0; kk < tc::BK; ++kk) {
        const float w = W[(size_t)(c * tc::BK + kk) * ldw + n0 + n];
        const float hi = tf32_hi(w), lo = w - hi;
        const size_t off = tc::sw64_offset(n, kk);
        std::memcpy(out.data() + (size_t)c * 2 * tile + off, &hi, 4);
        std::memcpy(out.data() + (size_t)c * 2 * tile + tile + off, &lo, 4);
      }
}

// fp16 scaled-split B-operand images (tc_common.cuh "fp16x3").  W is [K][ldw] row-major (in -> out);
// one 64-byte swizzled row per output feature, 32 consecutive k per chunk.
// Layout: [chunk][hi tile | lo tile], hi = fp16(w), lo = fp16((w - hi) * 2^11).
// `get(k, n)` returns the weight for contraction index k and output feature n.
// `gain(k)` (optional) is a relative weight correction g_k << 2^-11 folded into the lo image only:
// hi = fp16(w), lo = fp16((w (1 + g_k) - hi) * 2^11) -- the position-dependent compensation of the MP layers.
template <typename Get, typename Gain>
void pack_sw64_f16(Get get, Gain gain, int K, int rows_valid, int rows_tile, std::vector<uint8_t>& out,
                   double w_scale = 1.0, double lo_scale = (double)tc::LO_SCALE) {
  const int chunks = K / tc::HK;
  const size_t tile = (size_t)rows_tile * 64;
  out.assign((size_t)chunks * 2 * tile, 0);
  for (int c = 0; c < chunks; ++c)
    for (int n = 0; n < rows_valid; ++n)
      for (int kk = 0; kk < tc::HK; ++kk) {
        const double ws = (double)get(c * tc::HK + kk, n) * w_scale;      // (w_scale is a power of two: exact)
        const __half hi = __float2half_rn((float)ws);
        const double wg = ws * (1.0 + gain(c * tc::HK + kk));
        const __half lo = __float2half_rn((float)((wg - (double)__half2float(hi)) * lo_scale));
        const size_t off = (size_t)n * 64 + ((((size_t)kk >> 3) ^ (((size_t)n >> 1) & 3)) << 4) + (((size_t)kk & 7) << 1);
        std::memcpy(out.data() + (size_t)c * 2 * tile + off, &hi, 2);
        std::memcpy(out.data() + (size_t)c * 2 * tile + tile + off, &lo, 2);
      }
}

template <typename Get>
void pack_sw64_f16(Get get, int K, int rows_valid, int rows_tile, std::vector<uint8_t>& out) {
  pack_sw64_f16(get, [](int) { return 0.0; }, K, rows_valid, rows_tile, out);
}

// upper bound of |act(x)| for |x| <= b
float act_bound(float b, int act) {
  switch (act) {
    case ACT_SOFTPLUS: return b + 0.6931472f;
    case ACT_TANH: return 1.0f;
    default: return b;
  }
}
// smallest s >= 0 with bound * 2^-s <= 2^15 (fp16 operands stay finite with 2x margin)
int scale_exp_for(float bound) {
  int s = 0;
  while (bound > 32768.0f && s < 120) {
    bound *= 0.5f;
    ++s;
  }
  return s;
}
float max_abs(const float* w, size_t n) {
  float m = 0.f;
  for (size_t i = 0; i < n; ++i) m = std::fmax(m, std::fabs(w[i]));
  return m;
}
// max over output features n of sum_k |W[k][n]|, W [K][N] row-major
float max_col_abs_sum(const float* W, int K, int N) {
  std::vector<double> s(N, 0.0);
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) s[n] += std::fabs((double)W[(size_t)k * N + n]);
  double m = 0;
  for (double v : s) m = std::max(m, v);
  return (float)m;
}

// float32 grid exactly as tf.linspace evaluates it for float32 inputs:
// start + step*i with a separately rounded multiply and add; last point = stop.
void rbf_grid(float lo, float hi, int n, std::vector<float>& c, float& gap) {
  c.resize(n);
  if (n == 1) {
    c[0] = lo;
    gap = 0.f;
    return;
  }
  volatile float step = (hi - lo) / (float)(n - 1);
  for (int i = 0; i < n; ++i) {
    volatile float prod = step * (float)i;
    volatile float s = lo + prod;
    c[i] = s;
  }
  c[n - 1] = hi;
  volatile float g = c[1] - c[0];
  gap = g;
}

bool dims_ok(const nmrgnn_dims& d) {
  return d.num_elem >= 1 && d.num_elem <= 1024 && d.atom_features >= 2 && d.atom_features % 2 == 0 &&
         d.edge_features >= 1 && d.edge_hidden >= 1 && d.n_edge_fc >= 2 && d.n_edge_fc <= MAX_DENSE &&
         d.n_mp >= 0 && d.n_mp <= 64 && d.n_fc >= 1 && d.n_fc <= MAX_DENSE && d.mp_activation >= 0 &&
         d.mp_activation <= 3 && d.fc_activation >= 0 && d.fc_activation <= 3;
}

struct Io {  // resolves caller buffers to device pointers for one call
  nmrgnn_handle* h;
  int mem;
  cudaStream_t s;
  int in(const void* src, size_t bytes, DevBuf& stage, const void** out) {
    if (mem == NMRGNN_MEM_DEVICE || bytes == 0) {
      *out = src;
      return NMRGNN_OK;
    }
    int rc = ensure(h, stage, bytes);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(stage.p, src, bytes, cudaMemcpyHostToDevice, s));
    *out = stage.p;
    return NMRGNN_OK;
  }
  int out_buf(void* dst, size_t bytes, DevBuf& stage, void** out) {
    if (mem == NMRGNN_MEM_DEVICE) {
      *out = dst;
      return NMRGNN_OK;
    }
    int rc = ensure(h, stage, bytes);
    if (rc) return rc;
    *out = stage.p;
    return NMRGNN_OK;
  }
  int finish(void* dst, const void* dev, size_t bytes) {
    if (mem == NMRGNN_MEM_DEVICE || bytes == 0) return NMRGNN_OK;
    CUDA_TRY(h, cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, s));
    return NMRGNN_OK;
  }
};

int begin_call(nmrgnn_handle* h, int mem, void* stream, cudaStream_t* s) {
  if (!h) return NMRGNN_ERR_BAD_DIMS;
  if (mem != NMRGNN_MEM_HOST && mem != NMRGNN_MEM_DEVICE) return fail(h, NMRGNN_ERR_BAD_DIMS, "bad mem flag %d", mem);
  if (stream != nullptr && mem == NMRGNN_MEM_HOST)
    return fail(h, NMRGNN_ERR_BAD_DIMS, "host buffers require stream == NULL (synchronous call)");
  CUDA_TRY(h, cudaSetDevice(h->device));
  *s = stream ? (cudaStream_t)stream : h->stream;
  if (h->last_async && h->last_stream != *s) {   // the previous call may still be using the workspaces on its stream
    CUDA_TRY(h, cudaStreamWaitEvent(*s, h->ev_last, 0));
    h->last_async = false;
  }
  return NMRGNN_OK;
}

int end_call(nmrgnn_handle* h, void* stream, cudaStream_t s) {
  CUDA_TRY(h, cudaGetLastError());
  if (stream != nullptr) {                  // asynchronous: caller synchronises
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CUDA_TRY(h, cudaStreamIsCapturing(s, &cap));
    if (cap == cudaStreamCaptureStatusNone) {
      if (!h->ev_last) CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_last, cudaEventDisableTiming));
      CUDA_TRY(h, cudaEventRecord(h->ev_last, s));
      h->last_stream = s;
      h->last_async = true;
    }
    return NMRGNN_OK;
  }
  h->last_async = false;
  return nmrgnn_synchronize(h, nullptr);
}

// name of the compute path a full-size call takes with the current options (nmrgnn_compute_path)
void update_path(nmrgnn_handle* h) {
  const bool table = h->edge_table && h->edge_tab_ok;
  std::string p;
  if (!h->fast_path) p = "generic-fp32";
  else if (!h->tc_ok || h->force_ffma) p = "ffma";
  else {
    std::string blocks = table ? "" : "edge";
    if (h->mp_tc_ok) blocks += blocks.empty() ? "mp" : ",mp";
    if (h->fc_tc_ok) blocks += blocks.empty() ? "fc" : ",fc";
    p = "tcgen05-fp16x3(" + blocks + ")";
    if (!h->mp_tc_ok || !h->fc_tc_ok) p += "+ffma";
  }
  if (table) p = "edge-table-f64+" + p;
  h->path = p;
}

int grid_for(const nmrgnn_handle* h, int64_t tiles, int per_sm) {
  int64_t g = (int64_t)h->num_sms * per_sm;
  return (int)(tiles < g ? (tiles < 1 ? 1 : tiles) : g);
}

// the tensor-core kernels are instantiated per activation (common.cuh act_t)
#define ACT_DISPATCH(act, kernel, grid, threads, smem, stream, args)                         \
  do {                                                                                       \
    switch (act) {                                                                           \
      case ACT_SOFTPLUS: kernel<ACT_SOFTPLUS><<<grid, threads, smem, stream>>>(args); break; \
      case ACT_RELU: kernel<ACT_RELU><<<grid, threads, smem, stream>>>(args); break;         \
      case ACT_TANH: kernel<ACT_TANH><<<grid, threads, smem, stream>>>(args); break;         \
      default: kernel<ACT_LINEAR><<<grid, threads, smem, stream>>>(args); break;             \
    }                                                                                        \
  } while (0)
#define ACT_SET_SMEM(kernel, bytes)                                                                                  \
  do {                                                                                                               \
    CUDA_RC(cudaFuncSetAttribute(kernel<ACT_LINEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));    \
    CUDA_RC(cudaFuncSetAttribute(kernel<ACT_SOFTPLUS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));  \
    CUDA_RC(cudaFuncSetAttribute(kernel<ACT_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));      \
    CUDA_RC(cudaFuncSetAttribute(kernel<ACT_TANH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));      \
  } while (0)

// ------------------------------------------------------------------ launches
// generic-geometry route of the edge block: RBF tensor materialised, one dense kernel per layer
int launch_edge_generic(nmrgnn_handle* h, cudaStream_t s, const float* edges, int64_t n_edges, float* out,
                        const int32_t* nlist, int64_t n_atoms) {
  const int H = h->d.edge_hidden, E = h->d.edge_features, L = h->d.n_edge_fc;
  int rc;
  if (nlist != nullptr) {
    gen_index_check_kernel<<<(unsigned)((n_edges + 255) / 256), 256, 0, s>>>(nlist, n_edges, n_atoms, h->err_flag);
    h->launches++;
  }
  // process in slabs so that the [edges, H] activations stay below ~1 GB each
  const int64_t slab = std::max<int64_t>(1, ((int64_t)1 << 28) / std::max(H, 1));
  if ((rc = ensure(h, h->genA, (size_t)std::min(slab, n_edges) * H * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->genB, (size_t)std::min(slab, n_edges) * H * sizeof(float)))) return rc;
  for (int64_t e0 = 0; e0 < n_edges; e0 += slab) {
    const int64_t ne = std::min(slab, n_edges - e0);
    float* a = (float*)h->genA.p;
    float* b = (float*)h->genB.p;
    gen_rbf_kernel<<<(unsigned)((ne * H + 255) / 256), 256, 0, s>>>(edges + e0, h->centers, h->gap, a, ne, H);
    h->launches++;
    for (int l = 0; l < L; ++l) {
      const bool last = l == L - 1;
      gen_dense_kernel<<<(unsigned)ne, 256, 0, s>>>(a, h->edge_W[l], h->edge_b[l], nullptr, last ? edges + e0 : nullptr,
                                                    last ? out + e0 * E : b, ne, H, last ? E : H,
                                                    last ? ACT_LINEAR : h->d.fc_activation);
      h->launches++;
      std::swap(a, b);
    }
  }
  return NMRGNN_OK;
}

// ---------------------------------------------------------------------------- edge table (edge_table.cuh)
// Tabulates d -> EdgeFC(RBF(d)) in FP64 on the device, derives the per-interval cubics and accepts the table if the
// interpolation error measured at every interval midpoint stays below 2^-27 of the feature scale.
int build_edge_table(nmrgnn_handle* h) {
  const int H = h->d.edge_hidden, E = h->d.edge_features, act = h->d.fc_activation;
  h->edge_tab_ok = false;
  if (H > 256 || E > EDGE_TAB_MAX_E || act == ACT_RELU) return NMRGNN_OK;   // relu: f has kinks, a cubic does not follow them
  if (!(h->gap > 0.f)) return NMRGNN_OK;
  // beyond d_max every RBF has underflowed to zero in float32 (exp(-104) < 2^-149): f is the constant EdgeFC(0)
  const double d_max = (double)h->d.rbf_high + std::sqrt(104.0 * (double)h->gap);
  const double n_d = std::ceil(d_max * (double)EDGE_TAB_INV_H) + 1.0;
  if (!(n_d >= 1.0) || n_d > 1048576.0) return NMRGNN_OK;
  const int n = (int)n_d, n_nodes = n + 4;
  double *nodes = nullptr, *mids = nullptr, *err = nullptr;
  float4* tab = nullptr;
  auto cleanup = [&]() {
    if (nodes) cudaFree(nodes);
    if (mids) cudaFree(mids);
    if (err) cudaFree(err);
  };
  cudaStream_t s = h->stream;
  if (cudaMalloc(&nodes, (size_t)n_nodes * E * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&mids, (size_t)n_nodes * E * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&err, 2 * EDGE_TAB_MAX_E * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&tab, (size_t)n * E * sizeof(float4) + 16) != cudaSuccess) {
    cleanup();
    if (tab) cudaFree(tab);
    return fail(h, NMRGNN_ERR_OOM, "edge table allocation failed");
  }
  h->owned.push_back((float*)tab);
  EdgeTabBuildArgs a{};
  for (int i = 0; i < h->d.n_edge_fc; ++i) {
    a.W[i] = h->edge_W[i];
    a.b[i] = h->edge_b[i];
  }
  a.centers = h->centers;
  a.gap = h->gap;
  a.n_layers = h->d.n_edge_fc;
  a.H = H;
  a.E = E;
  a.act = act;
  a.n_nodes = n_nodes;
  a.nodes = nodes;
  a.mids = mids;
  cudaMemsetAsync(err, 0, 2 * EDGE_TAB_MAX_E * sizeof(double), s);
  edge_table_nodes_kernel<<<2 * n_nodes, 256, 0, s>>>(a);
  edge_table_coef_kernel<<<(n + 127) / 128, 128, 0, s>>>(nodes, mids, n, E, tab, err);
  double herr[2 * EDGE_TAB_MAX_E];
  std::vector<double> finf(E);
  cudaError_t e1 = cudaMemcpyAsync(herr, err, sizeof(herr), cudaMemcpyDeviceToHost, s);
  cudaError_t e2 = cudaMemcpyAsync(finf.data(), nodes + (size_t)(n_nodes - 1) * E, E * sizeof(double), cudaMemcpyDeviceToHost, s);
  cudaError_t e3 = cudaStreamSynchronize(s);
  cleanup();
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || cudaGetLastError() != cudaSuccess)
    return fail(h, NMRGNN_ERR_CUDA, "edge table build failed: %s", cudaGetErrorString(e3 != cudaSuccess ? e3 : e1));
  double rel = 0.0;
  for (int c = 0; c < E; ++c) rel = std::max(rel, herr[EDGE_TAB_MAX_E + c] > 0.0 ? herr[c] / herr[EDGE_TAB_MAX_E + c] : 0.0);
  h->edge_tab_err = rel;
  for (int c = 0; c < E; ++c) h->edge_tab_finf[c] = (float)finf[c];
  h->edge_tab = tab;
  h->edge_tab_n = n;
  h->edge_tab_ok = std::isfinite(rel) && rel <= 1.0 / 134217728.0;   // 2^-27
  return NMRGNN_OK;
}

bool rec_swizzled(int K) { return K == 8 || K == 16; }

int launch_edge_table(nmrgnn_handle* h, cudaStream_t s, const float* edges, int64_t n_edges, float* out, const int32_t* nlist,
                      int64_t n_atoms, float4* rec, int K, int64_t e0) {
  EdgeTabArgs t{};
  t.edges = edges;
  t.nlist = nlist;
  t.out = out;
  t.rec = rec;
  t.n_edges = n_edges;
  t.n_atoms = n_atoms;
  t.tab = h->edge_tab;
  t.n_intervals = h->edge_tab_n;
  for (int c = 0; c < EDGE_TAB_MAX_E; ++c) t.f_inf[c] = h->edge_tab_finf[c];
  t.E = h->d.edge_features;
  t.err_flag = h->err_flag;
  t.rec_k = (rec != nullptr && rec_swizzled(K)) ? K : 0;
  t.rec_e0 = e0;
  edge_table_kernel<<<(unsigned)((n_edges + 255) / 256), 256, 0, s>>>(t);
  h->launches++;
  return NMRGNN_OK;
}

int launch_edge(nmrgnn_handle* h, cudaStream_t s, const float* edges, int64_t n_edges, float* out,
                const int32_t* nlist, int64_t n_atoms, float4* rec = nullptr, int K = 0, int64_t e0 = 0) {
  if (n_edges == 0) return NMRGNN_OK;
  if (h->edge_table && h->edge_tab_ok && (rec == nullptr || h->d.edge_features <= 3))
    return launch_edge_table(h, s, edges, n_edges, out, nlist, n_atoms, rec, K, e0);
  if (!h->fast_path) return launch_edge_generic(h, s, edges, n_edges, out, nlist, n_atoms);
  if (h->tc_ok && !h->force_ffma) {
    EdgeTcArgs t{};
    t.edges = edges;
    t.out = out;
    t.rec = rec;
    t.n_edges = n_edges;
    for (int i = 0; i <= MAX_DENSE; ++i) {
      t.in_scale[i] = h->edge_in_scale[i];
      t.out_scale[i] = h->edge_out_scale[i];
    }
    t.centers = h->centers;
    t.gap = h->gap;
    t.rbf_c = (float)(-1.4426950408889634 / (double)h->gap);
    t.Wimg = h->edge_img;
    t.Wfimg = h->edge_f_img;
    t.bias = h->edge_bias;
    t.bias_f = h->edge_b[h->d.n_edge_fc - 1];
    t.Wf = h->edge_W[h->d.n_edge_fc - 1];
    t.n_hidden = h->d.n_edge_fc - 1;
    t.E = h->d.edge_features;
    t.act = h->d.fc_activation;
    t.nlist = nlist;
    t.n_atoms = n_atoms;
    t.err_flag = h->err_flag;
    t.dbg = h->mp_dbg;
    t.rec_k = (rec != nullptr && rec_swizzled(K)) ? K : 0;
    t.rec_e0 = e0;
    const int64_t tiles = (n_edges + 127) / 128;
    ACT_DISPATCH(t.act, edge_mlp_tc_kernel, grid_for(h, tiles, 1), ETC_THREADS, ETC_SMEM, s, t);
    h->launches++;
    return NMRGNN_OK;
  }
  EdgeArgs a{};
  a.edges = edges;
  a.out = out;
  a.n_edges = n_edges;
  a.centers = h->centers;
  a.gap = h->gap;
  a.n_layers = h->d.n_edge_fc;
  a.act = h->d.fc_activation;
  for (int i = 0; i < h->d.n_edge_fc; ++i) {
    a.W[i] = h->edge_W[i];
    a.b[i] = h->edge_b[i];
  }
  a.nlist = nlist;
  a.n_atoms = n_atoms;
  a.err_flag = h->err_flag;
  const int64_t tiles = (n_edges + 127) / 128;
  const int grid = grid_for(h, tiles, 2);
#define EDGE_CASE(EE)                                                                              \
  case EE:                                                                                         \
    edge_mlp_ffma_kernel<EE><<<grid, EDGE_THREADS, edge_smem_bytes<EE>(), s>>>(a);                 \
    break;
  switch (h->d.edge_features) {
    EDGE_CASE(1) EDGE_CASE(2) EDGE_CASE(3) EDGE_CASE(4)
    default: return fail(h, NMRGNN_ERR_BAD_DIMS, "edge_features %d unsupported", h->d.edge_features);
  }
#undef EDGE_CASE
  h->launches++;
  return NMRGNN_OK;
}

int launch_embed(nmrgnn_handle* h, cudaStream_t s, const float* atoms, int64_t n, float* nodes, float* hmax = nullptr) {
  if (n == 0) return NMRGNN_OK;
  const int C = h->d.num_elem, F = h->d.atom_features;
  if (F == 256 && (size_t)EMBED256_ATOMS * C * sizeof(float) <= 40 * 1024) {
    const int64_t blocks = (n + EMBED256_ATOMS - 1) / EMBED256_ATOMS;
    embed256_kernel<<<(unsigned)blocks, 256, EMBED256_ATOMS * C * sizeof(float), s>>>(atoms, h->embed, nodes, hmax, n, C);
    h->launches++;
    return NMRGNN_OK;
  }
  if (hmax != nullptr) return fail(h, NMRGNN_ERR_BAD_DIMS, "fused row maximum needs F = 256");
  const int64_t blocks = (n + EMBED_ATOMS - 1) / EMBED_ATOMS;
  embed_kernel<<<(unsigned)blocks, 256, EMBED_ATOMS * C * sizeof(float), s>>>(atoms, h->embed, nodes, n, C, F);
  h->launches++;
  return NMRGNN_OK;
}

int launch_mp(nmrgnn_handle* h, cudaStream_t s, int layer, const float* h_in, const int32_t* nlist,
              const float* efeat, const float* invdeg, int64_t n, int K, float* h_out, int raw = 0) {
  if (n == 0) return NMRGNN_OK;
  if (!h->fast_path) {
    const int F = h->d.atom_features, E = h->d.edge_features;
    if ((size_t)F * E * sizeof(float) > 200 * 1024) return fail(h, NMRGNN_ERR_BAD_DIMS, "atom_features * edge_features too large");
    CUDA_TRY(h, cudaFuncSetAttribute(gen_mp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(F * E * sizeof(float))));
    gen_mp_kernel<<<(unsigned)n, 256, F * E * sizeof(float), s>>>(h_in, nlist, efeat, invdeg, h->mp_W[layer], h_out, n, K, F, E,
                                                                h->d.mp_activation);
    h->launches++;
    return NMRGNN_OK;
  }
  MpArgs a{};
  a.h_in = h_in;
  a.h_out = h_out;
  a.nlist = nlist;
  a.efeat = efeat;
  a.inv_degree = invdeg;
  a.Wp = h->mp_Wp[layer];
  a.n_atoms = n;
  a.K = K;
  a.act = h->d.mp_activation;
  a.raw = raw;
  const int64_t tiles = (n + 127) / 128;
  const int grid = grid_for(h, tiles, 1);
#define MP_CASE(EE)                                                                                \
  case EE: {                                                                                       \
    const size_t smem = mp_smem_bytes<EE>(K);                                                      \
    if (smem > 227 * 1024) return fail(h, NMRGNN_ERR_BAD_DIMS, "neighbor_number %d too large", K); \
    CUDA_TRY(h, cudaFuncSetAttribute(mp_layer_ffma_kernel<EE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    mp_layer_ffma_kernel<EE><<<grid, MP_THREADS, smem, s>>>(a);                                    \
  } break;
  switch (h->d.edge_features) {
    MP_CASE(1) MP_CASE(2) MP_CASE(3) MP_CASE(4)
    default: return fail(h, NMRGNN_ERR_BAD_DIMS, "edge_features %d unsupported", h->d.edge_features);
  }
#undef MP_CASE
  h->launches++;
  return NMRGNN_OK;
}

bool mp_tc_usable(const nmrgnn_handle* h, int K) {
  return h->tc_ok && h->mp_tc_ok && !h->force_ffma && K >= 1 && K <= MTC_KMAX;
}

// Path policy of a whole call: below `tc_min_atoms` (a handful of 128-row tiles: latency-bound either way,
// 1-2 ms) the exact-FP32 kernels are used and carry no split/accumulation error; a 2 482-atom protein already
// runs 1.9x faster on the tensor-core kernels (tools/bench_md_stream.py).
struct PathScope {
  nmrgnn_handle* h;
  bool saved;
  PathScope(nmrgnn_handle* hh, int64_t n_atoms) : h(hh), saved(hh->force_ffma) {
    if (n_atoms < h->tc_min_atoms) h->force_ffma = true;
  }
  ~PathScope() { h->force_ffma = saved; }
};

int launch_absmax(nmrgnn_handle* h, cudaStream_t s, const float* nodes, int64_t n, float* hmax) {
  if (n == 0) return NMRGNN_OK;
  row_absmax256_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(nodes, hmax, n);
  h->launches++;
  return NMRGNN_OK;
}

int launch_pack_rec(nmrgnn_handle* h, cudaStream_t s, const int32_t* nlist, const float* efeat, float4* rec,
                    int64_t n_edges, int64_t n_atoms, int K) {
  if (n_edges == 0) return NMRGNN_OK;
  pack_edge_records_kernel<<<(unsigned)((n_edges + 255) / 256), 256, 0, s>>>(nlist, efeat, rec, n_edges,
                                                                              h->d.edge_features, n_atoms, h->err_flag,
                                                                              rec_swizzled(K) ? K : 0);
  h->launches++;
  return NMRGNN_OK;
}

// hmax_in: one row maximum per atom, or (in_pair) two partial maxima per atom as the column-split kernel writes them.
// Returns in *out_pair (optional) which of the two forms hmax_out has.
inline bool in_pair_unsupported(const nmrgnn_handle* h) { return h->num_sms < 2; }

int launch_mp_tc(nmrgnn_handle* h, cudaStream_t s, int layer, const float* h_in, const float* hmax_in,
                 const float4* rec, const float* invdeg, int64_t n, int K, float* h_out, float* hmax_out,
                 int raw = 0, int in_pair = 0, int* out_pair = nullptr) {
  if (n == 0) return NMRGNN_OK;
  MpTcArgs a{};
  a.h_in = h_in;
  a.hmax_in = hmax_in;
  a.h_out = h_out;
  a.hmax_out = hmax_out;
  a.rec = rec;
  a.inv_degree = invdeg;
  const bool one = h->mp_one && h->mp_nseg == 1 && !h->mp_nsplit;
  a.Wimg = one ? h->mp_img1[layer] : h->mp_img[layer];
  a.n_atoms = n;
  a.K = K;
  a.E = h->d.edge_features;
  a.act = h->d.mp_activation;
  a.corr = (h->compensate && !raw) ? (one ? h->mp_corr1[layer] : h->mp_corr[layer]) : 1.0f;
  if (one) a.corr *= h->mp_wscale_inv[layer];        // exact: a power of two
  a.raw = raw;
  a.swz = rec_swizzled(K) ? 1 : 0;
  a.nseg = h->mp_nseg;
  a.hmax_pair = in_pair;
  a.dbg = h->mp_dbg;
  a.l1_prefetch = h->mp_l1_prefetch ? 1 : 0;
  const int64_t tiles = (n + 127) / 128;
  const bool nsplit = h->mp_nsplit && a.nseg == 1 && !in_pair_unsupported(h);
  if (out_pair) *out_pair = nsplit ? 1 : 0;
  if (nsplit) {
    const int64_t pairs = std::min<int64_t>(tiles, h->num_sms / 2);
    ACT_DISPATCH(a.act, mp_layer_np_kernel, (unsigned)(2 * pairs), MNP_THREADS, MNP_SMEM, s, a);
  } else if (one) {
    ACT_DISPATCH(a.act, mp_layer_tc1_kernel, grid_for(h, tiles, 1), MTC_THREADS, MTC_SMEM, s, a);
  } else if (a.nseg > 1) ACT_DISPATCH(a.act, mp_layer_tc_seg_kernel, grid_for(h, tiles, 1), MTC_THREADS, MTC_SMEM, s, a);
  else {
    // less than one wave of 128-atom tiles: smaller tiles on more SMs (a tile's time follows its 32-row steps down to the
    // MMA + drain floor), e.g. one 2 482-atom protein = 78 tiles of 32 atoms instead of 20 of 128
    const int64_t per_sm = (n + h->num_sms - 1) / h->num_sms;
    const int64_t vt_rows = std::max<int64_t>(32, (per_sm + 31) / 32 * 32);
    if (h->mp_small_tiles && K > 8 && tiles < h->num_sms && vt_rows < 128) {
      a.vt_rows = (int)vt_rows;
      a.vt_tiles = (n + vt_rows - 1) / vt_rows;
      ACT_DISPATCH(a.act, mp_layer_tc_vt_kernel, grid_for(h, a.vt_tiles, 1), MTC_THREADS, MTC_SMEM, s, a);
    } else {
      ACT_DISPATCH(a.act, mp_layer_tc_kernel, grid_for(h, tiles, 1), MTC_THREADS, MTC_SMEM, s, a);
    }
  }
  h->launches++;
  return NMRGNN_OK;
}

// hmax: max |nodes row| per atom if the caller has it ([n][2] partial maxima if hmax_pair), else NULL
int launch_fc(nmrgnn_handle* h, cudaStream_t s, const float* nodes, const float* atoms, int64_t n, float* peaks,
              float* fc_nodes, const float* hmax = nullptr, int hmax_pair = 0) {
  if (n == 0) return NMRGNN_OK;
  if (!h->fast_path) {
    const int F = h->d.atom_features, F2 = F / 2, C = h->d.num_elem, L = h->d.n_fc;
    int rc;
    if ((rc = ensure(h, h->genA, (size_t)n * F * sizeof(float)))) return rc;
    if ((rc = ensure(h, h->genB, (size_t)n * F * sizeof(float)))) return rc;
    const float* x = nodes;
    float* a = (float*)h->genA.p;
    float* b = (float*)h->genB.p;
    for (int l = 0; l < L; ++l) {
      const bool last = l == L - 1;
      float* y = (last && fc_nodes != nullptr) ? fc_nodes : a;
      gen_dense_kernel<<<(unsigned)n, 256, 0, s>>>(x, h->fc_W[l], h->fc_b[l], last ? nullptr : x, nullptr, y, n, F,
                                                   last ? F2 : F, h->d.fc_activation);
      h->launches++;
      x = y;
      std::swap(a, b);
    }
    gen_readout_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, atoms, h->out_W, h->out_b, h->peak_std, h->peak_avg,
                                                                  peaks, n, F2, C);
    h->launches++;
    return NMRGNN_OK;
  }
  if (h->fc_tc_ok && !h->force_ffma && h->fc_pipe && h->fc_img1 && h->d.num_elem <= FPI_CMAX) {
    if (hmax == nullptr) {
      if (int rc = ensure(h, h->hmaxA, 2 * (size_t)n * sizeof(float))) return rc;
      if (int rc = launch_absmax(h, s, nodes, n, (float*)h->hmaxA.p)) return rc;
      hmax = (const float*)h->hmaxA.p;
      hmax_pair = 0;
    }
    FcPipeArgs t{};
    t.nodes = nodes;
    t.hmax = hmax;
    t.hmax_pair = hmax_pair;
    t.atoms = atoms;
    t.peaks = peaks;
    t.fc_nodes = fc_nodes;
    t.n_atoms = n;
    t.C = h->d.num_elem;
    t.Wimg = h->fc_img1;
    t.bias = h->fc_bias;
    for (int i = 0; i < h->d.n_fc; ++i) {
      t.g[i] = h->fc_g1[i];
      t.o[i] = h->fc_o1[i];
      t.wsinv[i] = h->fc_wsinv[i];
    }
    t.n_layers = h->d.n_fc;
    t.act = h->d.fc_activation;
    t.corr = h->compensate ? h->fc_rz : 1.0f;
    t.Wo = h->out_W;
    t.bo = h->out_b;
    t.peak_std = h->peak_std;
    t.peak_avg = h->peak_avg;
    t.dbg = h->fc_dbg;
    const int64_t tiles = (n + 127) / 128;
    if (h->fc_pair && h->fc_img1p) {
      t.Wimg = h->fc_img1p;
      const int64_t pairs = std::min<int64_t>((tiles + 1) / 2, h->num_sms / 2);
      ACT_DISPATCH(t.act, fc_readout_pipe_pair_kernel, (unsigned)(2 * pairs), FPI_THREADS, FPI_SMEM_PAIR, s, t);
    } else {
      ACT_DISPATCH(t.act, fc_readout_pipe_kernel, grid_for(h, tiles, 1), FPI_THREADS, FPI_SMEM, s, t);
    }
    h->launches++;
    return NMRGNN_OK;
  }
  if (h->fc_tc_ok && !h->force_ffma) {
    FcTcArgs t{};
    t.nodes = nodes;
    t.atoms = atoms;
    t.peaks = peaks;
    t.fc_nodes = fc_nodes;
    t.n_atoms = n;
    t.C = h->d.num_elem;
    t.Wimg = h->fc_img;
    t.bias = h->fc_bias;
    for (int i = 0; i < h->d.n_fc; ++i) {
      t.gain[i] = h->fc_gain[i];
      t.offs[i] = h->fc_offs[i];
    }
    t.n_layers = h->d.n_fc;
    t.act = h->d.fc_activation;
    t.corr = h->compensate ? h->fc_rz : 1.0f;
    t.Wo = h->out_W;
    t.bo = h->out_b;
    t.peak_std = h->peak_std;
    t.peak_avg = h->peak_avg;
    const int64_t tiles = (n + 127) / 128;
    ACT_DISPATCH(t.act, fc_readout_tc_kernel, grid_for(h, tiles, 1), FTC_THREADS, FTC_SMEM, s, t);
    h->launches++;
    return NMRGNN_OK;
  }
  FcArgs a{};
  a.nodes = nodes;
  a.atoms = atoms;
  a.peaks = peaks;
  a.fc_nodes = fc_nodes;
  a.n_atoms = n;
  a.C = h->d.num_elem;
  a.n_layers = h->d.n_fc;
  a.act = h->d.fc_activation;
  for (int i = 0; i < h->d.n_fc; ++i) {
    a.W[i] = h->fc_W[i];
    a.b[i] = h->fc_b[i];
  }
  a.Wo = h->out_W;
  a.bo = h->out_b;
  a.peak_std = h->peak_std;
  a.peak_avg = h->peak_avg;
  const int64_t tiles = (n + 127) / 128;
  fc_readout_ffma_kernel<<<grid_for(h, tiles, 1), FC_THREADS, fc_smem_bytes(), s>>>(a);
  h->launches++;
  return NMRGNN_OK;
}

// ---------------------------------------------------------------------------- calibration
// tcgen05 accumulates with round-toward-zero (2 guard bits), which shrinks a K-long contraction by a
// factor (1 - c) with c ~ 0.4 * (K/16 + 1) * 2^-24 plus a data-dependent part.  For the
// MP layers (48-instruction chains) c is measured once per model: a deterministic synthetic
// protein-like graph is pushed through the exact-FP32 FFMA kernels; at every layer the raw
// contraction D is computed by both paths on identical inputs and c = -<D_tc - D_ffma, D_ffma>_w /
// <D_ffma, D_ffma>_w, weighted by the squared activation slope.  The epilogue then multiplies by 1 + c.
// (Re)builds the MP layers' W' operand images.  Contraction index kg = (pass*E + n)*32 + ll  <->  input feature
// 32*pass + ll, edge channel n; one tcgen05 instruction covers 16 consecutive kg.
// Position-dependent compensation: the accumulator is truncated toward zero once per instruction, so a product that
// enters at instruction k of an n-instruction chain is shrunk n - k times more than the last one; to first order the
// expected loss of the chain is c' * sum_k (n - k + 1) P_k (P_k: the instruction's exact sum; measured c' = 0.5 x 2^-24,
// tools/sim_tc_accum.py) -- a LINEAR functional of the products, unlike the truncation itself.  It is folded into the
// weights: w_k (1 + c' (n - k + 1)); the factor is far below fp16 resolution, so only the lo image changes and the run
// time is untouched.  What a constant factor cannot follow (partial sums that overshoot the result) this does.
int pack_mp_images(nmrgnn_handle* h) {
  const int F = h->d.atom_features, E = h->d.edge_features, L = h->d.n_mp;
  const int n_instr = F * E / 16, chain = n_instr / h->mp_nseg;
  const double cpos = h->compensate ? (double)h->mp_pos_c / 16777216.0 : 0.0;
  const bool fresh = h->mp_img.empty();
  if (fresh) h->mp_img.assign(L, nullptr);
  std::vector<uint8_t> img;
  for (int l = 0; l < L; ++l) {
    const float* src = h->mp_w_host[l].data();   // w[l_in, m, n]
    pack_sw64_f16(
        [&](int kg, int m) {
          const int qch = kg / 32, ll = kg % 32, ps = qch / E, n = qch % E;
          return src[((size_t)(32 * ps + ll) * F + m) * E + n];
        },
        [&](int kg) { return cpos * (double)(chain - (kg / 16) % chain); }, F * E, F, F, img);
    if (fresh) {
      if (int rc = upload_bytes(h, img.data(), img.size(), &h->mp_img[l])) return rc;
    } else {
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      CUDA_TRY(h, cudaMemcpy(const_cast<uint8_t*>(h->mp_img[l]), img.data(), img.size(), cudaMemcpyHostToDevice));
    }
  }
  // ---- single-accumulator form: every one of the 6 instructions of a (pass, n) chunk -- main ks0, lo*hi ks0, main ks1,
  // lo*hi ks1, hi*lo ks0, hi*lo ks1 -- truncates the one accumulator, so a main product at position p of the 6*8*E
  // instructions is truncated 6*8*E - p times.  W' is pre-scaled by 2^s (s from max|w|: the unscaled lo image must stay
  // in the normal fp16 range); 2^-s goes into the epilogue's output scale.
  const bool fresh1 = h->mp_img1.empty();
  if (fresh1) {
    h->mp_img1.assign(L, nullptr);
    h->mp_wscale_inv.assign(L, 1.0f);
  }
  const int n_instr1 = 6 * (F / 32) * E;
  const double cpos1 = h->compensate ? (double)h->mp_pos_c1 / 16777216.0 : 0.0;
  for (int l = 0; l < L; ++l) {
    const float* src = h->mp_w_host[l].data();
    const float wmax = max_abs(src, (size_t)F * F * E);
    int sexp = 0;
    while (sexp < 24 && wmax * std::ldexp(1.0f, sexp + 1) <= 32768.0f) ++sexp;
    h->mp_wscale_inv[l] = std::ldexp(1.0f, -sexp);
    pack_sw64_f16(
        [&](int kg, int m) {
          const int qch = kg / 32, ll = kg % 32, ps = qch / E, n = qch % E;
          return src[((size_t)(32 * ps + ll) * F + m) * E + n];
        },
        [&](int kg) { return cpos1 * (double)(n_instr1 - (6 * (kg / 32) + 2 * ((kg % 32) / 16))); }, F * E, F, F, img,
        std::ldexp(1.0, sexp), 1.0);
    if (fresh1) {
      if (int rc = upload_bytes(h, img.data(), img.size(), &h->mp_img1[l])) return rc;
    } else {
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      CUDA_TRY(h, cudaMemcpy(const_cast<uint8_t*>(h->mp_img1[l]), img.data(), img.size(), cudaMemcpyHostToDevice));
    }
  }
  return NMRGNN_OK;
}

// (Re)builds the edge MLP's hidden-layer operand images [layer][4 chunks][hi 8192 | lo 8192] with the position-dependent
// compensation of their H/16-instruction chains (see pack_mp_images).
int pack_edge_images(nmrgnn_handle* h) {
  const int H = h->d.edge_hidden, n_instr = H / 16;
  const double cpos = h->compensate ? (double)h->edge_pos_c / 16777216.0 : 0.0;
  std::vector<uint8_t> img, all;
  for (const auto& Wv : h->edge_w_host) {
    const float* W = Wv.data();
    pack_sw64_f16([&](int k, int n) { return W[(size_t)k * H + n]; }, [&](int k) { return cpos * (double)(n_instr - k / 16); },
                  H, H, H, img);
    all.insert(all.end(), img.begin(), img.end());
  }
  if (!h->edge_img) return upload_bytes(h, all.data(), all.size(), &h->edge_img);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  CUDA_TRY(h, cudaMemcpy(const_cast<uint8_t*>(h->edge_img), all.data(), all.size(), cudaMemcpyHostToDevice));
  return NMRGNN_OK;
}

// (Re)builds the node MLP's operand images: per layer [8 chunks][2 halves][hi 8192 | lo 8192] (last layer: one half),
// with the position-dependent compensation of its 16-instruction chains in the lo images (see pack_mp_images).
int pack_fc_images(nmrgnn_handle* h) {
  const int F = h->d.atom_features, n_fc = h->d.n_fc;
  const int n_instr = F / 16;
  const double cpos = h->compensate ? (double)h->fc_pos_c / 16777216.0 : 0.0;
  std::vector<uint8_t> all;
  for (int i = 0; i < n_fc; ++i) {
    const bool last = (i == n_fc - 1);
    const int outw = last ? F / 2 : F;
    const float* W = h->fc_w_host[i].data();
    // pack each 128-column half as its own 8-chunk image, then interleave
    std::vector<uint8_t> half_img[2];
    const int halves = last ? 1 : 2;
    for (int hf = 0; hf < halves; ++hf)
      pack_sw64_f16([&](int k, int n) { return W[(size_t)k * outw + hf * 128 + n]; },
                    [&](int k) { return cpos * (double)(n_instr - k / 16); }, F, 128, 128, half_img[hf]);
    for (int c = 0; c < 8; ++c)
      for (int hf = 0; hf < halves; ++hf)
        all.insert(all.end(), half_img[hf].begin() + (size_t)c * 16384, half_img[hf].begin() + (size_t)(c + 1) * 16384);
  }
  // ---- layer-pipelined single-accumulator form (kernels_fc_pipe.cuh): the 6 instructions of a 32-feature chunk -- main
  // ks0, lo*hi ks0, main ks1, lo*hi ks1, hi*lo ks0, hi*lo ks1 -- all truncate the one accumulator, so a main product at
  // position p of the 6 F/32 instructions is truncated 6 F/32 - p times;
  // W is pre-scaled by 2^s (s from max|w|: the unscaled lo image must stay in the normal fp16 range), 2^-s goes into
  // the epilogue's output scale.  Every layer's image takes 16 slots of 16 KB (the last one uses 8).
  {
    const double cpos1 = h->compensate ? (double)h->fc_pos_c1 / 16777216.0 : 0.0;
    const int n_instr1 = 3 * (F / 16);
    std::vector<uint8_t> all1((size_t)n_fc * 16 * 16384, 0), all1p((size_t)n_fc * 16 * 16384, 0), img;
    for (int i = 0; i < n_fc; ++i) {
      const bool last = (i == n_fc - 1);
      const int outw = last ? F / 2 : F;
      const float* W = h->fc_w_host[i].data();
      const float wmax = max_abs(W, (size_t)F * outw);
      int sexp = 0;
      while (sexp < 24 && wmax * std::ldexp(1.0f, sexp + 1) <= 32768.0f) ++sexp;
      h->fc_wsinv[i] = std::ldexp(1.0f, -sexp);
      pack_sw64_f16([&](int k, int n) { return W[(size_t)k * outw + n]; },
                    [&](int k) { return cpos1 * (double)(n_instr1 - (6 * (k / 32) + 2 * ((k % 32) / 16))); }, F, outw, outw, img,
                    std::ldexp(1.0, sexp),
                    1.0);
      std::memcpy(all1.data() + (size_t)i * 16 * 16384, img.data(), img.size());
      // CTA-pair form: CTA r of a pair stages N rows [r outw/2, (r + 1) outw/2) of every tile
      for (int r = 0; r < 2; ++r) {
        const int half = outw / 2;
        pack_sw64_f16([&](int k, int n) { return W[(size_t)k * outw + r * half + n]; },
                      [&](int k) { return cpos1 * (double)(n_instr1 - (6 * (k / 32) + 2 * ((k % 32) / 16))); }, F, half, half, img,
                      std::ldexp(1.0, sexp), 1.0);
        std::memcpy(all1p.data() + (size_t)i * 16 * 16384 + (size_t)r * img.size(), img.data(), img.size());
      }
    }
    if (!h->fc_img1p) {
      if (int rc = upload_bytes(h, all1p.data(), all1p.size(), &h->fc_img1p)) return rc;
    } else {
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      CUDA_TRY(h, cudaMemcpy(const_cast<uint8_t*>(h->fc_img1p), all1p.data(), all1p.size(), cudaMemcpyHostToDevice));
    }
    if (!h->fc_img1) {
      if (int rc = upload_bytes(h, all1.data(), all1.size(), &h->fc_img1)) return rc;
    } else {
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      CUDA_TRY(h, cudaMemcpy(const_cast<uint8_t*>(h->fc_img1), all1.data(), all1.size(), cudaMemcpyHostToDevice));
    }
  }
  if (!h->fc_img) return upload_bytes(h, all.data(), all.size(), &h->fc_img);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  CUDA_TRY(h, cudaMemcpy(const_cast<uint8_t*>(h->fc_img), all.data(), all.size(), cudaMemcpyHostToDevice));
  return NMRGNN_OK;
}

int calibrate_mp(nmrgnn_handle* h) {
  const int N = 2048, K = 16;
  const int C = h->d.num_elem, F = h->d.atom_features, E = h->d.edge_features;
  cudaStream_t s = h->stream;
  std::vector<float> atoms((size_t)N * C, 0.f), edges((size_t)N * K), inv(N, 1.0f / K);
  std::vector<int32_t> nl((size_t)N * K);
  uint64_t st = 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() {  // splitmix64 -> [0,1)
    st += 0x9E3779B97F4A7C15ull;
    uint64_t z = st;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
  };
  for (int i = 0; i < N; ++i) {
    // element mix of a protein with explicit hydrogens, mapped onto the first columns that exist
    const double u = rnd();
    int col = u < 0.505 ? 4 : u < 0.824 ? 3 : u < 0.912 ? 2 : 5;
    atoms[(size_t)i * C + (col % C)] = 1.0f;
    for (int j = 0; j < K; ++j) {
      int o = 1 + (int)(rnd() * 48.0);
      if (rnd() < 0.5) o = -o;
      int t = i + o;
      if (t < 0) t = i + (o < 0 ? -o : o);
      if (t >= N) t = i - (o < 0 ? -o : o);
      nl[(size_t)i * K + j] = t;
      edges[(size_t)i * K + j] = rnd() < 0.02 ? 0.0f : (float)(0.09 + 0.26 * rnd());
      if (edges[(size_t)i * K + j] == 0.0f) nl[(size_t)i * K + j] = 0;
    }
  }
  int rc;
  if ((rc = ensure(h, h->atoms, atoms.size() * 4))) return rc;
  if ((rc = ensure(h, h->nlist, nl.size() * 4))) return rc;
  if ((rc = ensure(h, h->edges, edges.size() * 4))) return rc;
  if ((rc = ensure(h, h->invdeg, inv.size() * 4))) return rc;
  if ((rc = ensure(h, h->efeat, (size_t)N * K * E * 4))) return rc;
  if ((rc = ensure(h, h->rec, (size_t)N * K * sizeof(float4)))) return rc;
  if ((rc = ensure(h, h->hA, (size_t)N * F * 4))) return rc;
  if ((rc = ensure(h, h->hB, (size_t)N * F * 4))) return rc;
  if ((rc = ensure(h, h->tmp_in, (size_t)N * F * 4))) return rc;
  if ((rc = ensure(h, h->tmp_out, (size_t)N * F * 4))) return rc;
  if ((rc = ensure(h, h->hmaxA, (size_t)N * 8))) return rc;
  if ((rc = ensure(h, h->hmaxB, (size_t)N * 8))) return rc;
  CUDA_TRY(h, cudaMemcpyAsync(h->atoms.p, atoms.data(), atoms.size() * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(h->nlist.p, nl.data(), nl.size() * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(h->edges.p, edges.data(), edges.size() * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(h->invdeg.p, inv.data(), inv.size() * 4, cudaMemcpyHostToDevice, s));
  const bool saved = h->force_ffma;
  h->force_ffma = true;  // exact-FP32 edge features for both arms
  rc = launch_edge(h, s, (const float*)h->edges.p, (int64_t)N * K, (float*)h->efeat.p, (const int32_t*)h->nlist.p, N);
  h->force_ffma = saved;
  if (rc) return rc;
  if ((rc = launch_pack_rec(h, s, (const int32_t*)h->nlist.p, (const float*)h->efeat.p, (float4*)h->rec.p,
                            (int64_t)N * K, N, K)))
    return rc;
  std::vector<float> df((size_t)N * F), dt((size_t)N * F);
  const bool saved_one = h->mp_one, saved_nsplit = h->mp_nsplit;
  const int saved_nseg = h->mp_nseg;
  h->mp_corr1.assign(h->d.n_mp, 1.0f);
  struct Restore {
    nmrgnn_handle* h;
    bool one, nsplit;
    int nseg;
    ~Restore() {
      h->mp_one = one;
      h->mp_nsplit = nsplit;
      h->mp_nseg = nseg;
    }
  } restore{h, saved_one, saved_nsplit, saved_nseg};
  // two passes: the two-accumulator kernel (with the current chain segmentation), then the single-accumulator kernel
  for (int variant = 0; variant < 2; ++variant) {
  h->mp_one = variant == 1;
  h->mp_nsplit = false;
  if (variant == 1) h->mp_nseg = 1;
  float* ha = (float*)h->hA.p;
  float* hb = (float*)h->hB.p;
  if ((rc = launch_embed(h, s, (const float*)h->atoms.p, N, ha))) return rc;
  for (int l = 0; l < h->d.n_mp; ++l) {
    if ((rc = launch_mp(h, s, l, ha, (const int32_t*)h->nlist.p, (const float*)h->efeat.p, (const float*)h->invdeg.p, N,
                        K, (float*)h->tmp_in.p, 1)))
      return rc;
    if ((rc = launch_absmax(h, s, ha, N, (float*)h->hmaxA.p))) return rc;
    if ((rc = launch_mp_tc(h, s, l, ha, (const float*)h->hmaxA.p, (const float4*)h->rec.p, (const float*)h->invdeg.p, N,
                           K, (float*)h->tmp_out.p, (float*)h->hmaxB.p, 1)))
      return rc;
    CUDA_TRY(h, cudaMemcpyAsync(df.data(), h->tmp_in.p, df.size() * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, cudaMemcpyAsync(dt.data(), h->tmp_out.p, dt.size() * 4, cudaMemcpyDeviceToHost, s));
    if ((rc = launch_mp(h, s, l, ha, (const int32_t*)h->nlist.p, (const float*)h->efeat.p, (const float*)h->invdeg.p, N,
                        K, hb)))
      return rc;
    CUDA_TRY(h, cudaStreamSynchronize(s));
    // least squares weighted by the squared slope of the activation at r = inv_degree * D: what matters is
    // the error after the activation (softplus flattens the large negative entries that dominate |D|)
    const int act = h->d.mp_activation;
    double num = 0.0, den = 0.0;
    for (size_t i = 0; i < df.size(); ++i) {
      const double r = (double)df[i];
      double g = 1.0;
      if (act == ACT_SOFTPLUS) g = 1.0 / (1.0 + std::exp(-r));
      else if (act == ACT_RELU) g = r > 0.0 ? 1.0 : 0.0;
      else if (act == ACT_TANH) g = 1.0 - std::tanh(r) * std::tanh(r);
      const double w = g * g;
      num += w * ((double)dt[i] - r) * r;
      den += w * r * r;
    }
    // (the residual after the position-dependent part of the W' images; it may have either sign)
    double c = den > 0.0 ? -num / den : 0.0;
    if (!(c == c)) c = 0.0;                  // NaN
    c = std::fmin(std::fmax(c, -256.0 / 16777216.0), 256.0 / 16777216.0);
    (variant == 1 ? h->mp_corr1 : h->mp_corr)[l] = (float)(1.0 + c);
    std::swap(ha, hb);
  }
  }
  CUDA_TRY(h, cudaGetLastError());
  return nmrgnn_synchronize(h, nullptr);
}

void comm_release(nmrgnn_handle* h) {
  for (int r = 0; r < h->comm_world && r < PEER_MAX_WORLD; ++r) {
    if (r == h->comm_rank) continue;
    if (h->comm_peer_gbuf[r]) cudaIpcCloseMemHandle(h->comm_peer_gbuf[r]);
    if (h->comm_peer_flags[r]) cudaIpcCloseMemHandle(h->comm_peer_flags[r]);
  }
  for (int r = 0; r < PEER_MAX_WORLD; ++r) h->comm_peer_gbuf[r] = nullptr, h->comm_peer_flags[r] = nullptr;
  if (h->comm_gbuf) cudaFree(h->comm_gbuf);
  if (h->comm_flags) cudaFree(h->comm_flags);
  if (h->comm_done) cudaFree(h->comm_done);
  h->comm_gbuf = nullptr;
  h->comm_flags = nullptr;
  h->comm_done = nullptr;
  h->comm_ready = false;
  h->comm_world = 0;
  h->comm_rank = -1;
}

}  // namespace

// ============================================================================ C ABI
extern "C" {

int nmrgnn_abi_version(void) { return NMRGNN_ABI_VERSION; }

int nmrgnn_num_weights(const nmrgnn_dims* d) {
  if (!d) return NMRGNN_ERR_BAD_DIMS;
  return 2 * d->n_edge_fc + 1 + d->n_mp + 2 * d->n_fc + 2 + 2;
}

const char* nmrgnn_last_error(const nmrgnn_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int64_t nmrgnn_kernel_launches(const nmrgnn_handle* h) { return h ? h->launches : 0; }

const char* nmrgnn_compute_path(const nmrgnn_handle* h) { return h ? h->path.c_str() : ""; }

void nmrgnn_destroy(nmrgnn_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (float* p : h->owned) cudaFree(p);
  for (DevBuf* b : {&h->atoms, &h->nlist, &h->edges, &h->invdeg, &h->efeat, &h->hA, &h->hB, &h->peaks, &h->tmp_in,
                    &h->tmp_out, &h->pos, &h->offs, &h->knn_sorted, &h->knn_cells, &h->knn_grid, &h->rec, &h->hmaxA, &h->hmaxB, &h->genA, &h->genB})
    if (b->p) cudaFree(b->p);
  for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
  if (h->ev_last) cudaEventDestroy(h->ev_last);
  comm_release(h);
  if (h->err_flag) cudaFree(h->err_flag);
  if (h->err_flag_host) cudaFreeHost(h->err_flag_host);
  for (cudaEvent_t e : h->ev_copy)
    if (e) cudaEventDestroy(e);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int nmrgnn_create(const nmrgnn_dims* dims, const float* const* weights, int n_weights, int device,
                  nmrgnn_handle** out) {
  if (!dims || !weights || !out) return fail(nullptr, NMRGNN_ERR_BAD_DIMS, "null argument");
  *out = nullptr;
  if (!dims_ok(*dims)) return fail(nullptr, NMRGNN_ERR_BAD_DIMS, "inconsistent nmrgnn_dims");
  if (n_weights != nmrgnn_num_weights(dims))
    return fail(nullptr, NMRGNN_ERR_BAD_DIMS, "expected %d weight arrays, got %d", nmrgnn_num_weights(dims), n_weights);
  for (int i = 0; i < n_weights; ++i)
    if (!weights[i]) return fail(nullptr, NMRGNN_ERR_BAD_DIMS, "weights[%d] is null", i);

  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return fail(nullptr, NMRGNN_ERR_NO_DEVICE, "no CUDA device visible");
  }
  if (device < 0 || device >= n_dev) return fail(nullptr, NMRGNN_ERR_NO_DEVICE, "device %d out of range", device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, NMRGNN_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return fail(nullptr, NMRGNN_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);

  nmrgnn_handle* h = new (std::nothrow) nmrgnn_handle();
  if (!h) return fail(nullptr, NMRGNN_ERR_OOM, "host allocation failed");
  h->d = *dims;
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  int rc = NMRGNN_OK;
  auto bail = [&](int code) {
    g_create_error = h->err;
    nmrgnn_destroy(h);
    return code;
  };
#define TRY_RC(expr)            \
  do {                          \
    rc = (expr);                \
    if (rc) return bail(rc);    \
  } while (0)
#define CUDA_RC(expr)                                                        \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) {                                                 \
      fail(h, NMRGNN_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
      return bail(NMRGNN_ERR_CUDA);                                          \
    }                                                                        \
  } while (0)

  CUDA_RC(cudaSetDevice(device));
  CUDA_RC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CUDA_RC(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  for (auto& e : h->ev_copy) CUDA_RC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CUDA_RC(cudaMalloc(&h->err_flag, sizeof(int)));
  CUDA_RC(cudaMemset(h->err_flag, 0, sizeof(int)));
  CUDA_RC(cudaMallocHost(&h->err_flag_host, sizeof(int)));
  *h->err_flag_host = 0;

  const int C = dims->num_elem, F = dims->atom_features, E = dims->edge_features, H = dims->edge_hidden;
  const int F2 = F / 2;
  int w = 0;
  h->edge_W.resize(dims->n_edge_fc);
  h->edge_b.resize(dims->n_edge_fc);
  for (int i = 0; i < dims->n_edge_fc; ++i) {
    const int outw = (i == dims->n_edge_fc - 1) ? E : H;
    TRY_RC(upload(h, weights[w++], (size_t)H * outw, &h->edge_W[i]));
    TRY_RC(upload(h, weights[w++], (size_t)outw, &h->edge_b[i]));
  }
  TRY_RC(upload(h, weights[w++], (size_t)C * F, &h->embed));
  h->mp_Wp.resize(dims->n_mp);
  {
    // W'[(c, n, ll), m] = w[c*32+ll, m, n]: K-order matches the T slices the MP kernel builds
    std::vector<float> packed((size_t)F * E * F);
    for (int l = 0; l < dims->n_mp; ++l) {
      const float* src = weights[w++];
      if (F % 32 == 0) {
        for (int c = 0; c < F / 32; ++c)
          for (int n = 0; n < E; ++n)
            for (int ll = 0; ll < 32; ++ll) {
              const size_t krow = (size_t)c * 32 * E + (size_t)n * 32 + ll;
              const float* s = src + ((size_t)(c * 32 + ll) * F) * E + n;
              float* dst = packed.data() + krow * F;
              for (int m = 0; m < F; ++m) dst[m] = s[(size_t)m * E];
            }
      } else {
        for (int lf = 0; lf < F; ++lf)
          for (int n = 0; n < E; ++n)
            for (int m = 0; m < F; ++m) packed[((size_t)lf * E + n) * F + m] = src[((size_t)lf * F + m) * E + n];
      }
      TRY_RC(upload(h, packed.data(), packed.size(), &h->mp_Wp[l]));
      if (!(F == 256 && H == 128 && E >= 1 && E <= 4)) {
        h->mp_W.resize(dims->n_mp);
        TRY_RC(upload(h, src, (size_t)F * F * E, &h->mp_W[l]));
      }
    }
  }
  h->fc_W.resize(dims->n_fc);
  h->fc_b.resize(dims->n_fc);
  for (int i = 0; i < dims->n_fc; ++i) {
    const int outw = (i == dims->n_fc - 1) ? F2 : F;
    TRY_RC(upload(h, weights[w++], (size_t)F * outw, &h->fc_W[i]));
    TRY_RC(upload(h, weights[w++], (size_t)outw, &h->fc_b[i]));
  }
  TRY_RC(upload(h, weights[w++], (size_t)F2 * C, &h->out_W));
  TRY_RC(upload(h, weights[w++], (size_t)C, &h->out_b));
  TRY_RC(upload(h, weights[w++], (size_t)C, &h->peak_std));
  TRY_RC(upload(h, weights[w++], (size_t)C, &h->peak_avg));
  {
    std::vector<float> c;
    rbf_grid(dims->rbf_low, dims->rbf_high, H, c, h->gap);
    TRY_RC(upload(h, c.data(), c.size(), &h->centers));
  }
  TRY_RC(build_edge_table(h));

  h->fast_path = (F == 256 && H == 128 && E >= 1 && E <= 4);
  if (h->fast_path) {
    CUDA_RC(cudaFuncSetAttribute(fc_readout_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fc_smem_bytes()));
    CUDA_RC(cudaFuncSetAttribute(edge_mlp_ffma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)edge_smem_bytes<1>()));
    CUDA_RC(cudaFuncSetAttribute(edge_mlp_ffma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)edge_smem_bytes<2>()));
    CUDA_RC(cudaFuncSetAttribute(edge_mlp_ffma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)edge_smem_bytes<3>()));
    CUDA_RC(cudaFuncSetAttribute(edge_mlp_ffma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)edge_smem_bytes<4>()));
  }
  // tensor-core path: fp16 scaled-split operands need every weight below the fp16 range
  h->tc_ok = h->fast_path && E <= 4 && dims->n_edge_fc >= 2;
  if (h->tc_ok) {
    const int n_hidden = dims->n_edge_fc - 1;
    std::vector<uint8_t> img;
    std::vector<float> bias((size_t)n_hidden * 128);
    int wi = 0;
    float bound = 1.0f;  // RBF * mask <= 1
    h->edge_in_scale[0] = h->edge_out_scale[0] = 1.0f;
    for (int l = 0; l < n_hidden; ++l, wi += 2) {
      const float* W = weights[wi];
      if (max_abs(W, (size_t)H * H) > 60000.f) h->tc_ok = false;
      h->edge_w_host.emplace_back(W, W + (size_t)H * H);
      std::memcpy(bias.data() + (size_t)l * 128, weights[wi + 1], 128 * sizeof(float));
      bound = act_bound(bound * max_col_abs_sum(W, H, H) + max_abs(weights[wi + 1], H), dims->fc_activation);
      const int sx = scale_exp_for(bound);
      h->edge_in_scale[l + 1] = tc::pow2f_exact(-sx);
      h->edge_out_scale[l + 1] = tc::pow2f_exact(sx);
    }
    // round-toward-zero compensation of the H/16-instruction chains: position-dependent, in the lo images
    // (pack_edge_images); edge_rz is the constant alternative (option "edge_pos_comp_x100" = 0), folded into the scale
    // that the epilogue applies to the accumulator anyway
    TRY_RC(pack_edge_images(h));
    {
      const float* W = weights[wi];
      if (max_abs(W, (size_t)H * E) > 60000.f) h->tc_ok = false;
      pack_sw64_f16([&](int k, int n) { return W[(size_t)k * E + n]; }, H, E, 16, img);
    }
    TRY_RC(upload_bytes(h, img.data(), img.size(), &h->edge_f_img));
    TRY_RC(upload(h, bias.data(), bias.size(), &h->edge_bias));
    ACT_SET_SMEM(edge_mlp_tc_kernel, ETC_SMEM);
    CUDA_RC(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM));
    CUDA_RC(cudaFuncSetAttribute(tc_selftest_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STH_SMEM));
    CUDA_RC(cudaFuncSetAttribute(tc_selftest_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STH_SMEM));
  }
  h->mp_tc_ok = h->tc_ok && E <= 3 && dims->n_mp >= 1;
  if (h->mp_tc_ok) {
    const int w0 = 2 * dims->n_edge_fc + 1;
    for (int l = 0; l < dims->n_mp; ++l)
      if (max_abs(weights[w0 + l], (size_t)F * F * E) > 60000.f) h->mp_tc_ok = false;
  }
  if (h->mp_tc_ok) {
    const int w0 = 2 * dims->n_edge_fc + 1;
    h->mp_w_host.resize(dims->n_mp);
    for (int l = 0; l < dims->n_mp; ++l) h->mp_w_host[l].assign(weights[w0 + l], weights[w0 + l] + (size_t)F * F * E);
    TRY_RC(pack_mp_images(h));
    ACT_SET_SMEM(mp_layer_tc_kernel, MTC_SMEM);
    ACT_SET_SMEM(mp_layer_tc_vt_kernel, MTC_SMEM);
    CUDA_RC(cudaFuncSetAttribute(mp_layer_tc_vt_kernel<ACT_LINEAR>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    CUDA_RC(cudaFuncSetAttribute(mp_layer_tc_vt_kernel<ACT_SOFTPLUS>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    CUDA_RC(cudaFuncSetAttribute(mp_layer_tc_vt_kernel<ACT_RELU>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    CUDA_RC(cudaFuncSetAttribute(mp_layer_tc_vt_kernel<ACT_TANH>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    ACT_SET_SMEM(mp_layer_tc_seg_kernel, MTC_SMEM);
    ACT_SET_SMEM(mp_layer_tc1_kernel, MTC_SMEM);
    CUDA_RC(cudaFuncSetAttribute(mp_layer_tc1_kernel<ACT_LINEAR>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    CUDA_RC(cudaFuncSetAttribute(mp_layer_tc1_kernel<ACT_SOFTPLUS>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    CUDA_RC(cudaFuncSetAttribute(mp_layer_tc1_kernel<ACT_RELU>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    CUDA_RC(cudaFuncSetAttribute(mp_layer_tc1_kernel<ACT_TANH>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    ACT_SET_SMEM(mp_layer_np_kernel, MNP_SMEM);
    // 162 KB of shared memory: the 164 KB configuration leaves 92 KB of L1 for the gathers
    CUDA_RC(cudaFuncSetAttribute(mp_layer_np_kernel<ACT_LINEAR>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    CUDA_RC(cudaFuncSetAttribute(mp_layer_np_kernel<ACT_SOFTPLUS>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    CUDA_RC(cudaFuncSetAttribute(mp_layer_np_kernel<ACT_RELU>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    CUDA_RC(cudaFuncSetAttribute(mp_layer_np_kernel<ACT_TANH>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    // 196 KB of shared memory, 60 KB of L1 for the gathers (see MTC_SMEM)
    CUDA_RC(cudaFuncSetAttribute(mp_layer_tc_kernel<ACT_LINEAR>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    CUDA_RC(cudaFuncSetAttribute(mp_layer_tc_kernel<ACT_SOFTPLUS>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    CUDA_RC(cudaFuncSetAttribute(mp_layer_tc_kernel<ACT_RELU>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    CUDA_RC(cudaFuncSetAttribute(mp_layer_tc_kernel<ACT_TANH>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    CUDA_RC(cudaFuncSetAttribute(mp_layer_tc_seg_kernel<ACT_LINEAR>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    CUDA_RC(cudaFuncSetAttribute(mp_layer_tc_seg_kernel<ACT_SOFTPLUS>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    CUDA_RC(cudaFuncSetAttribute(mp_layer_tc_seg_kernel<ACT_RELU>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    CUDA_RC(cudaFuncSetAttribute(mp_layer_tc_seg_kernel<ACT_TANH>, cudaFuncAttributePreferredSharedMemoryCarveout, 85));
    h->mp_corr.assign(dims->n_mp, 1.0f);
    TRY_RC(calibrate_mp(h));
    h->launches = 0;
  }
  h->fc_tc_ok = h->tc_ok;
  if (h->fc_tc_ok) {
    const int w0 = 2 * dims->n_edge_fc + 1 + dims->n_mp;
    for (int i = 0; i < dims->n_fc; ++i) {
      const int outw = (i == dims->n_fc - 1) ? F2 : F;
      if (max_abs(weights[w0 + 2 * i], (size_t)F * outw) > 60000.f) h->fc_tc_ok = false;
    }
  }
  if (h->fc_tc_ok) {
    const int w0 = 2 * dims->n_edge_fc + 1 + dims->n_mp;
    std::vector<float> bias((size_t)dims->n_fc * 256, 0.f);
    float gain = 1.0f, offs = 0.0f;
    h->fc_w_host.resize(dims->n_fc);
    for (int i = 0; i < dims->n_fc; ++i) {
      const bool last = (i == dims->n_fc - 1);
      const int outw = last ? F2 : F;
      const float* W = weights[w0 + 2 * i];
      const float* b = weights[w0 + 2 * i + 1];
      std::memcpy(bias.data() + (size_t)i * 256, b, outw * sizeof(float));
      h->fc_w_host[i].assign(W, W + (size_t)F * outw);
      // |x_l| <= gain * max|x_0| + offs  (input of layer i), then through act(x W + b) + x
      h->fc_gain[i] = gain;
      h->fc_offs[i] = offs;
      const float ws = max_col_abs_sum(W, F, outw), bm = max_abs(b, outw);
      // one-layer growth bound for the pipelined kernel (residual layers): |act(x W + b) + x| <= g max|x| + o
      if (dims->fc_activation == ACT_TANH) {
        h->fc_g1[i] = 1.0f;
        h->fc_o1[i] = 1.0001f;
      } else {
        h->fc_g1[i] = (ws + 1.0f) * 1.0001f;
        h->fc_o1[i] = (bm + (dims->fc_activation == ACT_SOFTPLUS ? 0.6931472f : 0.0f)) * 1.0001f;
      }
      if (dims->fc_activation == ACT_TANH) {
        offs += 1.0f;
      } else {
        const float add = dims->fc_activation == ACT_SOFTPLUS ? 0.6931472f : 0.0f;
        offs = offs * (ws + 1.0f) + bm + add;
        gain = gain * (ws + 1.0f);
      }
    }
    TRY_RC(pack_fc_images(h));
    TRY_RC(upload(h, bias.data(), bias.size(), &h->fc_bias));
    h->fc_rz = 1.0f;
    ACT_SET_SMEM(fc_readout_tc_kernel, FTC_SMEM);
    ACT_SET_SMEM(fc_readout_pipe_kernel, FPI_SMEM);
    ACT_SET_SMEM(fc_readout_pipe_pair_kernel, FPI_SMEM_PAIR);
  }
  update_path(h);
#undef TRY_RC
#undef CUDA_RC
  *out = h;
  return NMRGNN_OK;
}

int nmrgnn_synchronize(nmrgnn_handle* h, void* stream) {
  if (!h) return NMRGNN_ERR_BAD_DIMS;
  CUDA_TRY(h, cudaSetDevice(h->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
  CUDA_TRY(h, cudaMemcpyAsync(h->err_flag_host, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(h, cudaStreamSynchronize(s));
  if (*h->err_flag_host != 0) {
    *h->err_flag_host = 0;
    CUDA_TRY(h, cudaMemsetAsync(h->err_flag, 0, sizeof(int), s));
    CUDA_TRY(h, cudaStreamSynchronize(s));
    return fail(h, NMRGNN_ERR_BAD_INDEX, "nlist holds an index outside [0, n_atoms)");
  }
  return NMRGNN_OK;
}

int nmrgnn_edge_features(nmrgnn_handle* h, const float* edges, int64_t n_edges, float* edge_features, int mem,
                         void* stream) {
  cudaStream_t s;
  int rc = begin_call(h, mem, stream, &s);
  if (rc) return rc;
  if (n_edges < 0 || (n_edges > 0 && (!edges || !edge_features))) return fail(h, NMRGNN_ERR_BAD_DIMS, "bad arguments");
  PathScope scope(h, n_edges / 16);
  Io io{h, mem, s};
  const void* d_edges;
  void* d_out;
  const size_t E = h->d.edge_features;
  if ((rc = io.in(edges, n_edges * sizeof(float), h->edges, &d_edges))) return rc;
  if ((rc = io.out_buf(edge_features, n_edges * E * sizeof(float), h->efeat, &d_out))) return rc;
  if ((rc = launch_edge(h, s, (const float*)d_edges, n_edges, (float*)d_out, nullptr, 0))) return rc;
  if ((rc = io.finish(edge_features, d_out, n_edges * E * sizeof(float)))) return rc;
  return end_call(h, stream, s);
}

int nmrgnn_embed(nmrgnn_handle* h, const float* atoms, int64_t n_atoms, float* nodes, int mem, void* stream) {
  cudaStream_t s;
  int rc = begin_call(h, mem, stream, &s);
  if (rc) return rc;
  if (n_atoms < 0 || (n_atoms > 0 && (!atoms || !nodes))) return fail(h, NMRGNN_ERR_BAD_DIMS, "bad arguments");
  Io io{h, mem, s};
  const void* d_atoms;
  void* d_out;
  const size_t C = h->d.num_elem, F = h->d.atom_features;
  if ((rc = io.in(atoms, n_atoms * C * sizeof(float), h->atoms, &d_atoms))) return rc;
  if ((rc = io.out_buf(nodes, n_atoms * F * sizeof(float), h->hA, &d_out))) return rc;
  if ((rc = launch_embed(h, s, (const float*)d_atoms, n_atoms, (float*)d_out))) return rc;
  if ((rc = io.finish(nodes, d_out, n_atoms * F * sizeof(float)))) return rc;
  return end_call(h, stream, s);
}

int nmrgnn_mp_layer(nmrgnn_handle* h, int32_t layer, const float* nodes_in, const int32_t* nlist,
                    const float* edge_features, const float* inv_degree, int64_t n_atoms, int32_t k,
                    float* nodes_out, int mem, void* stream) {
  cudaStream_t s;
  int rc = begin_call(h, mem, stream, &s);
  if (rc) return rc;
  if (layer < 0 || layer >= h->d.n_mp) return fail(h, NMRGNN_ERR_BAD_DIMS, "layer %d out of range", layer);
  if (n_atoms < 0 || k < 1 || (n_atoms > 0 && (!nodes_in || !nlist || !edge_features || !inv_degree || !nodes_out)))
    return fail(h, NMRGNN_ERR_BAD_DIMS, "bad arguments");
  if (nodes_in == nodes_out && n_atoms > 0) return fail(h, NMRGNN_ERR_BAD_DIMS, "nodes_out must not alias nodes_in");
  if (n_atoms >= ((int64_t)1 << 31)) return fail(h, NMRGNN_ERR_BAD_DIMS, "n_atoms exceeds int32 index range");
  PathScope scope(h, n_atoms);
  Io io{h, mem, s};
  const void *d_in, *d_nl, *d_ef, *d_inv;
  void* d_out;
  const size_t F = h->d.atom_features, E = h->d.edge_features;
  if ((rc = io.in(nodes_in, n_atoms * F * sizeof(float), h->hA, &d_in))) return rc;
  if ((rc = io.in(nlist, n_atoms * k * sizeof(int32_t), h->nlist, &d_nl))) return rc;
  if ((rc = io.in(edge_features, n_atoms * k * E * sizeof(float), h->efeat, &d_ef))) return rc;
  if ((rc = io.in(inv_degree, n_atoms * sizeof(float), h->invdeg, &d_inv))) return rc;
  if ((rc = io.out_buf(nodes_out, n_atoms * F * sizeof(float), h->hB, &d_out))) return rc;
  if (mp_tc_usable(h, k)) {
    if ((rc = ensure(h, h->rec, n_atoms * k * sizeof(float4)))) return rc;
    if ((rc = ensure(h, h->hmaxA, 2 * n_atoms * sizeof(float)))) return rc;
    if ((rc = ensure(h, h->hmaxB, 2 * n_atoms * sizeof(float)))) return rc;
    if ((rc = launch_pack_rec(h, s, (const int32_t*)d_nl, (const float*)d_ef, (float4*)h->rec.p, n_atoms * k, n_atoms, k)))
      return rc;
    if ((rc = launch_absmax(h, s, (const float*)d_in, n_atoms, (float*)h->hmaxA.p))) return rc;
    if ((rc = launch_mp_tc(h, s, layer, (const float*)d_in, (const float*)h->hmaxA.p, (const float4*)h->rec.p,
                           (const float*)d_inv, n_atoms, k, (float*)d_out, (float*)h->hmaxB.p)))
      return rc;
  } else {
    // the exact-FP32 / generic MP kernels clamp a neighbour index instead of flagging it (in the whole forward the edge
    // stage does the check); TF's GatherV2 raises, and so must this entry point on every route
    if (n_atoms > 0) {
      gen_index_check_kernel<<<(unsigned)((n_atoms * k + 255) / 256), 256, 0, s>>>((const int32_t*)d_nl, n_atoms * k, n_atoms,
                                                                                  h->err_flag);
      h->launches++;
    }
    if ((rc = launch_mp(h, s, layer, (const float*)d_in, (const int32_t*)d_nl, (const float*)d_ef,
                        (const float*)d_inv, n_atoms, k, (float*)d_out)))
      return rc;
  }
  if ((rc = io.finish(nodes_out, d_out, n_atoms * F * sizeof(float)))) return rc;
  return end_call(h, stream, s);
}

int nmrgnn_fc_readout(nmrgnn_handle* h, const float* nodes, const float* atoms, int64_t n_atoms, float* peaks,
                      float* fc_nodes, int mem, void* stream) {
  cudaStream_t s;
  int rc = begin_call(h, mem, stream, &s);
  if (rc) return rc;
  if (n_atoms < 0 || (n_atoms > 0 && (!nodes || !atoms || !peaks))) return fail(h, NMRGNN_ERR_BAD_DIMS, "bad arguments");
  PathScope scope(h, n_atoms);
  Io io{h, mem, s};
  const void *d_nodes, *d_atoms;
  void *d_peaks, *d_fc = nullptr;
  const size_t C = h->d.num_elem, F = h->d.atom_features, F2 = F / 2;
  if ((rc = io.in(nodes, n_atoms * F * sizeof(float), h->hA, &d_nodes))) return rc;
  if ((rc = io.in(atoms, n_atoms * C * sizeof(float), h->atoms, &d_atoms))) return rc;
  if ((rc = io.out_buf(peaks, n_atoms * sizeof(float), h->peaks, &d_peaks))) return rc;
  if (fc_nodes && (rc = io.out_buf(fc_nodes, n_atoms * F2 * sizeof(float), h->tmp_out, &d_fc))) return rc;
  if ((rc = launch_fc(h, s, (const float*)d_nodes, (const float*)d_atoms, n_atoms, (float*)d_peaks, (float*)d_fc)))
    return rc;
  if ((rc = io.finish(peaks, d_peaks, n_atoms * sizeof(float)))) return rc;
  if (fc_nodes && (rc = io.finish(fc_nodes, d_fc, n_atoms * F2 * sizeof(float)))) return rc;
  return end_call(h, stream, s);
}

// The whole forward.  dev_peaks != NULL: the peaks stay on the device at that address (no copy to `peaks`, which may
// be NULL); finish = false: the caller enqueues more work on the stream and ends the call itself.
static int forward_core(nmrgnn_handle* h, const float* atoms, const int32_t* nlist, const float* edges,
                        const float* inv_degree, int64_t n_atoms, int32_t k, float* peaks, int mem, void* stream,
                        float* dev_peaks, bool finish) {
  cudaStream_t s;
  int rc = begin_call(h, mem, stream, &s);
  if (rc) return rc;
  if (n_atoms < 0 || k < 1) return fail(h, NMRGNN_ERR_BAD_DIMS, "n_atoms=%lld k=%d invalid", (long long)n_atoms, (int)k);
  if (n_atoms == 0) return NMRGNN_OK;
  if (!atoms || !nlist || !edges || !inv_degree || (!peaks && !dev_peaks)) return fail(h, NMRGNN_ERR_BAD_DIMS, "null buffer");
  if (n_atoms >= ((int64_t)1 << 31)) return fail(h, NMRGNN_ERR_BAD_DIMS, "n_atoms exceeds int32 index range");
  PathScope scope(h, n_atoms);
  Io io{h, mem, s};
  const size_t C = h->d.num_elem, F = h->d.atom_features, E = h->d.edge_features;
  const void *d_atoms, *d_nl, *d_edges, *d_inv;
  void* d_peaks;
  // Host buffers, large call: the inputs go up on a second stream in chunks (edges + nlist first, in atom
  // ranges aligned to the 128-edge tiles), each chunk's edge kernel starts as soon as its copy has landed,
  // atoms / inv_degree follow while the edge kernels run.  Chunk c is guarded by ev_copy[c], the rest by [8].
  struct CopyGuard {   // an early error return must not leave uploads from the caller's buffers in flight
    cudaStream_t cs = nullptr;
    ~CopyGuard() {
      if (cs) cudaStreamSynchronize(cs);
    }
  } copy_guard;
  constexpr int MAX_CHUNKS = 8;
  int n_chunks = 1;
  int64_t chunk_atoms = n_atoms;
  if (mem == NMRGNN_MEM_HOST && n_atoms >= 32768) {
    n_chunks = n_atoms >= 131072 ? 8 : 4;    // only the first chunk's upload is not hidden behind an edge kernel
    chunk_atoms = (((n_atoms + n_chunks - 1) / n_chunks) + 127) / 128 * 128;
    n_chunks = (int)((n_atoms + chunk_atoms - 1) / chunk_atoms);
    if (n_chunks > MAX_CHUNKS) n_chunks = MAX_CHUNKS, chunk_atoms = n_atoms;
  }
  if (n_chunks > 1) {
    if ((rc = ensure(h, h->atoms, n_atoms * C * sizeof(float)))) return rc;
    if ((rc = ensure(h, h->nlist, n_atoms * k * sizeof(int32_t)))) return rc;
    if ((rc = ensure(h, h->edges, n_atoms * k * sizeof(float)))) return rc;
    if ((rc = ensure(h, h->invdeg, n_atoms * sizeof(float)))) return rc;
    d_atoms = h->atoms.p, d_nl = h->nlist.p, d_edges = h->edges.p, d_inv = h->invdeg.p;
    cudaStream_t cs = h->copy_stream;
    copy_guard.cs = cs;
    for (int c = 0; c < n_chunks; ++c) {
      const int64_t a0 = c * chunk_atoms, na = std::min<int64_t>(chunk_atoms, n_atoms - a0);
      CUDA_TRY(h, cudaMemcpyAsync((float*)h->edges.p + a0 * k, edges + a0 * k, na * k * sizeof(float), cudaMemcpyHostToDevice, cs));
      CUDA_TRY(h, cudaMemcpyAsync((int32_t*)h->nlist.p + a0 * k, nlist + a0 * k, na * k * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
      CUDA_TRY(h, cudaEventRecord(h->ev_copy[c], cs));
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->atoms.p, atoms, n_atoms * C * sizeof(float), cudaMemcpyHostToDevice, cs));
    CUDA_TRY(h, cudaMemcpyAsync(h->invdeg.p, inv_degree, n_atoms * sizeof(float), cudaMemcpyHostToDevice, cs));
    CUDA_TRY(h, cudaEventRecord(h->ev_copy[8], cs));
  } else {
  if ((rc = io.in(atoms, n_atoms * C * sizeof(float), h->atoms, &d_atoms))) return rc;
  if ((rc = io.in(nlist, n_atoms * k * sizeof(int32_t), h->nlist, &d_nl))) return rc;
  if ((rc = io.in(edges, n_atoms * k * sizeof(float), h->edges, &d_edges))) return rc;
  if ((rc = io.in(inv_degree, n_atoms * sizeof(float), h->invdeg, &d_inv))) return rc;
  }
  if (dev_peaks != nullptr) d_peaks = dev_peaks;
  else if ((rc = io.out_buf(peaks, n_atoms * sizeof(float), h->peaks, &d_peaks))) return rc;
  if ((rc = ensure(h, h->efeat, n_atoms * k * E * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->hA, n_atoms * F * sizeof(float)))) return rc;
  if ((rc = ensure(h, h->hB, n_atoms * F * sizeof(float)))) return rc;

  float* ef = (float*)h->efeat.p;
  float* ha = (float*)h->hA.p;
  float* hb = (float*)h->hB.p;
  if (h->profile && h->ev.empty()) {
    h->ev.resize(h->d.n_mp + 4);
    for (auto& e : h->ev) CUDA_TRY(h, cudaEventCreate(&e));
  }
  const float* fc_hmax = nullptr;
  int fc_hmax_pair = 0;
  int mk = 0;
  auto mark = [&]() {
    if (h->profile) cudaEventRecord(h->ev[mk++], s);
  };
  mark();
  if (mp_tc_usable(h, k) && h->d.n_mp > 0) {
    // tensor-core route: the edge kernel emits {e0,e1,e2,idx} records, the MP layers carry max|h| per atom
    if ((rc = ensure(h, h->rec, n_atoms * k * sizeof(float4)))) return rc;
    if ((rc = ensure(h, h->hmaxA, 2 * n_atoms * sizeof(float)))) return rc;
    if ((rc = ensure(h, h->hmaxB, 2 * n_atoms * sizeof(float)))) return rc;
    float4* rec = (float4*)h->rec.p;
    float* ma = (float*)h->hmaxA.p;
    float* mb = (float*)h->hmaxB.p;
    for (int c = 0; c < n_chunks; ++c) {
      const int64_t a0 = c * chunk_atoms, na = std::min<int64_t>(chunk_atoms, n_atoms - a0);
      if (n_chunks > 1) CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_copy[c], 0));
      if ((rc = launch_edge(h, s, (const float*)d_edges + a0 * k, na * k, nullptr, (const int32_t*)d_nl + a0 * k, n_atoms,
                            rec + a0 * k, k, a0 * k)))
        return rc;
    }
    mark();
    if (n_chunks > 1) CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_copy[8], 0));
    if ((rc = launch_embed(h, s, (const float*)d_atoms, n_atoms, ha, ma))) return rc;
    mark();
    int m_pair = 0;
    for (int l = 0; l < h->d.n_mp; ++l) {
      if ((rc = launch_mp_tc(h, s, l, ha, ma, rec, (const float*)d_inv, n_atoms, k, hb, mb, 0, m_pair, &m_pair))) return rc;
      mark();
      std::swap(ha, hb);
      std::swap(ma, mb);
    }
    fc_hmax = ma;                  // the last MP epilogue (or the embedding) left max |h row| of the node MLP's input
    fc_hmax_pair = m_pair;
  } else {
    for (int c = 0; c < n_chunks; ++c) {
      const int64_t a0 = c * chunk_atoms, na = std::min<int64_t>(chunk_atoms, n_atoms - a0);
      if (n_chunks > 1) CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_copy[c], 0));
      if ((rc = launch_edge(h, s, (const float*)d_edges + a0 * k, na * k, ef + a0 * k * E, (const int32_t*)d_nl + a0 * k, n_atoms)))
        return rc;
    }
    mark();
    if (n_chunks > 1) CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_copy[8], 0));
    if ((rc = launch_embed(h, s, (const float*)d_atoms, n_atoms, ha))) return rc;
    mark();
    for (int l = 0; l < h->d.n_mp; ++l) {
      if ((rc = launch_mp(h, s, l, ha, (const int32_t*)d_nl, ef, (const float*)d_inv, n_atoms, k, hb))) return rc;
      mark();
      std::swap(ha, hb);
    }
  }
  if ((rc = launch_fc(h, s, ha, (const float*)d_atoms, n_atoms, (float*)d_peaks, nullptr, fc_hmax, fc_hmax_pair))) return rc;
  mark();
  h->ev_valid = h->profile;
  if (dev_peaks == nullptr && (rc = io.finish(peaks, d_peaks, n_atoms * sizeof(float)))) return rc;
  if (!finish) return NMRGNN_OK;     // (copy_guard waits for the chunked uploads on the way out)
  return end_call(h, stream, s);
}

int nmrgnn_forward(nmrgnn_handle* h, const float* atoms, const int32_t* nlist, const float* edges,
                   const float* inv_degree, int64_t n_atoms, int32_t k, float* peaks, int mem, void* stream) {
  return forward_core(h, atoms, nlist, edges, inv_degree, n_atoms, k, peaks, mem, stream, nullptr, true);
}

// ---------------------------------------------------------------------------- multi-GPU (peer_gather.cuh)
int nmrgnn_comm_local(nmrgnn_handle* h, int64_t capacity, int32_t world, void* ipc_out) {
  if (!h || !ipc_out || world < 1 || world > PEER_MAX_WORLD || capacity < 1)
    return fail(h, NMRGNN_ERR_BAD_DIMS, "bad arguments (world must be 1..%d)", PEER_MAX_WORLD);
  CUDA_TRY(h, cudaSetDevice(h->device));
  comm_release(h);
  h->comm_capacity = (capacity + 3) / 4 * 4;
  h->comm_world = world;
  const size_t gbytes = (size_t)2 * world * h->comm_capacity * sizeof(float);
  CUDA_TRY(h, cudaMalloc(&h->comm_gbuf, gbytes));
  CUDA_TRY(h, cudaMalloc(&h->comm_flags, PEER_MAX_WORLD * sizeof(uint32_t)));
  CUDA_TRY(h, cudaMalloc(&h->comm_done, sizeof(unsigned int)));
  CUDA_TRY(h, cudaMemset(h->comm_gbuf, 0, gbytes));
  CUDA_TRY(h, cudaMemset(h->comm_flags, 0, PEER_MAX_WORLD * sizeof(uint32_t)));
  CUDA_TRY(h, cudaMemset(h->comm_done, 0, sizeof(unsigned int)));
  CUDA_TRY(h, cudaDeviceSynchronize());
  cudaIpcMemHandle_t hb, hf;
  std::memset(ipc_out, 0, NMRGNN_COMM_HANDLE_BYTES);
  if (world > 1) {
    CUDA_TRY(h, cudaIpcGetMemHandle(&hb, h->comm_gbuf));
    CUDA_TRY(h, cudaIpcGetMemHandle(&hf, h->comm_flags));
    static_assert(2 * sizeof(cudaIpcMemHandle_t) <= NMRGNN_COMM_HANDLE_BYTES, "handle blob too small");
    std::memcpy(ipc_out, &hb, sizeof(hb));
    std::memcpy((char*)ipc_out + sizeof(hb), &hf, sizeof(hf));
  }
  return NMRGNN_OK;
}

int nmrgnn_comm_init(nmrgnn_handle* h, int32_t rank, int32_t world, const void* all_ipc) {
  if (!h || world != h->comm_world || rank < 0 || rank >= world || !h->comm_gbuf || (world > 1 && !all_ipc))
    return fail(h, NMRGNN_ERR_BAD_DIMS, "nmrgnn_comm_init: call nmrgnn_comm_local with the same world first");
  CUDA_TRY(h, cudaSetDevice(h->device));
  h->comm_rank = rank;
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      h->comm_peer_gbuf[r] = h->comm_gbuf;
      h->comm_peer_flags[r] = h->comm_flags;
      continue;
    }
    cudaIpcMemHandle_t hb, hf;
    const char* blob = (const char*)all_ipc + (size_t)r * NMRGNN_COMM_HANDLE_BYTES;
    std::memcpy(&hb, blob, sizeof(hb));
    std::memcpy(&hf, blob + sizeof(hb), sizeof(hf));
    void *pb = nullptr, *pf = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&pb, hb, cudaIpcMemLazyEnablePeerAccess);
    if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&pf, hf, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(h, NMRGNN_ERR_COMM, "cannot map the gather buffer of rank %d (%s): peer access over NVLink / PCIe is "
                  "required between the GPUs of one node", r, cudaGetErrorString(e));
    }
    h->comm_peer_gbuf[r] = (float*)pb;
    h->comm_peer_flags[r] = (uint32_t*)pf;
  }
  h->comm_epoch = 0;
  h->comm_ready = true;
  return NMRGNN_OK;
}

int nmrgnn_forward_sharded(nmrgnn_handle* h, const float* atoms, const int32_t* nlist, const float* edges,
                           const float* inv_degree, int64_t n_local, int32_t k, float* gathered, int mem, void* stream) {
  cudaStream_t s;
  int rc = begin_call(h, mem, stream, &s);
  if (rc) return rc;
  if (!h->comm_ready) return fail(h, NMRGNN_ERR_BAD_DIMS, "nmrgnn_forward_sharded: nmrgnn_comm_init has not been called");
  if (n_local < 0 || n_local > h->comm_capacity)
    return fail(h, NMRGNN_ERR_BAD_DIMS, "n_local=%lld exceeds the slot capacity %lld", (long long)n_local, (long long)h->comm_capacity);
  const int world = h->comm_world, rank = h->comm_rank;
  const uint32_t epoch = ++h->comm_epoch;
  const int64_t cap = h->comm_capacity;
  const int64_t parity_off = (int64_t)(epoch & 1u) * world * cap;
  float* slot = h->comm_gbuf + parity_off + (int64_t)rank * cap;
  // the forward writes this rank's peaks straight into its own slot of its own gather buffer
  if (n_local > 0 &&
      (rc = forward_core(h, atoms, nlist, edges, inv_degree, n_local, k, nullptr, mem, stream, slot, false)))
    return rc;
  if (world > 1) {
    PeerArgs a{};
    a.src = slot;
    a.n = n_local;
    a.world = world;
    a.rank = rank;
    a.slot_off = parity_off + (int64_t)rank * cap;
    for (int r = 0; r < world; ++r) {
      a.gbuf[r] = h->comm_peer_gbuf[r];
      a.flags[r] = h->comm_peer_flags[r];
    }
    a.epoch = epoch;
    a.done_ctr = h->comm_done;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(2 * h->num_sms, (n_local / 4 + 255) / 256));
    peer_scatter_signal_kernel<<<blocks, 256, 0, s>>>(a);
    peer_wait_kernel<<<1, 32, 0, s>>>(h->comm_flags, world, epoch);
    h->launches += 2;
  }
  if (gathered != nullptr) {
    const size_t bytes = (size_t)world * cap * sizeof(float);
    CUDA_TRY(h, cudaMemcpyAsync(gathered, h->comm_gbuf + parity_off, bytes,
                                mem == NMRGNN_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  }
  return end_call(h, stream, s);
}

const float* nmrgnn_comm_buffer(nmrgnn_handle* h, int64_t* capacity) {
  if (!h || !h->comm_ready) return nullptr;
  if (capacity) *capacity = h->comm_capacity;
  return h->comm_gbuf + (int64_t)(h->comm_epoch & 1u) * h->comm_world * h->comm_capacity;
}

void nmrgnn_comm_destroy(nmrgnn_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  comm_release(h);
}

int nmrgnn_selftest_gemm(nmrgnn_handle* h, const float* A, const float* W, float* D, int mode) {
  cudaStream_t s;
  int rc = begin_call(h, NMRGNN_MEM_HOST, nullptr, &s);
  if (rc) return rc;
  if (!A || !W || !D) return fail(h, NMRGNN_ERR_BAD_DIMS, "null buffer");
  if (mode < 0 || mode > 6) return fail(h, NMRGNN_ERR_BAD_DIMS, "mode must be 0..6");
  if (!h->tc_ok) return fail(h, NMRGNN_ERR_BAD_DIMS, "tensor-core path not available for this geometry");
  std::vector<uint8_t> img;
  if (mode <= 1) pack_sw64(W, ST_K, 128, 0, 128, 128, img);
  else pack_sw64_f16([&](int k, int n) { return W[(size_t)k * 128 + n]; }, ST_K, 128, 128, img);
  if ((rc = ensure(h, h->tmp_in, 128 * ST_K * sizeof(float) + img.size()))) return rc;
  if ((rc = ensure(h, h->tmp_out, 128 * 128 * sizeof(float)))) return rc;
  uint8_t* d_img = (uint8_t*)h->tmp_in.p;
  float* d_A = (float*)(d_img + img.size());
  CUDA_TRY(h, cudaMemcpyAsync(d_img, img.data(), img.size(), cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaMemcpyAsync(d_A, A, 128 * ST_K * sizeof(float), cudaMemcpyHostToDevice, s));
  CUDA_TRY(h, cudaStreamSynchronize(s));  // img is a local vector
  if (mode <= 1) tc_selftest_kernel<<<1, 192, ST_SMEM, s>>>(d_A, d_img, (float*)h->tmp_out.p, mode);
  else if (mode == 4) tc_selftest_ts_kernel<<<1, 192, STH_SMEM, s>>>(d_A, d_img, (float*)h->tmp_out.p);
  else if (mode >= 5) tc_selftest_pair_kernel<<<2, 192, STP_SMEM, s>>>(d_A, d_img, (float*)h->tmp_out.p, mode - 5);
  else tc_selftest_f16_kernel<<<1, 192, STH_SMEM, s>>>(d_A, d_img, (float*)h->tmp_out.p, mode);
  h->launches++;
  CUDA_TRY(h, cudaMemcpyAsync(D, h->tmp_out.p, 128 * 128 * sizeof(float), cudaMemcpyDeviceToHost, s));
  return end_call(h, nullptr, s);
}

int nmrgnn_stage_times(nmrgnn_handle* h, float* ms, int cap) {
  if (!h || !ms) return NMRGNN_ERR_BAD_DIMS;
  if (!h->ev_valid) return fail(h, NMRGNN_ERR_BAD_DIMS, "no profiled forward recorded (set option \"profile\" = 1 first)");
  const int n = (int)h->ev.size() - 1;
  if (cap < n) return fail(h, NMRGNN_ERR_BAD_DIMS, "need room for %d stage times", n);
  CUDA_TRY(h, cudaSetDevice(h->device));
  CUDA_TRY(h, cudaEventSynchronize(h->ev[n]));
  for (int i = 0; i < n; ++i) CUDA_TRY(h, cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]));
  return n;
}

int nmrgnn_tc_compensation(nmrgnn_handle* h, float* c_ulp, int cap) {
  if (!h || !c_ulp) return NMRGNN_ERR_BAD_DIMS;
  const int n = (int)h->mp_corr.size();
  if (cap < n + 1) return fail(h, NMRGNN_ERR_BAD_DIMS, "need room for %d values", n + 1);
  c_ulp[0] = (h->edge_rz - 1.0f) * 16777216.0f;
  const bool one = h->mp_one && h->mp_nseg == 1 && !h->mp_nsplit && (int)h->mp_corr1.size() == n;
  for (int l = 0; l < n; ++l) c_ulp[1 + l] = ((one ? h->mp_corr1[l] : h->mp_corr[l]) - 1.0f) * 16777216.0f;
  return n + 1;
}

int nmrgnn_edge_table_info(nmrgnn_handle* h, int32_t* n_intervals, double* rel_error) {
  if (!h) return NMRGNN_ERR_BAD_DIMS;
  if (n_intervals) *n_intervals = h->edge_tab_n;
  if (rel_error) *rel_error = h->edge_tab_err;
  return h->edge_tab_ok ? 1 : 0;
}

int nmrgnn_set_option(nmrgnn_handle* h, const char* name, int value) {
  if (!h || !name) return NMRGNN_ERR_BAD_DIMS;
  if (std::strcmp(name, "profile") == 0) {
    h->profile = value != 0;
    h->ev_valid = false;
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "force_ffma") == 0) {
    h->force_ffma = value != 0;
    update_path(h);
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "mp_role_counters") == 0) {
    // value != 0: allocate / zero the per-CTA role counters and print their means on value == 2
    if (value && !h->mp_dbg) {
      CUDA_TRY(h, cudaMalloc(&h->mp_dbg, 1024 * 8 * sizeof(long long)));
      h->owned.push_back((float*)h->mp_dbg);
    }
    if (value == 1) CUDA_TRY(h, cudaMemset(h->mp_dbg, 0, 1024 * 8 * sizeof(long long)));
    if (value == 2 && h->mp_dbg) {
      std::vector<long long> c(1024 * 8);
      CUDA_TRY(h, cudaMemcpy(c.data(), h->mp_dbg, c.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      double m[8] = {0};
      int n = 0;
      for (int b = 0; b < h->num_sms; ++b)
        if (c[b * 8] > 0) {
          ++n;
          for (int i = 0; i < 8; ++i) m[i] += (double)c[b * 8 + i];
        }
      if (n)
        printf("mp roles (mean cycles per CTA over %d CTAs): mma total %.0f | mma waits: epilogue %.0f producers %.0f W' %.0f | "
               "epilogue busy %.0f | producer waits: a_empty %.0f rec %.0f | shipper idle %.0f\n",
               n, m[0] / n, m[1] / n, m[2] / n, m[3] / n, m[4] / n, m[5] / n, m[6] / n, m[7] / n);
    }
    if (value == 0) h->mp_dbg = nullptr;
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "edge_table") == 0) {
    h->edge_table = value != 0;
    update_path(h);
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "mp_comp_x10") == 0) {       // diagnostics: every MP layer's compensation = value / 10 x 2^-24
    if (value < 0) return calibrate_mp(h);               // negative: back to the calibrated constants
    for (auto& c : h->mp_corr) c = 1.0f + (float)value / 10.0f / 16777216.0f;
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "mp_comp_delta_x10") == 0) {  // diagnostics: shift every MP layer's constant by value / 10 x 2^-24
    for (auto& c : h->mp_corr) c += (float)value / 10.0f / 16777216.0f;
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "fc_comp_x10") == 0) {        // diagnostics: node-MLP compensation = value / 10 x 2^-24
    h->fc_rz = 1.0f + (float)value / 10.0f / 16777216.0f;
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "knn_warp") == 0) {
    h->knn_warp = value != 0;
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "knn_cells") == 0) {
    h->knn_cells_on = value != 0;
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "fc_pipe") == 0) {
    h->fc_pipe = value != 0;
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "fc_pair") == 0) {
    h->fc_pair = value != 0;
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "fc_pos_comp1_x100") == 0) {  // slope of the pipelined node MLP's position-dependent compensation
    if (value < 0 || value > 400) return fail(h, NMRGNN_ERR_BAD_DIMS, "fc_pos_comp1_x100 must be in 0..400");
    h->fc_pos_c1 = (float)value / 100.0f;
    return h->fc_tc_ok ? pack_fc_images(h) : NMRGNN_OK;
  }
  if (std::strcmp(name, "fc_role_counters") == 0) {
    if (value && !h->fc_dbg) {
      CUDA_TRY(h, cudaMalloc(&h->fc_dbg, 1024 * 8 * sizeof(long long)));
      h->owned.push_back((float*)h->fc_dbg);
    }
    if (value == 1) CUDA_TRY(h, cudaMemset(h->fc_dbg, 0, 1024 * 8 * sizeof(long long)));
    if (value == 2 && h->fc_dbg) {
      std::vector<long long> c(1024 * 8);
      CUDA_TRY(h, cudaMemcpy(c.data(), h->fc_dbg, c.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      double m[8] = {0};
      int n = 0;
      for (int b = 0; b < h->num_sms; ++b)
        if (c[b * 8] > 0) {
          ++n;
          for (int i = 0; i < 8; ++i) m[i] += (double)c[b * 8 + i];
        }
      if (n)
        printf("fc roles (mean cycles per CTA over %d CTAs): mma total %.0f | mma waits: operand chunks %.0f W %.0f | "
               "epilogue busy %.0f | stager waits for X %.0f | W slots waited for %.0f, issue -> arrival %.0f cycles\n",
               n, m[0] / n, m[1] / n, m[2] / n, m[3] / n, m[4] / n, m[6] / n, m[6] > 0 ? m[5] / m[6] : 0.0);
    }
    if (value == 0) h->fc_dbg = nullptr;
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "mp_single_acc") == 0) {
    h->mp_one = value != 0;
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "mp_pos_comp1_x100") == 0) {  // slope of the single-accumulator form's position-dependent compensation
    if (value < 0 || value > 400) return fail(h, NMRGNN_ERR_BAD_DIMS, "mp_pos_comp1_x100 must be in 0..400");
    h->mp_pos_c1 = (float)value / 100.0f;
    if (h->mp_tc_ok) {
      if (int rc = pack_mp_images(h)) return rc;
      return calibrate_mp(h);
    }
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "mp_small_tiles") == 0) {
    h->mp_small_tiles = value != 0;
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "mp_l1_prefetch") == 0) {
    h->mp_l1_prefetch = value != 0;
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "mp_nsplit") == 0) {
    h->mp_nsplit = value != 0;
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "mp_chain_segments") == 0) {
    if (value != 1 && value != 2 && value != 4 && value != 8) return fail(h, NMRGNN_ERR_BAD_DIMS, "mp_chain_segments must be 1, 2, 4 or 8");
    if (value != h->mp_nseg) {
      h->mp_nseg = value;
      if (h->mp_tc_ok) {                             // the compensation belongs to a chain length
        if (int rc = pack_mp_images(h)) return rc;
        return calibrate_mp(h);
      }
    }
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "edge_pos_comp_x100") == 0) {  // the same for the edge MLP (H/16-instruction chains)
    if (value < 0 || value > 400) return fail(h, NMRGNN_ERR_BAD_DIMS, "edge_pos_comp_x100 must be in 0..400");
    if (!h->tc_ok) return NMRGNN_OK;
    const int n_hidden = h->d.n_edge_fc - 1;
    const float rz = value ? 1.0f : 1.0f + 0.17f * (float)(h->d.edge_hidden / 16 + 1) / 16777216.0f;
    if (h->compensate)
      for (int l = 0; l <= n_hidden; ++l) h->edge_out_scale[l] = h->edge_out_scale[l] / h->edge_rz * rz;
    h->edge_rz = rz;
    h->edge_pos_c = (float)value / 100.0f;
    return pack_edge_images(h);
  }
  if (std::strcmp(name, "fc_pos_comp_x100") == 0) {  // the same for the node MLP (16-instruction chains)
    if (value < 0 || value > 400) return fail(h, NMRGNN_ERR_BAD_DIMS, "fc_pos_comp_x100 must be in 0..400");
    h->fc_pos_c = (float)value / 100.0f;
    // without the position-dependent part: the analytic constant of a linear trajectory
    h->fc_rz = value ? 1.0f : 1.0f + 0.17f * (float)(h->d.atom_features / 16 + 1) / 16777216.0f;
    return h->fc_tc_ok ? pack_fc_images(h) : NMRGNN_OK;
  }
  if (std::strcmp(name, "mp_pos_comp_x100") == 0) {  // position-dependent compensation slope c' = value / 100 x 2^-24
    if (value < 0 || value > 400) return fail(h, NMRGNN_ERR_BAD_DIMS, "mp_pos_comp_x100 must be in 0..400");
    h->mp_pos_c = (float)value / 100.0f;
    if (h->mp_tc_ok) {
      if (int rc = pack_mp_images(h)) return rc;
      return calibrate_mp(h);
    }
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "tc_min_atoms") == 0) {
    h->tc_min_atoms = value < 0 ? 0 : value;
    return NMRGNN_OK;
  }
  if (std::strcmp(name, "tc_compensate") == 0) {
    if (h->tc_ok && (value != 0) != h->compensate) {
      const int n_hidden = h->d.n_edge_fc - 1;
      for (int l = 0; l <= n_hidden; ++l) h->edge_out_scale[l] = value ? h->edge_out_scale[l] * h->edge_rz
                                                                       : h->edge_out_scale[l] / h->edge_rz;
    }
    const bool changed = h->compensate != (value != 0);
    h->compensate = value != 0;
    if (changed && h->tc_ok)                                 // the position-dependent part lives in the operand images
      if (int rc = pack_edge_images(h)) return rc;
    if (changed && h->fc_tc_ok)
      if (int rc = pack_fc_images(h)) return rc;
    if (changed && h->mp_tc_ok) return pack_mp_images(h);
    return NMRGNN_OK;
  }
  return fail(h, NMRGNN_ERR_BAD_DIMS, "unknown option '%s'", name);
}

int nmrgnn_knn_graph(nmrgnn_handle* h, const float* positions, const int64_t* graph_offsets, int64_t n_atoms,
                     int64_t n_graphs, int32_t k, float cutoff_nm, int32_t* nlist, float* edges, float* inv_degree,
                     int mem, void* stream) {
  cudaStream_t s;
  int rc = begin_call(h, mem, stream, &s);
  if (rc) return rc;
  if (n_atoms < 0 || n_graphs < 0 || k < 1 || k > KNN_KMAX || !graph_offsets)
    return fail(h, NMRGNN_ERR_BAD_DIMS, "bad arguments (k must be 1..%d)", KNN_KMAX);
  if (n_atoms == 0 || n_graphs == 0) return NMRGNN_OK;
  if (!positions || !nlist || !edges || !inv_degree) return fail(h, NMRGNN_ERR_BAD_DIMS, "null buffer");
  if (n_atoms >= ((int64_t)1 << 31) || n_graphs > 2147483647) return fail(h, NMRGNN_ERR_BAD_DIMS, "batch too large");
  int64_t max_n = 0;
  if (graph_offsets[0] != 0 || graph_offsets[n_graphs] != n_atoms) return fail(h, NMRGNN_ERR_BAD_DIMS, "graph_offsets must span [0, n_atoms]");
  for (int64_t g = 0; g < n_graphs; ++g) {
    const int64_t n = graph_offsets[g + 1] - graph_offsets[g];
    if (n < 0) return fail(h, NMRGNN_ERR_BAD_DIMS, "graph_offsets must be non-decreasing");
    if (n > max_n) max_n = n;
  }
  if (max_n == 0) return NMRGNN_OK;
  const int64_t chunks = (max_n + KNN_QPB - 1) / KNN_QPB;
  if (chunks > 65535) return fail(h, NMRGNN_ERR_BAD_DIMS, "graph of %lld atoms is too large for the kNN builder", (long long)max_n);
  Io io{h, mem, s};
  const void* d_pos;
  void *d_nl, *d_ed, *d_inv;
  if ((rc = io.in(positions, n_atoms * 3 * sizeof(float), h->pos, &d_pos))) return rc;
  {
    const void* before = h->offs.p;
    if ((rc = ensure(h, h->offs, (n_graphs + 1) * sizeof(int64_t)))) return rc;
    if (h->offs.p != before) h->offs_cached.clear();
  }
  if (h->offs_cached.size() != (size_t)(n_graphs + 1) ||
      std::memcmp(h->offs_cached.data(), graph_offsets, (n_graphs + 1) * sizeof(int64_t)) != 0) {
    // pageable host source: the runtime stages it before returning, so the caller may free it
    h->offs_cached.assign(graph_offsets, graph_offsets + n_graphs + 1);
    CUDA_TRY(h, cudaMemcpyAsync(h->offs.p, h->offs_cached.data(), (n_graphs + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    CUDA_TRY(h, cudaStreamSynchronize(s));   // different streams may use the cached copy from now on
  }
  if ((rc = io.out_buf(nlist, n_atoms * k * sizeof(int32_t), h->nlist, &d_nl))) return rc;
  if ((rc = io.out_buf(edges, n_atoms * k * sizeof(float), h->edges, &d_ed))) return rc;
  if ((rc = io.out_buf(inv_degree, n_atoms * sizeof(float), h->invdeg, &d_inv))) return rc;
  KnnArgs a{};
  a.pos = (const float*)d_pos;
  a.offsets = (const int64_t*)h->offs.p;
  a.nlist = (int32_t*)d_nl;
  a.edges = (float*)d_ed;
  a.inv_degree = (float*)d_inv;
  a.k = k;
  a.cutoff2 = cutoff_nm > 0.f ? cutoff_nm * cutoff_nm : 0.f;
  if (h->knn_cells_on) {
    if ((rc = ensure(h, h->knn_sorted, n_atoms * sizeof(float4)))) return rc;
    if ((rc = ensure(h, h->knn_cells, n_graphs * (KNN_MAX_CELLS + 1) * sizeof(uint32_t)))) return rc;
    if ((rc = ensure(h, h->knn_grid, n_graphs * sizeof(KnnGrid)))) return rc;
    KnnCellArgs b{};
    b.pos = a.pos;
    b.offsets = a.offsets;
    b.sorted = (float4*)h->knn_sorted.p;
    b.cell_start = (uint32_t*)h->knn_cells.p;
    b.grid = (KnnGrid*)h->knn_grid.p;
    b.k = k;
    knn_build_cells_kernel<<<(unsigned)n_graphs, KNN_BUILD_THREADS, 0, s>>>(b);
    KnnQueryArgs qa{};
    qa.out = a;
    qa.sorted = b.sorted;
    qa.cell_start = b.cell_start;
    qa.grid = b.grid;
    const int64_t chunks_w = (max_n + KNN_WARP_QPB - 1) / KNN_WARP_QPB;
    if (h->knn_warp && chunks_w <= 65535)      // one warp per query atom (grid.y limit: graphs up to 262 k atoms)
      knn_query_warp_kernel<<<dim3((unsigned)n_graphs, (unsigned)chunks_w), KNN_THREADS, 0, s>>>(qa);
    else
      knn_query_cells_kernel<<<dim3((unsigned)n_graphs, (unsigned)chunks), KNN_THREADS, 0, s>>>(qa);
    h->launches += 2;
  } else {
    knn_graph_kernel<<<dim3((unsigned)n_graphs, (unsigned)chunks), KNN_THREADS, 0, s>>>(a);
    h->launches++;
  }
  if ((rc = io.finish(nlist, d_nl, n_atoms * k * sizeof(int32_t)))) return rc;
  if ((rc = io.finish(edges, d_ed, n_atoms * k * sizeof(float)))) return rc;
  if ((rc = io.finish(inv_degree, d_inv, n_atoms * sizeof(float)))) return rc;
  return end_call(h, stream, s);
}

}  // extern "C"
