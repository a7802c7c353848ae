// Tensor-core (tcgen05, kind::tf32, 3xTF32 split) kernels of the GNN forward, sm_100a.
// Accumulators live in tensor memory; operands are K-major 64-byte-swizzled tiles in
// shared memory: activations are written there by the epilogue/producer warps, weights
// arrive pre-split and pre-swizzled from global memory through 1-D bulk async copies.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace nmr {

// ----------------------------------------------------------------------------------
// Self-test: D[128 x 128] = A[128 x 64] * W[64 x 128] on one CTA.  mode 0: 3xTF32,
// mode 1: hi*hi only (1xTF32).  Exercises descriptors, swizzle, TMEM addressing and
// the bulk-copy / commit barriers in isolation.
// ----------------------------------------------------------------------------------
constexpr int ST_K = 64;
constexpr int ST_CHUNKS = ST_K / tc::BK;
constexpr size_t ST_SMEM = 1024 + (size_t)ST_CHUNKS * (8192 * 2 + 16384) + 256;

__global__ void __launch_bounds__(192, 1) tc_selftest_kernel(const float* __restrict__ A, const uint8_t* __restrict__ Bimg,
                                                             float* __restrict__ D, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_hi = smem;                                  // [chunks][128 rows x 64 B]
  uint8_t* a_lo = a_hi + ST_CHUNKS * 8192;
  uint8_t* b = a_lo + ST_CHUNKS * 8192;                  // [chunks][hi 8192 | lo 8192]
  uint64_t* bars = reinterpret_cast<uint64_t*>(b + ST_CHUNKS * 16384);
  uint64_t* b_full = bars;
  uint64_t* d_full = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    tc::mbar_init(b_full, 1);
    tc::mbar_init(d_full, 1);
    tc::mbar_fence_init();
  }
  if (warp == 4) tc::tmem_alloc<128>(tmem_slot);
  __syncthreads();
  if (warp == 5 && lane == 0) {
    tc::mbar_expect_tx(b_full, ST_CHUNKS * 16384);
    for (int c = 0; c < ST_CHUNKS; ++c) tc::bulk_g2s(b + c * 16384, Bimg + (size_t)c * 16384, 16384, b_full);
  }
  if (tid < 128) {
    const int r = tid;
    for (int c = 0; c < ST_CHUNKS; ++c) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 x = *reinterpret_cast<const float4*>(A + r * ST_K + c * tc::BK + j * 4);
        float4 hi, lo;
        tc::split4(x, hi, lo);
        const uint32_t off = c * 8192 + tc::sw64_chunk_offset(r, j);
        *reinterpret_cast<float4*>(a_hi + off) = hi;
        *reinterpret_cast<float4*>(a_lo + off) = lo;
      }
    }
    tc::fence_proxy_async();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 4 && lane == 0) {
    tc::mbar_wait(b_full, 0);
    tc::tc_fence_after();
    const uint32_t idesc = tc::make_idesc_tf32(128, 128);
    for (int c = 0; c < ST_CHUNKS; ++c) {
      const uint64_t ah = tc::make_desc_sw64(tc::smem_u32(a_hi + c * 8192));
      const uint64_t al = tc::make_desc_sw64(tc::smem_u32(a_lo + c * 8192));
      const uint64_t bh = tc::make_desc_sw64(tc::smem_u32(b + c * 16384));
      const uint64_t bl = tc::make_desc_sw64(tc::smem_u32(b + c * 16384 + 8192));
#pragma unroll
      for (int ks = 0; ks < tc::BK / tc::UMMA_K; ++ks) {
        const uint64_t adv = (uint64_t)(ks * tc::UMMA_K * 4) >> 4;
        tc::umma_tf32(tmem_base, ah + adv, bh + adv, idesc, (c | ks) != 0);
        if (mode == 0) {
          tc::umma_tf32(tmem_base, al + adv, bh + adv, idesc, 1);
          tc::umma_tf32(tmem_base, ah + adv, bl + adv, idesc, 1);
        }
      }
    }
    tc::umma_commit(d_full);
  }
  if (tid < 128) {
    tc::mbar_wait(d_full, 0);
    tc::tc_fence_after();
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 8; ++c) {
      float v[16];
      tc::tmem_ld16(lane_base + c * 16, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) D[tid * 128 + c * 16 + i] = v[i];
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc<128>(tmem_base);
}

// ----------------------------------------------------------------------------------
// Edge MLP on tensor cores.  One CTA = 128 edges per tile (persistent over tiles).
//   warps 0-3 : thread r owns edge r: RBF prologue, then per layer the epilogue
//               (TMEM -> +bias -> softplus -> hi/lo split -> next layer's A operand in smem)
//   warp 4    : MMA issuer (one lane): D[128 x 128] (+)= X_chunk * W_chunk^T, 3 products
//   warp 5    : weight loader (one lane): 16 KB bulk copies into a 4-slot ring
// The activation operand X (hi and lo, 8 chunks of 16 features) is resident and updated
// in place chunk by chunk; the MMA of layer l+1 starts on chunk c as soon as the epilogue
// of layer l has produced it, while the two TMEM accumulators ping-pong between layers.
// The last (linear, 128 -> E) layer is one more MMA with N = 16 (E padded).
// ----------------------------------------------------------------------------------
struct EdgeTcArgs {
  const float* edges;        // [n_edges]
  float* out;                // [n_edges, E]
  int64_t n_edges;
  const float* centers;      // [128]
  float gap;
  const uint8_t* Wimg;       // hidden layers: [n_hidden][8 chunks][hi 8192 | lo 8192]
  const uint8_t* Wfimg;      // final layer:   [8 chunks][hi 1024 | lo 1024]   (16 rows, rows >= E are zero)
  const float* bias;         // [n_hidden][128]
  const float* bias_f;       // [E]
  int n_hidden;              // hidden (activated) layers, >= 1
  int E;
  int act;
  const int32_t* nlist;      // optional validation
  int64_t n_atoms;
  int* err_flag;
};

constexpr int ETC_THREADS = 192;
constexpr int ETC_SLOTS = 4;
constexpr int ETC_CHUNKS = 8;          // 128 / 16
constexpr size_t ETC_SMEM = 1024 + 2 * 65536 + ETC_SLOTS * 16384 + ETC_CHUNKS * 2048 + (MAX_DENSE * 128 + 128 + 16) * 4 + 512;

__global__ void __launch_bounds__(ETC_THREADS, 1) edge_mlp_tc_kernel(const EdgeTcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* x_hi = smem;                                   // [8][8192]
  uint8_t* x_lo = x_hi + 65536;
  uint8_t* ring = x_lo + 65536;                           // [SLOTS][hi 8192 | lo 8192]
  uint8_t* wf = ring + ETC_SLOTS * 16384;                 // [8][hi 1024 | lo 1024]
  float* bias_s = reinterpret_cast<float*>(wf + ETC_CHUNKS * 2048);   // [n_hidden][128]
  float* cen_s = bias_s + MAX_DENSE * 128;                // [128]
  float* bf_s = cen_s + 128;                              // [16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bf_s + 16);
  uint64_t* w_full = bars;                                // [SLOTS]
  uint64_t* w_empty = w_full + ETC_SLOTS;                 // [SLOTS]
  uint64_t* x_ready = w_empty + ETC_SLOTS;                // [8]
  uint64_t* d_full = x_ready + ETC_CHUNKS;                // [2]
  uint64_t* df_full = d_full + 2;
  uint64_t* wf_full = df_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wf_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < ETC_SLOTS; ++i) {
      tc::mbar_init(&w_full[i], 1);
      tc::mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < ETC_CHUNKS; ++i) tc::mbar_init(&x_ready[i], 4);
    tc::mbar_init(&d_full[0], 1);
    tc::mbar_init(&d_full[1], 1);
    tc::mbar_init(df_full, 1);
    tc::mbar_init(wf_full, 1);
    tc::mbar_fence_init();
  }
  for (int i = tid; i < p.n_hidden * 128; i += ETC_THREADS) bias_s[i] = p.bias[i];
  for (int i = tid; i < 128; i += ETC_THREADS) cen_s[i] = p.centers[i];
  if (tid < 16) bf_s[tid] = tid < p.E ? p.bias_f[tid] : 0.0f;
  if (warp == 4) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t n_tiles = (p.n_edges + 127) / 128;
  const int n_hidden = p.n_hidden;

  if (warp == 5) {
    // ===================== weight loader =====================
    if (lane == 0) {
      tc::mbar_expect_tx(wf_full, ETC_CHUNKS * 2048);
      tc::bulk_g2s(wf, p.Wfimg, ETC_CHUNKS * 2048, wf_full);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int l = 0; l < n_hidden; ++l)
          for (int c = 0; c < ETC_CHUNKS; ++c, ++it) {
            const uint32_t slot = it % ETC_SLOTS, ph = (it / ETC_SLOTS) & 1;
            tc::mbar_wait(&w_empty[slot], ph ^ 1);
            tc::mbar_expect_tx(&w_full[slot], 16384);
            tc::bulk_g2s(ring + slot * 16384, p.Wimg + ((size_t)l * ETC_CHUNKS + c) * 16384, 16384, &w_full[slot]);
          }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_h = tc::make_idesc_tf32(128, 128);
      const uint32_t idesc_f = tc::make_idesc_tf32(128, 16);
      tc::mbar_wait(wf_full, 0);
      uint32_t it = 0, pass = 0;   // every (tile, layer) pass completes one phase of each x_ready[c]
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l <= n_hidden; ++l, ++pass) {
          const bool fin = (l == n_hidden);
          const uint32_t d_tmem = tmem_base + (fin ? 256u : (uint32_t)(l & 1) * 128u);
          for (int c = 0; c < ETC_CHUNKS; ++c) {
            tc::mbar_wait(&x_ready[c], pass & 1);
            uint64_t bh, bl;
            uint32_t slot = 0;
            if (!fin) {
              slot = it % ETC_SLOTS;
              tc::mbar_wait(&w_full[slot], (it / ETC_SLOTS) & 1);
              bh = tc::make_desc_sw64(tc::smem_u32(ring + slot * 16384));
              bl = tc::make_desc_sw64(tc::smem_u32(ring + slot * 16384 + 8192));
            } else {
              bh = tc::make_desc_sw64(tc::smem_u32(wf + c * 2048));
              bl = tc::make_desc_sw64(tc::smem_u32(wf + c * 2048 + 1024));
            }
            tc::tc_fence_after();
            const uint64_t ah = tc::make_desc_sw64(tc::smem_u32(x_hi + c * 8192));
            const uint64_t al = tc::make_desc_sw64(tc::smem_u32(x_lo + c * 8192));
            const uint32_t idesc = fin ? idesc_f : idesc_h;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t adv = (uint64_t)(ks * 2);
              tc::umma_tf32(d_tmem, ah + adv, bh + adv, idesc, (c | ks) != 0);
              tc::umma_tf32(d_tmem, al + adv, bh + adv, idesc, 1);
              tc::umma_tf32(d_tmem, ah + adv, bl + adv, idesc, 1);
            }
            if (!fin) {
              tc::umma_commit(&w_empty[slot]);
              ++it;
            }
          }
          tc::umma_commit(fin ? df_full : &d_full[l & 1]);
        }
      }
    }
  } else {
    // ===================== producer / epilogue warps (thread r = edge r) =====================
    const int r = tid;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t ph_d[2] = {0, 0}, ph_f = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t e = tile * 128 + r;
      float d = 0.0f;
      if (e < p.n_edges) {
        d = p.edges[e];
        if (p.nlist != nullptr) {
          const int32_t idx = p.nlist[e];
          if (idx < 0 || idx >= p.n_atoms) atomicOr(p.err_flag, 1);
        }
      }
      const bool m = d > 0.0f;
      // pass 0: RBF expansion * mask  (layers.py:137-140, model.py:251-257)
      for (int c = 0; c < ETC_CHUNKS; ++c) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 x;
          float* xv = reinterpret_cast<float*>(&x);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float diff = d - cen_s[c * 16 + j * 4 + i];
            const float v = expf(__fdiv_rn(-__fmul_rn(diff, diff), p.gap));
            xv[i] = m ? v : 0.0f;
          }
          float4 hi, lo;
          tc::split4(x, hi, lo);
          const uint32_t off = c * 8192 + tc::sw64_chunk_offset(r, j);
          *reinterpret_cast<float4*>(x_hi + off) = hi;
          *reinterpret_cast<float4*>(x_lo + off) = lo;
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&x_ready[c]);
      }
      // hidden layers: X <- act(D + b), in place, chunk by chunk
      for (int l = 0; l < n_hidden; ++l) {
        tc::mbar_wait(&d_full[l & 1], ph_d[l & 1]);
        ph_d[l & 1] ^= 1;
        tc::tc_fence_after();
        const float* bl = bias_s + l * 128;
        for (int c = 0; c < ETC_CHUNKS; ++c) {
          float v[16];
          tc::tmem_ld16(lane_base + (uint32_t)(l & 1) * 128u + c * 16, v);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 x;
            x.x = apply_act(v[j * 4 + 0] + bl[c * 16 + j * 4 + 0], p.act);
            x.y = apply_act(v[j * 4 + 1] + bl[c * 16 + j * 4 + 1], p.act);
            x.z = apply_act(v[j * 4 + 2] + bl[c * 16 + j * 4 + 2], p.act);
            x.w = apply_act(v[j * 4 + 3] + bl[c * 16 + j * 4 + 3], p.act);
            float4 hi, lo;
            tc::split4(x, hi, lo);
            const uint32_t off = c * 8192 + tc::sw64_chunk_offset(r, j);
            *reinterpret_cast<float4*>(x_hi + off) = hi;
            *reinterpret_cast<float4*>(x_lo + off) = lo;
          }
          tc::fence_proxy_async();
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&x_ready[c]);
        }
      }
      // final linear layer (already in TMEM columns 256..271) * mask
      tc::mbar_wait(df_full, ph_f);
      ph_f ^= 1;
      tc::tc_fence_after();
      float v[8];
      tc::tmem_ld8(lane_base + 256u, v);
      tc::tc_fence_before();
      if (e < p.n_edges) {
        for (int n = 0; n < p.E; ++n) p.out[e * p.E + n] = m ? v[n] + bf_s[n] : 0.0f;
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc<512>(tmem_base);
}

}  // namespace nmr
