// Tensor-core (tcgen05) kernels of the GNN forward, sm_100a.
// Production scheme: FP16 scaled-split products (tc_common.cuh, "fp16x3") with separate
// main / correction accumulators in tensor memory; operands are K-major 64-byte-swizzled
// tiles in shared memory: activations are written there by the producer / epilogue warps,
// weights arrive pre-split and pre-swizzled from global memory through 1-D bulk async
// copies (UBLKCP) into mbarrier-guarded rings.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace nmr {

// ----------------------------------------------------------------------------------
// Self-tests: D[128 x 128] = A[128 x 64] * W[64 x 128] on one CTA.
//   tc_selftest_kernel      kind::tf32; mode 0: 3xTF32 (single accumulator), mode 1: 1xTF32
//   tc_selftest_f16_kernel  kind::f16;  mode 2: fp16x3 (main + corr accumulators), mode 3: hi*hi
// They exercise descriptors, swizzle, TMEM addressing and the bulk-copy / commit barriers
// in isolation, and measure the accumulation behaviour of the tensor core.
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

constexpr int STH_CHUNKS = ST_K / tc::HK;   // 2
constexpr size_t STH_SMEM = 1024 + (size_t)STH_CHUNKS * (8192 * 2 + 16384) + 256;

__global__ void __launch_bounds__(192, 1) tc_selftest_f16_kernel(const float* __restrict__ A,
                                                                 const uint8_t* __restrict__ Bimg,
                                                                 float* __restrict__ D, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_hi = smem;                                  // [chunks][128 rows x 64 B]
  uint8_t* a_lo = a_hi + STH_CHUNKS * 8192;
  uint8_t* b = a_lo + STH_CHUNKS * 8192;                 // [chunks][hi 8192 | lo 8192]
  uint64_t* bars = reinterpret_cast<uint64_t*>(b + STH_CHUNKS * 16384);
  uint64_t* b_full = bars;
  uint64_t* d_full = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    tc::mbar_init(b_full, 1);
    tc::mbar_init(d_full, 1);
    tc::mbar_fence_init();
  }
  if (warp == 4) tc::tmem_alloc<256>(tmem_slot);
  __syncthreads();
  if (warp == 5 && lane == 0) {
    tc::mbar_expect_tx(b_full, STH_CHUNKS * 16384);
    for (int c = 0; c < STH_CHUNKS; ++c) tc::bulk_g2s(b + c * 16384, Bimg + (size_t)c * 16384, 16384, b_full);
  }
  if (tid < 128) {
    const int r = tid;
    for (int c = 0; c < STH_CHUNKS; ++c) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x[8];
        const float4 x0 = *reinterpret_cast<const float4*>(A + r * ST_K + c * tc::HK + j * 8);
        const float4 x1 = *reinterpret_cast<const float4*>(A + r * ST_K + c * tc::HK + j * 8 + 4);
        x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w;
        x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
        uint4 hi, lo;
        tc::split8_f16(x, hi, lo);
        const uint32_t off = c * 8192 + tc::sw64_chunk_offset(r, j);
        *reinterpret_cast<uint4*>(a_hi + off) = hi;
        *reinterpret_cast<uint4*>(a_lo + off) = lo;
      }
    }
    tc::fence_proxy_async();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t d_main = tmem_base, d_corr = tmem_base + 128;
  if (warp == 4 && lane == 0) {
    tc::mbar_wait(b_full, 0);
    tc::tc_fence_after();
    const uint32_t idesc = tc::make_idesc_f16(128, 128);
    for (int c = 0; c < STH_CHUNKS; ++c) {
      const uint64_t ah = tc::make_desc_sw64(tc::smem_u32(a_hi + c * 8192));
      const uint64_t al = tc::make_desc_sw64(tc::smem_u32(a_lo + c * 8192));
      const uint64_t bh = tc::make_desc_sw64(tc::smem_u32(b + c * 16384));
      const uint64_t bl = tc::make_desc_sw64(tc::smem_u32(b + c * 16384 + 8192));
#pragma unroll
      for (int ks = 0; ks < tc::HK / tc::UMMA_K_F16; ++ks) {
        const uint64_t adv = (uint64_t)(ks * tc::UMMA_K_F16 * 2) >> 4;
        tc::umma_f16(d_main, ah + adv, bh + adv, idesc, (c | ks) != 0);
        if (mode == 2) {
          tc::umma_f16(d_corr, al + adv, bh + adv, idesc, (c | ks) != 0);
          tc::umma_f16(d_corr, ah + adv, bl + adv, idesc, 1);
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
      if (mode == 2) tc::tmem_ld16_combined(lane_base + c * 16, lane_base + 128 + c * 16, v);
      else tc::tmem_ld16(lane_base + c * 16, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) D[tid * 128 + c * 16 + i] = v[i];
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc<256>(tmem_base);
}

// mode 4: A operand (fp16(A), single product) from tensor memory (TS form), B from shared memory.
__global__ void __launch_bounds__(192, 1) tc_selftest_ts_kernel(const float* __restrict__ A, const uint8_t* __restrict__ Bimg,
                                                                float* __restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b = smem;                                     // [chunks][hi 8192 | lo 8192]
  uint64_t* bars = reinterpret_cast<uint64_t*>(b + STH_CHUNKS * 16384);
  uint64_t* b_full = bars;
  uint64_t* d_full = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    tc::mbar_init(b_full, 1);
    tc::mbar_init(d_full, 1);
    tc::mbar_fence_init();
  }
  if (warp == 4) tc::tmem_alloc<256>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t d_main = tmem_base, a_cols = tmem_base + 128;   // A: 64 k = 32 columns
  if (warp == 5 && lane == 0) {
    tc::mbar_expect_tx(b_full, STH_CHUNKS * 16384);
    for (int c = 0; c < STH_CHUNKS; ++c) tc::bulk_g2s(b + c * 16384, Bimg + (size_t)c * 16384, 16384, b_full);
  }
  if (tid < 128) {
    const uint32_t lane_base = a_cols + ((uint32_t)(warp * 32) << 16);
    for (int g = 0; g < 4; ++g) {            // 8 columns = 16 k per store
      uint32_t r[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 x = *reinterpret_cast<const float2*>(A + tid * ST_K + g * 16 + i * 2);
        const __half2 hh = __floats2half2_rn(x.x, x.y);
        r[i] = *reinterpret_cast<const uint32_t*>(&hh);
      }
      tc::tmem_st8(lane_base + g * 8, r);
    }
    tc::tmem_st_wait();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (warp == 4 && lane == 0) {
    tc::mbar_wait(b_full, 0);
    tc::tc_fence_after();
    const uint32_t idesc = tc::make_idesc_f16(128, 128);
    for (int c = 0; c < STH_CHUNKS; ++c) {
      const uint64_t bh = tc::make_desc_sw64(tc::smem_u32(b + c * 16384));
#pragma unroll
      for (int ks = 0; ks < tc::HK / tc::UMMA_K_F16; ++ks) {
        const uint64_t adv = (uint64_t)(ks * tc::UMMA_K_F16 * 2) >> 4;
        tc::umma_f16_ts(d_main, a_cols + (uint32_t)(c * 2 + ks) * 8u, bh + adv, idesc, (c | ks) != 0);
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
  if (warp == 4) tc::tmem_dealloc<256>(tmem_base);
}

// modes 5 / 6: CTA pair (cta_group::2), single product fp16(A) * fp16(W): one M = 256, N = 128 instruction per K-step
// over a cluster of two CTAs.  CTA r stages its own 128 rows of A (the peer's copy of A is rotated by one row so
// that the test can tell the two apart) and N rows [64 r, 64 r + 64) of W; the peer reports "operands ready" on the
// leader's barrier, the leader issues the MMAs and the commit is multicast to both CTAs.  mode 5 returns the leader's
// 128 x 128 block of D, mode 6 the peer's.
constexpr size_t STP_SMEM = 1024 + (size_t)STH_CHUNKS * (8192 + 4096) + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
    tc_selftest_pair_kernel(const float* __restrict__ A, const uint8_t* __restrict__ Bimg, float* __restrict__ D, int want_rank) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_hi = smem;                                  // [chunks][128 rows x 64 B]
  uint8_t* b = a_hi + STH_CHUNKS * 8192;                 // [chunks][64 rows x 64 B]: this CTA's half of W (hi image)
  uint64_t* bars = reinterpret_cast<uint64_t*>(b + STH_CHUNKS * 4096);
  uint64_t* b_full = bars;
  uint64_t* d_full = bars + 1;
  uint64_t* peer_ready = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = tc::cluster_ctarank();
  if (tid == 0) {
    tc::mbar_init(b_full, 1);
    tc::mbar_init(d_full, 1);
    tc::mbar_init(peer_ready, 1);
    tc::mbar_fence_init();
  }
  if (warp == 4) tc::tmem_alloc_pair<128>(tmem_slot);
  tc::tc_fence_before();
  tc::cluster_sync();        // barriers of both CTAs exist before any remote arrive / multicast commit
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 5 && lane == 0) {
    tc::mbar_expect_tx(b_full, STH_CHUNKS * 4096);
    for (int c = 0; c < STH_CHUNKS; ++c) tc::bulk_g2s(b + c * 4096, Bimg + (size_t)c * 16384 + rank * 4096, 4096, b_full);
  }
  if (tid < 128) {
    const int r = tid;
    const int src = (r + (int)rank) & 127;
    for (int c = 0; c < STH_CHUNKS; ++c) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x[8];
        const float4 x0 = *reinterpret_cast<const float4*>(A + src * ST_K + c * tc::HK + j * 8);
        const float4 x1 = *reinterpret_cast<const float4*>(A + src * ST_K + c * tc::HK + j * 8 + 4);
        x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w;
        x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
        uint4 hi, lo;
        tc::split8_f16(x, hi, lo);
        *reinterpret_cast<uint4*>(a_hi + c * 8192 + tc::sw64_chunk_offset(r, j)) = hi;
      }
    }
    tc::fence_proxy_async();
  }
  __syncthreads();
  if (rank == 1 && warp == 5 && lane == 0) {
    tc::mbar_wait(b_full, 0);
    tc::mbar_arrive_cluster(tc::map_to_cta(tc::smem_u32(peer_ready), 0));
  }
  if (rank == 0 && warp == 4 && lane == 0) {
    tc::mbar_wait(b_full, 0);
    tc::mbar_wait_cluster(peer_ready, 0);
    tc::tc_fence_after();
    const uint32_t idesc = tc::make_idesc_f16(256, 128);
    for (int c = 0; c < STH_CHUNKS; ++c) {
      const uint64_t ah = tc::make_desc_sw64(tc::smem_u32(a_hi + c * 8192));
      const uint64_t bh = tc::make_desc_sw64(tc::smem_u32(b + c * 4096));
#pragma unroll
      for (int ks = 0; ks < tc::HK / tc::UMMA_K_F16; ++ks) {
        const uint64_t adv = (uint64_t)(ks * tc::UMMA_K_F16 * 2) >> 4;
        tc::umma_f16_pair(tmem_base, ah + adv, bh + adv, idesc, (c | ks) != 0);
      }
    }
    tc::umma_commit_pair(d_full, 3);
  }
  if (tid < 128) {
    tc::mbar_wait(d_full, 0);
    tc::tc_fence_after();
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 8; ++c) {
      float v[16];
      tc::tmem_ld16(lane_base + c * 16, v);
      if ((int)rank == want_rank) {
#pragma unroll
        for (int i = 0; i < 16; ++i) D[tid * 128 + c * 16 + i] = v[i];
      }
    }
    tc::tc_fence_before();
  }
  tc::cluster_sync();        // both CTAs have drained their accumulators
  if (warp == 4) tc::tmem_dealloc_pair<128>(tmem_base);
}

// ----------------------------------------------------------------------------------
// Edge MLP on tensor cores (RBF -> EdgeFCBlock -> mask; model.py:251-261).
// One CTA keeps TWO 128-edge tiles in flight (slots 0/1) so that the CUDA-core epilogue
// of one tile (bias, softplus, split -> next layer's A operand) overlaps the MMAs of the
// other.  Per slot: X (activation operand, hi+lo, 4 K-chunks of 32) in shared memory,
// main/corr accumulators (2 x 128 columns) in tensor memory.
//   warp 0       : weight loader (one lane): 16 KB bulk copies into a 4-slot ring
//   warp 1       : MMA issuer (one lane) + TMEM owner
//   warps 2..17  : epilogue; all 16 warps work on slot 0, then slot 1, then slot 0 ... :
//                  warp%4 selects the TMEM lane quarter (32 edges), (warp-2)/4 the 32-column
//                  K-chunk: thread = (edge row, 32 features).
// The last (linear, 128 -> E, E <= 3) layer is not a tensor-core layer (it would be 16 near-empty
// instructions of ~105 cycles per tile): the epilogue of the last hidden layer forms its partial dot
// products from registers in FP32 and the four column quarters meet in shared memory.  (The TS-form
// kernel below still runs it as an N = 16 MMA against the resident final-layer image.)
// ----------------------------------------------------------------------------------
struct EdgeTcArgs {
  const float* edges;        // [n_edges]
  float* out;                // optional [n_edges, E]
  float4* rec;               // optional [n_edges] {e0, e1, e2, bits(idx)} records for the MP kernel (E <= 3)
  int64_t n_edges;
  const float* centers;      // [128]
  float gap;
  float rbf_c;               // -log2(e) / gap
  const uint8_t* Wimg;       // hidden layers: [n_hidden][4 chunks][hi 8192 | lo 8192]
  const uint8_t* Wfimg;      // final layer:   [4 chunks][hi 1024 | lo 1024]   (16 rows, rows >= E are zero)
  const float* bias;         // [n_hidden][128]
  const float* bias_f;       // [E]
  const float* Wf;           // final layer, fp32 [128][E] (the SS kernel evaluates it on the CUDA cores)
  float in_scale[MAX_DENSE + 1];   // power of two applied to the INPUT of layer l (a-priori range bound)
  float out_scale[MAX_DENSE + 1];  // its inverse, applied to the accumulator of layer l
  int n_hidden;              // hidden (activated) layers, >= 1
  int E;
  int act;
  const int32_t* nlist;      // optional [n_edges]: validated against n_atoms (and packed into rec)
  int64_t n_atoms;
  int* err_flag;
  long long* dbg;            // optional [gridDim][8] cycle counters (diagnostics)
  int rec_k;                 // K (8 or 16) if the records are slot-swizzled for the MP kernel (rec_slot), else 0
  int64_t rec_e0;            // edge index of this launch's first edge within the whole batch (chunked launches)
};

// Edge records of one atom are stored slot-swizzled: logical neighbour j of atom i sits in slot j ^ ((i & 3) << 1).
// The MP producers process 4 consecutive atoms per warp and read the same logical slot of all four in one shared
// memory instruction; with a 256-byte row stride those four 16-byte reads would hit the same banks (4-way conflict,
// measured: 2/3 of the kernel's L1 wavefronts).  The swizzle spreads them over 4 bank groups while every atom still
// accumulates its neighbours in logical order (bit-identical results, independent of the atom's position in a batch).
__host__ __device__ __forceinline__ int64_t rec_slot(int64_t e, int K) {   // K = 8 or 16
  const int sh = K == 16 ? 4 : 3;
  const int64_t atom = e >> sh;
  return e ^ (int64_t)((atom & 3) << 1);      // flips bits 1..2 of the slot index j = e & (K - 1)
}

constexpr int ETC_THREADS = 576;
constexpr int ETC_RING = 4;
constexpr int ETC_CHUNKS = 4;          // 128 / 32
constexpr int ETC_X_BYTES = ETC_CHUNKS * 16384;
constexpr size_t ETC_SMEM = 1024 + 2 * ETC_X_BYTES + ETC_RING * 16384 + 128 * 16 + 2 * 4 * 128 * 16 +
                            (MAX_DENSE * 128 + 128 + 16) * 4 + 512;

// SPLIT (option "edge_split"): the 16 epilogue warps form two groups of 8, one per slot, instead of all working on
// slot 0, then slot 1.  With every warp in the same phase at the same time the MUFU pipe (2 per element, 16 per
// clock and SM: 2.05 k cycles per tile and layer, the floor of this kernel) idles while all of them load
// accumulators or split / store operands; two groups half a period apart fill those gaps with each other's
// softplus.  Thread = (edge row, 64 features); same arithmetic in the same order (bit-identical output).
template <int ACT, bool SPLIT>
__device__ __forceinline__ void edge_mlp_tc_body(const EdgeTcArgs& p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xs = smem;                                     // [2 slots][4 chunks][hi 8192 | lo 8192]
  uint8_t* ring = xs + 2 * ETC_X_BYTES;                   // [RING][hi 8192 | lo 8192]
  float4* wf4 = reinterpret_cast<float4*>(ring + ETC_RING * 16384);   // [128] final-layer weights {w[k][0..2], 0}
  float4* part = wf4 + 128;                               // [2 slots][4 column quarters][128 rows] partial outputs
  float* bias_s = reinterpret_cast<float*>(part + 2 * 4 * 128);        // [n_hidden][128]
  float* cen_s = bias_s + MAX_DENSE * 128;                // [128]
  float* bf_s = cen_s + 128;                              // [16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bf_s + 16);
  uint64_t* w_full = bars;                                // [RING]
  uint64_t* w_empty = w_full + ETC_RING;                  // [RING]
  uint64_t* x_full = w_empty + ETC_RING;                  // [2]
  uint64_t* d_full = x_full + 2;                          // [2]
  uint64_t* wf_full = d_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wf_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < ETC_RING; ++i) {
      tc::mbar_init(&w_full[i], 1);
      tc::mbar_init(&w_empty[i], 1);
    }
    for (int g = 0; g < 2; ++g) {
      tc::mbar_init(&x_full[g], SPLIT ? 8 : 16);
      tc::mbar_init(&d_full[g], 1);
    }
    tc::mbar_init(wf_full, 1);
    tc::mbar_fence_init();
  }
  for (int i = tid; i < p.n_hidden * 128; i += ETC_THREADS) bias_s[i] = p.bias[i];
  for (int i = tid; i < 128; i += ETC_THREADS) cen_s[i] = p.centers[i];
  if (tid < 16) bf_s[tid] = tid < p.E ? p.bias_f[tid] : 0.0f;
  for (int i = tid; i < 128; i += ETC_THREADS)
    wf4[i] = make_float4(p.Wf[i * p.E], p.E > 1 ? p.Wf[i * p.E + 1] : 0.0f, p.E > 2 ? p.Wf[i * p.E + 2] : 0.0f,
                         p.E > 3 ? p.Wf[i * p.E + 3] : 0.0f);
  if (warp == 1) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t n_tiles = (p.n_edges + 127) / 128;
  const int n_hidden = p.n_hidden;
  // tiles of this CTA: blockIdx.x + n * gridDim.x, n = 0 .. n_my-1; tile n lives in slot n & 1
  const int n_my = blockIdx.x < n_tiles ? (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

  if (warp == 0) {
    // ===================== weight loader =====================
    if (lane == 0 && n_my > 0) {
      uint32_t it = 0;
      for (int pair = 0; pair * 2 < n_my; ++pair)
        for (int l = 0; l < n_hidden; ++l)
          for (int g = 0; g < 2; ++g) {
            if (pair * 2 + g >= n_my) continue;
            for (int c = 0; c < ETC_CHUNKS; ++c, ++it) {
              const uint32_t slot = it % ETC_RING, ph = (it / ETC_RING) & 1;
              tc::mbar_wait(&w_empty[slot], ph ^ 1);
              tc::mbar_expect_tx(&w_full[slot], 16384);
              tc::bulk_g2s(ring + slot * 16384, p.Wimg + ((size_t)l * ETC_CHUNKS + c) * 16384, 16384, &w_full[slot]);
            }
          }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && n_my > 0) {
      // A tcgen05.mma costs >= ~105 cycles however small it is (tools/microbench/mma_rate.cu), so the two products
      // that share the A operand are issued as ONE instruction: the hi and lo weight tiles are adjacent in the
      // image (256 rows) and main / corr adjacent in tensor memory, i.e. x_hi * [w_hi | w_lo] -> [main | corr].
      // For the same reason the last (128 -> E, E <= 3) layer is NOT a tensor-core layer: it would be 16
      // near-empty instructions per tile; the epilogue warps evaluate it in exact FP32 (below).
      const uint32_t idesc_h2 = tc::make_idesc_f16(128, 256), idesc_h = tc::make_idesc_f16(128, 128);
      uint32_t it = 0, px[2] = {0, 0};
      long long w_x = 0, w_w = 0, c0 = 0;
      const long long k0 = clock64();
      for (int pair = 0; pair * 2 < n_my; ++pair)
        for (int l = 0; l < n_hidden; ++l)
          for (int g = 0; g < 2; ++g) {
            if (pair * 2 + g >= n_my) continue;
            const uint32_t d_main = tmem_base + (uint32_t)g * 256u, d_corr = d_main + 128u;
            if (p.dbg) c0 = clock64();
            tc::mbar_wait(&x_full[g], px[g]);
            if (p.dbg) w_x += clock64() - c0;
            px[g] ^= 1;
            tc::tc_fence_after();
            for (int c = 0; c < ETC_CHUNKS; ++c, ++it) {
              const uint32_t slot = it % ETC_RING;
              if (p.dbg) c0 = clock64();
              tc::mbar_wait(&w_full[slot], (it / ETC_RING) & 1);
              if (p.dbg) w_w += clock64() - c0;
              tc::tc_fence_after();
              const uint64_t bh = tc::make_desc_sw64(tc::smem_u32(ring + slot * 16384));
              const uint8_t* xc = xs + g * ETC_X_BYTES + c * 16384;
              const uint64_t ah = tc::make_desc_sw64(tc::smem_u32(xc));
              const uint64_t al = tc::make_desc_sw64(tc::smem_u32(xc + 8192));
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 2);
                // [main | corr] (+)= x_hi * [w_hi | w_lo], then corr += x_lo * w_hi
                tc::umma_f16(d_main, ah + adv, bh + adv, idesc_h2, (c | ks) != 0);
                tc::umma_f16(d_corr, al + adv, bh + adv, idesc_h, 1);
              }
              tc::umma_commit(&w_empty[slot]);
            }
            tc::umma_commit(&d_full[g]);
          }
      if (p.dbg) {
        long long* o = p.dbg + (size_t)blockIdx.x * 8;
        o[0] = clock64() - k0;   // MMA thread total
        o[1] = w_x;              //   waiting for the epilogue warps (x_full)
        o[2] = w_w;              //   waiting for W (w_full)
      }
    }
  } else if (SPLIT) {
    // ===================== epilogue warps, one group of 8 per slot: thread = (edge row, 64 features) =====================
    const int we = warp - 2;                 // 0..15
    const int g = we >> 3;                   // this group's slot; its tiles are n = g, g + 2, g + 4, ...
    const int q = warp & 3;                  // TMEM lane quarter (fixed by the hardware: warp % 4)
    const int ch = (we & 7) >> 2;            // column half: K-chunks 2 ch and 2 ch + 1
    const int row = q * 32 + lane;
    const uint32_t xs_a = tc::smem_u32(xs) + (uint32_t)g * ETC_X_BYTES;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)g * 256u;
    uint32_t pd = 0;
    for (int tn = g; tn < n_my; tn += 2) {
      // ---- RBF expansion * mask (layers.py:137-140, model.py:251-257) -> operand X of this slot
      const int64_t tile = blockIdx.x + (int64_t)tn * gridDim.x;
      const int64_t e = tile * 128 + row;
      float d = 0.0f;
      int32_t idx = 0;
      if (e < p.n_edges) {
        d = __ldg(p.edges + e);
        if (p.nlist != nullptr && ch == 0) {
          idx = __ldg(p.nlist + e);
          if (idx < 0 || idx >= p.n_atoms) {
            atomicOr(p.err_flag, 1);
            idx = 0;
          }
        }
      }
      {
        const bool m = d > 0.0f;
        const float s_in = m ? p.in_scale[0] : 0.0f;
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const int cq = ch * 2 + c2, col0 = cq * 32;
          const uint32_t xg = xs_a + (uint32_t)cq * 16384u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float diff = d - cen_s[col0 + j * 8 + i];
              x[i] = tc::ex2_approx(diff * diff * p.rbf_c) * s_in;
            }
            uint4 hi, lo;
            tc::split8_f16(x, hi, lo);
            const uint32_t off = xg + tc::sw64_chunk_offset(row, j);
            tc::sts128(off, hi);
            tc::sts128(off + 8192u, lo);
          }
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&x_full[g]);
      }
      for (int l = 0; l < n_hidden; ++l) {
        const bool last = l == n_hidden - 1;
        const float s_out = p.out_scale[l], s_in = p.in_scale[l + 1];
        const long long q0 = p.dbg ? clock64() : 0;
        tc::mbar_wait(&d_full[g], pd);
        if (p.dbg && warp == 2 && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + (size_t)blockIdx.x * 8 + 3), (unsigned long long)(clock64() - q0));
        pd ^= 1;
        tc::tc_fence_after();
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const int cq = ch * 2 + c2, col0 = cq * 32;
          const uint32_t bl_a = tc::smem_u32(bias_s + l * 128 + col0);
          const uint32_t wf_a = tc::smem_u32(wf4 + col0);
          const uint32_t t_main = t_lane + col0, t_corr = t_main + 128u;
          const uint32_t xg = xs_a + (uint32_t)cq * 16384u;
          float o0 = 0.0f, o1 = 0.0f, o2 = 0.0f, o3 = 0.0f;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            float v[16];
            tc::tmem_ld16_combined(t_main + cc * 16, t_corr + cc * 16, v);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const float4 b0 = tc::lds128(bl_a + cc * 64 + hh * 32), b1 = tc::lds128(bl_a + cc * 64 + hh * 32 + 16);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              float x[8];
              if (!last) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  x[i] = act_t<ACT>(fmaf(v[hh * 8 + i], s_out, bb[i])) * s_in;
                uint4 hi, lo;
                tc::split8_f16(x, hi, lo);
                const uint32_t off = xg + tc::sw64_chunk_offset(row, cc * 2 + hh);
                tc::sts128(off, hi);
                tc::sts128(off + 8192u, lo);
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float y = act_t<ACT>(fmaf(v[hh * 8 + i], s_out, bb[i]));
                  const float4 w = tc::lds128(wf_a + (uint32_t)(cc * 16 + hh * 8 + i) * 16u);
                  o0 = fmaf(y, w.x, o0);
                  o1 = fmaf(y, w.y, o1);
                  o2 = fmaf(y, w.z, o2);
                  o3 = fmaf(y, w.w, o3);
                }
              }
            }
          }
          if (last) part[(g * 4 + cq) * 128 + row] = make_float4(o0, o1, o2, o3);
        }
        if (!last) {
          tc::fence_proxy_async();
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&x_full[g]);
        } else {
          // final linear layer * mask (model.py:128,261): sum the four column quarters in the same fixed order
          tc::tc_fence_before();
          if (g == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
          else asm volatile("bar.sync 2, 256;" ::: "memory");
          if (ch == 0 && e < p.n_edges) {
            const float4 p0 = part[(g * 4 + 0) * 128 + row], p1 = part[(g * 4 + 1) * 128 + row];
            const float4 p2 = part[(g * 4 + 2) * 128 + row], p3 = part[(g * 4 + 3) * 128 + row];
            const bool m = d > 0.0f;
            float o[4];
            o[0] = m ? ((p0.x + p1.x) + (p2.x + p3.x)) + bf_s[0] : 0.0f;
            o[1] = m ? ((p0.y + p1.y) + (p2.y + p3.y)) + bf_s[1] : 0.0f;
            o[2] = m ? ((p0.z + p1.z) + (p2.z + p3.z)) + bf_s[2] : 0.0f;
            o[3] = m ? ((p0.w + p1.w) + (p2.w + p3.w)) + bf_s[3] : 0.0f;
            if (p.out != nullptr)
              for (int i = 0; i < p.E; ++i) p.out[e * p.E + i] = o[i];
            if (p.rec != nullptr) {
              const int64_t er = p.rec_k ? rec_slot(p.rec_e0 + e, p.rec_k) - p.rec_e0 : e;
              p.rec[er] = make_float4(o[0], o[1], o[2], __int_as_float(idx));
            }
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps: thread = (edge row, 32 features = one K-chunk) =====================
    // All 16 warps work on one slot at a time (slot 0, slot 1, slot 0, ...), so that the MMAs of one slot
    // always overlap the CUDA-core work of the other and no warp group waits for its own slot's MMAs.
    const int we = warp - 2;                 // 0..15
    const int q = warp & 3;                  // TMEM lane quarter
    const int cq = we >> 2;                  // column quarter = K-chunk written by this thread
    const int row = q * 32 + lane;
    const int col0 = cq * 32;
    const uint32_t xs_a = tc::smem_u32(xs);
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t pd[2] = {0, 0};
    // pass 0 of a tile: RBF expansion * mask  (layers.py:137-140, model.py:251-257) -> operand X of slot g.
    // It is issued for the NEXT pair's tile right after the slot's last epilogue of the current pair, so the MMA
    // thread always has the other slot's (or the next tile's) layer to run while the CUDA cores work.
    float dd[2] = {0.0f, 0.0f};
    int32_t idxs[2] = {0, 0};
    int64_t eidx[2] = {0, 0};
    auto rbf_phase = [&](int g, int tile_n) {
      if (tile_n >= n_my) return;
      const int64_t tile = blockIdx.x + (int64_t)tile_n * gridDim.x;
      const int64_t e = tile * 128 + row;
      eidx[g] = e;
      float d = 0.0f;
      int32_t idx = 0;
      if (e < p.n_edges) {
        d = __ldg(p.edges + e);
        if (p.nlist != nullptr && cq == 0) {
          idx = __ldg(p.nlist + e);
          if (idx < 0 || idx >= p.n_atoms) {
            atomicOr(p.err_flag, 1);
            idx = 0;
          }
        }
      }
      idxs[g] = idx;
      dd[g] = d;
      const bool m = d > 0.0f;
      const float s_in = m ? p.in_scale[0] : 0.0f;
      const uint32_t xg = xs_a + (uint32_t)g * ETC_X_BYTES + (uint32_t)cq * 16384u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float diff = d - cen_s[col0 + j * 8 + i];
          // exp(-(d - mu)^2 / gap) = 2^(diff^2 * (-log2(e) / gap))
          x[i] = tc::ex2_approx(diff * diff * p.rbf_c) * s_in;
        }
        uint4 hi, lo;
        tc::split8_f16(x, hi, lo);
        const uint32_t off = xg + tc::sw64_chunk_offset(row, j);
        tc::sts128(off, hi);
        tc::sts128(off + 8192u, lo);
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&x_full[g]);
    };
    rbf_phase(0, 0);
    rbf_phase(1, 1);
    for (int pair = 0; pair * 2 < n_my; ++pair) {
      const int n_in_pair = min(2, n_my - pair * 2);
      // hidden layers: X <- act(D * 2^s + b) * 2^-s', in place.  The last hidden layer feeds the final linear
      // layer (H -> E) directly from registers: every thread forms the partial dot products of its 32 features,
      // the four column quarters meet in shared memory.
      for (int l = 0; l < n_hidden; ++l) {
        const bool last = l == n_hidden - 1;
        const uint32_t bl_a = tc::smem_u32(bias_s + l * 128 + col0);
        const uint32_t wf_a = tc::smem_u32(wf4 + col0);
        const float s_out = p.out_scale[l], s_in = p.in_scale[l + 1];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (g >= n_in_pair) continue;
          const long long q0 = p.dbg ? clock64() : 0;
          tc::mbar_wait(&d_full[g], pd[g]);
          if (p.dbg && warp == 2 && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + (size_t)blockIdx.x * 8 + 3), (unsigned long long)(clock64() - q0));
          pd[g] ^= 1;
          tc::tc_fence_after();
          const uint32_t t_main = t_lane + (uint32_t)g * 256u + col0, t_corr = t_main + 128u;
          const uint32_t xg = xs_a + (uint32_t)g * ETC_X_BYTES + (uint32_t)cq * 16384u;
          float o0 = 0.0f, o1 = 0.0f, o2 = 0.0f, o3 = 0.0f;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            float v[16];
            tc::tmem_ld16_combined(t_main + cc * 16, t_corr + cc * 16, v);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              // (explicit shared loads: a pointer carved out of the aligned buffer would compile to generic loads)
              const float4 b0 = tc::lds128(bl_a + cc * 64 + hh * 32), b1 = tc::lds128(bl_a + cc * 64 + hh * 32 + 16);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              float x[8];
              if (!last) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  x[i] = act_t<ACT>(fmaf(v[hh * 8 + i], s_out, bb[i])) * s_in;
                uint4 hi, lo;
                tc::split8_f16(x, hi, lo);
                const uint32_t off = xg + tc::sw64_chunk_offset(row, cc * 2 + hh);
                tc::sts128(off, hi);
                tc::sts128(off + 8192u, lo);
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float y = act_t<ACT>(fmaf(v[hh * 8 + i], s_out, bb[i]));
                  const float4 w = tc::lds128(wf_a + (uint32_t)(cc * 16 + hh * 8 + i) * 16u);
                  o0 = fmaf(y, w.x, o0);
                  o1 = fmaf(y, w.y, o1);
                  o2 = fmaf(y, w.z, o2);
                  o3 = fmaf(y, w.w, o3);
                }
              }
            }
          }
          if (!last) {
            tc::fence_proxy_async();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&x_full[g]);
          } else {
            // final linear layer * mask (model.py:128,261): sum the four column quarters in a fixed order
            tc::tc_fence_before();
            part[(g * 4 + cq) * 128 + row] = make_float4(o0, o1, o2, o3);
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (cq == 0) {
              const int64_t e = eidx[g];
              if (e < p.n_edges) {
                const float4 p0 = part[(g * 4 + 0) * 128 + row], p1 = part[(g * 4 + 1) * 128 + row];
                const float4 p2 = part[(g * 4 + 2) * 128 + row], p3 = part[(g * 4 + 3) * 128 + row];
                const bool m = dd[g] > 0.0f;
                float o[4];
                o[0] = m ? ((p0.x + p1.x) + (p2.x + p3.x)) + bf_s[0] : 0.0f;
                o[1] = m ? ((p0.y + p1.y) + (p2.y + p3.y)) + bf_s[1] : 0.0f;
                o[2] = m ? ((p0.z + p1.z) + (p2.z + p3.z)) + bf_s[2] : 0.0f;
                o[3] = m ? ((p0.w + p1.w) + (p2.w + p3.w)) + bf_s[3] : 0.0f;
                if (p.out != nullptr)
                  for (int i = 0; i < p.E; ++i) p.out[e * p.E + i] = o[i];
                if (p.rec != nullptr) {
                  const int64_t er = p.rec_k ? rec_slot(p.rec_e0 + e, p.rec_k) - p.rec_e0 : e;
                  p.rec[er] = make_float4(o[0], o[1], o[2], __int_as_float(idxs[g]));
                }
              }
            }
            rbf_phase(g, (pair + 1) * 2 + g);      // slot g is free (its MMAs completed before d_full): next tile
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<512>(tmem_base);
}

template <int ACT>
__global__ void __launch_bounds__(ETC_THREADS, 1) edge_mlp_tc_kernel(const EdgeTcArgs p) {
  edge_mlp_tc_body<ACT, false>(p);
}

}  // namespace nmr

namespace nmr {

// ----------------------------------------------------------------------------------
// Helpers for the tensor-core MP layer: edge records and per-atom feature maxima.
// ----------------------------------------------------------------------------------
// rec[e] = {e0, e1, e2, bits(idx)} from separate nlist / edge-feature arrays (E <= 3);
// out-of-range indices are flagged and replaced by 0 (the call then fails with BAD_INDEX).
__global__ void __launch_bounds__(256) pack_edge_records_kernel(const int32_t* __restrict__ nlist,
                                                                const float* __restrict__ efeat, float4* __restrict__ rec,
                                                                int64_t n_edges, int E, int64_t n_atoms, int* err_flag,
                                                                int rec_k) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= n_edges) return;
  int32_t idx = nlist[e];
  if (idx < 0 || idx >= n_atoms) {
    atomicOr(err_flag, 1);
    idx = 0;
  }
  float4 r = make_float4(0.f, 0.f, 0.f, __int_as_float(idx));
  r.x = efeat[e * E];
  if (E > 1) r.y = efeat[e * E + 1];
  if (E > 2) r.z = efeat[e * E + 2];
  rec[rec_k ? rec_slot(e, rec_k) : e] = r;
}

// hmax[i] = max_l |h[i, l]|, F = 256: one warp per atom
__global__ void __launch_bounds__(256) row_absmax256_kernel(const float* __restrict__ h, float* __restrict__ hmax,
                                                            int64_t n_atoms) {
  const int64_t atom = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (atom >= n_atoms) return;
  const int lane = threadIdx.x & 31;
  const float4 a = *reinterpret_cast<const float4*>(h + atom * 256 + lane * 4);
  const float4 b = *reinterpret_cast<const float4*>(h + atom * 256 + 128 + lane * 4);
  float m = fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
                  fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) hmax[atom] = m;
}

// ----------------------------------------------------------------------------------
// MP layer on tensor cores, F = 256, E <= 3, K <= 16 (layers.py:26-46 + model.py:167).
// One CTA = 128 atoms per tile, persistent over tiles.
//   T[i,(p,n,ll)] = sum_j e[i,j,n] * h[nl[i,j], 32p+ll]     gather-aggregate, fp32 FFMA (producers)
//   D[i, m]       = sum_k T[i,k] * W'[k, m]                 tcgen05 fp16x3, K = 256*E, N = 256
//   h_out[i, m]   = act(inv_degree[i] * D[i,m]) + h_in[i,m] epilogue
// The producers build T pass by pass (one pass = 32 input features x E = E operand chunks of
// 32 k), scale each row by a power of two from an a-priori bound (fp16 range), split it into
// the hi/lo operand tiles and hand it to the MMA warp through a 2-pass ring; W' streams
// through a 4-slot ring of 16 KB bulk copies (hi and lo images separately).  T never reaches HBM.
// The gather is the critical resource: every producer thread keeps 8-16 independent 16-byte
// row loads in flight (two register buffers of 8, the next half-step is issued before the
// current one is consumed, across step and pass boundaries), loads are unconditional (padded
// slots read row 0 with a zero edge feature) and edge records come from shared memory.
//   warp 0: W' loader   warp 1: MMA issuer + TMEM owner   warp 2: edge-record loader
//   warps 4-7: epilogue (thread = atom row)                warps 8-15: producers
// ----------------------------------------------------------------------------------
struct MpTcArgs {
  // Variable-tile form (mp_layer_tc_vt_kernel, calls of less than one wave of 128-atom tiles): every tile holds
  // tile_rows <= 128 atoms (a multiple of 32), so that a small call spreads over more SMs and a tile's producers only
  // run the 32-row steps it has.  The other kernels ignore these fields (tiles of 128).
  int64_t vt_tiles;
  int vt_rows;
  int l1_prefetch;           // option "mp_l1_prefetch": warp 2 prefetches the tile's own node rows into L1 one pass ahead
  const float* h_in;         // [n_atoms, 256]
  const float* hmax_in;      // [n_atoms]  max_l |h_in[i,l]|
  float* h_out;              // [n_atoms, 256]
  float* hmax_out;           // [n_atoms]
  const float4* rec;         // [n_atoms * K] {e0, e1, e2, bits(idx)}
  const float* inv_degree;   // [n_atoms]
  const uint8_t* Wimg;       // [8 passes][E][hi 16384 | lo 16384]
  int64_t n_atoms;
  int K;
  int E;
  int act;
  float corr;                // 1 + c: compensates the round-toward-zero accumulation of tcgen05
  int raw;                   // 1: h_out = inv_degree * D (no activation, no residual) -- calibration tap
  int swz;                   // records are slot-swizzled (rec_slot; K = 8 or 16)
  int nseg;                  // accumulation-chain segments per tile: 1, 2, 4 or 8 (see "chain segments" below)
  int hmax_pair;             // hmax_in holds two partial row maxima per atom (written by the column-split kernel)
  long long* dbg;            // optional [gridDim][8] cycle counters (diagnostics): see tools/diag_mp_roles.py
};

// Chain segments.  tcgen05 accumulates round-toward-zero; over the 48 instructions of the main product that is a
// trajectory-dependent error (it adds up coherently while a partial sum stays on one side of zero) which no constant
// factor removes, and it is the largest error source of the tensor-core path (profiles/r02_parity.md).  With nseg > 1 the
// K loop is cut into nseg chains of 8 / nseg feature passes: after each chain the epilogue warps (idle during a tile
// anyway) drain main + corr, apply the chain's compensation and add the partial sums in round-to-nearest FP32 -- the
// running sum is parked in the tile's own rows of h_out (written and re-read by the same thread; L2-resident) -- and
// the next chain starts from a cleared accumulator.  The producers keep filling the operand ring during a drain.
constexpr int MTC_THREADS = 512;
// Register budget per warpgroup (setmaxnreg).  The kernel as a whole sits at the 128 registers 512 threads allow, with no
// slack: two harmless additions (two griddepcontrol instructions; a variable trip count of the producers' step loop)
// each pushed a few bytes into local memory inside the gather loop and cost 3-4 %.  The light warpgroup (W' loader, MMA
// issuer, record loader, an idle warp) gives up most of its registers and the three heavy ones (epilogue, producers)
// take them: 4 x 32 x 40 + 12 x 32 x 152 = 63 488.  An increase is served only from registers released inside the CTA,
// every warp of a warpgroup must execute the SAME setmaxnreg instruction, and ptxas budgets the code a setmaxnreg
// dominates -- hence one at the head of each warpgroup's branch.
constexpr int MTC_REGS_LIGHT = 64, MTC_REGS_HEAVY = 144;
constexpr int MTC_PASSES = 8;          // 256 / 32
constexpr int MTC_BRING = 4;           // W' ring: slots of 16 KB, the hi and the lo image of a (pass, n) chunk are separate slots
constexpr int MTC_BSLOT = 16384;
constexpr int MTC_KMAX = 16;
// 64-byte-swizzled tiles need a 512-byte aligned base.  The kernel stays below 196 KB so that the SM can be
// configured with 60 KB of L1 (228 KB of shared memory would leave 28 KB): the neighbour rows gathered for one
// feature pass (~38 KB unique per tile) then mostly hit L1.
constexpr size_t MTC_SMEM = 512 + 2 * 3 * 16384 + MTC_BRING * MTC_BSLOT + 128 * MTC_KMAX * 16 + 2 * 2 * 128 * 4 + 256;
static_assert(MTC_SMEM + 1024 <= 196 * 1024, "MP tensor-core kernel no longer fits the 196 KB shared-memory configuration");
template <int ACT, bool SEG, bool ONE, bool VT = false>
__device__ __forceinline__ void mp_layer_tc_body(const MpTcArgs& p) {
  static_assert(!(SEG && ONE), "chain segments drain per segment: not combined with the single-accumulator form");
  static_assert(!VT || (!SEG && !ONE), "variable tiles: default kernel only");
  constexpr int BRING = MTC_BRING;
  constexpr int BSLOT = MTC_BSLOT;
  constexpr bool SPLIT = true;      // one slot = one image (hi or lo) of a (pass, n) chunk
  constexpr uint32_t AST = 2;      // operand stages
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 511) & ~uintptr_t(511));
  uint8_t* a_st = smem;                                   // [2 stages][3 chunks][hi 8192 | lo 8192]
  uint8_t* b_ring = a_st + AST * 3 * 16384;                 // [BRING][hi 16384 | lo 16384]
  float4* rec_s = reinterpret_cast<float4*>(b_ring + BRING * BSLOT);       // [128 * K]
  float* fscale = reinterpret_cast<float*>(rec_s + 128 * MTC_KMAX);        // [2][128]  2^-s
  float* oscale = fscale + 2 * 128;                                         // [2][128]  2^s * inv_degree
  uint64_t* bars = reinterpret_cast<uint64_t*>(oscale + 2 * 128);
  uint64_t* a_full = bars;            // [2]
  uint64_t* a_empty = a_full + AST;   // [AST]
  uint64_t* b_full = a_empty + AST;   // [BRING]
  uint64_t* b_empty = b_full + BRING;
  uint64_t* rec_full = b_empty + BRING;
  uint64_t* rec_empty = rec_full + 1;
  uint64_t* d_full = rec_empty + 1;   // [2] (the second one: single-accumulator form, accumulator set 1)
  uint64_t* d_empty = d_full + 2;     // [2]
  uint64_t* sc_full = d_empty + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sc_full + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < (int)AST; ++i) {
      tc::mbar_init(&a_full[i], 8);
      tc::mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) tc::mbar_init(&sc_full[i], 8);
    for (int i = 0; i < BRING; ++i) {
      tc::mbar_init(&b_full[i], 1);
      tc::mbar_init(&b_empty[i], 1);
    }
    tc::mbar_init(rec_full, 1);
    tc::mbar_init(rec_empty, 8);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&d_full[i], 1);
      tc::mbar_init(&d_empty[i], 4);
    }
    tc::mbar_fence_init();
  }
  if (warp == 1) {
    tc::tmem_alloc<512>(tmem_slot);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int K = p.K, E = p.E;
  const int64_t n_tiles = VT ? p.vt_tiles : (p.n_atoms + 127) / 128;
  const int TR = VT ? p.vt_rows : 128;      // atoms per tile
  // tile walk: the CTA takes tiles blockIdx.x, + gridDim.x, ...
  const int64_t tile_first = (int64_t)blockIdx.x;
  const int64_t tile_step = (int64_t)gridDim.x;
  const int64_t tile_end = n_tiles;

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MTC_REGS_LIGHT));
  if (warp == 0) {
    // ===================== W' loader =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = tile_first; tile < tile_end; tile += tile_step)
        for (int q = 0; q < MTC_PASSES * E * (SPLIT ? 2 : 1); ++q, ++it) {
          const uint32_t slot = it % BRING, ph = (it / BRING) & 1;
          tc::mbar_wait_relaxed(&b_empty[slot], ph ^ 1);
          tc::mbar_expect_tx(&b_full[slot], BSLOT);
          if (SPLIT) {  // [pass][n][hi | lo] is contiguous in 16 KB images
            tc::bulk_g2s(b_ring + slot * BSLOT, p.Wimg + (size_t)q * BSLOT, BSLOT, &b_full[slot]);
          } else {
            tc::bulk_g2s(b_ring + slot * BSLOT, p.Wimg + (size_t)q * 32768, 32768, &b_full[slot]);
          }
        }
    }
  } else if (warp == 2) {
    // ===================== edge-record loader + L2 prefetcher =====================
    // After handing tile t's records to the producers, the warp pulls the CTA's NEXT tile into L2: its
    // edge records and the node rows of its own atom range (for chain-ordered molecules that is where
    // most neighbours live), so that the gathers of the next tile hit L2 instead of paying DRAM latency
    // on the producers' critical path.  Pure hint: no effect on results.
    uint32_t t = 0;
    for (int64_t tile = tile_first; tile < tile_end; tile += tile_step, ++t) {
      if (lane == 0) {
        const int64_t a0 = tile * TR;
        const int rows = (int)max((int64_t)0, min((int64_t)TR, p.n_atoms - a0));
        const uint32_t bytes = (uint32_t)rows * (uint32_t)K * 16u;
        tc::mbar_wait_relaxed(rec_empty, (t & 1) ^ 1);
        tc::mbar_expect_tx(rec_full, bytes);
        if (bytes) tc::bulk_g2s(rec_s, p.rec + a0 * K, bytes, rec_full);
      }
      __syncwarp();
      const int64_t nt = tile + tile_step;
      if (nt < n_tiles) {
        const int64_t b0 = nt * TR;
        const int nrows = (int)min((int64_t)TR, p.n_atoms - b0);
        const char* hb = reinterpret_cast<const char*>(p.h_in + b0 * 256);
        for (int i = lane; i < nrows * 8; i += 32) tc::prefetch_l2(hb + (size_t)i * 128);
        const char* rb = reinterpret_cast<const char*>(p.rec + b0 * K);
        for (int i = lane; i * 128 < nrows * K * 16; i += 32) tc::prefetch_l2(rb + (size_t)i * 128);
      }
      if (p.l1_prefetch) {
        // L1 prefetch of the tile's OWN node rows, one feature pass ahead of the producers: for chain-ordered molecules
        // most neighbours of a tile's atoms are atoms of the tile, so these 128 lines per pass are the largest group of
        // compulsory L1 misses of the gather (every (row, pass) line is first touched exactly once).  Paced on the
        // producers' a_full barriers with a BOUNDED poll: this warp must never block on the producers, who wait for
        // its next record copy.
        const int64_t a0 = tile * TR;
        const int rows = (int)max((int64_t)0, min((int64_t)TR, p.n_atoms - a0));
        const char* hb = reinterpret_cast<const char*>(p.h_in + a0 * 256);
        for (int ps = 0; ps < MTC_PASSES; ++ps) {
          if (ps >= 2) {
            const uint32_t pk = t * MTC_PASSES + (uint32_t)ps - 2u;        // producers have finished pass ps - 2
            for (int spin = 0; spin < 64 && !tc::mbar_try_wait(&a_full[pk % AST], (pk / AST) & 1); ++spin) __nanosleep(200);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = lane + 32 * i;
            if (r < rows) tc::prefetch_l1(hb + (size_t)r * 1024 + (size_t)ps * 128);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_f16(128, 256);
      const uint32_t d_main = tmem_base, d_corr = tmem_base + 256u;
      uint32_t it = 0, pass = 0, t = 0, dph = 0;
      const int seg_passes = MTC_PASSES / p.nseg;
      long long w_d = 0, w_a = 0, w_b = 0, c0 = 0;
      const long long k0 = clock64();
      for (int64_t tile = tile_first; tile < tile_end; tile += tile_step, ++t) {
        for (int ps = 0; ps < MTC_PASSES; ++ps, ++pass) {
          if (ONE) {
            if (ps == 0) {                       // this accumulator set was drained two tiles ago
              if (p.dbg) c0 = clock64();
              tc::mbar_wait(&d_empty[t & 1], ((t >> 1) & 1) ^ 1);
              if (p.dbg) w_d += clock64() - c0;
              tc::tc_fence_after();
            }
          } else if (ps % seg_passes == 0) {     // a new accumulation chain: the epilogue has drained the previous one
            if (p.dbg) c0 = clock64();
            tc::mbar_wait(d_empty, (dph & 1) ^ 1);
            if (p.dbg) w_d += clock64() - c0;
            tc::tc_fence_after();
          }
          // single-accumulator form: main and correction products share ONE accumulator (256 columns), so tensor memory
          // holds two sets and tile t + 1 accumulates while tile t is drained
          const uint32_t dm = ONE ? tmem_base + (t & 1u) * 256u : d_main, dc = ONE ? dm : d_corr;
          const uint32_t st = pass % AST;
          if (p.dbg) c0 = clock64();
          tc::mbar_wait(&a_full[st], (pass / AST) & 1);
          if (p.dbg) w_a += clock64() - c0;
          tc::tc_fence_after();
          for (int n = 0; n < E; ++n, ++it) {
            uint32_t slot = it % BRING;
            if (p.dbg) c0 = clock64();
            tc::mbar_wait(&b_full[slot], (it / BRING) & 1);
            if (p.dbg) w_b += clock64() - c0;
            tc::tc_fence_after();
            const uint8_t* ac = a_st + (st * 3 + n) * 16384;
            const uint64_t ah = tc::make_desc_sw64(tc::smem_u32(ac));
            const uint64_t al = tc::make_desc_sw64(tc::smem_u32(ac + 8192));
            const uint64_t bh = tc::make_desc_sw64(tc::smem_u32(b_ring + slot * BSLOT));
            // products against the hi image of W': main = hi * hi, corr = lo * hi
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t adv = (uint64_t)(ks * 2);
              // main restarts with every chain; the correction accumulator (its truncation is scaled by 2^-11)
              // runs through the whole tile and is read once, by the last chain's epilogue
              const uint32_t acc = ((ps % seg_passes) | n | ks) != 0, acc_c = ONE ? 1u : (uint32_t)((ps | n | ks) != 0);
              tc::umma_f16(dm, ah + adv, bh + adv, idesc, acc);
              tc::umma_f16(dc, al + adv, bh + adv, idesc, acc_c);
            }
            uint64_t bl;
            if (SPLIT) {     // the lo image is the next slot of the ring
              tc::umma_commit(&b_empty[slot]);
              ++it;
              slot = it % BRING;
              if (p.dbg) c0 = clock64();
              tc::mbar_wait(&b_full[slot], (it / BRING) & 1);
              if (p.dbg) w_b += clock64() - c0;
              tc::tc_fence_after();
              bl = tc::make_desc_sw64(tc::smem_u32(b_ring + slot * BSLOT));
            } else {
              bl = tc::make_desc_sw64(tc::smem_u32(b_ring + slot * BSLOT + BSLOT / 2));
            }
            // corr += hi * lo
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t adv = (uint64_t)(ks * 2);
              tc::umma_f16(dc, ah + adv, bl + adv, idesc, 1);
            }
            tc::umma_commit(&b_empty[slot]);
          }
          tc::umma_commit(&a_empty[st]);
          if (ONE) {
            if (ps + 1 == MTC_PASSES) tc::umma_commit(&d_full[t & 1]);
          } else if ((ps + 1) % seg_passes == 0) {      // chain complete: hand the accumulators to the epilogue
            tc::umma_commit(d_full);
            ++dph;
          }
        }
      }
      if (p.dbg) {
        long long* o = p.dbg + (size_t)blockIdx.x * 8;
        o[0] = clock64() - k0;   // MMA thread: total
        o[1] = w_d;              //   waiting for the epilogue (d_empty)
        o[2] = w_a;              //   waiting for the producers (a_full)
        o[3] = w_b;              //   waiting for W' (b_full)
      }
    }
  }
  } else if (warp >= 4 && warp < 8) {
    // (the single-accumulator kernel runs best with 160 registers in the epilogue: 0.330 against 0.340 ms; the default
    //  kernel spills there)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(ONE ? 160 : MTC_REGS_HEAVY));
    // ===================== epilogue =====================
    // TMEM hands every thread one atom row; global memory wants 128 contiguous bytes per row and
    // instruction.  Each group of 8 lanes therefore transposes its 8 rows x 8 float4 block with
    // shuffles: afterwards lane i of the group holds the i-th 16 bytes of all 8 rows, so that the residual
    // loads and the stores of h_out cover full 128-byte lines (4 L1 wavefronts per instruction, not 32).
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int gi = lane & 7;                       // position inside the 8-lane group = float4 chunk after transposition
    const int grow = q * 32 + (lane & 24);         // first row of the group
    const uint32_t t_main = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t t_corr = t_main + 256u;
    uint32_t t = 0, dph = 0;
    float osc_one = 0.0f;          // single-accumulator form: this tile's output scale, fetched one tile ahead (see below)
    if (ONE && tile_first < tile_end) {
      tc::mbar_wait(&sc_full[0], 0);
      osc_one = oscale[row] * p.corr;
    }
    for (int64_t tile = tile_first; tile < tile_end; tile += tile_step, ++t) {
      const int64_t a0 = tile * TR;
      const int rows = (int)max((int64_t)0, min((int64_t)TR, p.n_atoms - a0));
      if (!ONE) tc::mbar_wait(&sc_full[t & 1], (t >> 1) & 1);
      float hm[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) hm[k] = 0.0f;
      if constexpr (!SEG) {
        // one accumulation chain per tile (the default): drain, compensate, activate, add the residual
        const float osc = ONE ? osc_one : oscale[(t & 1) * 128 + row] * p.corr;   // (1 + c) once per row: one chain
        const uint32_t t_set = t_main + (ONE ? (t & 1u) * 256u : 0u);
        // rows grow .. grow+7 of this group, clamped for the loads (stores are predicated)
        const float* hin = p.h_in + (rows > 0 ? a0 + min(grow, rows - 1) : 0) * 256 + gi * 4;
        float* hout = p.h_out + (a0 + grow) * 256 + gi * 4;
        // (rows past the end of a partial tile re-read the tile's last row; their stores are predicated off)
        const int rlast = rows > 0 ? rows - 1 - min(grow, rows - 1) : 0;
        float4 res[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) res[k] = p.raw ? make_float4(0.f, 0.f, 0.f, 0.f) : tc::ldg128(hin + min(k, rlast) * 256);
        if (ONE) {
          tc::mbar_wait(&d_full[t & 1], (t >> 1) & 1);
        } else {
          tc::mbar_wait(d_full, dph & 1);
          ++dph;
        }
        const long long e0 = p.dbg ? clock64() : 0;
        tc::tc_fence_after();
        if (ONE) {
          // The next tile's scale is picked up BEFORE this accumulator set is handed back: the producers can overwrite
          // that scale buffer (for tile t + 3) only after the MMAs of tile t + 2 have started, which wait for this set.
          // (It was written long ago: the producers finished tile t before its MMAs completed.)
          if (tile + tile_step < tile_end) {
            tc::mbar_wait(&sc_full[(t + 1) & 1], ((t + 1) >> 1) & 1);
            osc_one = oscale[((t + 1) & 1) * 128 + row] * p.corr;
          }
        }
#pragma unroll 1
        for (int cc = 0; cc < 8; ++cc) {
          float4 x[8];
          {
            float v[16];
            if (ONE) tc::tmem_ld16(t_set + cc * 32, v);
            else tc::tmem_ld16_combined(t_main + cc * 32, t_corr + cc * 32, v);
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = make_float4(v[4 * j] * osc, v[4 * j + 1] * osc, v[4 * j + 2] * osc, v[4 * j + 3] * osc);
            if (ONE) tc::tmem_ld16(t_set + cc * 32 + 16, v);
            else tc::tmem_ld16_combined(t_main + cc * 32 + 16, t_corr + cc * 32 + 16, v);
#pragma unroll
            for (int j = 0; j < 4; ++j) x[4 + j] = make_float4(v[4 * j] * osc, v[4 * j + 1] * osc, v[4 * j + 2] * osc, v[4 * j + 3] * osc);
          }
          if (ONE && cc == 7) {    // every accumulator of the set is in registers: the MMAs of tile t + 2 may overwrite it
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&d_empty[t & 1]);
          }
          // 8 x 8 transpose of float4 inside the 8-lane group (3 butterfly stages)
#pragma unroll
          for (int m = 1; m < 8; m <<= 1) {
            const bool up = (lane & m) != 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (j & m) continue;
              const float4 lo4 = x[j], hi4 = x[j | m];
              float4 snd = up ? lo4 : hi4, rcv;
              rcv.x = __shfl_xor_sync(0xffffffffu, snd.x, m);
              rcv.y = __shfl_xor_sync(0xffffffffu, snd.y, m);
              rcv.z = __shfl_xor_sync(0xffffffffu, snd.z, m);
              rcv.w = __shfl_xor_sync(0xffffffffu, snd.w, m);
              x[j] = up ? rcv : lo4;
              x[j | m] = up ? hi4 : rcv;
            }
          }
          // now x[k] = columns cc*32 + gi*4 .. +3 of row grow + k
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float4 o;
            if (p.raw) {
              o = x[k];
            } else {
              o.x = act_t<ACT>(x[k].x) + res[k].x;
              o.y = act_t<ACT>(x[k].y) + res[k].y;
              o.z = act_t<ACT>(x[k].z) + res[k].z;
              o.w = act_t<ACT>(x[k].w) + res[k].w;
            }
            hm[k] = fmaxf(hm[k], fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
            if (grow + k < rows) *reinterpret_cast<float4*>(hout + k * 256 + cc * 32) = o;
            // the residual of the next 32 columns is in flight during the next accumulator read + transposition
            if (cc + 1 < 8 && !p.raw) res[k] = tc::ldg128(hin + min(k, rlast) * 256 + (cc + 1) * 32);
          }
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0 && !ONE) {
          tc::mbar_arrive(d_empty);
        }
        if (p.dbg && warp == 4 && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + (size_t)blockIdx.x * 8 + 4), (unsigned long long)(clock64() - e0));
      } else {
        const float* oscr = oscale + (t & 1) * 128 + grow;   // 2^s * inv_degree of rows grow .. grow+7 (the rows this thread finishes)
        // rows grow .. grow+7 of this group, clamped for the loads (stores are predicated)
        const float* hin = p.h_in + (rows > 0 ? a0 + min(grow, rows - 1) : 0) * 256 + gi * 4;
        float* hout = p.h_out + (a0 + grow) * 256 + gi * 4;
        // the same rows clamped like `hin`, for re-reading the running sum of a partial tile's trailing groups
        const float* hpart = p.h_out + (rows > 0 ? a0 + min(grow, rows - 1) : 0) * 256 + gi * 4;
        // (rows past the end of a partial tile re-read the tile's last row; their stores are predicated off)
        const int rlast = rows > 0 ? rows - 1 - min(grow, rows - 1) : 0;
        float4 buf[8];                  // in flight: partial sums of the previous chains, then the residual rows
        const float cf = p.corr;        // chain compensation, applied per segment (oscale carries 2^s * inv_degree only)
        for (int seg = 0; seg < p.nseg; ++seg, ++dph) {
          const bool first = seg == 0, last = seg == p.nseg - 1;
          if (!first) {                 // running sum of the previous chains, columns 0..31: this thread wrote it
#pragma unroll
            for (int k = 0; k < 8; ++k) buf[k] = *reinterpret_cast<const float4*>(hpart + min(k, rlast) * 256);
          }
          tc::mbar_wait(d_full, dph & 1);
          const long long e0 = p.dbg ? clock64() : 0;
          tc::tc_fence_after();
#pragma unroll 1
          for (int cc = 0; cc < 8; ++cc) {
            float4 x[8];
            {
              float v[16];
              if (last) tc::tmem_ld16_combined(t_main + cc * 32, t_corr + cc * 32, v);
              else tc::tmem_ld16(t_main + cc * 32, v);             // the correction accumulator keeps running
#pragma unroll
              for (int j = 0; j < 4; ++j) x[j] = make_float4(v[4 * j] * cf, v[4 * j + 1] * cf, v[4 * j + 2] * cf, v[4 * j + 3] * cf);
              if (last) tc::tmem_ld16_combined(t_main + cc * 32 + 16, t_corr + cc * 32 + 16, v);
              else tc::tmem_ld16(t_main + cc * 32 + 16, v);
#pragma unroll
              for (int j = 0; j < 4; ++j) x[4 + j] = make_float4(v[4 * j] * cf, v[4 * j + 1] * cf, v[4 * j + 2] * cf, v[4 * j + 3] * cf);
            }
            // 8 x 8 transpose of float4 inside the 8-lane group (3 butterfly stages)
#pragma unroll
            for (int m = 1; m < 8; m <<= 1) {
              const bool up = (lane & m) != 0;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (j & m) continue;
                const float4 lo4 = x[j], hi4 = x[j | m];
                float4 snd = up ? lo4 : hi4, rcv;
                rcv.x = __shfl_xor_sync(0xffffffffu, snd.x, m);
                rcv.y = __shfl_xor_sync(0xffffffffu, snd.y, m);
                rcv.z = __shfl_xor_sync(0xffffffffu, snd.z, m);
                rcv.w = __shfl_xor_sync(0xffffffffu, snd.w, m);
                x[j] = up ? rcv : lo4;
                x[j | m] = up ? hi4 : rcv;
              }
            }
            // now x[k] = columns cc*32 + gi*4 .. +3 of row grow + k (accumulator units, this chain only)
            if (!first) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                x[k].x += buf[k].x;
                x[k].y += buf[k].y;
                x[k].z += buf[k].z;
                x[k].w += buf[k].w;
              }
            }
            if (!last) {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (grow + k < rows) *reinterpret_cast<float4*>(hout + k * 256 + cc * 32) = x[k];
            } else {
              // the residual rows travel while the activation is evaluated
              if (!p.raw) {
#pragma unroll
                for (int k = 0; k < 8; ++k) buf[k] = tc::ldg128(hin + min(k, rlast) * 256 + cc * 32);
              }
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float osk = oscr[k];
                x[k].x *= osk;
                x[k].y *= osk;
                x[k].z *= osk;
                x[k].w *= osk;
                if (!p.raw) {
                  x[k].x = act_t<ACT>(x[k].x);
                  x[k].y = act_t<ACT>(x[k].y);
                  x[k].z = act_t<ACT>(x[k].z);
                  x[k].w = act_t<ACT>(x[k].w);
                }
              }
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                float4 o = x[k];
                if (!p.raw) {
                  o.x += buf[k].x;
                  o.y += buf[k].y;
                  o.z += buf[k].z;
                  o.w += buf[k].w;
                }
                hm[k] = fmaxf(hm[k], fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
                if (grow + k < rows) *reinterpret_cast<float4*>(hout + k * 256 + cc * 32) = o;
              }
            }
            if (!first && cc + 1 < 8) {   // the running sum of the next 32 columns, in flight during the next accumulator read
#pragma unroll
              for (int k = 0; k < 8; ++k) buf[k] = *reinterpret_cast<const float4*>(hpart + min(k, rlast) * 256 + (cc + 1) * 32);
            }
          }
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            tc::mbar_arrive(d_empty);
          }
          if (p.dbg && warp == 4 && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + (size_t)blockIdx.x * 8 + 4), (unsigned long long)(clock64() - e0));
        }
      }
      // row maxima: reduce over the 8 lanes of the group, lane k writes row grow + k
      float mine = 0.0f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float v = hm[k];
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
        if (gi == k) mine = v;
      }
      if (grow + gi < rows) p.hmax_out[a0 + grow + gi] = mine;
    }
  } else if (warp >= 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(MTC_REGS_HEAVY));
    // ===================== producers: gather-aggregate, scale, split =====================
    const int pw = warp - 8;                 // 0..7
    const int ptid = tid - 256;              // 0..255
    const int q8 = lane & 7;                 // 4-feature group inside the 32-feature pass
    const int rsub = lane >> 3;              // row inside the warp's 4-row group
    const uint32_t rec_a = tc::smem_u32(rec_s);
    const uint32_t ast_a = tc::smem_u32(a_st);
    const float* hq = p.h_in + q8 * 4;
    uint32_t pass = 0, t = 0;
    for (int64_t tile = tile_first; tile < tile_end; tile += tile_step, ++t) {
      const int64_t a0 = tile * TR;
      const int rows = (int)max((int64_t)0, min((int64_t)TR, p.n_atoms - a0));
      float* fs = fscale + (t & 1) * 128;
      float* os = oscale + (t & 1) * 128;
      const uint32_t fs_a = tc::smem_u32(fs);
      const long long r0c = p.dbg ? clock64() : 0;
      tc::mbar_wait(rec_full, t & 1);
      if (p.dbg && warp == 8 && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + (size_t)blockIdx.x * 8 + 6), (unsigned long long)(clock64() - r0c));

      // 8 independent row loads of half-step (row, half) at feature pass ps
      // (no branches around the loads: a branch makes the compiler drain every outstanding load first;
      //  out-of-range rows / slots read an in-bounds shared address and are neutralised by selects)
      // FULL = full tile and K == 16: no range predicates at all
      const bool full = rows == 128 && K == 16;
      // neighbour indices are fetched one half-step ahead of the row loads that use them (`nidx`), so the
      // address computation never waits for shared memory
      uint32_t nidx[8];
      const uint32_t sw = p.swz ? (uint32_t)(rsub << 1) : 0u;   // row & 3 == rsub for every row this thread touches
      auto load_idx = [&](int row, int half, bool FULL) {
        const bool rv = row < rows;
        const uint32_t ra = rec_a + (uint32_t)(row * K) * 16u + 12u;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          uint32_t idx = tc::lds32(ra + (((uint32_t)(half * 8 + u)) ^ sw) * 16u);
          if (!FULL) idx = (rv && half * 8 + u < K) ? idx : 0u;
          nidx[u] = idx;
        }
      };
      auto issue = [&](float4 (&hv)[8], int ps) {
        const float* hp = hq + ps * 32;
#pragma unroll
        for (int u = 0; u < 8; ++u) hv[u] = tc::ldg128(hp + (size_t)nidx[u] * 256);
      };
      auto consume = [&](const float4 (&hv)[8], int row, int half, float (&acc)[3][4], bool FULL) {
        const bool rv = row < rows;
        const uint32_t ra = rec_a + (uint32_t)(row * K) * 16u;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float4 r = tc::lds128(ra + (((uint32_t)(half * 8 + u)) ^ sw) * 16u);
          if (!FULL) {
            const bool ok = rv && half * 8 + u < K;
            r.x = ok ? r.x : 0.0f;
            r.y = ok ? r.y : 0.0f;
            r.z = ok ? r.z : 0.0f;
          }
          acc[0][0] = fmaf(r.x, hv[u].x, acc[0][0]);
          acc[0][1] = fmaf(r.x, hv[u].y, acc[0][1]);
          acc[0][2] = fmaf(r.x, hv[u].z, acc[0][2]);
          acc[0][3] = fmaf(r.x, hv[u].w, acc[0][3]);
          acc[1][0] = fmaf(r.y, hv[u].x, acc[1][0]);
          acc[1][1] = fmaf(r.y, hv[u].y, acc[1][1]);
          acc[1][2] = fmaf(r.y, hv[u].z, acc[1][2]);
          acc[1][3] = fmaf(r.y, hv[u].w, acc[1][3]);
          acc[2][0] = fmaf(r.z, hv[u].x, acc[2][0]);
          acc[2][1] = fmaf(r.z, hv[u].y, acc[2][1]);
          acc[2][2] = fmaf(r.z, hv[u].z, acc[2][2]);
          acc[2][3] = fmaf(r.z, hv[u].w, acc[2][3]);
        }
      };

      // K <= 8: the second half-step of every row is all padding -- it is skipped (adding 0 * h is the identity, so the
      // results are the same bits) and the two register buffers alternate between consecutive rows instead
      const bool k8 = K <= 8;
      float4 hvA[8], hvB[8];
      load_idx(pw * 4 + rsub, 0, false);
      issue(hvA, 0);                            // first half-step of the tile, in flight during the scale pass
      if (k8) load_idx(32 + pw * 4 + rsub, 0, false);
      else load_idx(pw * 4 + rsub, 1, false);

      // per-row bound |T[i,.]| <= sum_j max_n|e_ijn| * hmax[nl_ij]  ->  power-of-two scale
      {
        const int row = ptid >> 1, hf = ptid & 1;
        float b = 0.0f;
        if (row < rows) {
          for (int j = hf; j < K; j += 2) {
            const float4 r = tc::lds128(rec_a + (uint32_t)(row * K + j) * 16u);
            const float em = fmaxf(fmaxf(fabsf(r.x), fabsf(r.y)), fabsf(r.z));
            if (em != 0.0f) b = fmaf(em, __ldg(p.hmax_in + __float_as_int(r.w)), b);
          }
        }
        b += __shfl_xor_sync(0xffffffffu, b, 1);
        if (hf == 0) {
          // exponent of b (0 for b < 2^15): scale so that |T| * 2^-s < 2^15
          int s = ((__float_as_int(b) >> 23) & 0xff) - 127 - 14;
          s = b > 0.0f ? max(s, 0) : 0;
          s = min(s, 100);
          fs[row] = tc::pow2f_exact(-s);
          os[row] = row < rows ? tc::pow2f_exact(s) * p.inv_degree[a0 + row] : 0.0f;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (lane == 0) tc::mbar_arrive(&sc_full[t & 1]);

      for (int ps = 0; ps < MTC_PASSES; ++ps, ++pass) {
        const uint32_t st = pass % AST;
        const long long p0 = p.dbg ? clock64() : 0;
        tc::mbar_wait(&a_empty[st], ((pass / AST) & 1) ^ 1);
        if (p.dbg && warp == 8 && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + (size_t)blockIdx.x * 8 + 5), (unsigned long long)(clock64() - p0));
        const uint32_t ab = ast_a + st * 3 * 16384;
        // scale, split and store one row's 4 features x E channels into the operand stage
        auto store_row = [&](int row, const float (&acc)[3][4], float sc) {
          const uint32_t off = (uint32_t)row * 64u + ((((uint32_t)q8 >> 1) ^ (((uint32_t)row >> 1) & 3u)) << 4) +
                               (((uint32_t)q8 & 1u) << 3);
#pragma unroll
          for (int n = 0; n < 3; ++n) {
            if (n < E) {
              uint2 hi, lo;
              if (ONE) {       // lo = x - hi, not scaled by 2^11: it meets the main products in the same accumulator
                tc::split2_f16_plain(acc[n][0] * sc, acc[n][1] * sc, hi.x, lo.x);
                tc::split2_f16_plain(acc[n][2] * sc, acc[n][3] * sc, hi.y, lo.y);
              } else {
                tc::split2_f16(acc[n][0] * sc, acc[n][1] * sc, hi.x, lo.x);
                tc::split2_f16(acc[n][2] * sc, acc[n][3] * sc, hi.y, lo.y);
              }
              tc::sts64(ab + n * 16384 + off, hi.x, hi.y);
              tc::sts64(ab + n * 16384 + 8192 + off, lo.x, lo.y);
            }
          }
        };
        if (k8) {
          // entering: hvA holds (row of step 0, slots 0..7) in flight, nidx the indices of the row of step 1
#pragma unroll 1
          for (int step = 0; VT ? step * 32 < rows : step < 4; step += 2) {
            const int row0 = step * 32 + pw * 4 + rsub, row1 = row0 + 32;
            const bool last = VT ? (step + 2) * 32 >= rows : step == 2;
            const int nrow = last ? pw * 4 + rsub : row1 + 32;          // the row after row1 ...
            const int nps = min(last ? ps + 1 : ps, MTC_PASSES - 1);    // ... in this pass or the next one
            const int nrow2 = nrow + 32;                                // and the one after that (indices only)
            float acc[3][4];
#pragma unroll
            for (int n = 0; n < 3; ++n)
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[n][i] = 0.0f;
            float sc = tc::lds32f(fs_a + (uint32_t)row0 * 4u);
            issue(hvB, ps);                     // row1
            load_idx(nrow, 0, false);
            consume(hvA, row0, 0, acc, false);
            store_row(row0, acc, sc);
#pragma unroll
            for (int n = 0; n < 3; ++n)
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[n][i] = 0.0f;
            sc = tc::lds32f(fs_a + (uint32_t)row1 * 4u);
            issue(hvA, nps);                    // nrow
            load_idx(nrow2, 0, false);
            consume(hvB, row1, 0, acc, false);
            store_row(row1, acc, sc);
          }
        } else {
#pragma unroll 1
        for (int step = 0; VT ? step * 32 < rows : step < 4; ++step) {
          const int row = step * 32 + pw * 4 + rsub;
          float acc[3][4];
#pragma unroll
          for (int n = 0; n < 3; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[n][i] = 0.0f;
          const float sc = tc::lds32f(fs_a + (uint32_t)row * 4u);   // needed only after the FMAs: latency hidden
          // next half-step: next row group of this pass, or the first one of the next pass
          const bool last = VT ? (step + 1) * 32 >= rows : step == 3;
          const int nrow = last ? pw * 4 + rsub : row + 32;
          const int nps = min(last ? ps + 1 : ps, MTC_PASSES - 1);   // (the tile's very last prefetch is unused)
          // entering: hvA holds (row, half 0) in flight, nidx the indices of (row, half 1)
          if (full) {
            issue(hvB, ps);
            load_idx(nrow, 0, true);
            consume(hvA, row, 0, acc, true);
            issue(hvA, nps);
            load_idx(nrow, 1, true);
            consume(hvB, row, 1, acc, true);
          } else {
            issue(hvB, ps);
            load_idx(nrow, 0, false);
            consume(hvA, row, 0, acc, false);
            issue(hvA, nps);
            load_idx(nrow, 1, false);
            consume(hvB, row, 1, acc, false);
          }
          store_row(row, acc, sc);
        }
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tc::mbar_arrive(&a_full[st]);
        }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(rec_empty);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tmem_dealloc<512>(tmem_base);
  }
}

template <int ACT>
__global__ void __launch_bounds__(MTC_THREADS, 1) mp_layer_tc_kernel(const MpTcArgs p) {
  mp_layer_tc_body<ACT, false, false>(p);
}
// Single-accumulator form (option "mp_single_acc"): main and correction products accumulate into ONE 256-column
// accumulator (the lo operand images are not scaled by 2^11; W' is pre-scaled by a power of two so that its lo image
// stays in the normal fp16 range), so tensor memory holds two accumulator sets and the epilogue of tile t runs under
// the MMAs of tile t + 1.  Price: 144 instead of 48 round-toward-zero steps at full magnitude per output (the
// position-dependent compensation follows the 144-instruction order), error at the level of the exact-FP32 kernels.
// Variable-tile form for calls of less than one wave (see MpTcArgs): same arithmetic per atom, same bits.
template <int ACT>
__global__ void __launch_bounds__(MTC_THREADS, 1) mp_layer_tc_vt_kernel(const MpTcArgs p) {
  mp_layer_tc_body<ACT, false, false, true>(p);
}

template <int ACT>
__global__ void __launch_bounds__(MTC_THREADS, 1) mp_layer_tc1_kernel(const MpTcArgs p) {
  mp_layer_tc_body<ACT, false, true>(p);
}
// the same with the K loop cut into p.nseg accumulation chains (option "mp_chain_segments" > 1)
template <int ACT>
__global__ void __launch_bounds__(MTC_THREADS, 1) mp_layer_tc_seg_kernel(const MpTcArgs p) {
  mp_layer_tc_body<ACT, true, false>(p);
}

}  // namespace nmr

namespace nmr {

// ----------------------------------------------------------------------------------
// Node MLP + readout on tensor cores, F = 256 (FCBlock + out_layer + peak standardisation;
// model.py:191-196, 268-273).  One CTA = 128 atoms per tile, persistent over tiles.
// The node tile X lives in shared memory as the fp16x3 A operand (hi | lo, 8 K-chunks of
// 32 features); every residual layer X <- act(X W + b) + X is
//   MMA:      D[128 x 256] = X * W     two N = 128 halves per K-chunk, main / corr accumulators in TMEM
//   epilogue: thread = (atom row, 64-column quarter): D -> RZ compensation -> bias -> act -> + x_old
//             (x_old is re-assembled from the hi/lo pair it is about to overwrite: 22 mantissa bits,
//              round-to-nearest) -> split -> written back in place as the next layer's operand.
// The last layer (256 -> 128, no residual) leaves Z in fp32 in shared memory (overlaying X, whose
// MMAs have completed) for the warp-per-atom readout.  W streams through a ring of 16 KB bulk
// copies (one K-chunk x one N-half).  Rows are pre-scaled by a power of two from an a-priori
// affine bound of |x| per layer (fp16 range); the scale is exact to undo.
//   warp 0: W loader   warp 1: MMA issuer + TMEM owner   warps 2-17: load / epilogue / readout
// ----------------------------------------------------------------------------------
struct FcTcArgs {
  const float* nodes;        // [n_atoms, 256]
  const float* atoms;        // [n_atoms, C]
  float* peaks;              // [n_atoms]
  float* fc_nodes;           // optional [n_atoms, 128]
  int64_t n_atoms;
  int C;
  const uint8_t* Wimg;       // layers 0..n-2: [8 chunks][2 halves][hi 8192 | lo 8192]; last: [8 chunks][hi | lo]
  const float* bias;         // [n_layers][256]
  float gain[MAX_DENSE];     // |x_l| <= gain[l] * max|x_0| + offs[l]   (input of layer l)
  float offs[MAX_DENSE];
  int n_layers;
  int act;
  float corr;                // 1 + c: round-toward-zero compensation for a 16-instruction chain
  const float* Wo;           // [128, C]
  const float* bo;           // [C]
  const float* peak_std;     // [C]
  const float* peak_avg;     // [C]
};

constexpr int FTC_THREADS = 576;
constexpr int FTC_RING = 5;
constexpr int FTC_LDZ = 132;
constexpr size_t FTC_X_BYTES = 8 * 16384;
constexpr size_t FTC_SMEM = 1024 + FTC_X_BYTES + FTC_RING * 16384 + MAX_DENSE * 256 * 4 + MAX_DENSE * 128 * 4 + 256;
static_assert(FTC_SMEM <= 227 * 1024, "node-MLP tensor-core kernel exceeds the 227 KB shared-memory limit");
static_assert(128 * FTC_LDZ * 4 <= FTC_X_BYTES, "Z overlays the X operand");

template <int ACT>
__device__ __forceinline__ void fc_readout_tc_body(const FcTcArgs& p) {
  constexpr int SLOT = 16384;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xs = smem;                                      // [8 chunks][hi 8192 | lo 8192]; later Z [128][FTC_LDZ] fp32
  uint8_t* ring = xs + FTC_X_BYTES;                        // [RING][hi 8192 | lo 8192]
  float* bias_s = reinterpret_cast<float*>(ring + FTC_RING * 16384);   // [n_layers][256]
  float* rs = bias_s + MAX_DENSE * 256;                    // [n_layers][128] per-row 2^-s of the layer's input
  uint64_t* bars = reinterpret_cast<uint64_t*>(rs + MAX_DENSE * 128);
  uint64_t* w_full = bars;                                 // [RING]
  uint64_t* w_empty = w_full + FTC_RING;                   // [RING]
  uint64_t* x_full = w_empty + FTC_RING;
  uint64_t* d_full = x_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < FTC_RING; ++i) {
      tc::mbar_init(&w_full[i], 1);
      tc::mbar_init(&w_empty[i], 1);
    }
    tc::mbar_init(x_full, 16);
    tc::mbar_init(d_full, 1);
    tc::mbar_fence_init();
  }
  const int nl = p.n_layers;
  for (int i = tid; i < nl * 256; i += FTC_THREADS) bias_s[i] = p.bias[i];
  if (warp == 1) {
    tc::tmem_alloc<512>(tmem_slot);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t n_tiles = (p.n_atoms + 127) / 128;
  const int64_t tile_first = (int64_t)blockIdx.x;
  const int64_t tile_step = (int64_t)gridDim.x;
  const int64_t tile_end = n_tiles;

  if (warp == 0) {
    // ===================== W loader =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = tile_first; tile < tile_end; tile += tile_step)
        for (int l = 0; l < nl; ++l) {
          const int n_slots = (l == nl - 1) ? 8 : 16;
          const uint8_t* src = p.Wimg + (size_t)l * 16 * 16384;
          for (int q = 0; q < n_slots; ++q, ++it) {
            const uint32_t slot = it % FTC_RING, ph = (it / FTC_RING) & 1;
            tc::mbar_wait(&w_empty[slot], ph ^ 1);
            tc::mbar_expect_tx(&w_full[slot], SLOT);
            tc::bulk_g2s(ring + slot * SLOT, src + (size_t)q * 16384, 16384, &w_full[slot]);
          }
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // per 128-column half h: tensor-memory columns [256h, 256h+128) = main, [256h+128, 256h+256) = corr, so that
      // x_hi * [w_hi | w_lo] -> [main | corr] is ONE N = 256 instruction (the slot holds the hi and lo tiles adjacently)
      const uint32_t idesc = tc::make_idesc_f16(128, 128), idesc2 = tc::make_idesc_f16(128, 256);
      uint32_t it = 0, px = 0;
      for (int64_t tile = tile_first; tile < tile_end; tile += tile_step)
        for (int l = 0; l < nl; ++l) {
          const int halves = (l == nl - 1) ? 1 : 2;
          tc::mbar_wait(x_full, px);
          px ^= 1;
          tc::tc_fence_after();
          for (int c = 0; c < 8; ++c) {
            const uint64_t ah = tc::make_desc_sw64(tc::smem_u32(xs + c * 16384));
            const uint64_t al = tc::make_desc_sw64(tc::smem_u32(xs + c * 16384 + 8192));
            for (int hf = 0; hf < halves; ++hf, ++it) {
              const uint32_t slot = it % FTC_RING;
              tc::mbar_wait(&w_full[slot], (it / FTC_RING) & 1);
              tc::tc_fence_after();
              const uint64_t bh = tc::make_desc_sw64(tc::smem_u32(ring + slot * SLOT));
              const uint32_t d_main = tmem_base + (uint32_t)hf * 256u, d_corr = d_main + 128u;
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 2);
                tc::umma_f16(d_main, ah + adv, bh + adv, idesc2, (c | ks) != 0);
                tc::umma_f16(d_corr, al + adv, bh + adv, idesc, 1);
              }
              tc::umma_commit(&w_empty[slot]);
            }
          }
          tc::umma_commit(d_full);
        }
    }
  } else {
    // ===================== load / epilogue / readout warps (16) =====================
    const int we = warp - 2;                 // 0..15
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    const int cq = we >> 2;                  // column quarter: 64 columns of the residual layers, 32 of the last
    const int row = q * 32 + lane;
    const uint32_t xs_a = tc::smem_u32(xs);
    const uint32_t rs_a = tc::smem_u32(rs);
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    float* Z = reinterpret_cast<float*>(xs);
    uint32_t pd = 0;
    for (int64_t tile = tile_first; tile < tile_end; tile += tile_step) {
      const int64_t a0 = tile * 128;
      const int rows = (int)min((int64_t)128, p.n_atoms - a0);
      // ---- stage the node tile: warp `we` loads rows we, we+16, ... (1 KB coalesced per row),
      //      lane = 8 consecutive features = one 16-byte piece of the hi tile and one of the lo tile
      asm volatile("bar.sync 1, 512;" ::: "memory");     // previous tile's readout has finished with Z
#pragma unroll 4
      for (int i = 0; i < 8; ++i) {
        const int r = we + i * 16;
        float x[8];
        if (r < rows) {
          const float* src = p.nodes + (a0 + r) * 256 + lane * 8;
          const float4 x0 = tc::ldg128(src), x1 = tc::ldg128(src + 4);
          x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w;
          x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) x[u] = 0.0f;
        }
        float m = 0.0f;
#pragma unroll
        for (int u = 0; u < 8; ++u) m = fmaxf(m, fabsf(x[u]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        // per-layer power-of-two input scale from the affine bound of |x_l|
        float s0 = 1.0f;
        if (lane < nl) {
          const float b = fmaf(p.gain[lane], m, p.offs[lane]);
          int s = ((__float_as_int(b) >> 23) & 0xff) - 127 - 14;
          s = (b > 0.0f && b < 3.0e38f) ? min(max(s, 0), 100) : 0;
          rs[lane * 128 + r] = tc::pow2f_exact(-s);
          if (lane == 0) s0 = tc::pow2f_exact(-s);
        }
        s0 = __shfl_sync(0xffffffffu, s0, 0);
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] *= s0;
        uint4 hi, lo;
        tc::split8_f16(x, hi, lo);
        const uint32_t off = xs_a + (uint32_t)(lane >> 2) * 16384u + tc::sw64_chunk_offset(r, lane & 3);
        tc::sts128(off, hi);
        tc::sts128(off + 8192u, lo);
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(x_full);
      }

      // ---- residual layers: thread = (row, 64 columns = K-chunks 2cq, 2cq+1 of the next layer's operand)
      for (int l = 0; l + 1 < nl; ++l) {
        tc::mbar_wait(d_full, pd);
        pd ^= 1;
        tc::tc_fence_after();
        const uint32_t bl_a = tc::smem_u32(bias_s + l * 256 + cq * 64);
        const float s_in = tc::lds32f(rs_a + (uint32_t)(l * 128 + row) * 4u);        // 2^-s of this layer's input
        const float s_nx = tc::lds32f(rs_a + (uint32_t)((l + 1) * 128 + row) * 4u);  // 2^-s' of the next layer's input
        const float s_old = __fdiv_rn(1.0f, s_in);         // exact: s_in is a power of two
        const float s_out = p.corr * s_old;                // accumulator -> true scale, with the RZ compensation
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          float v[16];
          const int col = cq * 64 + cc * 16;
          const uint32_t t_main = t_lane + (uint32_t)(col >> 7) * 256u + (uint32_t)(col & 127);
          tc::tmem_ld16_combined(t_main, t_main + 128u, v);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int k0 = col + hh * 8;
            const uint32_t off = xs_a + (uint32_t)(k0 >> 5) * 16384u + tc::sw64_chunk_offset(row, (k0 & 31) >> 3);
            uint4 ohi, olo;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(ohi.x), "=r"(ohi.y), "=r"(ohi.z), "=r"(ohi.w) : "r"(off));
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(olo.x), "=r"(olo.y), "=r"(olo.z), "=r"(olo.w) : "r"(off + 8192u));
            const float4 b0 = tc::lds128(bl_a + (cc * 16 + hh * 8) * 4), b1 = tc::lds128(bl_a + (cc * 16 + hh * 8) * 4 + 16);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            const uint32_t oh[4] = {ohi.x, ohi.y, ohi.z, ohi.w}, ol[4] = {olo.x, olo.y, olo.z, olo.w};
            float x[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&oh[i]));
              const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&ol[i]));
              const float old0 = fmaf(fl.x, tc::LO_UNSCALE, fh.x) * s_old;
              const float old1 = fmaf(fl.y, tc::LO_UNSCALE, fh.y) * s_old;
              x[2 * i] = (act_t<ACT>(fmaf(v[hh * 8 + 2 * i], s_out, bb[2 * i])) + old0) * s_nx;
              x[2 * i + 1] = (act_t<ACT>(fmaf(v[hh * 8 + 2 * i + 1], s_out, bb[2 * i + 1])) + old1) * s_nx;
            }
            uint4 hi, lo;
            tc::split8_f16(x, hi, lo);
            tc::sts128(off, hi);
            tc::sts128(off + 8192u, lo);
          }
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          tc::mbar_arrive(x_full);
        }
      }
      // ---- last layer: Z = act(D + b), 128 columns; each (row, quarter) thread takes 32 of them
      {
        const int l = nl - 1;
        tc::mbar_wait(d_full, pd);
        pd ^= 1;
        tc::tc_fence_after();
        const float s_out = p.corr * __fdiv_rn(1.0f, tc::lds32f(rs_a + (uint32_t)(l * 128 + row) * 4u));
        const uint32_t bl_a = tc::smem_u32(bias_s + l * 256 + cq * 32);
        float* zr = Z + row * FTC_LDZ + cq * 32;
        // all MMAs that read X have completed (d_full); every warp must be past its own X reads too
        asm volatile("bar.sync 1, 512;" ::: "memory");
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          float v[16];
          const int col = cq * 32 + cc * 16;
          tc::tmem_ld16_combined(t_lane + col, t_lane + 128u + col, v);
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b4 = tc::lds128(bl_a + (cc * 16 + i) * 4);
            float4 o;
            o.x = act_t<ACT>(fmaf(v[i + 0], s_out, b4.x));
            o.y = act_t<ACT>(fmaf(v[i + 1], s_out, b4.y));
            o.z = act_t<ACT>(fmaf(v[i + 2], s_out, b4.z));
            o.w = act_t<ACT>(fmaf(v[i + 3], s_out, b4.w));
            *reinterpret_cast<float4*>(zr + cc * 16 + i) = o;
          }
        }
        tc::tc_fence_before();
        asm volatile("bar.sync 1, 512;" ::: "memory");
      }
      if (p.fc_nodes != nullptr) {
        for (int i = tid - 64; i < rows * 128; i += 512) {
          const int r = i >> 7, k = i & 127;
          p.fc_nodes[(a0 + r) * 128 + k] = Z[r * FTC_LDZ + k];
        }
      }
      // ---- readout: peaks = sum_c (z . Wo[:,c] + bo[c]) * a[c] * std[c] + a[c] * avg[c]; warp per atom,
      //      skipping classes with a[c] == 0 (exact for finite activations).  The warp's 8 atom rows are
      //      fetched up front (lane = class, all loads in flight together); classes beyond 32 take the slow loop.
      {
        float av[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = we * 8 + i;
          av[i] = (r < rows && lane < p.C) ? __ldg(p.atoms + (a0 + r) * p.C + lane) : 0.0f;
        }
        const float std_l = lane < p.C ? __ldg(p.peak_std + lane) : 0.0f;
        const float avg_l = lane < p.C ? __ldg(p.peak_avg + lane) : 0.0f;
        const float bo_l = lane < p.C ? __ldg(p.bo + lane) : 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = we * 8 + i;
          if (r >= rows) continue;               // warp-uniform
          const float* zr = Z + r * FTC_LDZ;
          const float4 z4 = *reinterpret_cast<const float4*>(zr + lane * 4);   // this lane's 4 features
          float peak = 0.0f;
          unsigned nzmask = __ballot_sync(0xffffffffu, av[i] != 0.0f);
          while (nzmask) {
            const int c = __ffs(nzmask) - 1;
            nzmask &= nzmask - 1;
            const float* wo = p.Wo + (lane * 4) * p.C + c;
            float dot = z4.x * __ldg(wo);
            dot = fmaf(z4.y, __ldg(wo + p.C), dot);
            dot = fmaf(z4.z, __ldg(wo + 2 * p.C), dot);
            dot = fmaf(z4.w, __ldg(wo + 3 * p.C), dot);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            const float a = __shfl_sync(0xffffffffu, av[i], c);
            const float full = dot + __shfl_sync(0xffffffffu, bo_l, c);
            peak += full * a * __shfl_sync(0xffffffffu, std_l, c) + a * __shfl_sync(0xffffffffu, avg_l, c);
          }
          for (int c = 32; c < p.C; ++c) {       // (num_elem > 32 only)
            const float a = p.atoms[(a0 + r) * p.C + c];
            if (a != 0.0f) {
              float dot = 0.0f;
              for (int k = lane; k < 128; k += 32) dot = fmaf(zr[k], __ldg(p.Wo + k * p.C + c), dot);
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
              peak += (dot + p.bo[c]) * a * p.peak_std[c] + a * p.peak_avg[c];
            }
          }
          if (lane == 0) p.peaks[a0 + r] = peak;
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tmem_dealloc<512>(tmem_base);
  }
}

template <int ACT>
__global__ void __launch_bounds__(FTC_THREADS, 1) fc_readout_tc_kernel(const FcTcArgs p) {
  fc_readout_tc_body<ACT>(p);
}

}  // namespace nmr
