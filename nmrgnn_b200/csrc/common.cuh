// Shared device helpers for the nmrgnn_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nmr {

enum : int { ACT_LINEAR = 0, ACT_SOFTPLUS = 1, ACT_RELU = 2, ACT_TANH = 3 };

// softplus(x) = max(x,0) + log1p(t), t = exp(-|x|) in (0, 1].
// t comes from MUFU ex2 (relative error 2^-22, harmless); log1p(t) = t * P9(t) is a degree-9 near-minimax polynomial
// for log1p(t)/t on [0, 1] (approximation error 4.8e-9, measured relative error of the fp32 evaluation <= 1.7e-7, mean
// 1e-8).  The obvious MUFU form lg2.approx(1 + t) is NOT good enough here: near 1 lg2.approx carries an absolute error
// of ~2^-23 with a constant positive offset (+1.2 x 2^-24 measured on B200, tools/microbench/softplus_err.cu), i.e. every
// small activation (x < -2) came out too large by the same 7e-8 -- up to 100 % of its value -- and the next layer sums
// that offset coherently over hundreds of inputs; that bias was the largest single contribution to the peak error of
// both compute paths (profiles/r02_softplus_bias.md).  The reference's TF kernel computes log(exp(x) + 1) with
// thresholds (tensorflow/core/kernels/softplus_op.h); this form is closer to the exact function than that one.
__device__ __forceinline__ float softplus_f(float x) {
  float t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-1.4426950408889634f * fabsf(x)));
  float p = -3.256378463e-03f;
  p = fmaf(p, t, 1.990716159e-02f);
  p = fmaf(p, t, -5.706420168e-02f);
  p = fmaf(p, t, 1.061426476e-01f);
  p = fmaf(p, t, -1.531186253e-01f);
  p = fmaf(p, t, 1.967811733e-01f);
  p = fmaf(p, t, -2.495455891e-01f);
  p = fmaf(p, t, 3.333000541e-01f);
  p = fmaf(p, t, -4.999990463e-01f);
  p = fmaf(p, t, 1.0f);
  return fmaf(p, t, fmaxf(x, 0.0f));
}

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ACT_SOFTPLUS: return softplus_f(x);
    case ACT_RELU: return fmaxf(x, 0.0f);
    case ACT_TANH: return tanhf(x);
    default: return x;
  }
}

// Compile-time activation: the hot epilogues are instantiated per activation, because a runtime switch inside
// the per-element code gets if-converted and every element then pays for softplus AND tanh (4 MUFU, ~35
// instructions instead of ~15).
template <int ACT>
__device__ __forceinline__ float act_t(float x) {
  if (ACT == ACT_SOFTPLUS) return softplus_f(x);
  if (ACT == ACT_RELU) return fmaxf(x, 0.0f);
  if (ACT == ACT_TANH) return tanhf(x);
  return x;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// --------------------------------------------------------------------------------
// Block-level FP32 GEMM core used by the edge-MLP, MP-layer and node-MLP kernels:
//     acc[128 x BN] += A[128 x KT] * W[KT x BN]
// A lives in shared memory (row-major, stride `lda` floats with lda % 4 == 2 so the
// two row groups of a warp hit different banks); W is streamed from global/L2 in
// 16-row chunks through a 2-stage cp.async ring.  NT = 2*BN threads; warp (wm, wn)
// owns a 16 x 128 sub-tile, each thread an 8 x 8 register tile made of rows
// r0..r0+7 and columns {c0..c0+3, c0+64..c0+67}.
// --------------------------------------------------------------------------------
template <int BN, int NT>
struct TileGemm {
  static constexpr int KC = 16;
  static constexpr int STAGE_FLOATS = KC * BN;
  static constexpr int SMEM_FLOATS = 2 * STAGE_FLOATS;

  __device__ static __forceinline__ void thread_origin(int tid, int& r0, int& c0) {
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp & 7, wn = warp >> 3;
    r0 = wm * 16 + (lane >> 4) * 8;
    c0 = wn * 128 + (lane & 15) * 4;
  }

  __device__ static __forceinline__ void load_chunk(float* stage, const float* __restrict__ W, int ldw,
                                                    int k0, int tid) {
    constexpr int F4_PER_ROW = BN / 4;
    constexpr int TOTAL = KC * F4_PER_ROW;
#pragma unroll
    for (int i = tid; i < TOTAL; i += NT) {
      const int r = i / F4_PER_ROW, c4 = i % F4_PER_ROW;
      cp_async16(stage + r * BN + c4 * 4, W + (size_t)(k0 + r) * ldw + c4 * 4);
    }
  }

  // All NT threads must call (they cooperate on the W stream); `compute` = false lets
  // a warp help with the loads only.  KT must be a multiple of 16.
  __device__ static __forceinline__ void run(float (&acc)[8][8], const float* __restrict__ As, int lda,
                                             const float* __restrict__ W, int ldw, int KT, float* Bs,
                                             bool compute) {
    const int tid = threadIdx.x;
    int r0, c0;
    thread_origin(tid, r0, c0);
    const int nch = KT / KC;
    load_chunk(Bs, W, ldw, 0, tid);
    cp_async_commit();
    for (int ch = 0; ch < nch; ++ch) {
      cp_async_wait_all();
      __syncthreads();
      if (ch + 1 < nch) {
        load_chunk(Bs + ((ch + 1) & 1) * STAGE_FLOATS, W, ldw, (ch + 1) * KC, tid);
        cp_async_commit();
      }
      if (compute) {
        const float* a_ptr = As + r0 * lda + ch * KC;
        const float* b_ptr = Bs + (ch & 1) * STAGE_FLOATS + c0;
#pragma unroll
        for (int kk = 0; kk < KC; kk += 2) {
          float2 a[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float2*>(a_ptr + i * lda + kk);
          const float4 b00 = *reinterpret_cast<const float4*>(b_ptr + kk * BN);
          const float4 b01 = *reinterpret_cast<const float4*>(b_ptr + kk * BN + 64);
          const float4 b10 = *reinterpret_cast<const float4*>(b_ptr + (kk + 1) * BN);
          const float4 b11 = *reinterpret_cast<const float4*>(b_ptr + (kk + 1) * BN + 64);
          const float b0[8] = {b00.x, b00.y, b00.z, b00.w, b01.x, b01.y, b01.z, b01.w};
          const float b1[8] = {b10.x, b10.y, b10.z, b10.w, b11.x, b11.y, b11.z, b11.w};
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              acc[i][j] = fmaf(a[i].x, b0[j], acc[i][j]);
              acc[i][j] = fmaf(a[i].y, b1[j], acc[i][j]);
            }
        }
      }
    }
    __syncthreads();  // Bs (and As) may be overwritten by the caller after this
  }
};

}  // namespace nmr
