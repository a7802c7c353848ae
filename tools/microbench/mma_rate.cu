// Microbenchmark (run under gpurun): cycles per tcgen05.mma.kind::f16 instruction (M = 128, K = 16) issued
// back to back by one thread, for N = 16 / 64 / 128 / 256, A operand from shared memory (SS) or tensor memory (TS),
// alone on the SM (no other shared-memory traffic).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../nmrgnn_b200/csrc/tc_common.cuh"
using namespace nmr;

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int ts, int n_mma, int n_acc, int style, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a = smem;                  // 128 rows x 64 B
  uint8_t* b = smem + 8192;           // 256 rows x 64 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192 + 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (8192 + 16384) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) {
    tc::mbar_init(bar, 1);
    tc::mbar_fence_init();
  }
  if (threadIdx.x < 32) tc::tmem_alloc<512>(slot);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = __shfl_sync(0xffffffffu, *slot, 0);   // warp-uniform by construction
  if (style == 0 && threadIdx.x == 0) {
    // style 0: one thread owns the loop (what the round-1 kernels did first)
    const uint32_t idesc = tc::make_idesc_f16(128, N);
    const uint64_t ad = tc::make_desc_sw64(tc::smem_u32(a)), bd = tc::make_desc_sw64(tc::smem_u32(b));
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t d = tm + (uint32_t)(i & (n_acc - 1)) * (uint32_t)N;       // rotate accumulators
      if (ts) tc::umma_f16_ts(d, tm + 448u + (uint32_t)(i & 1) * 8u, bd, idesc, 1);
      else tc::umma_f16(d, ad + (uint64_t)((i & 1) * 2), bd + (uint64_t)((i & 1) * 2), idesc, 1);
    }
    tc::umma_commit(bar);
    tc::mbar_wait(bar, 0);
    out[0] = (unsigned long long)(clock64() - t0);
  }
  if (style == 1 && threadIdx.x < 32) {
    // style 1: the whole warp runs the loop, one elected lane issues
    const uint32_t idesc = tc::make_idesc_f16(128, N);
    const uint64_t ad = tc::make_desc_sw64(tc::smem_u32(a)), bd = tc::make_desc_sw64(tc::smem_u32(b));
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t d = tm + (uint32_t)(i & (n_acc - 1)) * (uint32_t)N;
      if (tc::elect_one()) {
        if (ts) tc::umma_f16_ts(d, tm + 448u + (uint32_t)(i & 1) * 8u, bd, idesc, 1);
        else tc::umma_f16(d, ad + (uint64_t)((i & 1) * 2), bd + (uint64_t)((i & 1) * 2), idesc, 1);
      }
      __syncwarp();
    }
    if (tc::elect_one()) tc::umma_commit(bar);
    __syncwarp();
    tc::mbar_wait(bar, 0);
    if (threadIdx.x == 0) out[0] = (unsigned long long)(clock64() - t0);
  }
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc<512>(tm);
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  for (int style = 0; style < 2; ++style)
    for (int ts = 0; ts < 2; ++ts)
      for (int N : {16, 64, 128, 256})
        for (int n_acc : {1, N <= 128 ? 2 : 1}) {
          const int n_mma = 4096;
          rate_kernel<<<1, 128, 40960>>>(N, ts, 64, n_acc, style, d);
          rate_kernel<<<1, 128, 40960>>>(N, ts, n_mma, n_acc, style, d);
          cudaDeviceSynchronize();
          unsigned long long c;
          cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          printf("%s %s N=%3d accumulators=%d: %7.1f cycles per MMA  (%s)\n", style ? "elect_one in a converged warp" : "single-thread loop          ",
                 ts ? "TS" : "SS", N, n_acc, (double)c / n_mma, cudaGetErrorString(cudaGetLastError()));
        }
  return 0;
}
