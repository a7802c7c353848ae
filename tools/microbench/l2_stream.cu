// Microbenchmark (run under gpurun): aggregate L2 -> shared-memory bandwidth of 1-D bulk async copies when every SM
// re-streams the same small (L2-resident) weight image, as the persistent tensor-core kernels do; and the latency
// of one dependent bulk copy.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_stream l2_stream.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../nmrgnn_b200/csrc/tc_common.cuh"
using namespace nmr;

template <int SLOT_BYTES, int SLOTS>
__global__ void __launch_bounds__(64, 1) stream_kernel(const uint8_t* src, size_t img_bytes, int iters, unsigned long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)SLOT_BYTES * SLOTS);
  if (threadIdx.x == 0) {
    for (int i = 0; i < SLOTS; ++i) tc::mbar_init(&full[i], 1);
    tc::mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    const int per_img = (int)(img_bytes / SLOT_BYTES);
    // keep SLOTS copies in flight; re-issue a slot as soon as its data has landed (consumer = nobody)
    for (int i = 0; i < SLOTS && i < iters; ++i) {
      tc::mbar_expect_tx(&full[i], SLOT_BYTES);
      tc::bulk_g2s(smem + (size_t)i * SLOT_BYTES, src + (size_t)(i % per_img) * SLOT_BYTES, SLOT_BYTES, &full[i]);
    }
    for (int i = 0; i < iters; ++i) {
      const int s = i % SLOTS;
      tc::mbar_wait(&full[s], (i / SLOTS) & 1);
      const int nx = i + SLOTS;
      if (nx < iters) {
        tc::mbar_expect_tx(&full[s], SLOT_BYTES);
        tc::bulk_g2s(smem + (size_t)s * SLOT_BYTES, src + (size_t)(nx % per_img) * SLOT_BYTES, SLOT_BYTES, &full[s]);
      }
    }
    cyc[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
}

template <int SLOT_BYTES, int SLOTS>
void run(const uint8_t* d, size_t img, int grid, unsigned long long* dc) {
  const int iters = 4096 * (16384 / SLOT_BYTES > 0 ? 16384 / SLOT_BYTES : 1) / 2;
  size_t smem = (size_t)SLOT_BYTES * SLOTS + 2048;
  cudaFuncSetAttribute(stream_kernel<SLOT_BYTES, SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  stream_kernel<SLOT_BYTES, SLOTS><<<grid, 64, smem>>>(d, img, 64, dc);
  cudaEventRecord(e0);
  stream_kernel<SLOT_BYTES, SLOTS><<<grid, 64, smem>>>(d, img, iters, dc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long c0;
  cudaMemcpy(&c0, dc, 8, cudaMemcpyDeviceToHost);
  double bytes = (double)grid * iters * SLOT_BYTES;
  printf("grid %3d slot %6d B x %d in flight: %8.1f GB/s aggregate, %6.1f B/clk/SM, %7.1f cycles per slot (err %s)\n", grid,
         SLOT_BYTES, SLOTS, bytes / ms / 1e6, bytes / grid / (double)c0, (double)c0 / iters, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const size_t img = 786432;  // one MP layer's W' image
  uint8_t* d;
  cudaMalloc(&d, img);
  cudaMemset(d, 1, img);
  unsigned long long* dc;
  cudaMalloc(&dc, 8 * 1024);
  for (int grid : {1, 148}) {
    run<16384, 1>(d, img, grid, dc);
    run<16384, 2>(d, img, grid, dc);
    run<16384, 5>(d, img, grid, dc);
    run<16384, 8>(d, img, grid, dc);
    run<32768, 3>(d, img, grid, dc);
    run<32768, 6>(d, img, grid, dc);
    run<8192, 8>(d, img, grid, dc);
    run<4096, 16>(d, img, grid, dc);
  }
  return 0;
}
