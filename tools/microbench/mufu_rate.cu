// MUFU (ex2 / lg2) issue rate per SM on this GPU: 16 warps per SM, 8 independent chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(512, 1) mufu_kernel(float* out, long long* cyc, int iters) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 0.001f * (float)(threadIdx.x + i + 1);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      else if (OP == 1) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      else {   // softplus-like pair: ex2 then lg2 of 1 + t (adds one FADD per pair)
        float t;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x[i]));
        asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(x[i]) : "f"(1.0f + t));
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  long long* cyc;
  cudaMalloc(&out, sms * 512 * sizeof(float));
  cudaMallocManaged(&cyc, sms * sizeof(long long));
  const int iters = 4096;
  for (int op = 0; op < 3; ++op) {
    for (int rep = 0; rep < 2; ++rep) {
      if (op == 0) mufu_kernel<0><<<sms, 512>>>(out, cyc, iters);
      else if (op == 1) mufu_kernel<1><<<sms, 512>>>(out, cyc, iters);
      else mufu_kernel<2><<<sms, 512>>>(out, cyc, iters);
      cudaDeviceSynchronize();
    }
    double mean = 0;
    for (int b = 0; b < sms; ++b) mean += (double)cyc[b];
    mean /= sms;
    const double ops = 512.0 * 8 * iters * (op == 2 ? 2 : 1);
    printf("%s: %.0f cycles per CTA, %.2f MUFU results per clock per SM (err %s)\n",
           op == 0 ? "ex2.approx" : op == 1 ? "lg2.approx" : "ex2 + add + lg2", mean, ops / mean,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
