// Error of the kernels' MUFU-based softplus (common.cuh softplus_f) against fp64, per range of x:
// mean signed and max absolute error, in units of 2^-24 (and relative to the result).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o softplus_err softplus_err.cu
#include <cstdio>
#include <cmath>
#include <vector>
#include "../../nmrgnn_b200/csrc/common.cuh"

__device__ __forceinline__ float softplus_v2(float x) {
  // candidate: t = e^-|x| (MUFU), log1p(t) by the compensated form  lg(u) + (t - (u - 1)) / u  for u = fl(1 + t)
  float t, l;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-1.4426950408889634f * fabsf(x)));
  const float u = 1.0f + t;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u));
  const float c = __fdividef(t - (u - 1.0f), u);
  return fmaf(l, 0.6931471805599453f, fmaxf(x, 0.0f)) + c;
}

__global__ void eval(const float* x, float* y0, float* y1, float* y2, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  y0[i] = nmr::softplus_f(x[i]);
  y1[i] = softplus_v2(x[i]);
  y2[i] = fmaxf(x[i], 0.0f) + log1pf(expf(-fabsf(x[i])));
}

int main() {
  const int n = 1 << 22;
  std::vector<float> x(n), y0(n), y1(n), y2(n);
  for (int i = 0; i < n; ++i) x[i] = -30.0f + 60.0f * (float)i / (float)(n - 1) + 1e-3f * sinf((float)i);
  float *dx, *d0, *d1, *d2;
  cudaMalloc(&dx, n * 4); cudaMalloc(&d0, n * 4); cudaMalloc(&d1, n * 4); cudaMalloc(&d2, n * 4);
  cudaMemcpy(dx, x.data(), n * 4, cudaMemcpyHostToDevice);
  eval<<<(n + 255) / 256, 256>>>(dx, d0, d1, d2, n);
  cudaMemcpy(y0.data(), d0, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(y1.data(), d1, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(y2.data(), d2, n * 4, cudaMemcpyDeviceToHost);
  const double edges[] = {-30, -20, -14, -10, -7, -5, -3, -2, -1, -0.5, 0, 0.5, 1, 2, 3, 5, 7, 10, 14, 20, 30};
  const char* names[3] = {"softplus_f (MUFU ex2+lg2)", "compensated 1+t rounding", "expf/log1pf"};
  std::vector<float>* ys[3] = {&y0, &y1, &y2};
  for (int v = 0; v < 3; ++v) {
    printf("%s\n   x range        mean signed err  rms err   max |err|  (x 2^-24)   mean rel err   max rel err\n", names[v]);
    for (int b = 0; b + 1 < (int)(sizeof(edges) / sizeof(double)); ++b) {
      double s = 0, s2 = 0, mx = 0, sr = 0, mr = 0;
      long cnt = 0;
      for (int i = 0; i < n; ++i) {
        if (x[i] < edges[b] || x[i] >= edges[b + 1]) continue;
        const double xe = (double)x[i];
        const double ref = fmax(xe, 0.0) + log1p(exp(-fabs(xe)));
        const double e = (double)(*ys[v])[i] - ref;
        s += e; s2 += e * e; mx = fmax(mx, fabs(e)); sr += e / ref; mr = fmax(mr, fabs(e / ref)); ++cnt;
      }
      const double u = 16777216.0;
      printf("  [%5.1f, %5.1f)   %+10.4f     %8.4f   %8.3f              %+.3e    %.3e\n", edges[b], edges[b + 1], s / cnt * u,
             sqrt(s2 / cnt) * u, mx * u, sr / cnt, mr);
    }
  }
  return 0;
}
