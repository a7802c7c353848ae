// Reassembly of the per-rank peaks over NVLink peer memory, without a collective library call.
//
// Graphs never interact (tf.gather indexes inside one graph, nmrgnn/layers.py:33), so a batch is sharded by whole
// graphs over one process per GPU and the only exchange of the path is "every rank gets every rank's peaks".  Each
// rank owns a gather buffer [2 parities][world][capacity] floats and a flag word per source rank, both exported to the
// other processes as CUDA IPC handles (nmrgnn_comm_local / nmrgnn_comm_init).  After its forward a rank
//   1. stores its peaks into slot `rank` of EVERY rank's buffer (16-byte stores, straight over NVLink / NVSwitch),
//   2. fences at system scope; the last CTA to finish publishes the call's epoch in flags[rank] of every rank
//      (st.release.sys),
//   3. a one-warp wait kernel on the same stream spins (ld.acquire.sys) until every source's flag has reached the epoch.
// The buffers alternate by epoch parity: a rank may already scatter epoch e+1 while a slower rank still reads epoch e;
// epoch e+2 reuses the parity of e only after that rank has published e+1, which it enqueues after consuming e.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace nmr {

constexpr int PEER_MAX_WORLD = 16;

struct PeerArgs {
  const float* src;                  // this rank's peaks (its own slot of its own buffer)
  int64_t n;                         // number of floats
  int world, rank;
  int64_t slot_off;                  // float offset of (parity, rank) inside every gather buffer
  float* gbuf[PEER_MAX_WORLD];       // gather buffer of rank r as mapped into this process
  uint32_t* flags[PEER_MAX_WORLD];   // flag words [world] of rank r
  uint32_t epoch;
  unsigned int* done_ctr;            // device counter, zero between launches
};

__global__ void __launch_bounds__(256) peer_scatter_signal_kernel(const PeerArgs p) {
  const int64_t n4 = p.n >> 2;
  const float4* s4 = reinterpret_cast<const float4*>(p.src);
  for (int r = 0; r < p.world; ++r) {
    if (r == p.rank) continue;       // the forward already wrote the local slot
    float4* d4 = reinterpret_cast<float4*>(p.gbuf[r] + p.slot_off);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
      d4[i] = s4[i];
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x)
      p.gbuf[r][p.slot_off + i] = p.src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(p.done_ctr, 1u);
    if (done == gridDim.x - 1) {
      __threadfence_system();
      for (int r = 0; r < p.world; ++r)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.flags[r] + p.rank), "r"(p.epoch) : "memory");
      *p.done_ctr = 0u;
    }
  }
}

// one thread per source rank; epochs compare modulo 2^32
__global__ void peer_wait_kernel(const uint32_t* flags, int world, uint32_t epoch) {
  const int r = threadIdx.x;
  if (r < world) {
    const long long t0 = clock64();
    for (;;) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
      if ((int32_t)(v - epoch) >= 0) break;
      __nanosleep(200);
      if (clock64() - t0 > 40000000000LL) {   // ~20 s: a peer never published this epoch
        printf("nmrgnn_b200: peer gather timed out waiting for rank %d (flag %u, epoch %u)\n", r, v, epoch);
        __trap();
      }
    }
  }
}

}  // namespace nmr
