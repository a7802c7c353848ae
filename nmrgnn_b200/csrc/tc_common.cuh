// Blackwell (sm_100a) building blocks for the tensor-core path: mbarrier, bulk async
// copy (TMA 1-D), tcgen05 MMA / TMEM / commit wrappers, UMMA shared-memory descriptors
// for K-major 64-byte-swizzled operands, and the fp32 -> (tf32 hi, fp32 lo) split used
// by the 3xTF32 products  x*w ~= hi_x*hi_w + lo_x*hi_w + hi_x*lo_w  (fp32-level accuracy,
// accumulation in fp32 in tensor memory).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nmr {
namespace tc {

constexpr int BK = 16;                 // K elements (tf32) per operand chunk = one 64-byte swizzled row
constexpr int ROW_BYTES = BK * 4;      // 64
constexpr int UMMA_K = 8;              // tf32: 32 bytes per instruction
constexpr uint32_t LAYOUT_SW64 = 4;    // UMMA::LayoutType::SWIZZLE_64B

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// try_wait suspends the warp in hardware until the phase completes or an implementation-defined time
// limit expires.  HINT_NS > 0 passes an explicit suspend-time hint: used by the loader threads, which
// run far ahead of their consumers and would otherwise burn issue slots of the working warps.
template <uint32_t HINT_NS>
__device__ __forceinline__ bool mbar_try_wait_t(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  if (HINT_NS == 0) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(HINT_NS)
        : "memory");
  }
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) { return mbar_try_wait_t<0>(bar, parity); }
// Bounded wait: a protocol bug must trap (-> cudaErrorLaunchFailure) instead of hanging the GPU.
template <uint32_t HINT_NS>
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_t<HINT_NS>(bar, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait_t<HINT_NS>(bar, parity)) {
    __nanosleep(HINT_NS ? 256 : 40);       // back off: polling warps must not take issue slots from working warps
    if ((++spins & 255u) == 0u) {          // look at the clock only every 256 polls
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000LL) {  // ~4 s at 2 GHz
        printf("nmrgnn_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x,
               threadIdx.x, smem_u32(bar), parity);
        __trap();
      }
    }
  }
}
// (the latency-critical form polls with a clock read per iteration: measured faster on the edge kernel
//  than back-to-back polls, nanosleep back-off or a suspend hint)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) {  // ~4 s at 2 GHz
      printf("nmrgnn_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x,
             threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) { mbar_wait_t<4000>(bar, parity); }
// Hardware-suspended wait: try_wait with a long suspend-time hint parks the warp until the phase completes -- no polling
// traffic on the shared-memory pipe (many warps spinning on try_wait take bandwidth from tensor-core operand reads and
// bulk copies) and no nanosleep granularity on the wake-up.  Bounded like the others.
__device__ __forceinline__ void mbar_wait_susp(uint64_t* bar, uint32_t parity) {
  uint32_t n = 0;
  while (!mbar_try_wait_t<100000>(bar, parity)) {
    if (++n > 40000u) {                    // ~4 s of 100 us suspensions
      printf("nmrgnn_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x,
             smem_u32(bar), parity);
      __trap();
    }
  }
}


// One lane of a fully converged warp.  Issuing tcgen05.mma / commit under `if (elect_one())` inside warp-uniform
// control flow lets the compiler keep descriptors and addresses in uniform registers; issuing them from an
// `if (lane == 0)` region instead makes it wrap every UTCHMMA in an elect / R2UR.BROADCAST / branch "waterfall"
// loop (measured: 200 cycles per MMA instruction regardless of its shape, tools/microbench/mma_rate.cu).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 1-D bulk async copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// explicit shared-window accesses (32-bit shared addresses): pointers carved out of the aligned dynamic
// shared buffer go through an integer cast, after which the compiler falls back to generic LD/ST
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32f(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32f(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// read-only 16-byte global load
__device__ __forceinline__ float4 ldg128(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ------------------------------------------------------------------ tcgen05 / TMEM
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all MMAs previously issued by this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 16 consecutive fp32 columns: thread i of warp q gets TMEM lane 32q+i
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  __syncwarp();            // .sync.aligned needs a converged warp (see tmem_ld16_nowait)
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  __syncwarp();
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------ descriptors
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, both K-major.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Shared-memory matrix descriptor for a K-major operand tile of `rows` x 16 tf32 stored as
// rows of 64 bytes with the 64-byte swizzle (8-row groups are 512 bytes apart = SBO).
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(512u >> 4) << 32;                // stride byte offset: 8 rows x 64 B
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  d |= (uint64_t)LAYOUT_SW64 << 61;
  return d;
}
// byte offset of element (row, k) inside a K-major SW64 tile (tile base 512-B aligned)
__host__ __device__ __forceinline__ uint32_t sw64_offset(int row, int k) {
  return (uint32_t)row * 64u + ((((uint32_t)k >> 2) ^ (((uint32_t)row >> 1) & 3u)) << 4) + (((uint32_t)k & 3u) << 2);
}
// 16-byte chunk j (0..3) of `row`
__host__ __device__ __forceinline__ uint32_t sw64_chunk_offset(int row, int j) {
  return (uint32_t)row * 64u + ((((uint32_t)j) ^ (((uint32_t)row >> 1) & 3u)) << 4);
}

// ------------------------------------------------------------------ 3xTF32 split
// hi = x rounded to tf32 (low 13 mantissa bits zero, so any tf32 conversion in the tensor
// core reproduces it exactly); lo = x - hi is exact in fp32.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h & 0xFFFFE000u);
  lo = x - hi;
}
__device__ __forceinline__ void split4(const float4 x, float4& hi, float4& lo) {
  split_tf32(x.x, hi.x, lo.x);
  split_tf32(x.y, hi.y, lo.y);
  split_tf32(x.z, hi.z, lo.z);
  split_tf32(x.w, hi.w, lo.w);
}


// ==================================================================================
// FP16 scaled-split products ("fp16x3"): the production tensor-core scheme.
//   x = hi + lo * 2^-11,  hi = fp16(x),  lo = fp16((x - hi) * 2^11)      (22+ mantissa bits)
//   x*w ~= hi_x*hi_w  +  2^-11 * (lo_x*hi_w + hi_x*lo_w)                 (dropped term ~2^-22)
// Every fp16 x fp16 product is exact in fp32.  The main products accumulate in one TMEM
// accumulator, the correction products in a second one; they are combined with a
// round-to-nearest FFMA in the epilogue.  Why two accumulators: tcgen05 accumulation
// truncates (measured on B200: -0.31 ulp per MMA instruction), so the
// chain into the accumulator that carries the result's magnitude must be as short as
// possible: K/16 instructions instead of the 3*K/8 of a single-accumulator 3xTF32 chain.
// Range: operands must stay below 65504 -> activations are pre-scaled by a power of two
// derived from an a-priori bound (exact to undo); the 2^11 scaling of `lo` keeps fp32-level
// absolute accuracy down to 2^-35.
// ==================================================================================
constexpr int HK = 32;                 // fp16 K elements per 64-byte swizzled row
constexpr int UMMA_K_F16 = 16;         // 32 bytes per instruction
constexpr float LO_SCALE = 2048.0f;
constexpr float LO_UNSCALE = 1.0f / 2048.0f;

// Instruction descriptor: D=F32, A=B=F16, both K-major.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- CTA pairs (cta_group::2): two CTAs of a cluster on one TPC execute ONE M = 256 instruction; each provides its
// 128 rows of A and half of the N rows of B from its own shared memory, accumulators land in both tensor memories.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same location in CTA `rank`
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {   // release at cluster scope
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a local barrier whose arrivals come from the peer CTA (acquire at cluster scope); bounded like mbar_wait
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 8000000000LL) {
      printf("nmrgnn_b200: cluster mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x,
             threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result) {  // one whole warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
// issued by ONE thread of the leader CTA (rank 0); descriptors address the same offsets in both CTAs
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this offset in every CTA of `cta_mask` once all prior MMAs of the pair have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// TS form: the A operand comes from tensor memory (lane = row, every 32-bit column holds two consecutive k
// as packed fp16: a K = 16 instruction consumes 8 columns), B from shared memory.  The four mask words
// (disable-output-lane) are zero.
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// registers -> tensor memory: this thread's lane (row), 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// two floats -> packed (hi, hi) and (lo, lo) half2 words
__device__ __forceinline__ void split2_f16(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn((x0 - hf.x) * LO_SCALE, (x1 - hf.y) * LO_SCALE);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// the same without the 2^11 scaling of lo (single-accumulator form of the MP layer)
__device__ __forceinline__ void split2_f16_plain(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// 8 consecutive k -> one 16-byte piece of the hi tile and one of the lo tile
__device__ __forceinline__ void split8_f16(const float (&x)[8], uint4& hi, uint4& lo) {
  split2_f16(x[0], x[1], hi.x, lo.x);
  split2_f16(x[2], x[3], hi.y, lo.y);
  split2_f16(x[4], x[5], hi.z, lo.z);
  split2_f16(x[6], x[7], hi.w, lo.w);
}

// 32 lanes x 16 columns without the wait (issue several, then tmem_ld_wait once).
// tcgen05.ld / st / wait are .sync.aligned: every lane of the warp must execute them TOGETHER.  A warp that has just left
// an mbarrier spin loop (or any data-dependent branch) is not guaranteed to have reconverged -- observed on B200 in a
// round-2 experiment: lanes 16..31 of a warp read garbage from tensor memory -- hence the explicit __syncwarp().
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  __syncwarp();
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// main + corr * 2^-11 for 16 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16_combined(uint32_t t_main, uint32_t t_corr, float (&v)[16]) {
  uint32_t a[16], b[16];
  tmem_ld16_nowait(t_main, a);
  tmem_ld16_nowait(t_corr, b);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = fmaf(__uint_as_float(b[i]), LO_UNSCALE, __uint_as_float(a[i]));
}

// exact power of two 2^e as float (e in [-126, 127])
__host__ __device__ __forceinline__ float pow2f_exact(int e) {
#ifdef __CUDA_ARCH__
  return __int_as_float((e + 127) << 23);
#else
  union { uint32_t u; float f; } c;
  c.u = (uint32_t)(e + 127) << 23;
  return c.f;
#endif
}

}  // namespace tc
}  // namespace nmr
