// MP layer on tensor cores, column-split CTA pairs (F = 256, E <= 3, K <= 16): the default MP kernel on sm_100a.
//
// Why pairs.  The one-CTA kernel (mp_layer_tc_kernel, kernels_tc.cuh) keeps main + correction accumulators for 256
// output columns: 512 columns = ALL of tensor memory, so the epilogue of a tile cannot overlap the MMAs of the next
// one (a third of the launch).  Here the two CTAs of a cluster work on the SAME 128-atom tile and split the OUTPUT
// columns: CTA r owns columns 128 r .. 128 r + 127.  Its accumulators (main 128 + corr 128 columns) fit tensor memory
// twice, so tile t is drained while tile t + 1 accumulates.  Everything else halves per SM as well:
//   * W' : each CTA streams only its 128 N rows of the hi / lo images (8 KB slots instead of 16 KB),
//   * gather : CTA r aggregates atom rows 64 r .. 64 r + 63 of every feature pass (half the gather work and half the
//     L1 footprint per SM) and SHIPS its half of the operand stage to the peer with bulk shared->shared::cluster
//     copies (UBLKCP.S.S) that complete on the peer's a_full barrier; both CTAs then hold the full 128-row A operand,
//   * epilogue : 128 rows x 128 columns per CTA.
// Every MMA is a plain cta_group::1 instruction (M = 128, N = 128) issued by each CTA on its own; the only cross-CTA
// traffic is the operand halves, the 64 output scales per tile, and "stage free" commits that are multicast to both CTAs.
//   warp 0: W' loader   warp 1: MMA issuer + TMEM owner   warp 2: edge-record loader + L2 prefetch
//   warp 3: operand shipper   warps 4-7: epilogue (thread = atom row)   warps 8-15: producers
// Row maxima of h_out (needed by the next layer's fp16 range scaling) are written per column half: hmax_out[2 i + r].
// Arithmetic, operand images, scaling and compensation are those of the one-CTA kernel: results are bit-identical.
#pragma once
#include "kernels_tc.cuh"

namespace nmr {

constexpr int MNP_THREADS = 512;
constexpr int MNP_AST = 3;             // operand stages: [3 chunks][hi 8192 | lo 8192] for 128 rows
constexpr int MNP_BRING = 4;           // W' ring: 8 KB slots = this CTA's 128 N rows of one image (hi or lo) of a (pass, n) chunk
constexpr int MNP_BSLOT = 8192;
// 64-byte-swizzled tiles need a 512-byte aligned base.  162 KB: the SM can run in its 164 KB configuration (92 KB of L1
// for the gathers).
constexpr size_t MNP_SMEM = 512 + MNP_AST * 3 * 16384 + MNP_BRING * MNP_BSLOT + 64 * MTC_KMAX * 16 + 2 * (64 + 128) * 4 + 256;
static_assert(MNP_SMEM + 1024 <= 196 * 1024, "column-split MP kernel no longer fits the 196 KB shared-memory configuration");

namespace tc {
// bulk async copy from this CTA's shared memory into the shared memory of a CTA of the cluster; the bytes are counted
// on an mbarrier of the DESTINATION CTA (both given as shared::cluster addresses, see map_to_cta)
__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
// arrives (once all MMAs previously issued by this thread have completed) on the barrier at this offset in every CTA of
// `cta_mask`; the MMAs themselves are cta_group::1
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void sts32f_cluster(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
}  // namespace tc

template <int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MNP_THREADS, 1) mp_layer_np_kernel(const MpTcArgs p) {
  constexpr uint32_t AST = MNP_AST;
  constexpr int BRING = MNP_BRING;
  constexpr uint32_t STAGE = 3 * 16384;
  extern __shared__ uint8_t smem_raw[];
  // (the dynamic shared window starts at the same offset in both CTAs, so the aligned layout is the same too)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 511) & ~uintptr_t(511));
  uint8_t* a_st = smem;                                          // [AST][3 chunks][hi 8192 | lo 8192]
  uint8_t* b_ring = a_st + AST * STAGE;                          // [BRING][8192]
  float4* rec_s = reinterpret_cast<float4*>(b_ring + BRING * MNP_BSLOT);   // [64 * K] this CTA's atom rows
  float* fscale = reinterpret_cast<float*>(rec_s + 64 * MTC_KMAX);         // [2][64]   2^-s of this CTA's rows
  float* oscale = fscale + 2 * 64;                                          // [2][128]  2^s * inv_degree, all rows of the tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(oscale + 2 * 128);
  uint64_t* a_half = bars;             // [AST] this CTA's rows of the stage are written (8 producer warps)
  uint64_t* a_full = a_half + AST;     // [AST] ... and the peer's rows have landed (shipper's arrive + bytes from the peer)
  uint64_t* a_empty = a_full + AST;    // [AST] both CTAs' MMAs have consumed the stage (2 multicast commits)
  uint64_t* b_full = a_empty + AST;    // [BRING]
  uint64_t* b_empty = b_full + BRING;  // [BRING]
  uint64_t* rec_full = b_empty + BRING;
  uint64_t* rec_empty = rec_full + 1;
  uint64_t* d_full = rec_empty + 1;    // [2] accumulator sets
  uint64_t* d_empty = d_full + 2;      // [2]
  uint64_t* sc_full = d_empty + 2;     // [2] output scales of all 128 rows present (8 local + 8 remote producer warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sc_full + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = tc::cluster_ctarank(), peer = rank ^ 1u;
  if (tid == 0) {
    for (int i = 0; i < (int)AST; ++i) {
      tc::mbar_init(&a_half[i], 8);
      tc::mbar_init(&a_full[i], 1);
      tc::mbar_init(&a_empty[i], 2);
    }
    for (int i = 0; i < BRING; ++i) {
      tc::mbar_init(&b_full[i], 1);
      tc::mbar_init(&b_empty[i], 1);
    }
    tc::mbar_init(rec_full, 1);
    tc::mbar_init(rec_empty, 8);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&d_full[i], 1);
      tc::mbar_init(&d_empty[i], 4);
      tc::mbar_init(&sc_full[i], 16);
    }
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  tc::cluster_sync();                  // both CTAs' barriers exist before any remote arrive / copy / multicast commit
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int K = p.K, E = p.E;
  const int64_t n_tiles = (p.n_atoms + 127) / 128;
  const int64_t tile_first = (int64_t)(blockIdx.x >> 1), tile_step = (int64_t)(gridDim.x >> 1);
  const int col0 = (int)rank * 128;    // this CTA's output columns
  const int row0 = (int)rank * 64;     // this CTA's rows of the operand tile

  if (warp == 0) {
    // ===================== W' loader: this CTA's 128 N rows of every image =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = tile_first; tile < n_tiles; tile += tile_step)
        for (int q = 0; q < MTC_PASSES * E * 2; ++q, ++it) {
          const uint32_t slot = it % BRING, ph = (it / BRING) & 1;
          tc::mbar_wait_relaxed(&b_empty[slot], ph ^ 1);
          tc::mbar_expect_tx(&b_full[slot], MNP_BSLOT);
          tc::bulk_g2s(b_ring + slot * MNP_BSLOT, p.Wimg + (size_t)q * 16384 + rank * 8192, MNP_BSLOT, &b_full[slot]);
        }
    }
  } else if (warp == 2) {
    // ===================== edge-record loader + L2 prefetcher (see mp_layer_tc_body) =====================
    uint32_t t = 0;
    for (int64_t tile = tile_first; tile < n_tiles; tile += tile_step, ++t) {
      if (lane == 0) {
        const int64_t a0 = tile * 128 + row0;
        const int rows = (int)max((int64_t)0, min((int64_t)64, p.n_atoms - a0));
        const uint32_t bytes = (uint32_t)rows * (uint32_t)K * 16u;
        tc::mbar_wait_relaxed(rec_empty, (t & 1) ^ 1);
        tc::mbar_expect_tx(rec_full, bytes);
        if (bytes) tc::bulk_g2s(rec_s, p.rec + a0 * K, bytes, rec_full);
      }
      __syncwarp();
      const int64_t nt = tile + tile_step;
      if (nt < n_tiles) {
        const int64_t b0 = nt * 128 + row0;
        const int nrows = (int)max((int64_t)0, min((int64_t)64, p.n_atoms - b0));
        const char* hb = reinterpret_cast<const char*>(p.h_in + b0 * 256);
        for (int i = lane; i < nrows * 8; i += 32) tc::prefetch_l2(hb + (size_t)i * 128);
        const char* rb = reinterpret_cast<const char*>(p.rec + b0 * K);
        for (int i = lane; i * 128 < nrows * K * 16; i += 32) tc::prefetch_l2(rb + (size_t)i * 128);
      }
    }
  } else if (warp == 3) {
    // ===================== operand shipper: this CTA's rows of every stage -> the peer =====================
    if (lane == 0) {
      const uint32_t ast_a = tc::smem_u32(a_st);
      const uint32_t ast_peer = tc::map_to_cta(ast_a, peer);
      const uint32_t a_full_peer = tc::map_to_cta(tc::smem_u32(a_full), peer);
      const uint32_t bytes = (uint32_t)(2 * E) * 4096u;
      uint32_t pass = 0;
      for (int64_t tile = tile_first; tile < n_tiles; tile += tile_step)
        for (int ps = 0; ps < MTC_PASSES; ++ps, ++pass) {
          const uint32_t st = pass % AST;
          const long long s0 = p.dbg ? clock64() : 0;
          tc::mbar_wait(&a_half[st], (pass / AST) & 1);
          if (p.dbg) p.dbg[(size_t)blockIdx.x * 8 + 7] += clock64() - s0;    // shipper idle: waiting for the producers
          tc::mbar_expect_tx(&a_full[st], bytes);        // my arrival + the bytes the peer is going to send me
          for (int img = 0; img < 2 * E; ++img) {          // rows 64 r .. 64 r + 63 = one 4 KB block of every 8 KB image
            const uint32_t off = st * STAGE + (uint32_t)img * 8192u + rank * 4096u;
            tc::bulk_s2c(ast_peer + off, ast_a + off, 4096u, a_full_peer + st * 8u);
          }
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_f16(128, 128);
      uint32_t it = 0, pass = 0, t = 0;
      long long w_d = 0, w_a = 0, w_b = 0, c0 = 0;
      const long long k0 = clock64();
      for (int64_t tile = tile_first; tile < n_tiles; tile += tile_step, ++t) {
        const uint32_t set = t & 1;
        const uint32_t d_main = tmem_base + set * 256u, d_corr = d_main + 128u;
        if (p.dbg) c0 = clock64();
        tc::mbar_wait(&d_empty[set], ((t >> 1) & 1) ^ 1);    // the epilogue has drained this set (two tiles ago)
        if (p.dbg) w_d += clock64() - c0;
        tc::tc_fence_after();
        for (int ps = 0; ps < MTC_PASSES; ++ps, ++pass) {
          const uint32_t st = pass % AST;
          if (p.dbg) c0 = clock64();
          tc::mbar_wait(&a_full[st], (pass / AST) & 1);
          if (p.dbg) w_a += clock64() - c0;
          tc::tc_fence_after();
          for (int n = 0; n < E; ++n) {
            const uint32_t a_base = tc::smem_u32(a_st + st * STAGE + n * 16384);
            const uint64_t ah = tc::make_desc_sw64(a_base), al = tc::make_desc_sw64(a_base + 8192);
            {
              const uint32_t slot = it % BRING;
              if (p.dbg) c0 = clock64();
              tc::mbar_wait(&b_full[slot], (it / BRING) & 1);
              if (p.dbg) w_b += clock64() - c0;
              tc::tc_fence_after();
              const uint64_t bh = tc::make_desc_sw64(tc::smem_u32(b_ring + slot * MNP_BSLOT));
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 2);
                const uint32_t acc = (ps | n | ks) != 0;
                tc::umma_f16(d_main, ah + adv, bh + adv, idesc, acc);
                tc::umma_f16(d_corr, al + adv, bh + adv, idesc, acc);
              }
              tc::umma_commit(&b_empty[slot]);
              ++it;
            }
            {
              const uint32_t slot = it % BRING;
              tc::mbar_wait(&b_full[slot], (it / BRING) & 1);
              tc::tc_fence_after();
              const uint64_t bl = tc::make_desc_sw64(tc::smem_u32(b_ring + slot * MNP_BSLOT));
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 2);
                tc::umma_f16(d_corr, ah + adv, bl + adv, idesc, 1);
              }
              tc::umma_commit(&b_empty[slot]);
              ++it;
            }
          }
          tc::umma_commit_mc(&a_empty[st], 3);               // the stage is mirrored: free in both CTAs only when both are done
        }
        tc::umma_commit(&d_full[set]);
      }
      if (p.dbg) {
        long long* o = p.dbg + (size_t)blockIdx.x * 8;
        o[0] = clock64() - k0;   // MMA thread: total
        o[1] = w_d;              //   waiting for the epilogue (d_empty)
        o[2] = w_a;              //   waiting for the operand stage (a_full: local rows + the peer's copy)
        o[3] = w_b;              //   waiting for W' (b_full, hi slots)
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== epilogue: 128 rows x this CTA's 128 columns (see mp_layer_tc_body) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int gi = lane & 7;                       // position inside the 8-lane group = float4 chunk after transposition
    const int grow = q * 32 + (lane & 24);         // first row of the group
    uint32_t t = 0;
    float osc = 0.0f;                              // 2^s * inv_degree * (1 + c) of this thread's row
    if (tile_first < n_tiles) {
      tc::mbar_wait_cluster(&sc_full[0], 0);
      osc = oscale[row] * p.corr;
    }
    for (int64_t tile = tile_first; tile < n_tiles; tile += tile_step, ++t) {
      const int64_t a0 = tile * 128;
      const int rows = (int)min((int64_t)128, p.n_atoms - a0);
      const uint32_t set = t & 1;
      const uint32_t t_main = tmem_base + ((uint32_t)(q * 32) << 16) + set * 256u;
      const uint32_t t_corr = t_main + 128u;
      // rows grow .. grow+7 of this group, clamped for the loads (stores are predicated)
      const float* hin = p.h_in + (a0 + min(grow, rows - 1)) * 256 + col0 + gi * 4;
      float* hout = p.h_out + (a0 + grow) * 256 + col0 + gi * 4;
      const int rlast = rows - 1 - min(grow, rows - 1);
      float4 res[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) res[k] = p.raw ? make_float4(0.f, 0.f, 0.f, 0.f) : tc::ldg128(hin + min(k, rlast) * 256);
      tc::mbar_wait(&d_full[set], (t >> 1) & 1);
      const long long e0 = p.dbg ? clock64() : 0;
      tc::tc_fence_after();
      // The next tile's scale is picked up BEFORE this set is handed back: the producers can overwrite the scale
      // buffer of tile t + 1 (for tile t + 3) only after the MMAs of tile t + 2 have started, which wait for this
      // set.  (It has been written long ago: the producers of both CTAs finished tile t before its MMAs completed.)
      float osc_next = 0.0f;
      if (tile + tile_step < n_tiles) {
        tc::mbar_wait_cluster(&sc_full[(t + 1) & 1], ((t + 1) >> 1) & 1);
        osc_next = oscale[((t + 1) & 1) * 128 + row] * p.corr;
      }
      float hm[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) hm[k] = 0.0f;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        float4 x[8];
        {
          float v[16];
          tc::tmem_ld16_combined(t_main + cc * 32, t_corr + cc * 32, v);
#pragma unroll
          for (int j = 0; j < 4; ++j) x[j] = make_float4(v[4 * j] * osc, v[4 * j + 1] * osc, v[4 * j + 2] * osc, v[4 * j + 3] * osc);
          tc::tmem_ld16_combined(t_main + cc * 32 + 16, t_corr + cc * 32 + 16, v);
#pragma unroll
          for (int j = 0; j < 4; ++j) x[4 + j] = make_float4(v[4 * j] * osc, v[4 * j + 1] * osc, v[4 * j + 2] * osc, v[4 * j + 3] * osc);
        }
        if (cc == 3) {     // the accumulators are in registers: hand the set back before the arithmetic of the last chunk
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&d_empty[set]);
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
        // now x[k] = columns col0 + cc*32 + gi*4 .. +3 of row grow + k
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
          if (cc + 1 < 4 && !p.raw) res[k] = tc::ldg128(hin + min(k, rlast) * 256 + (cc + 1) * 32);
        }
      }
      // row maxima over this CTA's columns: reduce over the 8 lanes of the group, lane k writes row grow + k
      float mine = 0.0f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float v = hm[k];
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
        if (gi == k) mine = v;
      }
      if (grow + gi < rows) p.hmax_out[(a0 + grow + gi) * 2 + rank] = mine;
      if (p.dbg && warp == 4 && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + (size_t)blockIdx.x * 8 + 4), (unsigned long long)(clock64() - e0));
      osc = osc_next;
    }
  } else if (warp >= 8) {
    // ===================== producers: gather-aggregate, scale, split (rows 64 r .. 64 r + 63 of the tile) =====================
    const int pw = warp - 8;                 // 0..7
    const int ptid = tid - 256;              // 0..255
    const int q8 = lane & 7;                 // 4-feature group inside the 32-feature pass
    const int rsub = lane >> 3;              // row inside the warp's 4-row group
    const uint32_t rec_a = tc::smem_u32(rec_s);
    const uint32_t ast_a = tc::smem_u32(a_st);
    const uint32_t os_peer = tc::map_to_cta(tc::smem_u32(oscale), peer);
    const uint32_t sc_full_peer = tc::map_to_cta(tc::smem_u32(sc_full), peer);
    const float* hq = p.h_in + q8 * 4;
    uint32_t pass = 0, t = 0;
    for (int64_t tile = tile_first; tile < n_tiles; tile += tile_step, ++t) {
      const int64_t a0 = tile * 128 + row0;                                             // first atom of this CTA's rows
      const int rows = (int)max((int64_t)0, min((int64_t)64, p.n_atoms - a0));         // valid local rows
      float* fs = fscale + (t & 1) * 64;
      const uint32_t fs_a = tc::smem_u32(fs);
      const long long r0c = p.dbg ? clock64() : 0;
      tc::mbar_wait(rec_full, t & 1);
      if (p.dbg && warp == 8 && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + (size_t)blockIdx.x * 8 + 6), (unsigned long long)(clock64() - r0c));

      // (see mp_layer_tc_body for the structure of the gather pipeline)
      const bool full = rows == 64 && K == 16;
      uint32_t nidx[8];
      const uint32_t sw = p.swz ? (uint32_t)(rsub << 1) : 0u;   // row & 3 == rsub for every row this thread touches
      auto load_idx = [&](int lr, int half, bool FULL) {
        const bool rv = lr < rows;
        const uint32_t ra = rec_a + (uint32_t)(lr * K) * 16u + 12u;
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
      auto consume = [&](const float4 (&hv)[8], int lr, int half, float (&acc)[3][4], bool FULL) {
        const bool rv = lr < rows;
        const uint32_t ra = rec_a + (uint32_t)(lr * K) * 16u;
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

      float4 hvA[8], hvB[8];
      load_idx(pw * 4 + rsub, 0, false);
      issue(hvA, 0);                            // first half-step of the tile, in flight during the scale pass
      load_idx(pw * 4 + rsub, 1, false);

      // per-row bound |T[i,.]| <= sum_j max_n|e_ijn| * hmax[nl_ij]  ->  power-of-two scale; the output scale of every
      // row goes to both CTAs (each epilogue drains all 128 rows)
      if (ptid < 128) {
        const int lr = ptid >> 1, hf = ptid & 1;
        float b = 0.0f;
        if (lr < rows) {
          for (int j = hf; j < K; j += 2) {
            const float4 r = tc::lds128(rec_a + (uint32_t)(lr * K + j) * 16u);
            const float em = fmaxf(fmaxf(fabsf(r.x), fabsf(r.y)), fabsf(r.z));
            if (em != 0.0f) {
              const int idx = __float_as_int(r.w);
              const float hmx = p.hmax_pair ? fmaxf(__ldg(p.hmax_in + 2 * (size_t)idx), __ldg(p.hmax_in + 2 * (size_t)idx + 1))
                                            : __ldg(p.hmax_in + idx);
              b = fmaf(em, hmx, b);
            }
          }
        }
        b += __shfl_xor_sync(0xffffffffu, b, 1);
        if (hf == 0) {
          // exponent of b (0 for b < 2^15): scale so that |T| * 2^-s < 2^15
          int s = ((__float_as_int(b) >> 23) & 0xff) - 127 - 14;
          s = b > 0.0f ? max(s, 0) : 0;
          s = min(s, 100);
          fs[lr] = tc::pow2f_exact(-s);
          const float o = lr < rows ? tc::pow2f_exact(s) * p.inv_degree[a0 + lr] : 0.0f;
          const uint32_t oo = (uint32_t)((t & 1) * 128 + row0 + lr) * 4u;
          oscale[(t & 1) * 128 + row0 + lr] = o;
          tc::sts32f_cluster(os_peer + oo, o);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (lane == 0) {
        tc::mbar_arrive(&sc_full[t & 1]);
        tc::mbar_arrive_cluster(sc_full_peer + (t & 1) * 8u);
      }

      for (int ps = 0; ps < MTC_PASSES; ++ps, ++pass) {
        const uint32_t st = pass % AST;
        const long long p0 = p.dbg ? clock64() : 0;
        tc::mbar_wait_cluster(&a_empty[st], ((pass / AST) & 1) ^ 1);
        if (p.dbg && warp == 8 && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + (size_t)blockIdx.x * 8 + 5), (unsigned long long)(clock64() - p0));
        const uint32_t ab = ast_a + st * STAGE;
#pragma unroll 1
        for (int step = 0; step < 2; ++step) {
          const int lr = step * 32 + pw * 4 + rsub;          // local row; operand row = row0 + lr
          float acc[3][4];
#pragma unroll
          for (int n = 0; n < 3; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[n][i] = 0.0f;
          const float sc = tc::lds32f(fs_a + (uint32_t)lr * 4u);   // needed only after the FMAs: latency hidden
          // next half-step: next row group of this pass, or the first one of the next pass
          const bool last = step == 1;
          const int nlr = last ? pw * 4 + rsub : lr + 32;
          const int nps = min(last ? ps + 1 : ps, MTC_PASSES - 1);   // (the tile's very last prefetch is unused)
          // entering: hvA holds (lr, half 0) in flight, nidx the indices of (lr, half 1)
          if (full) {
            issue(hvB, ps);
            load_idx(nlr, 0, true);
            consume(hvA, lr, 0, acc, true);
            issue(hvA, nps);
            load_idx(nlr, 1, true);
            consume(hvB, lr, 1, acc, true);
          } else {
            issue(hvB, ps);
            load_idx(nlr, 0, false);
            consume(hvA, lr, 0, acc, false);
            issue(hvA, nps);
            load_idx(nlr, 1, false);
            consume(hvB, lr, 1, acc, false);
          }
          const uint32_t orow = (uint32_t)(row0 + lr);
          const uint32_t off = orow * 64u + ((((uint32_t)q8 >> 1) ^ ((orow >> 1) & 3u)) << 4) + (((uint32_t)q8 & 1u) << 3);
#pragma unroll
          for (int n = 0; n < 3; ++n) {
            if (n < E) {
              uint2 hi, lo;
              tc::split2_f16(acc[n][0] * sc, acc[n][1] * sc, hi.x, lo.x);
              tc::split2_f16(acc[n][2] * sc, acc[n][3] * sc, hi.y, lo.y);
              tc::sts64(ab + n * 16384 + off, hi.x, hi.y);
              tc::sts64(ab + n * 16384 + 8192 + off, lo.x, lo.y);
            }
          }
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&a_half[st]);
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(rec_empty);
    }
  }
  tc::tc_fence_before();
  tc::cluster_sync();                  // the peer's copies into this CTA and this CTA's multicast commits have all landed
  if (warp == 1) tc::tmem_dealloc<512>(tmem_base);
}

}  // namespace nmr
