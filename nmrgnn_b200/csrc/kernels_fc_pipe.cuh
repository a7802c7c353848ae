// Node MLP + readout on tensor cores, layer-pipelined form (FCBlock + out_layer + peak standardisation;
// nmrgnn/model.py:191-196, 268-273).  Same contract as fc_readout_tc_kernel (kernels_tc.cuh); what changes is the
// schedule.  There, MMA phases and epilogue phases strictly alternate on the CTA's one resident 128-atom tile (its
// operand X is 128 KB: a second tile does not fit) and the tile load, four MMA phases, four epilogues and the readout
// add up to ~94 k cycles per tile against a tensor floor of ~21 k.  Here
//   * main and correction products of a layer go into ONE 256-column accumulator (lo images not scaled by 2^11, the
//     weights pre-scaled by a power of two so that their lo image stays in the normal fp16 range), so tensor memory
//     holds TWO accumulator sets and layer l + 1 accumulates while layer l is drained;
//   * the epilogue of layer l produces the operand of layer l + 1 K-chunk by K-chunk (32 columns, in place) and hands
//     every chunk to the MMA thread as soon as all 128 rows are written: the MMAs of layer l + 1 run UNDER the
//     epilogue of layer l, one chunk behind;
//   * four stager warps load, scale, split and store the NEXT tile's operand as soon as the last layer's MMAs have
//     released X, so layer 0 of tile t + 1 runs under the last epilogue + readout of tile t;
//   * the readout is formed from the last epilogue's registers (partial dot products per column quarter meet in
//     shared memory): no fp32 copy of Z in shared memory, no extra CTA-wide barriers.
// The epilogue warps are the only role that is never idle; a tile costs ~4 epilogues.
// Row scaling: the input of layer l + 1 is bounded from the ACTUAL row maximum of layer l's input (carried through the
// epilogues) and one layer's growth bound, so the scaled operand sits within 2^-7 of the fp16 range and the unscaled lo
// part keeps fp32-level absolute accuracy.
//   warp 0: W loader   warp 1: MMA issuer + TMEM owner   warps 2-17: epilogue / readout   warps 18-21: stagers
#pragma once
#include "kernels_tc.cuh"

namespace nmr {

struct FcPipeArgs {
  const float* nodes;        // [n_atoms, 256]
  const float* hmax;         // [n_atoms] max |nodes row|  ([n_atoms][2] partial maxima if hmax_pair)
  int hmax_pair;
  const float* atoms;        // [n_atoms, C]
  float* peaks;              // [n_atoms]
  float* fc_nodes;           // optional [n_atoms, 128]
  int64_t n_atoms;
  int C;                     // <= FPI_CMAX
  const uint8_t* Wimg;       // residual layers: [8 chunks][hi 16384 | lo 16384]; last: [8 chunks][hi 8192 | lo 8192]; layer stride 262144
  const float* bias;         // [n_layers][256]
  float g[MAX_DENSE];        // max |x_{l+1} row| <= g[l] * max |x_l row| + o[l]
  float o[MAX_DENSE];
  float wsinv[MAX_DENSE];    // 2^-s of the layer's weight image
  int n_layers;
  int act;
  float corr;                // residual constant of the round-toward-zero compensation (1 by default)
  const float* Wo;           // [128, C]
  const float* bo;           // [C]
  const float* peak_std;     // [C]
  const float* peak_avg;     // [C]
  long long* dbg;            // optional [grid][8] role cycle counters
};

constexpr int FPI_THREADS = 704;
// W ring.  One CTA: 2 slots of 32 KB = the hi and lo tiles (256 N rows each) of one 32-feature chunk, ONE bulk copy each
// (a bulk copy costs its latency whatever its size, and under load that latency is thousands of cycles).  CTA pair:
// every CTA stages only its 128 N rows of both tiles, 16 KB per chunk, so the same bytes hold 5 chunks.
constexpr int FPI_RING = 2;
constexpr int FPI_SLOT = 32768;
constexpr int FPI_RING_PAIR = 5;
constexpr int FPI_SLOT_PAIR = 16384;
constexpr int FPI_CMAX = 16;
constexpr size_t FPI_X_BYTES = 8 * 16384;
// (Wo and the per-class constants of the readout are read through L1 from global memory: that leaves room for a fifth
//  W slot)
constexpr size_t FPI_SMEM = 1024 + FPI_X_BYTES + FPI_RING * FPI_SLOT + MAX_DENSE * 256 * 4 + 2 * 4 * 128 * 4 +
                            2 * 4 * 128 * 4 + 512;
constexpr size_t FPI_SMEM_PAIR = FPI_SMEM + (FPI_RING_PAIR * FPI_SLOT_PAIR - FPI_RING * FPI_SLOT);
static_assert(FPI_SMEM_PAIR <= 227 * 1024 && FPI_SMEM <= 227 * 1024, "pipelined node-MLP kernel exceeds the 227 KB shared-memory limit");

// exponent s with bound * 2^-s in [2^14, 2^15): fp16 operands stay finite with 2x margin, as large as possible
__device__ __forceinline__ int fpi_scale_exp(float bound) {
  const int s = ((__float_as_int(bound) >> 23) & 0xff) - 127 - 14;
  return (bound > 0.0f && bound < 3.0e38f) ? min(max(s, -60), 100) : 0;
}

// 32 lanes x 8 consecutive fp32 columns, load and wait in ONE statement: the destination registers of an asynchronous
// tcgen05.ld must not be touched before the wait, and the compiler does not know that -- with the wait as a separate
// statement and the registers live across a loop back edge it inserted moves between the two (observed: garbage rows)
__device__ __forceinline__ void fpi_tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  __syncwarp();
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void fpi_split8_plain(const float (&x)[8], uint4& hi, uint4& lo) {
  tc::split2_f16_plain(x[0], x[1], hi.x, lo.x);
  tc::split2_f16_plain(x[2], x[3], hi.y, lo.y);
  tc::split2_f16_plain(x[4], x[5], hi.z, lo.z);
  tc::split2_f16_plain(x[6], x[7], hi.w, lo.w);
}

template <int ACT, bool PAIR>
__device__ __forceinline__ void fc_readout_pipe_body(const FcPipeArgs& p) {
  constexpr int SLOT = PAIR ? FPI_SLOT_PAIR : FPI_SLOT;
  constexpr int RING = PAIR ? FPI_RING_PAIR : FPI_RING;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xs = smem;                                       // [8 chunks][hi 8192 | lo 8192]
  uint8_t* ring = xs + FPI_X_BYTES;                         // [RING][16384]
  float* bias_s = reinterpret_cast<float*>(ring + RING * SLOT);         // [n_layers][256]
  float* pm = bias_s + MAX_DENSE * 256;                     // [2][4][128] partial row maxima of a layer's output
  float* part = pm + 2 * 4 * 128;                           // [2][4][128] partial readout dot products
  uint64_t* bars = reinterpret_cast<uint64_t*>(part + 2 * 4 * 128);
  uint64_t* w_full = bars;                                  // [RING]
  uint64_t* w_empty = w_full + RING;                        // [RING]
  uint64_t* w_peer = w_empty + RING;                        // [RING] (pair, leader: the peer's half of the slot has landed)
  uint64_t* xs_full = w_peer + RING;                        // [8]  chunk staged (4 stager warps; pair: of both CTAs, on the leader)
  uint64_t* xe_full = xs_full + 8;                          // [8]  chunk rewritten by an epilogue (16 warps)
  uint64_t* d_full = xe_full + 8;                           // [2]  accumulator set complete
  uint64_t* x_free = d_full + 2;                            //      last layer's MMAs have read X
  uint64_t* d_free = x_free + 1;                            //      last layer's accumulators drained (odd layer counts)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_free + 1);
  uint32_t* t_issue = tmem_slot + 2;                        // [RING] diagnostics: clock at which a slot's copy was issued

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nl = p.n_layers, C = p.C;
  const uint32_t rank = PAIR ? tc::cluster_ctarank() : 0u;
  if (tid == 0) {
    for (int i = 0; i < RING; ++i) {
      tc::mbar_init(&w_full[i], 1);
      tc::mbar_init(&w_empty[i], 1);
      tc::mbar_init(&w_peer[i], 1);
    }
    for (int i = 0; i < 8; ++i) {          // pair: the warps of both CTAs arrive on the leader's barriers
      tc::mbar_init(&xs_full[i], PAIR ? 8 : 4);
      tc::mbar_init(&xe_full[i], PAIR ? 32 : 16);
    }
    tc::mbar_init(&d_full[0], 1);
    tc::mbar_init(&d_full[1], 1);
    tc::mbar_init(x_free, 1);
    tc::mbar_init(d_free, PAIR ? 32 : 16);
    tc::mbar_fence_init();
  }
  for (int i = tid; i < nl * 256; i += FPI_THREADS) bias_s[i] = p.bias[i];
  if (warp == 1) {
    if (PAIR) tc::tmem_alloc_pair<512>(tmem_slot);
    else tc::tmem_alloc<512>(tmem_slot);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (PAIR) tc::cluster_sync();            // both CTAs' barriers exist before any remote arrive / multicast commit
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t n_tiles = (p.n_atoms + 127) / 128;
  // tile walk: a single CTA takes tiles blockIdx.x, + gridDim.x, ...; a pair takes tile pairs and CTA r the r-th tile of
  // each pair (a trailing odd tile leaves the peer with an empty tile: it still runs the whole protocol)
  const int64_t tile_first = PAIR ? (int64_t)(blockIdx.x >> 1) * 2 + rank : (int64_t)blockIdx.x;
  const int64_t tile_step = (int64_t)gridDim.x;
  const int64_t tile_end = PAIR ? ((n_tiles + 1) / 2) * 2 : n_tiles;
  // the leader's hand-off barriers as seen from either CTA of the pair
  const uint32_t xs_full_a = PAIR ? tc::map_to_cta(tc::smem_u32(xs_full), 0) : tc::smem_u32(xs_full);
  const uint32_t xe_full_a = PAIR ? tc::map_to_cta(tc::smem_u32(xe_full), 0) : tc::smem_u32(xe_full);
  const uint32_t d_free_a = PAIR ? tc::map_to_cta(tc::smem_u32(d_free), 0) : tc::smem_u32(d_free);
  auto arrive = [&](uint32_t addr) {       // one arrival on a hand-off barrier (pair: release at cluster scope)
    if (PAIR) tc::mbar_arrive_cluster(addr);
    else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
  };
  // accumulator set of layer l: the last layer always takes set 1, so that with an even layer count layer 0 of the
  // next tile (set 0) never meets the set the last epilogue is still reading
  const int set0 = (nl & 1) ? 1 : 0;                        // set of layer 0; layer l: (set0 + l) & 1

  if (warp == 0) {
    // ===================== W loader =====================
    // one CTA: chunk = [hi 16384 | lo 16384] (last layer: two chunks of [hi 8192 | lo 8192]) per 32 KB copy;
    // pair: this CTA's 128 (64) N rows: chunk = [hi 8192 | lo 8192] (last layer: two chunks of [hi 4096 | lo 4096]) per 16 KB copy
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = tile_first; tile < tile_end; tile += tile_step)
        for (int l = 0; l < nl; ++l) {
          const int n_ops = (l == nl - 1) ? 4 : 8;
          const uint8_t* src = p.Wimg + (size_t)l * 262144 + (PAIR ? (size_t)rank * (l == nl - 1 ? 65536 : 131072) : 0);
          for (int q = 0; q < n_ops; ++q, ++it) {
            const uint32_t slot = it % RING, ph = (it / RING) & 1;
            tc::mbar_wait_susp(&w_empty[slot], ph ^ 1);
            tc::mbar_expect_tx(&w_full[slot], SLOT);
            if (p.dbg) *reinterpret_cast<volatile uint32_t*>(&t_issue[slot]) = (uint32_t)clock64();
            tc::bulk_g2s(ring + slot * SLOT, src + (size_t)q * SLOT, SLOT, &w_full[slot]);
          }
        }
    }
  } else if (warp == 1 && PAIR && rank != 0) {
    // ===================== peer CTA: relay "my half of the slot has landed" to the leader =====================
    if (lane == 0) {
      const uint32_t w_peer_ldr = tc::map_to_cta(tc::smem_u32(w_peer), 0);
      uint32_t it = 0;
      for (int64_t tile = tile_first; tile < tile_end; tile += tile_step)
        for (int l = 0; l < nl; ++l) {
          const int n_ops = (l == nl - 1) ? 4 : 8;
          for (int q = 0; q < n_ops; ++q, ++it) {
            const uint32_t slot = it % RING;
            tc::mbar_wait_susp(&w_full[slot], (it / RING) & 1);
            tc::mbar_arrive_cluster(w_peer_ldr + slot * 8u);
          }
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair: the leader CTA only) =====================
    if (lane == 0) {
      const uint32_t idesc256 = tc::make_idesc_f16(PAIR ? 256 : 128, 256), idesc128 = tc::make_idesc_f16(PAIR ? 256 : 128, 128);
      auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
        if (PAIR) tc::umma_f16_pair(d, a, b, idesc, acc);
        else tc::umma_f16(d, a, b, idesc, acc);
      };
      auto commit = [&](uint64_t* bar) {       // pair: the arrival goes to the same barrier of both CTAs
        if (PAIR) tc::umma_commit_pair(bar, 3);
        else tc::umma_commit(bar);
      };
      auto wait = [&](uint64_t* bar, uint32_t parity) {   // barriers with arrivals from the peer CTA
        if (PAIR) tc::mbar_wait_cluster(bar, parity);
        else tc::mbar_wait(bar, parity);
      };
      uint32_t it = 0, t = 0, ne = 0;
      long long w_x = 0, w_w = 0, c0 = 0, lat_sum = 0, lat_n = 0;
      const long long k0 = clock64();
      for (int64_t tile = tile_first; tile < tile_end; tile += tile_step, ++t) {
        if ((nl & 1) && t > 0) {          // odd layer count: layer 0 shares its set with the previous tile's last layer
          wait(d_free, (t - 1) & 1);
          tc::tc_fence_after();
        }
        for (int l = 0; l < nl; ++l) {
          const bool last = l == nl - 1;
          const uint32_t d = tmem_base + (uint32_t)((set0 + l) & 1) * 256u;
          for (int c = 0; c < 8; ++c) {
            if (p.dbg) c0 = clock64();
            if (l == 0) wait(&xs_full[c], t & 1);
            else wait(&xe_full[c], ne & 1);
            if (p.dbg) w_x += clock64() - c0;
            tc::tc_fence_after();
            const uint64_t ah = tc::make_desc_sw64(tc::smem_u32(xs + c * 16384));
            const uint64_t al = tc::make_desc_sw64(tc::smem_u32(xs + c * 16384 + 8192));
            const uint32_t slot = it % RING;
            if (!last || (c & 1) == 0) {
              if (p.dbg) c0 = clock64();
              tc::mbar_wait(&w_full[slot], (it / RING) & 1);
              if (PAIR) tc::mbar_wait_cluster(&w_peer[slot], (it / RING) & 1);
              if (p.dbg) {
                const long long now = clock64();
                w_w += now - c0;
                if (now - c0 > 100) {
                  lat_sum += (uint32_t)((uint32_t)now - *reinterpret_cast<volatile uint32_t*>(&t_issue[slot]));
                  ++lat_n;
                }
              }
              tc::tc_fence_after();
            }
            // instruction order of a chunk: hi*hi, lo*hi (ks 0), hi*hi, lo*hi (ks 1), hi*lo (ks 0, 1)
            if (!last) {
              const uint64_t bh = tc::make_desc_sw64(tc::smem_u32(ring + slot * SLOT));
              const uint64_t bl = tc::make_desc_sw64(tc::smem_u32(ring + slot * SLOT + SLOT / 2));
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 2);
                mma(d, ah + adv, bh + adv, idesc256, (c | ks) != 0);
                mma(d, al + adv, bh + adv, idesc256, 1);
              }
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 2);
                mma(d, ah + adv, bl + adv, idesc256, 1);
              }
              commit(&w_empty[slot]);
              it += 1;
            } else {
              const uint8_t* wb = ring + slot * SLOT + (c & 1) * (SLOT / 2);
              const uint64_t bh = tc::make_desc_sw64(tc::smem_u32(wb));
              const uint64_t bl = tc::make_desc_sw64(tc::smem_u32(wb + SLOT / 4));
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 2);
                mma(d, ah + adv, bh + adv, idesc128, (c | ks) != 0);
                mma(d, al + adv, bh + adv, idesc128, 1);
              }
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 2);
                mma(d, ah + adv, bl + adv, idesc128, 1);
              }
              if (c & 1) {
                commit(&w_empty[slot]);
                it += 1;
              }
            }
          }
          commit(&d_full[(set0 + l) & 1]);
          if (last) commit(x_free);
          if (l > 0) ++ne;
        }
      }
      if (p.dbg) {
        long long* o = p.dbg + (size_t)blockIdx.x * 8;
        o[0] = clock64() - k0;   // MMA thread: total
        o[1] = w_x;              //   waiting for operand chunks (stagers / epilogue)
        o[2] = w_w;              //   waiting for W
        o[5] = lat_sum;          //   issue -> arrival of the W hi slots it had to wait for (sum, count)
        o[6] = lat_n;
      }
    }
  } else if (warp < 18) {
    // ===================== epilogue / readout warps (16) =====================
    const int q = warp & 3;                  // TMEM lane quarter this warp may read
    const int j = (warp - 2) >> 2;           // 8-column piece of a 32-column chunk (residual layers); 32-column quarter (last)
    const int row = q * 32 + lane;
    const uint32_t xs_a = tc::smem_u32(xs);
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t dphase = 0, t = 0;
    long long e_busy = 0;
    for (int64_t tile = tile_first; tile < tile_end; tile += tile_step, ++t) {
      const int64_t a0 = tile * 128;
      const int rows = (int)max((int64_t)0, min((int64_t)128, p.n_atoms - a0));   // 0: the pair's trailing empty tile
      // this row's input maximum and element classes (loads in flight while layer 0 accumulates)
      float m_in = 0.0f;
      uint32_t cmask = 0;
      if (row < rows) {
        m_in = p.hmax_pair ? fmaxf(__ldg(p.hmax + 2 * (a0 + row)), __ldg(p.hmax + 2 * (a0 + row) + 1)) : __ldg(p.hmax + a0 + row);
        const float* ar = p.atoms + (a0 + row) * C;
        for (int c = 0; c < C; ++c) cmask |= (__ldg(ar + c) != 0.0f) ? (1u << c) : 0u;
      }
      int e_in = fpi_scale_exp(m_in);        // the stagers scale the tile's rows by 2^-e_in (same function, same input)
      // ---- residual layers: step s rewrites K-chunk s of the operand; thread = (row, columns 32 s + 8 j .. + 7)
      for (int l = 0; l + 1 < nl; ++l) {
        const int set = (set0 + l) & 1;
        const uint32_t t_set = t_lane + (uint32_t)set * 256u + (uint32_t)(8 * j);
        const uint32_t bl_a = tc::smem_u32(bias_s + l * 256 + 8 * j);
        float mx = 0.0f;
        tc::mbar_wait_susp(&d_full[set], (dphase >> set) & 1);
        dphase ^= 1u << set;
        const long long e0 = p.dbg ? clock64() : 0;
        tc::tc_fence_after();
        if (l > 0) {       // (behind the wait: every warp has published its partial maxima of the previous epilogue)
          const float* pmr = pm + (l & 1) * 512 + row;
          m_in = fmaxf(fmaxf(pmr[0], pmr[128]), fmaxf(pmr[256], pmr[384]));
        }
        const int e_nx = fpi_scale_exp(fmaf(p.g[l], m_in, p.o[l]));
        const float s_old = tc::pow2f_exact(e_in);
        const float s_nx = tc::pow2f_exact(-e_nx);
        const float s_out = p.corr * s_old * p.wsinv[l];
        auto step = [&](int s) {
          uint32_t cur[8];
          fpi_tmem_ld8(t_set + (uint32_t)(32 * s), cur);
          const uint32_t off = xs_a + (uint32_t)s * 16384u + tc::sw64_chunk_offset(row, j);
          uint4 ohi, olo;
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(ohi.x), "=r"(ohi.y), "=r"(ohi.z), "=r"(ohi.w) : "r"(off));
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(olo.x), "=r"(olo.y), "=r"(olo.z), "=r"(olo.w) : "r"(off + 8192u));
          const float4 b0 = tc::lds128(bl_a + (uint32_t)s * 128u), b1 = tc::lds128(bl_a + (uint32_t)s * 128u + 16u);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          const uint32_t oh[4] = {ohi.x, ohi.y, ohi.z, ohi.w}, ol[4] = {olo.x, olo.y, olo.z, olo.w};
          float x[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&oh[i]));
            const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&ol[i]));
            const float y0 = act_t<ACT>(fmaf(__uint_as_float(cur[2 * i]), s_out, bb[2 * i])) + (fh.x + fl.x) * s_old;
            const float y1 = act_t<ACT>(fmaf(__uint_as_float(cur[2 * i + 1]), s_out, bb[2 * i + 1])) + (fh.y + fl.y) * s_old;
            mx = fmaxf(mx, fmaxf(fabsf(y0), fabsf(y1)));
            x[2 * i] = y0 * s_nx;
            x[2 * i + 1] = y1 * s_nx;
          }
          uint4 hi, lo;
          fpi_split8_plain(x, hi, lo);
          tc::sts128(off, hi);
          tc::sts128(off + 8192u, lo);
          // (the row maximum must be published in front of the step's release, not after it: the readers of the next
          //  layer are ordered behind this warp only through the barrier chain xe_full -> MMA -> d_full)
          if (s == 7) pm[((l + 1) & 1) * 512 + j * 128 + row] = mx;
          tc::fence_proxy_async();
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive(xe_full_a + (uint32_t)s * 8u);
        };
#pragma unroll 1
        for (int s = 0; s < 8; ++s) step(s);
        e_in = e_nx;
        if (p.dbg) e_busy += clock64() - e0;
      }
      // ---- last layer: z = act(D + b), this thread's 32 of the 128 columns; readout from the registers
      {
        const int l = nl - 1;
        const int set = 1;
        const float s_out = p.corr * tc::pow2f_exact(e_in) * p.wsinv[l];
        const uint32_t t_set = t_lane + (uint32_t)set * 256u + (uint32_t)(32 * j);
        const uint32_t bl_a = tc::smem_u32(bias_s + l * 256 + 32 * j);
        tc::mbar_wait_susp(&d_full[set], (dphase >> set) & 1);
        dphase ^= 1u << set;
        const long long e0 = p.dbg ? clock64() : 0;
        tc::tc_fence_after();
        float z[32];
        {
          uint32_t r0[16], r1[16];
          tc::tmem_ld16_nowait(t_set, r0);
          tc::tmem_ld16_nowait(t_set + 16u, r1);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            z[i] = __uint_as_float(r0[i]);
            z[16 + i] = __uint_as_float(r1[i]);
          }
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive(d_free_a);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = tc::lds128(bl_a + (uint32_t)i * 4u);
          z[i + 0] = act_t<ACT>(fmaf(z[i + 0], s_out, b4.x));
          z[i + 1] = act_t<ACT>(fmaf(z[i + 1], s_out, b4.y));
          z[i + 2] = act_t<ACT>(fmaf(z[i + 2], s_out, b4.z));
          z[i + 3] = act_t<ACT>(fmaf(z[i + 3], s_out, b4.w));
        }
        if (p.fc_nodes != nullptr && row < rows) {
          float4* dst = reinterpret_cast<float4*>(p.fc_nodes + (a0 + row) * 128 + 32 * j);
#pragma unroll
          for (int i = 0; i < 8; ++i) dst[i] = make_float4(z[4 * i], z[4 * i + 1], z[4 * i + 2], z[4 * i + 3]);
        }
        // peaks = sum_c ((z . Wo[:,c] + bo[c]) * a[c] * std[c] + a[c] * avg[c])   (model.py:268-273): this thread's quarter
        // of the dot products of the row's non-zero classes.  One-hot rows (the reference's contract) keep the plain dot
        // product so that the final expression is evaluated in the reference's order.
        const bool onehot = __popc(cmask) == 1;
        float partial = 0.0f;
        uint32_t mk = cmask;
        const float* ar = p.atoms + (a0 + row) * C;
        while (mk) {
          const int c = __ffs(mk) - 1;
          mk &= mk - 1;
          const float* w = p.Wo + (32 * j) * C + c;
          float dot = 0.0f;
#pragma unroll
          for (int k = 0; k < 32; ++k) dot = fmaf(z[k], __ldg(w + k * C), dot);
          partial = onehot ? dot : fmaf(dot, __ldg(ar + c) * __ldg(p.peak_std + c), partial);
        }
        float* pt = part + (t & 1) * 512;
        pt[j * 128 + row] = partial;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (j == 0 && row < rows) {
          const float dsum = (pt[row] + pt[128 + row]) + (pt[256 + row] + pt[384 + row]);
          float peak;
          if (onehot) {
            const int c = __ffs(cmask) - 1;
            const float a = __ldg(ar + c);
            peak = (dsum + __ldg(p.bo + c)) * a * __ldg(p.peak_std + c) + a * __ldg(p.peak_avg + c);
          } else {
            peak = dsum;
            mk = cmask;
            while (mk) {
              const int c = __ffs(mk) - 1;
              mk &= mk - 1;
              const float a = __ldg(ar + c);
              peak += __ldg(p.bo + c) * a * __ldg(p.peak_std + c) + a * __ldg(p.peak_avg + c);
            }
          }
          p.peaks[a0 + row] = peak;
        }
        if (p.dbg) e_busy += clock64() - e0;
      }
    }
    if (p.dbg && warp == 2 && lane == 0) p.dbg[(size_t)blockIdx.x * 8 + 3] = e_busy;   // epilogue warp: busy after its waits
  } else {
    // ===================== stagers (4 warps): next tile's operand =====================
    // lane = (row r8 of an 8-row group, 8-feature piece of the 32-feature chunk): full 32-byte sectors per lane,
    // whole 128-byte lines per 4 lanes; chunk c + 1 is in flight while chunk c is scaled, split and stored
    const int sw = warp - 18;
    const int r8 = lane >> 2, piece = lane & 3;
    const uint32_t xs_a = tc::smem_u32(xs);
    uint32_t t = 0;
    long long s_wait = 0;
    for (int64_t tile = tile_first; tile < tile_end; tile += tile_step, ++t) {
      const int64_t a0 = tile * 128;
      const int rows = (int)max((int64_t)0, min((int64_t)128, p.n_atoms - a0));
      float sc[4];
      const int r0 = 32 * sw + r8;            // this lane's rows: r0 + 8 g
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int r = r0 + 8 * g;
        float m = 0.0f;
        if (r < rows)
          m = p.hmax_pair ? fmaxf(__ldg(p.hmax + 2 * (a0 + r)), __ldg(p.hmax + 2 * (a0 + r) + 1)) : __ldg(p.hmax + a0 + r);
        sc[g] = tc::pow2f_exact(-fpi_scale_exp(m));
      }
      const float* src = p.nodes + (a0 + r0) * 256 + piece * 8;
      float4 buf[2][8];
      auto load = [&](int c, float4 (&b)[8]) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (r0 + 8 * g < rows) {
            b[2 * g] = tc::ldg128(src + g * 2048 + c * 32);
            b[2 * g + 1] = tc::ldg128(src + g * 2048 + c * 32 + 4);
          } else {
            b[2 * g] = make_float4(0.f, 0.f, 0.f, 0.f);
            b[2 * g + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      };
      load(0, buf[0]);
      if (t > 0) {                         // the previous tile's last layer has read X
        const long long c0 = p.dbg ? clock64() : 0;
        tc::mbar_wait_susp(x_free, (t - 1) & 1);
        if (p.dbg) s_wait += clock64() - c0;
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c + 1 < 8) load(c + 1, buf[(c + 1) & 1]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 x0 = buf[c & 1][2 * g], x1 = buf[c & 1][2 * g + 1];
          const float x[8] = {x0.x * sc[g], x0.y * sc[g], x0.z * sc[g], x0.w * sc[g],
                              x1.x * sc[g], x1.y * sc[g], x1.z * sc[g], x1.w * sc[g]};
          uint4 hi, lo;
          fpi_split8_plain(x, hi, lo);
          const uint32_t off = xs_a + (uint32_t)c * 16384u + tc::sw64_chunk_offset(r0 + 8 * g, piece);
          tc::sts128(off, hi);
          tc::sts128(off + 8192u, lo);
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) arrive(xs_full_a + (uint32_t)c * 8u);
      }
      // pull the CTA's next tile into L2 (pure hint)
      const int64_t nt = tile + tile_step;
      if (nt < n_tiles) {
        const int nrows = (int)min((int64_t)128, p.n_atoms - nt * 128);
        const char* hb = reinterpret_cast<const char*>(p.nodes + nt * 128 * 256);
        for (int i = sw * 32 + lane; i < nrows * 8; i += 128) tc::prefetch_l2(hb + (size_t)i * 128);
      }
    }
    if (p.dbg && warp == 18 && lane == 0) p.dbg[(size_t)blockIdx.x * 8 + 4] = s_wait;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (PAIR) tc::cluster_sync();            // the leader's MMAs read the peer's shared memory until the very end
  if (warp == 1) {
    if (PAIR) tc::tmem_dealloc_pair<512>(tmem_base);
    else tc::tmem_dealloc<512>(tmem_base);
  }
}

template <int ACT>
__global__ void __launch_bounds__(FPI_THREADS, 1) fc_readout_pipe_kernel(const FcPipeArgs p) {
  fc_readout_pipe_body<ACT, false>(p);
}

// CTA-pair form (cta_group::2, option "fc_pair"): the two CTAs of a cluster work on two neighbouring tiles with ONE
// M = 256 instruction stream issued by the leader; each CTA stages only its 128 N rows of every W tile.  Per SM that is
// a third less tensor-core operand traffic and half the W traffic through the shared-memory pipe -- the pipe that
// bounds the one-CTA kernel -- and one instruction stream for two tiles.
template <int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FPI_THREADS, 1) fc_readout_pipe_pair_kernel(const FcPipeArgs p) {
  fc_readout_pipe_body<ACT, true>(p);
}

}  // namespace nmr
