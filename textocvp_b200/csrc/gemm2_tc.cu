// CTA-pair tcgen05 GEMM (cta_group::2): 256 x BN output tiles, one tile per pair of SMs at a time.
//     C[M,N] = A[M,K] . W[N,K]^T  (+bias) (ReLU) (+residual)   -> fp32 and/or fp16          (and the implicit-conv mode)
//
// Why: with 128 x 128 single-CTA tiles every SM pulls 32 KB of operands per 256 tensor-clocks = 128 B/clk/SM, and the
// L2 -> SM fabric of the chip sustains ~42 B/clk/SM, so that kernel plateaus near 600-900 TFLOP/s (ncu r1: lts-bound).
// A CTA pair sharing one 256 x 256 tile halves the operand traffic twice over: each CTA loads only its 128 rows of A
// and HALF of the W tile (the tensor core reads the other half from the peer's shared memory), 32 KB per 512 clocks
// = 64 B/clk/SM.
//
// Structure (both CTAs run the same code; rank 0 is the leader):
//   warp 0   TMA producer: own A half (128 x 64) + own W half (BN/2 x 64) per stage, completion bytes are signalled on
//            the LEADER's full barrier (cp.async.bulk.tensor ... cta_group::2)
//   warp 1   TMEM allocator (cta_group::2, both CTAs); in the leader: the single-thread tcgen05.mma.cta_group::2 issuer;
//            tcgen05.commit ... multicast releases the smem stage in BOTH CTAs and publishes the accumulator to BOTH
//   warps 2..9  epilogue on the CTA's own 128 accumulator rows (TMEM lanes); the peer's warps arrive remotely on the
//            leader's tmem-empty barrier.  Results leave through per-warp 128B-swizzled staging tiles and TMA STORES
//            (cp.async.bulk.tensor global <- shared): a thread owns a row, so direct stores would touch 32 different
//            rows per instruction and half-fill every 32-byte sector; with K = 512 the output is as large as the
//            operands and that store pattern, not the tensor pipe, set the kernel's pace (r1 measurement).
#include "gemm.h"
#include "host_util.h"
#include "ptx.cuh"

namespace tocvp {

constexpr int G2_BM = 128;          // rows per CTA (256 per pair)
constexpr int G2_BK = 64;
// TMA warp, MMA warp, ew epilogue warps.  8 is the measured optimum: 16 warps (4 per TMEM lane quarter, one staging
// tile each, 96 registers) ran 25% SLOWER on every predictor shape in the same process (r1 A/B), so it is not built.
constexpr int g2_threads(int ew) { return 64 + 32 * ew; }

struct Gemm2Args {
  int M, N, K;
  const float* bias;
  const float* residual;
  int ldr;
  int res_div, res_mod;
  int relu;
  float* out32;
  int ld32;
  __half* out16;
  int ld16;
  ConvMap cm;
  // LayerNorm folded into the GEMM (LN(x).W^T = rstd*(x.(W*gamma)^T - mu*c) + d, see predictor.cu):
  const float* ln_stats;  // consumer: per-row partial [sum, sumsq] pairs (ln_slots of them) of the fp32 tensor whose f16
                          // copy is A, or null
  int ln_slots;           // partial pairs per row = (producer N) / 64: one per (128-wide tile, 64-column warp half)
  const float* ln_c;      // consumer: per-column c_n = sum_k W'[n,k] (bias then holds d_n)
  float ln_inv_k, ln_eps; // 1 / (LayerNorm width), eps
  float* stats_out;       // producer: writes its [sum, sumsq] of 64 output columns to slot n/64 of the row (no atomics)
  int rev;                // 1: walk the row blocks from the last to the first (see g_tile_rev, host_util.h)
};

// WRES ("W resident", K <= 512, 256-wide tiles): a pair keeps ITS half of the W tile for the whole K extent in shared memory
// (8 k-blocks x 16 KB) and owns one column block nb for its entire life, so only A streams through the ring: the per-SM
// operand traffic drops from 64 to 32 B/clk.  At K = 512 the kernel is bound by the L2 -> SM fabric (operand re-reads plus the
// output stores, ~10 TB/s in total; profiles/gemm2_r1_summary.md), not by the tensor pipe, so halving the re-reads is time.
template <int BN, int EWN, bool WRES>
struct G2Smem {
  static constexpr int A_BYTES = G2_BM * G2_BK * 2;          // 16 KB
  static constexpr int B_BYTES = (BN / 2) * G2_BK * 2;       // 16 KB (BN = 256) / 8 KB (BN = 128)
  static constexpr int STAGE_BYTES = WRES ? A_BYTES : A_BYTES + B_BYTES;
  static constexpr int STAGES = WRES ? 4 : ((BN == 256) ? 4 : 5);
  static constexpr int W_KB = 8;                             // k-blocks held by the resident W half (K <= 512)
  static constexpr int W_BYTES = WRES ? W_KB * B_BYTES : 0;
  static constexpr int EW = EWN;
  static constexpr int STG_TILES = WRES ? 1 : ((BN == 128) ? 3 : (EWN == 16 ? 1 : 2));   // staging tiles per epilogue warp
  static constexpr int STG_BYTES = EW * STG_TILES * 4096;
  static constexpr int BAR_BYTES = 512;                     // barriers (bias / ln_c are read through L1, not staged)
  static constexpr int OFF_W = STAGES * STAGE_BYTES;
  static constexpr int OFF_STG = OFF_W + W_BYTES;
  static constexpr int OFF_BAR = OFF_STG + STG_BYTES;
  static constexpr int TOTAL = OFF_BAR + BAR_BYTES + 1024;
};

// TMA store of one staging tile (32 rows x 128 B, 128B-swizzled) to global; clipped at the tensor bounds by the TMA unit
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// EPI = 2 is the producer side (fp32 residual in and out through TMA, f16 copy + row statistics; 128-wide tiles), EPI = 0
// the general epilogue.
// EPI = 1 specialises the epilogue for f16-only output without a residual (ReLU or nothing; with or without a folded
// LayerNorm: the QKV / cross-q / MLP up projections of the predictor, the hoisted text K|V projection): the other variants'
// code and registers are compiled out (ncu r2: the all-in-one epilogue spilled and took instruction-cache misses on its branch-over code).
template <int BN, bool CONV, int EWN, bool WRES, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(g2_threads(EWN), 1)
gemm2_f16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC16, const __grid_constant__ CUtensorMap tmC32,
                 const __grid_constant__ CUtensorMap tmR, const __grid_constant__ Gemm2Args g) {
  using S = G2Smem<BN, EWN, WRES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* stage_base = smem_align1024(smem_raw);
  uint8_t* w_base = stage_base + S::OFF_W;                                            // WRES: resident W half, [kb][BN/2 x 64]
  uint8_t* stg_base = stage_base + S::OFF_STG;                                        // 1024B aligned
  uint64_t* full = reinterpret_cast<uint64_t*>(stage_base + S::OFF_BAR);              // used in the leader only
  uint64_t* empty = full + S::STAGES;                                                 // per CTA (multicast commit)
  uint64_t* tfull = empty + S::STAGES;                                                // per CTA (multicast commit)
  uint64_t* tempty = tfull + 2;                                                       // leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint64_t* resbar = tempty + 4;                                                      // [8 warps][2]: residual tiles landed
  uint64_t* wfull = resbar + 16;                                                      // WRES, leader: resident W landed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int tiles_m = (g.M + 2 * G2_BM - 1) / (2 * G2_BM);
  const int tiles_n = (g.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (g.K + G2_BK - 1) / G2_BK;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  // i-th tile of this pair.  Default: round-robin over (mb, nb).  WRES: the pair keeps column block nb = pair / ppn and
  // walks the row blocks slot, slot + ppn, ... (ppn = pairs per column block).
  const int ppn = WRES ? npairs / tiles_n : 1;
  auto tile_at = [&](int i, int& mb, int& nb) -> bool {
    bool ok;
    if constexpr (WRES) {
      nb = pair / ppn;
      mb = pair % ppn + i * ppn;
      ok = nb < tiles_n && mb < tiles_m;
    } else {
      const int t = pair + i * npairs;
      mb = t / tiles_n;
      nb = t % tiles_n;
      ok = t < num_tiles;
    }
    if (g.rev) mb = tiles_m - 1 - mb;      // same tiles, opposite order: most recently produced rows first
    return ok;
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 2 * S::EW);   // epilogue warps x 2 CTAs
    }
    for (int i = 0; i < 16; ++i) mbar_init(&resbar[i], 1);
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();              // barrier inits + TMEM allocation visible to both CTAs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above overlapped the previous kernel's tail.  The W-resident producer lane
  // additionally issues its (constant) weight loads before it waits for the previous grid's activations.
  if (!(WRES && threadIdx.x == 0)) pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (each CTA loads its halves)
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int mb, nb;
      if constexpr (WRES) {
        if (tile_at(0, mb, nb)) {                          // this pair's W half for every k-block, once
          const uint32_t wfull_leader = mapa_rank(smem_u32(wfull), 0);
          if (rank == 0) mbar_expect_tx(wfull, 2 * num_kb * S::B_BYTES);
          for (int kb = 0; kb < num_kb; ++kb)
            tma_load_2d_pair(&tmB, wfull_leader, w_base + kb * S::B_BYTES, kb * G2_BK, nb * BN + int(rank) * (BN / 2));
        }
        pdl_wait();
      }
      for (int i = 0; tile_at(i, mb, nb); ++i) {
        const int m0 = mb * 2 * G2_BM + int(rank) * G2_BM;
        const int n0 = nb * BN + int(rank) * (BN / 2);
        const int phase = (CONV && g.cm.tiles_per_phase > 0) ? nb / g.cm.tiles_per_phase : 0;
        const int kpt = CONV ? g.cm.cin / G2_BK : num_kb;
        for (int kb = 0; kb < num_kb; ++kb) {
          int arow = m0, acol = kb * G2_BK;
          if constexpr (CONV) {
            const int tap = kb / kpt;
            arow = m0 + g.cm.off[phase][tap];
            acol = (kb - tap * kpt) * G2_BK;
          }
          mbar_wait(&empty[s], ph ^ 1);
          const uint32_t full_leader = mapa_rank(smem_u32(&full[s]), 0);
          if (rank == 0) mbar_expect_tx(&full[s], 2 * S::STAGE_BYTES);   // both CTAs' bytes land on this barrier
          uint8_t* st = stage_base + s * S::STAGE_BYTES;
          tma_load_2d_pair(&tmA, full_leader, st, acol, arow);
          if constexpr (!WRES) tma_load_2d_pair(&tmB, full_leader, st + S::A_BYTES, kb * G2_BK, n0);
          if (++s == S::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(2 * G2_BM, BN, 0);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t leader = elect_one_sync();
      int s = 0;
      uint32_t ph = 0;
      int mb, nb;
      if constexpr (WRES) {
        if (tile_at(0, mb, nb)) mbar_wait(wfull, 0);
      }
      for (int it = 0; tile_at(it, mb, nb); ++it) {
        const int b = it & 1;
        const uint32_t bph = (it >> 1) & 1;
        mbar_wait(&tempty[b], bph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + uint32_t(b * BN);
#pragma unroll 1
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(stage_base + s * S::STAGE_BYTES);
          const uint64_t da = make_desc_sw128(a0, 1024);
          const uint64_t db = make_desc_sw128(WRES ? smem_u32(w_base + kb * S::B_BYTES) : a0 + S::A_BYTES, 1024);
#pragma unroll
          for (int k = 0; k < G2_BK / 16; ++k)
            umma_f16_pair(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0, leader);
          umma_commit_pair(&empty[s], leader);
          if (++s == S::STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit_pair(&tfull[b], leader);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9) on this CTA's 128 rows
    const int ew = warp - 2;
    const int q = warp & 3;               // TMEM lane quarter
    const int half = ew >> 2;             // which column slice of the tile (BN / (EW/4) columns each)
    constexpr int CW = BN / (S::EW / 4);  // columns per warp: 64 (two 32-column chunks)
    constexpr int NCH = CW / 32;
    uint8_t* stg = stg_base + ew * (S::STG_TILES * 4096);  // this warp's staging tiles
    int sbuf = 0;
    // fp32 residual + fp32 output with two chunks per warp (the N = 512 projections of the predictor): the residual
    // tiles are TMA-LOADED into the staging tiles while the main loop of the tile runs, the accumulator is added in
    // place and the same tile is TMA-stored -- both directions in full 128-byte lines instead of 16 bytes per row
    // (a producer whose fp32 result nobody reads -- only its f16 copy + statistics -- passes out32 = nullptr: no store)
    constexpr bool E1 = EPI == 1, E2 = EPI == 2;
    const bool has_res = E2 || (!E1 && g.residual != nullptr);
    const bool has_out32 = !E1 && g.out32 != nullptr;
    const bool has_out16 = E1 || g.out16 != nullptr;
    const bool has_ln = !E2 && g.ln_stats != nullptr;
    const bool has_bias = g.bias != nullptr;
    const int act = E1 ? (g.relu & 1) : (E2 ? 0 : g.relu);
    const bool tma_res = E2 || (!E1 && !CONV && S::STG_TILES == 3 && has_res && g.res_mod == 0 &&
                                (has_out32 || g.stats_out != nullptr));
    const bool reg_res = has_res && !tma_res;        // residual rows through registers
    const uint32_t sw = uint32_t(lane & 7);
    // Row statistics of a folded LayerNorm (consumer): [sum, sumsq] partials -> rstd, -mu * rstd of this thread's row.
    // With the predictor's 8 partial pairs per row the NEXT tile's are fetched while the last chunk of the current tile
    // is processed, so their L2 latency is off the per-tile critical path (ncu r2: ~1000 clk per tile exposed before).
    const bool ln_pf = has_ln && g.ln_slots == 8;
    auto ln_finish = [&](float sm, float sq, float& rstd, float& nmr) {
      const float mu = sm * g.ln_inv_k;
      rstd = rsqrtf(fmaxf(sq * g.ln_inv_k - mu * mu, 0.f) + g.ln_eps);
      nmr = -mu * rstd;
    };
    auto tile_row = [&](int mbx) { return mbx * 2 * G2_BM + int(rank) * G2_BM + q * 32 + lane; };
    float nx_rstd = 1.f, nx_nmr = 0.f;
    int mb, nb;
    if (ln_pf && tile_at(0, mb, nb)) {
      const int row0 = tile_row(mb);
      if (row0 < g.M) {
        const float4* sp = reinterpret_cast<const float4*>(g.ln_stats + size_t(row0) * 16);
        const float4 a0 = __ldg(sp), a1 = __ldg(sp + 1), a2 = __ldg(sp + 2), a3 = __ldg(sp + 3);
        ln_finish((a0.x + a0.z) + (a1.x + a1.z) + (a2.x + a2.z) + (a3.x + a3.z),
                  (a0.y + a0.w) + (a1.y + a1.w) + (a2.y + a2.w) + (a3.y + a3.w), nx_rstd, nx_nmr);
      }
    }
    for (int it = 0; tile_at(it, mb, nb); ++it) {
      const int b = it & 1;
      const uint32_t bph = (it >> 1) & 1;
      const int row = mb * 2 * G2_BM + int(rank) * G2_BM + q * 32 + lane;
      bool row_ok = row < g.M;
      const int ncol0 = nb * BN + half * CW;
      const float* res_row = nullptr;
      size_t cbase = 0;
      if constexpr (CONV) {
        const int hw = g.cm.Hp * g.cm.Wp;
        const int img = row / hw, rem = row - img * hw;
        const int yp = rem / g.cm.Wp, xp = rem - yp * g.cm.Wp;
        row_ok = row_ok && yp >= 1 && yp <= g.cm.Hp - 2 && xp >= 1 && xp <= g.cm.Wp - 2;
        const int sc = g.cm.up ? 2 : 1;
        cbase = (size_t(img) * g.cm.Hop + size_t((yp - 1) * sc + g.cm.pad)) * g.cm.Wop + size_t((xp - 1) * sc + g.cm.pad);
      }
      if (has_res && row_ok) {
        const int rr = g.res_mod ? (row / g.res_div) % g.res_mod : row;
        res_row = g.residual + size_t(rr) * g.ldr;
      }
      float ln_rstd = nx_rstd, ln_nmr = nx_nmr;   // consumer of a folded LayerNorm: rstd and -mu*rstd of this thread's row
      if (has_ln && !ln_pf && row < g.M) {
        const float4* sp = reinterpret_cast<const float4*>(g.ln_stats + size_t(row) * 2 * g.ln_slots);
        float sm = 0.f, sq = 0.f;
        for (int i = 0; i < g.ln_slots / 2; ++i) {
          const float4 v4 = __ldg(sp + i);
          sm += v4.x + v4.z;
          sq += v4.y + v4.w;
        }
        ln_finish(sm, sq, ln_rstd, ln_nmr);
      }
      float4 pf0, pf1, pf2, pf3;              // next tile's statistics in flight (ln_pf)
      bool pf_ok = false;
      float st_sum = 0.f, st_sq = 0.f;        // producer: row statistics of the fp32 output
      float4 rbuf[2][8];
      auto load_res = [&](int c, float4 (&dst)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int n = ncol0 + c * 32 + j * 4;
          dst[j] = (res_row != nullptr && n < g.N) ? __ldg(reinterpret_cast<const float4*>(res_row + n))
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      if (tma_res) {
        if (lane == 0) {
          bulk_wait_read<0>();                          // the previous tile's stores have drained both staging tiles
          const int trow0 = mb * 2 * G2_BM + int(rank) * G2_BM + q * 32;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            mbar_expect_tx(&resbar[ew * 2 + c], 4096);
            tma_load_2d(&tmR, &resbar[ew * 2 + c], stg + c * 4096, ncol0 + c * 32, trow0);
          }
        }
        __syncwarp();
      } else if (reg_res) {
        load_res(0, rbuf[0]);
      }
      mbar_wait(&tfull[b], bph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(b * BN + half * CW);
      uint32_t v[2][32];
      uint4 hold[4];
      tmem_ld32(t_addr, v[0]);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        tmem_ld_wait();                                            // chunk c is in v[c & 1]
        if (c + 1 < NCH) {
          tmem_ld32(t_addr + uint32_t((c + 1) * 32), v[(c + 1) & 1]);   // next chunk in flight while this one is stored
          if (reg_res) load_res(c + 1, rbuf[(c + 1) & 1]);
        } else {
          // all accumulator columns of this warp are in registers: hand the TMEM buffer back to the leader's MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(mapa_rank(smem_u32(&tempty[b]), 0));
          if (ln_pf) {                           // the second accumulator buffer's registers are free from here on
            int mb2, nb2;
            if (tile_at(it + 1, mb2, nb2)) {
              const int row2 = tile_row(mb2);
              if (row2 < g.M) {
                const float4* sp = reinterpret_cast<const float4*>(g.ln_stats + size_t(row2) * 16);
                pf0 = __ldg(sp); pf1 = __ldg(sp + 1); pf2 = __ldg(sp + 2); pf3 = __ldg(sp + 3);
                pf_ok = true;
              }
            }
          }
        }
        const int n0 = ncol0 + c * 32;
        if constexpr (!CONV) {
          // ---- bias / ReLU / residual in registers, then out through swizzled staging tiles + TMA stores
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[c & 1][j]);
          // bias / c_n / d_n come straight from global memory: the same 16 bytes for every lane (one L1 broadcast hit),
          // which needs neither a staging pass nor an epilogue-wide barrier per tile
          auto col4 = [&](const float* p, int n) {
            return n < g.N ? __ldg(reinterpret_cast<const float4*>(p + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
          };
          if (has_ln) {
            // y = rstd * (acc - mu * c_n) + d_n
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 dd = col4(g.bias, n0 + j4 * 4);
              const float4 cc = col4(g.ln_c, n0 + j4 * 4);
              f[4 * j4] = fmaf(ln_rstd, f[4 * j4], fmaf(ln_nmr, cc.x, dd.x));
              f[4 * j4 + 1] = fmaf(ln_rstd, f[4 * j4 + 1], fmaf(ln_nmr, cc.y, dd.y));
              f[4 * j4 + 2] = fmaf(ln_rstd, f[4 * j4 + 2], fmaf(ln_nmr, cc.z, dd.z));
              f[4 * j4 + 3] = fmaf(ln_rstd, f[4 * j4 + 3], fmaf(ln_nmr, cc.w, dd.w));
            }
          } else if (has_bias) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 bb = col4(g.bias, n0 + j4 * 4);
              f[4 * j4] += bb.x; f[4 * j4 + 1] += bb.y; f[4 * j4 + 2] += bb.z; f[4 * j4 + 3] += bb.w;
            }
          }
          const bool relu_in_cvt = act == 1 && !has_out32 && !has_res;   // f16-only output: fused into the conversion
          if (act == 1 && !relu_in_cvt) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          } else if (act == 2) {             // exact GELU (ViT MLP)
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
          }
          if (reg_res) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 r = rbuf[c & 1][j4];
              f[4 * j4] += r.x; f[4 * j4 + 1] += r.y; f[4 * j4 + 2] += r.z; f[4 * j4 + 3] += r.w;
            }
          }
          const int trow = mb * 2 * G2_BM + int(rank) * G2_BM + q * 32;
          if (tma_res) {
            mbar_wait(&resbar[ew * 2 + c], uint32_t(it & 1));
            uint8_t* dst = stg + c * 4096 + lane * 128;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              float4* sp = reinterpret_cast<float4*>(dst + ((uint32_t(j4) ^ sw) << 4));
              const float4 r = *sp;
              f[4 * j4] += r.x; f[4 * j4 + 1] += r.y; f[4 * j4 + 2] += r.z; f[4 * j4 + 3] += r.w;
              if (g.out32 != nullptr) *sp = make_float4(f[4 * j4], f[4 * j4 + 1], f[4 * j4 + 2], f[4 * j4 + 3]);
            }
            if (g.out32 != nullptr) {
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tmC32, stg + c * 4096, n0, trow);
                bulk_commit();
              }
            } else {
              __syncwarp();                              // every lane has read the residual tile before it is reloaded
            }
            if (g.stats_out != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                st_sum += f[j];
                st_sq = fmaf(f[j], f[j], st_sq);
              }
              if (c == NCH - 1 && row < g.M)
                *reinterpret_cast<float2*>(g.stats_out + (size_t(row) * (g.N / 64) + size_t(ncol0 / 64)) * 2) =
                    make_float2(st_sum, st_sq);
            }
            if (g.out16 != nullptr) {
              // f16 copy of the fp32 result (the A operand of the GEMM that consumes the folded LayerNorm): third tile
              if constexpr (S::STG_TILES == 3) {
                uint4 pk[4];
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                  pk[j8].x = pack_half2(f[8 * j8], f[8 * j8 + 1]);
                  pk[j8].y = pack_half2(f[8 * j8 + 2], f[8 * j8 + 3]);
                  pk[j8].z = pack_half2(f[8 * j8 + 4], f[8 * j8 + 5]);
                  pk[j8].w = pack_half2(f[8 * j8 + 6], f[8 * j8 + 7]);
                }
                uint8_t* hd = stg + 2 * 4096 + lane * 128;
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8)
                  *reinterpret_cast<uint4*>(hd + ((uint32_t((c & 1) * 4 + j8) ^ sw) << 4)) = pk[j8];
                if (c & 1) {
                  fence_proxy_async();
                  __syncwarp();
                  if (lane == 0) {
                    tma_store_2d(&tmC16, stg + 2 * 4096, n0 - 32, trow);
                    bulk_commit();
                  }
                }
              }
            }
          } else if (has_out32) {
            if (lane == 0) {                             // the store that last read this staging tile has drained
              if (S::STG_TILES >= 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
            }
            __syncwarp();
            uint8_t* dst = stg + sbuf * 4096 + lane * 128;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
              *reinterpret_cast<float4*>(dst + ((uint32_t(j4) ^ sw) << 4)) =
                  make_float4(f[4 * j4], f[4 * j4 + 1], f[4 * j4 + 2], f[4 * j4 + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC32, stg + sbuf * 4096, n0, trow);
              bulk_commit();
            }
            if (S::STG_TILES >= 2) sbuf ^= 1;
          }
          if (has_out16 && !tma_res) {
            // two 32-column chunks make one 64-column (128 B) staging row: the even chunk waits in registers
            uint4 p[4];
            if (relu_in_cvt) {
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                p[j8].x = pack_half2_relu(f[8 * j8], f[8 * j8 + 1]);
                p[j8].y = pack_half2_relu(f[8 * j8 + 2], f[8 * j8 + 3]);
                p[j8].z = pack_half2_relu(f[8 * j8 + 4], f[8 * j8 + 5]);
                p[j8].w = pack_half2_relu(f[8 * j8 + 6], f[8 * j8 + 7]);
              }
            } else {
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                p[j8].x = pack_half2(f[8 * j8], f[8 * j8 + 1]);
                p[j8].y = pack_half2(f[8 * j8 + 2], f[8 * j8 + 3]);
                p[j8].z = pack_half2(f[8 * j8 + 4], f[8 * j8 + 5]);
                p[j8].w = pack_half2(f[8 * j8 + 6], f[8 * j8 + 7]);
              }
            }
            if (NCH == 4 && S::STG_TILES == 2 && !has_out32) {
              // f16-only output of a 256-wide tile: the warp's 128 columns fill exactly its two staging tiles, so ONE
              // proxy fence and one bulk group per tile cover both TMA stores (the fence, not the math, is what an
              // epilogue warp waits on: r1 A/B in profiles/gemm2_r1_summary.md)
              if (c == 0) {
                if (lane == 0) bulk_wait_read<0>();       // last tile's stores (issued a whole main loop ago) have drained
                __syncwarp();
              }
              uint8_t* dst = stg + (c >> 1) * 4096 + lane * 128;
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8)
                *reinterpret_cast<uint4*>(dst + ((uint32_t((c & 1) * 4 + j8) ^ sw) << 4)) = p[j8];
              if (c == NCH - 1) {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                  tma_store_2d(&tmC16, stg, ncol0, trow);
                  tma_store_2d(&tmC16, stg + 4096, ncol0 + 64, trow);
                  bulk_commit();
                }
              }
            } else if ((c & 1) == 0) {
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) hold[j8] = p[j8];
            } else {
              if (lane == 0) {
                if (S::STG_TILES >= 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
              }
              __syncwarp();
              uint8_t* dst = stg + sbuf * 4096 + lane * 128;
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                *reinterpret_cast<uint4*>(dst + ((uint32_t(j8) ^ sw) << 4)) = hold[j8];
                *reinterpret_cast<uint4*>(dst + ((uint32_t(4 + j8) ^ sw) << 4)) = p[j8];
              }
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tmC16, stg + sbuf * 4096, n0 - 32, trow);
                bulk_commit();
              }
              if (S::STG_TILES >= 2) sbuf ^= 1;
            }
          }
        } else if (row_ok && n0 < g.N) {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            const int n = n0 + j8 * 8;
            if (n < g.N) {  // N is a multiple of 8
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[c & 1][j8 * 8 + j]);
              if (g.bias != nullptr) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(g.bias + n));        // n + 8 <= N (N % 8 == 0)
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(g.bias + n + 4));
                f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
              }
              if (g.relu == 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
              } else if (g.relu == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = gelu_erf(f[j]);
              }
              // column n -> (phase, channel); a 4-column group never straddles a phase (cpp % 4 == 0)
#pragma unroll
              for (int h4 = 0; h4 < 2; ++h4) {
                const int n4 = n + 4 * h4;
                const int phs = g.cm.up ? n4 / g.cm.cpp : 0;
                const int cc = n4 - phs * g.cm.cpp;
                const size_t pix = cbase + size_t(phs >> 1) * g.cm.Wop + size_t(phs & 1);
                if (phs < 4) {
                  if (g.out32 != nullptr)
                    *reinterpret_cast<float4*>(g.out32 + pix * g.ld32 + cc) =
                        make_float4(f[4 * h4], f[4 * h4 + 1], f[4 * h4 + 2], f[4 * h4 + 3]);
                  if (g.out16 != nullptr) {
                    uint2 p2;
                    p2.x = pack_half2(f[4 * h4], f[4 * h4 + 1]);
                    p2.y = pack_half2(f[4 * h4 + 2], f[4 * h4 + 3]);
                    *reinterpret_cast<uint2*>(g.out16 + pix * g.ld16 + cc) = p2;
                  }
                }
              }
            }
          }
        }
      }
      if (ln_pf) {
        nx_rstd = 1.f; nx_nmr = 0.f;
        if (pf_ok)
          ln_finish((pf0.x + pf0.z) + (pf1.x + pf1.z) + (pf2.x + pf2.z) + (pf3.x + pf3.z),
                    (pf0.y + pf0.w) + (pf1.y + pf1.w) + (pf2.y + pf2.w) + (pf3.y + pf3.w), nx_rstd, nx_nmr);
      }
    }
    if (lane == 0) bulk_wait_read<0>();   // staging tiles must outlive the TMA stores that read them
    __syncwarp();
  }

  tc_fence_before();
  cluster_sync_all();              // both CTAs done with TMEM and with each other's barriers / shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

template <int BN, bool CONV, int EWN, bool WRES, int EPI>
static int launch_gemm2_epi(const CUtensorMap& tmA, const CUtensorMap& tmB, const Gemm2Args& g, cudaStream_t stream) {
  // output maps for the TMA-store epilogue: 32-row x 128-byte boxes (64 f16 / 32 fp32 columns), 128B swizzle
  CUtensorMap tmC16 = tmA, tmC32 = tmA, tmR = tmA;     // placeholders when absent (never dereferenced)
  if (!CONV && g.residual != nullptr && g.res_mod == 0) {
    TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(g.residual) & 15) == 0);
    const uint64_t dims[2] = {uint64_t(g.N), uint64_t(g.M)};
    const uint64_t str[1] = {uint64_t(g.ldr) * 4};
    const uint32_t box[2] = {32, 32};
    TOCVP_TRY(encode_tmap(&tmR, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g.residual, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  if (!CONV && g.out16 != nullptr) {
    const uint64_t dims[2] = {uint64_t(g.N), uint64_t(g.M)};
    const uint64_t str[1] = {uint64_t(g.ld16) * 2};
    const uint32_t box[2] = {64, 32};
    TOCVP_TRY(encode_tmap(&tmC16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, g.out16, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  if (!CONV && g.out32 != nullptr) {
    const uint64_t dims[2] = {uint64_t(g.N), uint64_t(g.M)};
    const uint64_t str[1] = {uint64_t(g.ld32) * 4};
    const uint32_t box[2] = {32, 32};
    TOCVP_TRY(encode_tmap(&tmC32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g.out32, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  using S = G2Smem<BN, EWN, WRES>;
  static SmemAttrOnce attr_once;
  TOCVP_TRY(ensure_smem_attr(attr_once, gemm2_f16_kernel<BN, CONV, EWN, WRES, EPI>, S::TOTAL));
  const int tiles_m = (g.M + 2 * G2_BM - 1) / (2 * G2_BM), tiles_n = (g.N + BN - 1) / BN;
  const int tiles = tiles_m * tiles_n;
  const int pairs = num_sms() / 2;
  int grid = 2 * (tiles < pairs ? tiles : pairs);
  if (WRES) {                                          // tiles_n column blocks x ppn pairs each
    int ppn = pairs / tiles_n;
    ppn = ppn < tiles_m ? ppn : tiles_m;
    grid = 2 * tiles_n * ppn;
  }
  TOCVP_CUDA(launch_pdl(gemm2_f16_kernel<BN, CONV, EWN, WRES, EPI>, dim3(grid), dim3(g2_threads(EWN)), S::TOTAL, stream, tmA,
                        tmB, tmC16, tmC32, tmR, g));
  count_launch();
  return TOCVP_OK;
}

template <int BN, bool CONV, int EWN, bool WRES = false>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const Gemm2Args& g, cudaStream_t stream) {
  if constexpr (!CONV) {
    const bool ln_ok = g.ln_stats == nullptr || (g.ln_c != nullptr && g.bias != nullptr);
    if (ln_ok && g.residual == nullptr && g.out32 == nullptr && g.out16 != nullptr && g.relu != 2 && g.stats_out == nullptr)
      return launch_gemm2_epi<BN, CONV, EWN, WRES, 1>(tmA, tmB, g, stream);
    if constexpr (BN == 128 && !WRES) {
      if (g.ln_stats == nullptr && g.residual != nullptr && g.res_mod == 0 && g.relu == 0 &&
          (g.out32 != nullptr || g.stats_out != nullptr))
        return launch_gemm2_epi<BN, CONV, EWN, WRES, 2>(tmA, tmB, g, stream);
    }
  }
  return launch_gemm2_epi<BN, CONV, EWN, WRES, 0>(tmA, tmB, g, stream);
}

// Tile width for the pair kernel, or 0 if the problem should stay on the single-CTA kernel.
// cost model: waves of pair-tiles x tile width, with the narrower tile paying ~10% for its higher L2 traffic per FLOP
// (measured r1: N = 512 -> 128-wide tiles (5 waves beat 3 of twice the width), N = 1536 / 2048 -> 256-wide).
int gemm2_pick_bn(int M, int N, int force) {
  if (force == 128 || force == 256) return force;
  if (M < 1024 || N < 128) return 0;   // ragged N is fine: TMA zero-fills W rows >= N and clips the stores
  // wide outputs: 256-wide tiles at every row count the rollout meets (r2 sweep, tools/ab_bn_sweep.py: the wave model below
  // picked 128 at M = 8192 / 10240 and lost 17-21 % there once the 256-wide epilogue had been specialised)
  if (N >= 1024 && N % 256 == 0 && M >= 2048) return 256;
  const int pairs = num_sms() / 2;
  const int tm = (M + 255) / 256;
  auto waves = [&](int bn) { return (tm * ((N + bn - 1) / bn) + pairs - 1) / pairs; };
  const double c256 = waves(256) * 256.0;
  const double c128 = waves(128) * 128.0 * 1.1;
  return c256 <= c128 ? 256 : 128;
}

int gemm2_f16(int bn, const __half* A, int lda, const __half* W, int ldw, int M, int N, int K, const float* bias, int relu,
              const float* residual, int ldr, int res_div, int res_mod, float* out32, int ld32, __half* out16, int ld16,
              cudaStream_t stream, const GemmLn* ln) {
  // the fused f16 copy + row statistics live on the TMA-residual path: 128-wide tiles, plain residual rows, fp32 output
  if (ln != nullptr && ln->stats_out != nullptr)
    TOCVP_CHECK_ARG(bn == 128 && residual != nullptr && res_mod == 0 && (out32 != nullptr || out16 != nullptr) && N % 128 == 0);
  if (ln != nullptr && ln->stats != nullptr) TOCVP_CHECK_ARG(ln->slots >= 2 && ln->slots % 2 == 0);
  // bias / c_n are read as float4 by the epilogue
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(bias) & 15) == 0 && (ln == nullptr || (reinterpret_cast<uintptr_t>(ln->c) & 15) == 0));
  CUtensorMap tmA, tmB;
  TOCVP_TRY(encode_tmap_2d_f16(&tmA, A, M, K, lda, G2_BM, G2_BK));
  TOCVP_TRY(encode_tmap_2d_f16(&tmB, W, N, K, ldw, bn / 2, G2_BK));
  const int rev = tile_order_reversed();
  Gemm2Args g{M, N, K, bias, residual, ldr, res_div, res_mod, relu, out32, ld32, out16, ld16, ConvMap{},
              ln ? ln->stats : nullptr, ln ? ln->slots : 0, ln ? ln->c : nullptr, ln ? ln->inv_k : 0.f, ln ? ln->eps : 0.f,
              ln ? ln->stats_out : nullptr, rev};
  if (bn == 256) {
    // W-resident variant: K <= 512 and a column-block-stationary schedule that needs no more rounds than round-robin
    const int pairs = num_sms() / 2;
    const int tiles_m = (M + 255) / 256, tiles_n = (N + 255) / 256;
    if (!opts().gemm_no_wres && K <= 512 && K % G2_BK == 0 && tiles_n <= pairs) {
      int ppn = pairs / tiles_n;
      ppn = ppn < tiles_m ? ppn : tiles_m;
      const int rounds_w = (tiles_m + ppn - 1) / ppn, rounds_rr = (tiles_m * tiles_n + pairs - 1) / pairs;
      if (rounds_w <= rounds_rr) return launch_gemm2<256, false, 8, true>(tmA, tmB, g, stream);
    }
    return launch_gemm2<256, false, 8>(tmA, tmB, g, stream);
  }
  return launch_gemm2<128, false, 8>(tmA, tmB, g, stream);
}

int gemm2_conv_f16(int bn, const __half* X, const __half* W, int M, int N, int K, const ConvMap& cm, const float* bias,
                   int relu, float* out32, __half* out16, int ldo, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(bias) & 15) == 0);
  TOCVP_TRY(encode_tmap_2d_f16(&tmA, X, M, cm.cin, cm.cin, G2_BM, G2_BK));
  TOCVP_TRY(encode_tmap_2d_f16(&tmB, W, N, K, K, bn / 2, G2_BK));
  Gemm2Args g{M, N, K, bias, nullptr, 0, 1, 0, relu, out32, ldo, out16, ldo, cm, nullptr, 0, nullptr, 0.f, 0.f, nullptr, 0};
  if (bn == 256) return launch_gemm2<256, true, 8>(tmA, tmB, g, stream);
  return launch_gemm2<128, true, 8>(tmA, tmB, g, stream);
}

}  // namespace tocvp
