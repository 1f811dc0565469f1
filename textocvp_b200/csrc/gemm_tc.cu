// tcgen05 GEMM for the dense projections of the rollout path:
//     C[M,N] = A[M,K] . W[N,K]^T  (+bias) (ReLU) (+residual)   -> fp32 and/or fp16
// A and W are fp16, K-major (torch.nn.Linear weight layout [out,in] is already K-major), fp32 accumulate
// in TMEM.  One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer (128B-swizzled 128x64 A tile + BNx64 W tile per stage)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16)
//   warps 2..9  epilogue: tcgen05.ld -> bias/ReLU/residual -> global (double-buffered accumulators, so the
//               epilogue of tile i overlaps the main loop of tile i+1)
// Replaces the cuBLAS calls behind nn.Linear at reference src/models/Blocks/attention.py:167-175,255,296-300,
// 352-356 and src/models/Predictors/text_cond_OCVP.py:47-48, src/models/SAVi.py:117-119.
#include "gemm.h"
#include "host_util.h"
#include "ptx.cuh"

namespace tocvp {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 320;   // TMA warp, MMA warp, 8 epilogue warps

struct GemmArgs {
  int M, N, K;
  const float* bias;      // [N] or null
  const float* residual;  // fp32 rows of length N (leading dim ldr) or null
  int ldr;
  int res_div, res_mod;   // residual row = res_mod ? (row / res_div) % res_mod : row
  int relu;
  float* out32;
  int ld32;
  __half* out16;
  int ld16;
  ConvMap cm;             // implicit-convolution mode (CONV instantiations only)
};

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN <= 64) ? 8 : (BN <= 128 ? 6 : 4);
  static constexpr int BAR_BYTES = 256 + 2 * BN * 4;   // mbarriers + tmem slot, then bias staging [2][BN]
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024 /*align slack*/;
};

template <int BN, bool CONV>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ GemmArgs g) {
  using S = GemmSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE_BYTES);
  uint64_t* empty = full + S::STAGES;
  uint64_t* tfull = empty + S::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sbias = reinterpret_cast<float*>(smem + S::STAGES * S::STAGE_BYTES + 256);   // [2][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (g.M + GEMM_BM - 1) / GEMM_BM;
  const int tiles_n = (g.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (g.K + GEMM_BK - 1) / GEMM_BK;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 8);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();        // everything above overlapped the previous kernel's tail (launch_pdl)
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int mb = t / tiles_n, nb = t % tiles_n;
        if constexpr (CONV) {
          // K-blocks are grouped by filter tap; the A tile of a tap is the same pixel rows shifted by a constant
          const int phase = g.cm.tiles_per_phase > 0 ? nb / g.cm.tiles_per_phase : 0;
          const int kpt = g.cm.cin / GEMM_BK;
          int kb = 0;
          for (int tap = 0; tap < g.cm.taps; ++tap) {
            const int arow = mb * GEMM_BM + g.cm.off[phase][tap];
            for (int kc = 0; kc < kpt; ++kc, ++kb) {
              mbar_wait(&empty[s], ph ^ 1);
              mbar_expect_tx(&full[s], S::STAGE_BYTES);
              uint8_t* st = smem + s * S::STAGE_BYTES;
              tma_load_2d(&tmA, &full[s], st, kc * GEMM_BK, arow);
              tma_load_2d(&tmB, &full[s], st + S::A_BYTES, kb * GEMM_BK, nb * BN);
              if (++s == S::STAGES) { s = 0; ph ^= 1; }
            }
          }
        } else {
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect_tx(&full[s], S::STAGE_BYTES);
            uint8_t* st = smem + s * S::STAGE_BYTES;
            tma_load_2d(&tmA, &full[s], st, kb * GEMM_BK, mb * GEMM_BM);
            tma_load_2d(&tmB, &full[s], st + S::A_BYTES, kb * GEMM_BK, nb * BN);
            if (++s == S::STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, elected lane issues)
    {
      constexpr uint32_t idesc = make_idesc_f16(GEMM_BM, BN, 0);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t leader = elect_one_sync();
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int b = it & 1;
        const uint32_t bph = (it >> 1) & 1;
        mbar_wait(&tempty[b], bph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + uint32_t(b * BN);
#pragma unroll 1
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + s * S::STAGE_BYTES);
          const uint64_t da = make_desc_sw128(a0, 1024);
          const uint64_t db = make_desc_sw128(a0 + S::A_BYTES, 1024);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128B swizzle atom: +2 in the (addr>>4) field
            umma_f16(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0, leader);
          }
          umma_commit(&empty[s], leader);
          if (++s == S::STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull[b], leader);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    // Two warps per TMEM lane quarter, each owning half of the tile's columns.  Everything that does not depend on the
    // accumulator (bias -> smem, the fp32 residual row segment -> registers) is fetched BEFORE waiting for the MMA,
    // and the residual of chunk c+1 is in flight while chunk c is processed: for K = 512 the main loop of a tile is
    // only ~2k cycles, so an epilogue that exposes global-load latency would set the pace of the whole kernel.
    const int ew = warp - 2;
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int half = ew >> 2;             // which half of the BN columns
    const int et = threadIdx.x - 64;      // 0..255
    constexpr int CW = BN / 2;            // columns per warp
    constexpr int NCH = CW / 32;          // 32-column chunks per warp
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int mb = t / tiles_n, nb = t % tiles_n;
      const int b = it & 1;
      const uint32_t bph = (it >> 1) & 1;
      if (g.bias != nullptr && et < BN) {
        const int n = nb * BN + et;
        sbias[b * BN + et] = (n < g.N) ? __ldg(g.bias + n) : 0.f;
      }
      const int row = mb * GEMM_BM + q * 32 + lane;
      bool row_ok = row < g.M;
      const int ncol0 = nb * BN + half * CW;
      const float* res_row = nullptr;
      size_t cbase = 0;       // CONV: output pixel index of phase (0,0) for this row
      if constexpr (CONV) {
        const int hw = g.cm.Hp * g.cm.Wp;
        const int img = row / hw, rem = row - img * hw;
        const int yp = rem / g.cm.Wp, xp = rem - yp * g.cm.Wp;
        row_ok = row_ok && yp >= 1 && yp <= g.cm.Hp - 2 && xp >= 1 && xp <= g.cm.Wp - 2;   // border rows: not stored
        const int sc = g.cm.up ? 2 : 1;
        cbase = (size_t(img) * g.cm.Hop + size_t((yp - 1) * sc + g.cm.pad)) * g.cm.Wop + size_t((xp - 1) * sc + g.cm.pad);
      }
      if (g.residual != nullptr && row_ok) {
        const int rr = g.res_mod ? (row / g.res_div) % g.res_mod : row;
        res_row = g.residual + size_t(rr) * g.ldr;
      }
      float4 rbuf[2][8];
      auto load_res = [&](int c, float4 (&dst)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int n = ncol0 + c * 32 + j * 4;
          dst[j] = (res_row != nullptr && n < g.N) ? __ldg(reinterpret_cast<const float4*>(res_row + n))
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      load_res(0, rbuf[0]);
      asm volatile("bar.sync 1, 256;" ::: "memory");   // bias staged (epilogue warps only)
      mbar_wait(&tfull[b], bph);
      tc_fence_after();
      uint32_t v[NCH][32];
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(b * BN + half * CW + c * 32), v[c]);
      tmem_ld_wait();
      // accumulators are in registers: release the TMEM buffer to the MMA warp as early as possible
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[b]);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (c + 1 < NCH) load_res(c + 1, rbuf[(c + 1) & 1]);
        const int n0 = ncol0 + c * 32;
        if (row_ok && n0 < g.N) {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            const int n = n0 + j8 * 8;
            if (n < g.N) {  // N is a multiple of 8
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[c][j8 * 8 + j]);
              if (g.bias != nullptr) {
                const float4 b0 = *reinterpret_cast<const float4*>(&sbias[b * BN + half * CW + c * 32 + j8 * 8]);
                const float4 b1 = *reinterpret_cast<const float4*>(&sbias[b * BN + half * CW + c * 32 + j8 * 8 + 4]);
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
              if (res_row != nullptr) {
                const float4 r0 = rbuf[c & 1][j8 * 2], r1 = rbuf[c & 1][j8 * 2 + 1];
                f[0] += r0.x; f[1] += r0.y; f[2] += r0.z; f[3] += r0.w;
                f[4] += r1.x; f[5] += r1.y; f[6] += r1.z; f[7] += r1.w;
              }
              if constexpr (CONV) {
                // column n -> (phase, channel); a 4-column group never straddles a phase (cpp % 4 == 0)
#pragma unroll
                for (int h4 = 0; h4 < 2; ++h4) {
                  const int n4 = n + 4 * h4;
                  const int phs = g.cm.up ? n4 / g.cm.cpp : 0;
                  const int c = n4 - phs * g.cm.cpp;
                  const size_t pix = cbase + size_t(phs >> 1) * g.cm.Wop + size_t(phs & 1);
                  if (phs < 4) {
                    if (g.out32 != nullptr)
                      *reinterpret_cast<float4*>(g.out32 + pix * g.ld32 + c) =
                          make_float4(f[4 * h4], f[4 * h4 + 1], f[4 * h4 + 2], f[4 * h4 + 3]);
                    if (g.out16 != nullptr) {
                      uint2 p2;
                      p2.x = pack_half2(f[4 * h4], f[4 * h4 + 1]);
                      p2.y = pack_half2(f[4 * h4 + 2], f[4 * h4 + 3]);
                      *reinterpret_cast<uint2*>(g.out16 + pix * g.ld16 + c) = p2;
                    }
                  }
                }
              } else {
              if (g.out32 != nullptr) {
                float* o = g.out32 + size_t(row) * g.ld32 + n;
                *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
                *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
              }
              if (g.out16 != nullptr) {
                uint4 p;
                p.x = pack_half2(f[0], f[1]);
                p.y = pack_half2(f[2], f[3]);
                p.z = pack_half2(f[4], f[5]);
                p.w = pack_half2(f[6], f[7]);
                *reinterpret_cast<uint4*>(g.out16 + size_t(row) * g.ld16 + n) = p;
              }
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BN, bool CONV>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g, cudaStream_t stream) {
  using S = GemmSmem<BN>;
  static SmemAttrOnce attr_once;
  TOCVP_TRY(ensure_smem_attr(attr_once, gemm_f16_kernel<BN, CONV>, S::TOTAL));
  const int tiles = ((g.M + GEMM_BM - 1) / GEMM_BM) * ((g.N + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  TOCVP_CUDA(launch_pdl(gemm_f16_kernel<BN, CONV>, dim3(grid), dim3(GEMM_THREADS), S::TOTAL, stream, tmA, tmB, g));
  count_launch();
  return TOCVP_OK;
}

int gemm2_pick_bn(int M, int N, int force);
int gemm2_f16(int bn, const __half* A, int lda, const __half* W, int ldw, int M, int N, int K, const float* bias, int relu,
              const float* residual, int ldr, int res_div, int res_mod, float* out32, int ld32, __half* out16, int ld16,
              cudaStream_t stream, const GemmLn* ln);
int gemm2_conv_f16(int bn, const __half* X, const __half* W, int M, int N, int K, const ConvMap& cm, const float* bias,
                   int relu, float* out32, __half* out16, int ldo, cudaStream_t stream);

// tocvp_tuning.gemm_mode (per call): 0 = automatic kernel choice, 1 = single-CTA kernel only,
// 128 / 256 = CTA-pair kernel with that tile width wherever it is applicable.
static inline int gemm_mode() {
  const int m = opts().gemm_mode;
  return (m == 1 || m == 128 || m == 256) ? m : 0;
}

// Internal entry used by the stage drivers and by the public tocvp_gemm_f16.
int gemm_f16(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K, const float* bias, int relu,
             const float* residual, int ldr, int res_div, int res_mod, float* out32, int ld32, __half* out16,
             int ld16, cudaStream_t stream) {
  TOCVP_CHECK_ARG(A != nullptr && W != nullptr);
  TOCVP_CHECK_ARG(M > 0 && N > 0 && K > 0);
  TOCVP_CHECK_ARG(N % 8 == 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0);
  TOCVP_CHECK_ARG(out32 != nullptr || out16 != nullptr);
  TOCVP_CHECK_ARG(out32 == nullptr || ld32 % 4 == 0);
  TOCVP_CHECK_ARG(out16 == nullptr || ld16 % 8 == 0);
  TOCVP_CHECK_ARG(residual == nullptr || ldr % 4 == 0);
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0);
  const int g_gemm_mode = gemm_mode();
  if (g_gemm_mode != 1) {
    // CTA-pair kernel (256-row tiles, half the L2 operand traffic) for everything large enough to fill the chip
    const int bn2 = (g_gemm_mode >= 128 && N >= g_gemm_mode) ? g_gemm_mode : gemm2_pick_bn(M, N, 0);
    if (bn2 != 0)
      return gemm2_f16(bn2, A, lda, W, ldw, M, N, K, bias, relu, residual, ldr, res_div, res_mod, out32, ld32, out16, ld16,
                       stream, nullptr);
  }
  // Tile width: wide tiles for wide outputs; 64 keeps enough tiles in flight for narrow / short problems.
  const int tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
  int bn = 128;
  if (N <= 64 || (N % 128 != 0 && N % 64 == 0 && N < 512)) bn = 64;
  else if (tiles_m * ((N + 127) / 128) < num_sms() && N >= 128) bn = 64;
  CUtensorMap tmA, tmB;
  TOCVP_TRY(encode_tmap_2d_f16(&tmA, A, M, K, lda, GEMM_BM, GEMM_BK));
  TOCVP_TRY(encode_tmap_2d_f16(&tmB, W, N, K, ldw, bn, GEMM_BK));
  GemmArgs g{M, N, K, bias, residual, ldr, res_div, res_mod, relu, out32, ld32, out16, ld16, ConvMap{}};
  if (bn == 64) return launch_gemm<64, false>(tmA, tmB, g, stream);
  return launch_gemm<128, false>(tmA, tmB, g, stream);
}

bool gemm_ln_supported(int M, int N) { return gemm_mode() != 1 && gemm2_pick_bn(M, N, 0) != 0 && N % 64 == 0; }

int gemm_f16_ln(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K, const float* bias, int relu,
                const float* residual, int ldr, float* out32, int ld32, __half* out16, int ld16, const GemmLn& ln,
                cudaStream_t stream) {
  TOCVP_CHECK_ARG(A && W && gemm_ln_supported(M, N));
  TOCVP_CHECK_ARG(N % 8 == 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && (out32 || out16));
  // producers need the 128-wide tile (3 staging tiles per warp); consumers take the usual choice
  const int bn = ln.stats_out != nullptr ? 128 : gemm2_pick_bn(M, N, gemm_mode() >= 128 ? gemm_mode() : 0);
  return gemm2_f16(bn, A, lda, W, ldw, M, N, K, bias, relu, residual, ldr, 1, 0, out32, ld32, out16, ld16, stream, &ln);
}

int gemm_conv_f16(const __half* X, const __half* W, int n_img, int N, const ConvMap& cm, const float* bias, int relu,
                  float* out32, __half* out16, int ldo, cudaStream_t stream) {
  TOCVP_CHECK_ARG(X && W && n_img > 0 && N > 0 && N % 8 == 0 && (out32 || out16));
  TOCVP_CHECK_ARG(cm.taps >= 1 && cm.taps <= 9 && cm.cin % GEMM_BK == 0 && cm.cpp % 4 == 0 && ldo % 4 == 0);
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0);
  const long long M64 = (long long)n_img * cm.Hp * cm.Wp;
  TOCVP_CHECK_ARG(M64 < (1ll << 31));
  const int M = int(M64), K = cm.taps * cm.cin;
  const bool phased = cm.tiles_per_phase != 0;     // each phase (cpp columns) has its own tap set: tiles must not straddle
  ConvMap cmx = cm;
  const int g_gemm_mode = gemm_mode();
  if (g_gemm_mode != 1 && M >= 1024 && N >= 128) {
    int bn2 = 0;
    const int unit = phased ? cm.cpp : N;
    if (unit % 256 == 0 && gemm2_pick_bn(M, N, 0) == 256 && g_gemm_mode != 128) bn2 = 256;
    else if (unit % 128 == 0) bn2 = 128;
    if (bn2 != 0) {
      cmx.tiles_per_phase = phased ? cm.cpp / bn2 : 0;
      return gemm2_conv_f16(bn2, X, W, M, N, K, cmx, bias, relu, out32, out16, ldo, stream);
    }
  }
  const int bn = (N >= 128 && (!phased || cm.cpp % 128 == 0)) ? 128 : 64;
  TOCVP_CHECK_ARG(!phased || cm.cpp % bn == 0);
  cmx.tiles_per_phase = phased ? cm.cpp / bn : 0;
  CUtensorMap tmA, tmB;
  TOCVP_TRY(encode_tmap_2d_f16(&tmA, X, M, cm.cin, cm.cin, GEMM_BM, GEMM_BK));
  // W must be allocated with at least `bn` rows (narrow heads are zero-padded by the packer)
  TOCVP_TRY(encode_tmap_2d_f16(&tmB, W, N < bn ? bn : N, K, K, bn, GEMM_BK));
  GemmArgs g{M, N, K, bias, nullptr, 0, 1, 0, relu, out32, ldo, out16, ldo, cmx};
  if (bn == 64) return launch_gemm<64, true>(tmA, tmB, g, stream);
  return launch_gemm<128, true>(tmA, tmB, g, stream);
}

}  // namespace tocvp

extern "C" size_t tocvp_sizeof_tuning(void) { return sizeof(tocvp_tuning); }

extern "C" int tocvp_gemm_f16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                              int relu, const float* residual, int ldr, float* out_f32, int ld32, void* out_f16,
                              int ld16, const tocvp_tuning* tuning, void* stream) {
  tocvp::OptsScope scope(tuning);
  return tocvp::gemm_f16(static_cast<const __half*>(A), lda, static_cast<const __half*>(W), ldw, M, N, K, bias, relu,
                         residual, ldr, 1, 0, out_f32, ld32, static_cast<__half*>(out_f16), ld16,
                         static_cast<cudaStream_t>(stream));
}
