// 5x5 stride-1 "same" convolution + bias + ReLU as a tcgen05 implicit GEMM (NHWC f16 in/out, fp32 accumulate).
//
// The workhorse of the decoder (reference src/models/EncodersDecoders/decoders.py:96-119, 3 x conv5x5 64->64 per
// slot-image = 2.5 GFLOP) and of the encoder (encoders.py:141-153, 3 x conv5x5 32->32).  Replaces cuDNN.
//
// Mapping.  A CTA owns an output tile of 16 rows x 8G columns of one image = G UMMA M-tiles of 128 pixels
// (M index = row*8 + col inside a 16x8 sub-tile).  The input halo (20 rows x (8G+8) cols x CIN) is fetched ONCE
// per tile by a single 4-D TMA box load -- out-of-image coordinates are zero-filled by the TMA unit, which is the
// convolution's zero padding -- into a swizzled, pixel-major smem tile (one CIN*2-byte row per pixel).  Every filter
// tap (ty,tx) then reads the SAME smem tile through a K-major UMMA descriptor whose start address is shifted by
// (ty*WBUF + tx) pixel rows and whose 8-row-group stride (SBO) is one halo row: no im2col copy, 25x on-chip reuse.
// The per-tap [COUT x CIN] weight slices stream through a small TMA ring (they are L2 resident).
//   warp 0: TMA producer   warp 1: TMEM alloc + MMA issuer   warps 2..5: epilogue (bias+ReLU, f16 NHWC store)
// Accumulators are double-buffered in TMEM (2 x G x COUT columns) so the epilogue overlaps the next tile.
#include "host_util.h"
#include "ptx.cuh"

namespace tocvp {

template <int CIN, int COUT, int G, int KS>
struct ConvCfg {
  static constexpr int KB = CIN * 2;                 // bytes per pixel row in smem (64 or 128)
  static constexpr int TILE_H = 16;
  static constexpr int TILE_W = 8 * G;
  static constexpr int WBUF = TILE_W + 8;            // halo width, multiple of 8 (>= TILE_W + 4)
  static constexpr int HROWS = TILE_H + KS - 1;
  static constexpr int TAPS = KS * KS;
  static constexpr int A_BYTES = HROWS * WBUF * KB;  // multiple of 1024 for G in {2,4}
  static constexpr int W_BYTES = COUT * KB;
  static constexpr int WSTAGES = (W_BYTES >= 8192) ? 3 : 8;
  static constexpr int KSTEPS = CIN / 16;
  static constexpr int ACC_COLS = G * COUT;          // per accumulator buffer
  static constexpr int TMEM_COLS = (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128 ? 128 : (2 * ACC_COLS <= 256 ? 256 : 512));
  static constexpr int SMEM = 2 * ((HROWS * WBUF * KB + 1023) & ~1023) + WSTAGES * W_BYTES + 256 + 1024;
  static constexpr int LAYOUT = (KB == 128) ? 2 : 4;  // UMMA layout type: SWIZZLE_128B / SWIZZLE_64B
  static_assert(KB == 128 || KB == 64, "CIN must be 32 or 64");
  static constexpr int A_STRIDE = (A_BYTES + 1023) & ~1023;   // keeps the second buffer 1024B aligned
  static_assert(2 * ACC_COLS <= 512, "TMEM overflow");
};

__device__ __forceinline__ uint64_t make_desc_kmajor(uint32_t saddr, uint32_t sbo_bytes, int layout) {
  return uint64_t((saddr >> 4) & 0x3FFF) | (uint64_t(1) << 16) | (uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32) |
         (uint64_t(1) << 46) | (uint64_t(layout) << 61);
}

struct ConvArgs {
  int n_img, H, W;
  const float* bias;  // [COUT]
  __half* out;        // EPI 0: f16 NHWC [n_img, H, W, COUT]
  float* out4;        // EPI 1: fp32 NHWC [n_img, H, W, 4] (first 4 output channels, no activation)
  int relu;
  // optional fused tail of the SAVi encoder (pair kernel, COUT = 32): y = LayerNorm(relu(conv + b) + posemb[pixel]) in
  // fp32 straight from the accumulators (SoftPositionEmbed + encoder_mlp.0, reference src/models/SAVi.py:232-237, 116)
  const float* ln_posemb;   // [H*W, COUT] or null
  const float* ln_g;
  const float* ln_b;
  float ln_eps;
};

template <int CIN, int COUT, int G, int KS, int EPI>
__global__ void __launch_bounds__(192, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, ConvArgs a) {
  using C = ConvCfg<CIN, COUT, G, KS>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sA = smem;                       // 2 halo buffers
  uint8_t* sW = smem + 2 * C::A_STRIDE;     // weight ring
  uint64_t* w_full = reinterpret_cast<uint64_t*>(sW + C::WSTAGES * C::W_BYTES);
  uint64_t* w_empty = w_full + C::WSTAGES;
  uint64_t* a_full = w_empty + C::WSTAGES;
  uint64_t* a_empty = a_full + 2;
  uint64_t* t_full = a_empty + 2;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_x = a.W / C::TILE_W, tiles_y = a.H / C::TILE_H;
  const int tiles_per_img = tiles_x * tiles_y;
  const int num_tiles = a.n_img * tiles_per_img;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < C::WSTAGES; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&a_full[b], 1);
      mbar_init(&a_empty[b], 1);
      mbar_init(&t_full[b], 1);
      mbar_init(&t_empty[b], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      auto load_halo = [&](int t, int it) {
        const int buf = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const int img = t / tiles_per_img, r = t % tiles_per_img;
        const int y0 = (r / tiles_x) * C::TILE_H, x0 = (r % tiles_x) * C::TILE_W;
        mbar_wait(&a_empty[buf], ph ^ 1);
        mbar_expect_tx(&a_full[buf], C::A_BYTES);
        tma_load_4d(&tmX, &a_full[buf], sA + buf * C::A_STRIDE, 0, x0 - KS / 2, y0 - KS / 2, img);
      };
      int s = 0;
      uint32_t wph = 0;
      int it = 0;
      if (int(blockIdx.x) < num_tiles) load_halo(blockIdx.x, 0);
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        if (t + int(gridDim.x) < num_tiles) load_halo(t + gridDim.x, it + 1);  // prefetch the next tile's halo
        for (int tap = 0; tap < C::TAPS; ++tap) {
          mbar_wait(&w_empty[s], wph ^ 1);
          mbar_expect_tx(&w_full[s], C::W_BYTES);
          tma_load_2d(&tmW, &w_full[s], sW + s * C::W_BYTES, 0, tap * COUT);
          if (++s == C::WSTAGES) { s = 0; wph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (whole warp converged, one elected lane issues)
    {
      constexpr uint32_t idesc = make_idesc_f16(128, COUT, 0);
      constexpr uint32_t SBO = C::WBUF * C::KB;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);   // provably warp-uniform
      const uint32_t leader = elect_one_sync();
      int s = 0;
      uint32_t wph = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(&t_empty[buf], ph ^ 1);
        mbar_wait(&a_full[buf], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + buf * C::A_STRIDE);
        const uint32_t d_base = tmem_u + uint32_t(buf * C::ACC_COLS);
#pragma unroll 1
        for (int tap = 0; tap < C::TAPS; ++tap) {
          const int ty = tap / KS, tx = tap % KS;
          mbar_wait(&w_full[s], wph);
          tc_fence_after();
          const uint64_t db = make_desc_kmajor(smem_u32(sW + s * C::W_BYTES), 8 * C::KB, C::LAYOUT);
#pragma unroll
          for (int j = 0; j < G; ++j) {
            const uint32_t a_addr = a_base + uint32_t((ty * C::WBUF + tx + 8 * j) * C::KB);
            const uint64_t da = make_desc_kmajor(a_addr, SBO, C::LAYOUT);
#pragma unroll
            for (int k = 0; k < C::KSTEPS; ++k) {
              umma_f16(d_base + uint32_t(j * COUT), da + uint64_t(2 * k), db + uint64_t(2 * k), idesc,
                       (tap | k) != 0, leader);
            }
          }
          umma_commit(&w_empty[s], leader);
          if (++s == C::WSTAGES) { s = 0; wph ^= 1; }
        }
        umma_commit(&a_empty[buf], leader);
        umma_commit(&t_full[buf], leader);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int q = warp & 3;
    const int m = q * 32 + lane;          // TMEM lane = pixel index inside the 16x8 sub-tile
    const int pr = m >> 3, pc = m & 7;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int img = t / tiles_per_img, r = t % tiles_per_img;
      const int y = (r / tiles_x) * C::TILE_H + pr;
      const int x0 = (r % tiles_x) * C::TILE_W + pc;
      mbar_wait(&t_full[buf], ph);
      tc_fence_after();
      if constexpr (EPI == 0) {
#pragma unroll 1
        for (int j = 0; j < G; ++j) {
          __half* o = a.out + (size_t(img) * a.H * a.W + size_t(y) * a.W + (x0 + 8 * j)) * COUT;
#pragma unroll
          for (int c = 0; c < COUT / 32; ++c) {
            uint32_t v[32];
            tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(buf * C::ACC_COLS + j * COUT + c * 32), v);
            tmem_ld_wait();
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
              const int n = c * 32 + j8 * 8;
              const float4 b0 = *reinterpret_cast<const float4*>(a.bias + n);
              const float4 b1 = *reinterpret_cast<const float4*>(a.bias + n + 4);
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j8 * 8 + e]);
              f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
              f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
              if (a.relu) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
              }
              uint4 p;
              p.x = pack_half2(f[0], f[1]);
              p.y = pack_half2(f[2], f[3]);
              p.z = pack_half2(f[4], f[5]);
              p.w = pack_half2(f[6], f[7]);
              *reinterpret_cast<uint4*>(o + n) = p;
            }
          }
        }
      } else {
        // narrow head (COUT = 16, 4 real channels): one float4 per pixel, no activation
        const float4 b0 = *reinterpret_cast<const float4*>(a.bias);
#pragma unroll
        for (int j = 0; j < G; ++j) {
          uint32_t v[4];
          tmem_ld4(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(buf * C::ACC_COLS + j * COUT), v);
          tmem_ld_wait();
          float* o = a.out4 + (size_t(img) * a.H * a.W + size_t(y) * a.W + (x0 + 8 * j)) * 4;
          *reinterpret_cast<float4*>(o) = make_float4(__uint_as_float(v[0]) + b0.x, __uint_as_float(v[1]) + b0.y,
                                                      __uint_as_float(v[2]) + b0.z, __uint_as_float(v[3]) + b0.w);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): the two CTAs of a cluster own two different output tiles and the leader issues
// M = 256 MMAs covering sub-tile j of both.  Each CTA stages only HALF of a tap's weight slice (the tensor core reads the
// peer's half from the peer's shared memory): per 128x64x16 MMA-half the smem operand read drops from 4 KB (A) + 2 KB (B)
// to 4 KB + 1 KB per SM -- the smem read port is what bounds the single-CTA kernel at N = 64 (ncu, profiles/) -- and
// the per-SM L2 traffic for weights halves as well.
// ------------------------------------------------------------------------------------------------
//
// XP ("x pairs", the encoder's 32 -> 32 layers): a 32-channel MMA with N = 32 is bound by the A-operand fetch (128 rows x
// 32 B per instruction, ~50 clocks for 16 clocks of math).  The NHWC activation is therefore read as PIXEL PAIRS -- one
// 128-byte row = pixels (2j, 2j+1) x 32 channels, which is the same memory -- and one instruction computes BOTH pixels of
// the pair: N = 64 = [32 outputs of pixel 2j | 32 outputs of pixel 2j+1].  Output 2j needs inputs 2j-2 .. 2j+2 and output
// 2j+1 inputs 2j-1 .. 2j+3, so a filter row becomes KS + 1 horizontal shifts u = 0..5 of one input pixel (32 channels =
// two k-steps, at byte (u & 1) * 64 of pair row u >> 1) against a [64 x 32] weight block that holds tap u for the even
// pixel and tap u - 1 for the odd one (zeros where the tap does not exist; packed by the host).  30 x 2 instructions per
// 256 pixels instead of 25 x 2 per 128: 0.6 x the instruction count for 1.2 x the flops.  With CIN = COUT = 64 as template
// arguments the halo tile, the TMEM layout and the store addressing are exactly the 64-channel kernel's.
template <int CIN, int COUT, int G, int KS, bool GEN = false, bool XP = false>
struct ConvCfg2 : ConvCfg<CIN, COUT, G, KS> {
  using Base = ConvCfg<CIN, COUT, G, KS>;
  static_assert(!XP || (CIN == 64 && COUT == 64 && !GEN), "XP is the 32-channel layer seen as 64-wide pixel pairs");
  static constexpr int WKB = XP ? 64 : Base::KB;            // bytes per weight row = K extent of one tap
  static constexpr int WLAYOUT = (WKB == 128) ? 2 : 4;
  static constexpr int KSTEPS2 = WKB / 32;
  static constexpr int NTAPS = XP ? KS * (KS + 1) : Base::TAPS;
  static constexpr int W_HALF = (COUT / 2) * WKB;
  // 5 (not 6) stages at 64 channels: 230912 + 1024 static + 1024 reserved bytes per CTA left no room for the 1 KB a second,
  // shared-memory-free CTA needs on the SM, so the decoder's layer-1 kernel (chunk pipeline, decoder.cu) could never be
  // co-resident with this kernel (r1 timeline: the convolution waited ~250 us per chunk for layer-1 CTAs to exit).
  static constexpr int WSTAGES = GEN ? 4 : ((W_HALF >= 4096) ? 5 : 8);
  static constexpr int GEN_BYTES = GEN ? 25 * CIN * 4 : 0;   // this tile's 5x5 border-pattern sums, fp32
  // no alignment slack: the dynamic shared array is declared __align__(1024) and checked at run time.  The 32-channel
  // configuration (111104 B, 161 registers x 192 threads, 256 TMEM columns) then fits TWO CTAs per SM.
  static constexpr int LN_BYTES = XP ? 512 : 0;             // XP: bias | gamma | beta of the fused LayerNorm epilogue (3 x 32 floats)
  static constexpr int SMEM = 2 * Base::A_STRIDE + WSTAGES * W_HALF + GEN_BYTES + LN_BYTES + 512;
  static constexpr int CTAS_PER_SM = (2 * (SMEM + 2048) <= 233472 && 2 * Base::TMEM_COLS <= 512 && !GEN) ? 2 : 1;
};

// Generated input (decoder layer 2 only): instead of TMA-loading layer 1's activation, four extra warps COMPUTE the halo
// tile in place -- layer 1 of the spatial-broadcast decoder is relu(P[y,x,:] + S[pattern(y,x),:]) with P batch
// independent and S the 25 border-pattern tap sums of the slot (decoder.cu) -- so that activation (1 GiB written and
// read back per 256-frame chunk) never exists in HBM.
struct ConvGen {
  const float* P;     // [H*W, CIN] fp32
  const float* S;     // [n_img, 25, CIN] fp32
};
__device__ __forceinline__ int border_pat(int v, int n) { return v < 2 ? v : (v >= n - 2 ? v - (n - 5) : 2); }

// VP ("vertical pairs", encoder conv 1): the input rows already hold an x-im2col of TWO image rows (encoder.cu,
// enc_pack_vp_kernel), so the filter collapses to (KS+1)/2 vertical taps two rows apart, all at the centre column; the
// packed tensor has one extra row on top (stored row r = image row r-1).
template <int CIN, int COUT, int G, int KS, bool GEN, bool VP = false, bool XP = false>
__global__ void __cluster_dims__(2, 1, 1)
__launch_bounds__((GEN || XP) ? 320 : 192, ConvCfg2<CIN, COUT, G, KS, GEN, XP>::CTAS_PER_SM)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, ConvArgs a, ConvGen gen) {
  using C = ConvCfg2<CIN, COUT, G, KS, GEN, XP>;
  constexpr int NTAPS = VP ? (KS + 1) / 2 : C::NTAPS;
  constexpr int HALO_X = XP ? 1 : KS / 2;               // XP: the halo starts one pixel PAIR left of the tile
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();     // swizzled TMA / UMMA tiles need 1024-byte alignment
  uint8_t* sA = smem;
  uint8_t* sW = smem + 2 * C::A_STRIDE;
  uint64_t* w_full = reinterpret_cast<uint64_t*>(sW + C::WSTAGES * C::W_HALF);   // leader: bytes of both halves
  uint64_t* w_empty = w_full + C::WSTAGES;                                        // per CTA (multicast commit)
  uint64_t* a_full = w_empty + C::WSTAGES;                                        // leader: both halos
  uint64_t* a_empty = a_full + 2;                                                 // per CTA
  uint64_t* t_full = a_empty + 2;                                                 // per CTA
  uint64_t* t_empty = t_full + 2;                                                 // leader: 4 warps x 2 CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  float* sS = reinterpret_cast<float*>(sW + C::WSTAGES * C::W_HALF + 512);        // GEN: [25][CIN]; XP: [3][32] LN vectors

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int tiles_x = a.W / C::TILE_W, tiles_y = a.H / C::TILE_H;
  const int tiles_per_img = tiles_x * tiles_y;
  const int num_ptiles = (a.n_img * tiles_per_img) / 2;   // pair-tiles: tile 2*pt + rank belongs to CTA `rank`

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < C::WSTAGES; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&a_full[b], GEN ? 8 : 1);     // GEN: one arrival per generator warp of both CTAs
      mbar_init(&a_empty[b], 1);
      mbar_init(&t_full[b], 1);
      mbar_init(&t_empty[b], XP ? 16 : 8);   // epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
  if constexpr (XP) {
    // the fused LayerNorm epilogue reads bias / gamma / beta for every pixel: as 64 + 8 scalar global loads per pixel they
    // throttled the load / store queue (48 % of the kernel's stall samples, ncu r2); staged once, they are broadcast LDS.128
    if (a.ln_g != nullptr && threadIdx.x >= 64 && threadIdx.x < 64 + 96) {
      const int e = threadIdx.x - 64;
      sS[e] = e < 32 ? a.bias[e] : (e < 64 ? a.ln_g[e - 32] : a.ln_b[e - 64]);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (own halo, own half of the weights)
    if (lane == 0) {
      auto load_halo = [&](int pt, int it) {
        const int buf = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const int t = 2 * pt + int(rank);
        const int img = t / tiles_per_img, r = t % tiles_per_img;
        const int y0 = (r / tiles_x) * C::TILE_H, x0 = (r % tiles_x) * C::TILE_W;
        mbar_wait(&a_empty[buf], ph ^ 1);
        if (rank == 0) mbar_expect_tx(&a_full[buf], 2 * C::A_BYTES);
        tma_load_4d_pair(&tmX, mapa_rank(smem_u32(&a_full[buf]), 0), sA + buf * C::A_STRIDE, 0, x0 - HALO_X,
                         y0 - KS / 2 + (VP ? 1 : 0), img);
      };
      int s = 0;
      uint32_t wph = 0;
      int it = 0;
      if (!GEN && pair < num_ptiles) load_halo(pair, 0);
      for (int pt = pair; pt < num_ptiles; pt += npairs, ++it) {
        if (!GEN && pt + npairs < num_ptiles) load_halo(pt + npairs, it + 1);
        for (int tap = 0; tap < NTAPS; ++tap) {
          mbar_wait(&w_empty[s], wph ^ 1);
          if (rank == 0) mbar_expect_tx(&w_full[s], 2 * C::W_HALF);
          tma_load_2d_pair(&tmW, mapa_rank(smem_u32(&w_full[s]), 0), sW + s * C::W_HALF, 0,
                           tap * COUT + int(rank) * (COUT / 2));
          if (++s == C::WSTAGES) { s = 0; wph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (leader CTA)
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(256, COUT, 0);
      constexpr uint32_t SBO = C::WBUF * C::KB;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t leader = elect_one_sync();
      int s = 0;
      uint32_t wph = 0;
      int it = 0;
      for (int pt = pair; pt < num_ptiles; pt += npairs, ++it) {
        const int buf = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(&t_empty[buf], ph ^ 1);
        if (GEN) mbar_wait_cluster(&a_full[buf], ph);     // tiles written by generator warps of BOTH CTAs
        else mbar_wait(&a_full[buf], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + buf * C::A_STRIDE);
        const uint32_t d_base = tmem_u + uint32_t(buf * C::ACC_COLS);
#pragma unroll 1
        for (int tap = 0; tap < NTAPS; ++tap) {
          // XP: tap = (ty, u), u = 0..KS: input pixel u of the row = pair row u >> 1, byte (u & 1) * 64
          const int ty = VP ? 2 * tap : (XP ? tap / (KS + 1) : tap / KS);
          const int tx = VP ? KS / 2 : (XP ? (tap % (KS + 1)) >> 1 : tap % KS);
          const uint32_t sub = XP ? uint32_t((tap % (KS + 1)) & 1) * 64u : 0u;
          mbar_wait(&w_full[s], wph);
          tc_fence_after();
          const uint64_t db = make_desc_kmajor(smem_u32(sW + s * C::W_HALF), 8 * C::WKB, C::WLAYOUT);
#pragma unroll
          for (int j = 0; j < G; ++j) {
            const uint32_t a_addr = a_base + uint32_t((ty * C::WBUF + tx + 8 * j) * C::KB) + sub;
            const uint64_t da = make_desc_kmajor(a_addr, SBO, C::LAYOUT);
#pragma unroll
            for (int k = 0; k < C::KSTEPS2; ++k)
              umma_f16_pair(d_base + uint32_t(j * COUT), da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (tap | k) != 0,
                            leader);
          }
          umma_commit_pair(&w_empty[s], leader);
          if (++s == C::WSTAGES) { s = 0; wph ^= 1; }
        }
        umma_commit_pair(&a_empty[buf], leader);
        umma_commit_pair(&t_full[buf], leader);
      }
    }
  } else if (GEN && warp >= 6) {
    // ---------------------------------------------------------------- generator warps (GEN): compute the halo tile
    if constexpr (GEN) {
      static_assert(CIN == 64, "generator is written for 64-channel rows (128 B, SWIZZLE_128B)");
      const int gt = threadIdx.x - 192;                       // 0..127
      int it = 0;
      for (int pt = pair; pt < num_ptiles; pt += npairs, ++it) {
        const int buf = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const int t = 2 * pt + int(rank);
        const int img = t / tiles_per_img, r = t % tiles_per_img;
        const int y0 = (r / tiles_x) * C::TILE_H - KS / 2, x0 = (r % tiles_x) * C::TILE_W - KS / 2;
        asm volatile("bar.sync 2, 128;" ::: "memory");       // everyone is done reading the previous tile's patterns
        const float* Sg = gen.S + size_t(img) * 25 * CIN;
        for (int e = gt; e < 25 * CIN / 4; e += 128)
          reinterpret_cast<float4*>(sS)[e] = __ldg(reinterpret_cast<const float4*>(Sg) + e);
        asm volatile("bar.sync 2, 128;" ::: "memory");
        mbar_wait(&a_empty[buf], ph ^ 1);                      // the MMAs that read this buffer have completed
        uint8_t* dstA = sA + buf * C::A_STRIDE;
        // 16 halo pixels per pass (8 threads x 16 B per pixel row); GB passes are issued together so that their P loads
        // (L2 hits, ~300+ cycles each) are in flight at once instead of serialising the generator behind the MMAs
        constexpr int PASSES = C::HROWS * C::WBUF / 16;
        constexpr int GB = 10;
        static_assert(PASSES % GB == 0, "halo passes must split into equal batches");
        const int j = gt & 7;
        for (int p0 = 0; p0 < PASSES; p0 += GB) {
          float4 pv[GB][2];
          int pat[GB];
#pragma unroll
          for (int u = 0; u < GB; ++u) {
            const int prow = (p0 + u) * 16 + (gt >> 3);
            const int hy = prow / C::WBUF, hx = prow - hy * C::WBUF;
            const int y = y0 + hy, x = x0 + hx;
            pat[u] = -1;
            if (y >= 0 && y < a.H && x >= 0 && x < a.W) {
              const float* pp = gen.P + (size_t(y) * a.W + x) * CIN + j * 8;
              pv[u][0] = __ldg(reinterpret_cast<const float4*>(pp));
              pv[u][1] = __ldg(reinterpret_cast<const float4*>(pp + 4));
              pat[u] = border_pat(y, a.H) * 5 + border_pat(x, a.W);
            }
          }
#pragma unroll
          for (int u = 0; u < GB; ++u) {
            const int prow = (p0 + u) * 16 + (gt >> 3);
            uint4 pk = make_uint4(0u, 0u, 0u, 0u);             // outside the image: the convolution's zero padding
            if (pat[u] >= 0) {
              const float* ss = sS + pat[u] * CIN + j * 8;
              const float4 s0 = *reinterpret_cast<const float4*>(ss);
              const float4 s1 = *reinterpret_cast<const float4*>(ss + 4);
              pk.x = pack_half2_relu(pv[u][0].x + s0.x, pv[u][0].y + s0.y);
              pk.y = pack_half2_relu(pv[u][0].z + s0.z, pv[u][0].w + s0.w);
              pk.z = pack_half2_relu(pv[u][1].x + s1.x, pv[u][1].y + s1.y);
              pk.w = pack_half2_relu(pv[u][1].z + s1.z, pv[u][1].w + s1.w);
            }
            *reinterpret_cast<uint4*>(dstA + prow * 128 + ((j ^ (prow & 7)) << 4)) = pk;
          }
        }
        fence_proxy_async();                                   // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_rank(smem_u32(&a_full[buf]), 0));
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue (own tile)
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;     // XP: 0 = warps 2-5, 1 = warps 6-9
    const int m = q * 32 + lane;
    const int pr = m >> 3, pc = m & 7;
    int it = 0;
    for (int pt = pair; pt < num_ptiles; pt += npairs, ++it) {
      const int buf = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int t = 2 * pt + int(rank);
      const int img = t / tiles_per_img, r = t % tiles_per_img;
      const int y = (r / tiles_x) * C::TILE_H + pr;
      const int x0 = (r % tiles_x) * C::TILE_W + pc;
      mbar_wait(&t_full[buf], ph);
      tc_fence_after();
      // the tile's (sub-tile j, 32-column slice c) pieces; XP runs two epilogue warps per TMEM lane quarter (warps 2-5
      // take the even pixel of the pair, warps 6-9 the odd one): with one warp per scheduler the LayerNorm epilogue of
      // encoder conv 4 was a dependent-issue chain longer than the tile's MMAs
      constexpr int HP = COUT / 32, NIT = G * HP;
      const size_t pix0 = size_t(y) * a.W + x0;
      bool with_ln = false;
      if constexpr (COUT == 32 || XP) with_ln = a.ln_g != nullptr;
#pragma unroll 1
      for (int i = XP ? part : 0; i < NIT; i += XP ? 2 : 1) {
        const int j = i / HP, c = i % HP;
        uint32_t v[32];
        tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(buf * C::ACC_COLS + j * COUT + c * 32), v);
        tmem_ld_wait();
        __half* o = a.out + (size_t(img) * a.H * a.W + pix0 + size_t(8 * j)) * COUT + c * 32;
        const float* bias = a.bias + (XP ? 0 : c * 32);            // XP: both pixels of the pair share the 32 biases
        if (with_ln) {
          // ---- relu(conv + b) + posemb -> LayerNorm over the pixel's 32 channels, all in this thread's registers
          // (requesting the next piece's embedding ahead of the current one's arithmetic measured no gain: 1.267 vs 1.256 ms)
          const float* pe = a.ln_posemb + (pix0 + size_t(8 * j)) * COUT + c * 32;
          float f[32];
          float sum = 0.f;
          float4 pv[8];
#pragma unroll
          for (int e8 = 0; e8 < 4; ++e8) ldg_nc_256(pe + 8 * e8, pv[2 * e8], pv[2 * e8 + 1]);
#pragma unroll
          for (int e4 = 0; e4 < 8; ++e4) {
            const float4 bb = XP ? reinterpret_cast<const float4*>(sS)[e4] : __ldg(reinterpret_cast<const float4*>(bias) + e4);
            const float4 pp = pv[e4];
            f[4 * e4 + 0] = fmaxf(__uint_as_float(v[4 * e4 + 0]) + bb.x, 0.f) + pp.x;
            f[4 * e4 + 1] = fmaxf(__uint_as_float(v[4 * e4 + 1]) + bb.y, 0.f) + pp.y;
            f[4 * e4 + 2] = fmaxf(__uint_as_float(v[4 * e4 + 2]) + bb.z, 0.f) + pp.z;
            f[4 * e4 + 3] = fmaxf(__uint_as_float(v[4 * e4 + 3]) + bb.w, 0.f) + pp.w;
            sum += (f[4 * e4] + f[4 * e4 + 1]) + (f[4 * e4 + 2] + f[4 * e4 + 3]);
          }
          const float mean = sum * (1.f / 32.f);
          float var = 0.f;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            f[e] -= mean;
            var = fmaf(f[e], f[e], var);
          }
          const float rstd = rsqrtf(var * (1.f / 32.f) + a.ln_eps);
          uint4 pk[4];
#pragma unroll
          for (int e8 = 0; e8 < 4; ++e8) {
            float g[8], gam[8], bet[8];
            if constexpr (XP) {
              const float4 g0 = reinterpret_cast<const float4*>(sS + 32)[2 * e8], g1 = reinterpret_cast<const float4*>(sS + 32)[2 * e8 + 1];
              const float4 b0 = reinterpret_cast<const float4*>(sS + 64)[2 * e8], b1 = reinterpret_cast<const float4*>(sS + 64)[2 * e8 + 1];
              gam[0] = g0.x; gam[1] = g0.y; gam[2] = g0.z; gam[3] = g0.w; gam[4] = g1.x; gam[5] = g1.y; gam[6] = g1.z; gam[7] = g1.w;
              bet[0] = b0.x; bet[1] = b0.y; bet[2] = b0.z; bet[3] = b0.w; bet[4] = b1.x; bet[5] = b1.y; bet[6] = b1.z; bet[7] = b1.w;
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) { gam[e] = __ldg(a.ln_g + 8 * e8 + e); bet[e] = __ldg(a.ln_b + 8 * e8 + e); }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) g[e] = f[8 * e8 + e] * rstd * gam[e] + bet[e];
            pk[e8].x = pack_half2(g[0], g[1]);
            pk[e8].y = pack_half2(g[2], g[3]);
            pk[e8].z = pack_half2(g[4], g[5]);
            pk[e8].w = pack_half2(g[6], g[7]);
          }
          stg_256(o, pk[0], pk[1]);
          stg_256(o + 16, pk[2], pk[3]);
        } else {
          uint4 pk[4];
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            const float4 b0 = *reinterpret_cast<const float4*>(bias + j8 * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(bias + j8 * 8 + 4);
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j8 * 8 + e]);
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
            f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
            if (a.relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
            }
            pk[j8].x = pack_half2(f[0], f[1]);
            pk[j8].y = pack_half2(f[2], f[3]);
            pk[j8].z = pack_half2(f[4], f[5]);
            pk[j8].w = pack_half2(f[6], f[7]);
          }
          stg_256(o, pk[0], pk[1]);
          stg_256(o + 16, pk[2], pk[3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(mapa_rank(smem_u32(&t_empty[buf]), 0));
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
  }
}

template <int CIN, int COUT, int G, int KS, bool GEN = false, bool VP = false, bool XP = false>
static int launch_conv2(const __half* x, const __half* wpacked, const float* bias, __half* out, int n_img, int H, int W,
                        int relu, cudaStream_t stream, ConvGen gen = ConvGen{nullptr, nullptr},
                        const float* ln_posemb = nullptr, const float* ln_g = nullptr, const float* ln_b = nullptr,
                        float ln_eps = 0.f) {
  using C = ConvCfg2<CIN, COUT, G, KS, GEN, XP>;
  TOCVP_CHECK_ARG(H % C::TILE_H == 0 && W % C::TILE_W == 0);
  // the epilogue stores each lane's 64 bytes with 256-bit accesses (and reads the positional embedding the same way)
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 31) == 0 && (reinterpret_cast<uintptr_t>(ln_posemb) & 31) == 0);
  static SmemAttrOnce attr_once;
  TOCVP_TRY(ensure_smem_attr(attr_once, conv_tc2_kernel<CIN, COUT, G, KS, GEN, VP, XP>, C::SMEM));
  const CUtensorMapSwizzle sw = (C::KB == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap tmX, tmW;
  {
    const uint64_t Hin = uint64_t(H) + (VP ? 1 : 0);    // VP: one extra packed row on top of every image
    const uint64_t dims[4] = {uint64_t(CIN), uint64_t(W), Hin, uint64_t(n_img)};
    const uint64_t str[3] = {uint64_t(CIN) * 2, uint64_t(W) * CIN * 2, Hin * W * CIN * 2};
    const uint32_t box[4] = {uint32_t(CIN), uint32_t(C::WBUF), uint32_t(C::HROWS), 1};
    TOCVP_TRY(encode_tmap(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x, dims, str, box, sw));
  }
  {
    constexpr int WK = C::WKB / 2;                        // K extent of one tap's weight rows (XP: 32 of the 64 'channels')
    const uint64_t dims[2] = {uint64_t(WK), uint64_t((VP ? (KS + 1) / 2 : C::NTAPS) * COUT)};
    const uint64_t str[1] = {uint64_t(WK) * 2};
    const uint32_t box[2] = {uint32_t(WK), uint32_t(COUT / 2)};
    TOCVP_TRY(encode_tmap(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, wpacked, dims, str, box,
                          C::WKB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B));
  }
  const int num_ptiles = n_img * (H / C::TILE_H) * (W / C::TILE_W) / 2;
  const int pairs = (num_sms() / 2) * C::CTAS_PER_SM;
  const int grid = 2 * (num_ptiles < pairs ? num_ptiles : pairs);
  ConvArgs a{n_img, H, W, bias, out, nullptr, relu, ln_posemb, ln_g, ln_b, ln_eps};
  conv_tc2_kernel<CIN, COUT, G, KS, GEN, VP, XP><<<grid, (GEN || XP) ? 320 : 192, C::SMEM, stream>>>(tmX, tmW, a, gen);
  TOCVP_LAUNCHED();
  return TOCVP_OK;
}

// tocvp_tuning.conv_mode (per call): 0 = automatic (pair kernel when the tile count is even), 1 = single-CTA kernel only

template <int CIN, int COUT, int G, int KS, int EPI>
static int launch_conv(const __half* x, const __half* wpacked, const float* bias, __half* out, float* out4, int n_img,
                       int H, int W, int relu, cudaStream_t stream) {
  using C = ConvCfg<CIN, COUT, G, KS>;
  TOCVP_CHECK_ARG(H % C::TILE_H == 0 && W % C::TILE_W == 0);
  static SmemAttrOnce attr_once;
  TOCVP_TRY(ensure_smem_attr(attr_once, conv_tc_kernel<CIN, COUT, G, KS, EPI>, C::SMEM));
  const CUtensorMapSwizzle sw = (C::KB == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap tmX, tmW;
  {
    const uint64_t dims[4] = {uint64_t(CIN), uint64_t(W), uint64_t(H), uint64_t(n_img)};
    const uint64_t str[3] = {uint64_t(CIN) * 2, uint64_t(W) * CIN * 2, uint64_t(H) * W * CIN * 2};
    const uint32_t box[4] = {uint32_t(CIN), uint32_t(C::WBUF), uint32_t(C::HROWS), 1};
    TOCVP_TRY(encode_tmap(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x, dims, str, box, sw));
  }
  {
    const uint64_t dims[2] = {uint64_t(CIN), uint64_t(C::TAPS * COUT)};
    const uint64_t str[1] = {uint64_t(CIN) * 2};
    const uint32_t box[2] = {uint32_t(CIN), uint32_t(COUT)};
    TOCVP_TRY(encode_tmap(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, wpacked, dims, str, box, sw));
  }
  const int num_tiles = n_img * (H / C::TILE_H) * (W / C::TILE_W);
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  ConvArgs a{n_img, H, W, bias, out, out4, relu, nullptr, nullptr, nullptr, 0.f};
  conv_tc_kernel<CIN, COUT, G, KS, EPI><<<grid, 192, C::SMEM, stream>>>(tmX, tmW, a);
  TOCVP_LAUNCHED();
  return TOCVP_OK;
}

// x: f16 NHWC [n_img,H,W,cin]; wpacked: f16 [25, cout, cin] (tap-major, tap = ky*5+kx); out: f16 NHWC [n_img,H,W,cout]
int conv5x5_f16(const __half* x, const __half* wpacked, const float* bias, __half* out, int n_img, int H, int W, int cin,
                int cout, int relu, cudaStream_t stream) {
  TOCVP_CHECK_ARG(x && wpacked && bias && out && n_img > 0);
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const bool pair_ok = opts().conv_mode == 0 && H % 16 == 0 && W % 32 == 0 && ((n_img * (H / 16) * (W / 32)) % 2 == 0);
  if (cin == 64 && cout == 64) {
    if (pair_ok) return launch_conv2<64, 64, 4, 5>(x, wpacked, bias, out, n_img, H, W, relu, stream);
    return launch_conv<64, 64, 4, 5, 0>(x, wpacked, bias, out, nullptr, n_img, H, W, relu, stream);
  }
  if (cin == 32 && cout == 32) {
    if (pair_ok) return launch_conv2<32, 32, 4, 5>(x, wpacked, bias, out, n_img, H, W, relu, stream);
    return launch_conv<32, 32, 4, 5, 0>(x, wpacked, bias, out, nullptr, n_img, H, W, relu, stream);
  }
  set_last_error(__FILE__, __LINE__, "conv5x5_f16: only 64->64 and 32->32 channels are instantiated");
  return TOCVP_ERR_BAD_ARG;
}

// Encoder conv 4 with the positional-embedding add and LayerNorm(32) fused into the epilogue (pair kernel only).
int conv5x5_ln_f16(const __half* x, const __half* wpacked, const float* bias, const float* posemb, const float* ln_g,
                   const float* ln_b, float ln_eps, __half* out, int n_img, int H, int W, cudaStream_t stream) {
  TOCVP_CHECK_ARG(x && wpacked && bias && posemb && ln_g && ln_b && out && n_img > 0);
  TOCVP_CHECK_ARG(H % 16 == 0 && W % 32 == 0 && ((n_img * (H / 16) * (W / 32)) % 2 == 0));
  return launch_conv2<32, 32, 4, 5>(x, wpacked, bias, out, n_img, H, W, 1, stream, ConvGen{nullptr, nullptr}, posemb, ln_g,
                                    ln_b, ln_eps);
}

// Encoder 32 -> 32 layers in pixel-pair form (XP, see ConvCfg2).  x / out: f16 NHWC [n_img, H, W, 32] (W % 64 == 0);
// wxp: f16 [30, 64, 32] = [(ky, u)][pixel parity * 32 + cout][cin] with W[ky][u - parity] or zeros (host-packed).
// With ln_g != null the epilogue also adds the positional embedding and applies LayerNorm(32) (encoder conv 4).
int conv5x5_xp_f16(const __half* x, const __half* wxp, const float* bias, const float* posemb, const float* ln_g,
                   const float* ln_b, float ln_eps, __half* out, int n_img, int H, int W, cudaStream_t stream) {
  TOCVP_CHECK_ARG(x && wxp && bias && out && n_img > 0);
  TOCVP_CHECK_ARG(H % 16 == 0 && W % 64 == 0 && ((n_img * (H / 16) * (W / 64)) % 2 == 0));
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  return launch_conv2<64, 64, 4, 5, false, false, true>(x, wxp, bias, out, n_img, H, W / 2, 1, stream, ConvGen{nullptr, nullptr},
                                                        posemb, ln_g, ln_b, ln_eps);
}

// Encoder conv 1 on the x-im2col packed input (encoder.cu): xp f16 [n_img, H+1, W, 32], wpacked f16 [3, 32, 32].
int conv5x5_vp_f16(const __half* xp, const __half* wpacked, const float* bias, __half* out, int n_img, int H, int W,
                   cudaStream_t stream) {
  TOCVP_CHECK_ARG(xp && wpacked && bias && out && n_img > 0);
  TOCVP_CHECK_ARG(H % 16 == 0 && W % 32 == 0 && ((n_img * (H / 16) * (W / 32)) % 2 == 0));
  return launch_conv2<32, 32, 4, 5, false, true>(xp, wpacked, bias, out, n_img, H, W, 1, stream);
}

// Decoder layer 2 with layer 1 generated in the kernel (see ConvGen): P fp32 [H*W,64], S fp32 [n_img,25,64].
// `x_dummy` only anchors the (unused) input tensor map: any valid device pointer.
int conv5x5_gen_f16(const float* P, const float* S, const __half* x_dummy, const __half* wpacked, const float* bias,
                    __half* out, int n_img, int H, int W, cudaStream_t stream) {
  TOCVP_CHECK_ARG(P && S && x_dummy && wpacked && bias && out && n_img > 0);
  TOCVP_CHECK_ARG(H % 16 == 0 && W % 32 == 0 && ((n_img * (H / 16) * (W / 32)) % 2 == 0));
  return launch_conv2<64, 64, 4, 5, true>(x_dummy, wpacked, bias, out, n_img, H, W, 1, stream, ConvGen{P, S});
}

// Decoder head: conv3x3 64 -> 4 (weights zero-padded to 16 output channels: the narrowest UMMA N for M = 128), no
// activation.  wpacked: f16 [9, 16, 64]; bias fp32 [4]; out4: fp32 NHWC [n_img, H, W, 4].
int conv3x3_head_f16(const __half* x, const __half* wpacked, const float* bias, float* out4, int n_img, int H, int W,
                     cudaStream_t stream) {
  TOCVP_CHECK_ARG(x && wpacked && bias && out4 && n_img > 0);
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out4) & 15) == 0);
  return launch_conv<64, 16, 4, 3, 1>(x, wpacked, bias, nullptr, out4, n_img, H, W, 0, stream);
}

}  // namespace tocvp

extern "C" int tocvp_conv5x5_f16(const void* x, const void* w_packed, const float* bias, void* out, int n_img, int H,
                                 int W, int cin, int cout, int relu, const tocvp_tuning* tuning, void* stream) {
  tocvp::OptsScope scope(tuning);
  return tocvp::conv5x5_f16(static_cast<const __half*>(x), static_cast<const __half*>(w_packed), bias,
                            static_cast<__half*>(out), n_img, H, W, cin, cout, relu, static_cast<cudaStream_t>(stream));
}
