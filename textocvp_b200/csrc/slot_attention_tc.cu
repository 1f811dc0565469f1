// SlotAttention streaming pass on the 5th-gen tensor cores (f16 features, the pipeline's format).
// Same math and same outputs as sa_stream_kernel (slot_attention.cu) -- reference src/models/Blocks/attention.py:86-103
// with K/V folded away -- but the two location-axis contractions run as tcgen05 MMAs over ONE smem copy of the features:
//
//   per tile of 128 locations (TMA, 2 x [128 loc x 64 ch] 128B-swizzled boxes = 32 KB):
//     MMA1  D1[128 loc, 32] = X[128 loc, 128 ch] . Gt          A = X, K-major.  Gt rows: 0-7 g_hi, 8-15 g_lo (g split in
//                                                               two f16 so the logits keep ~fp32 accuracy), 16 = ones
//                                                               (-> sum_ch x for the LayerNorm mean), 17-31 zero
//     softmax warps: thread = location (= TMEM lane): tcgen05.ld the 17 useful columns, sum x^2 from the smem tile,
//                    LayerNorm statistics, 8-way softmax over slots IN REGISTERS (no shuffles), a = softmax + eps,
//                    w = a * rstd split in hi/lo f16 -> W[16, 128 loc] written K-major / swizzled to smem
//     MMA2  U[128 ch, 16] += Xt[128 ch, 128 loc] . Wt           A = the SAME smem tile read MN-major; accumulates in TMEM
//                                                               over all tiles of the CTA (16 columns)
//   end: U (hi + lo columns) -> partial[b][chunk][slot][ch]; A_i = sum a, Mw_i = sum w*mu reduced from registers.
//
// warp 0: TMA producer (3-stage ring)   warp 1: TMEM alloc + MMA issue (MMA1 of tile t+1 is issued before MMA2 of tile t)
// warps 2-5 / 6-9: two softmax groups working on even / odd tiles (each owns one D1 and one Wt buffer), so the
// per-tile register-level latency chain of one group overlaps the other's.  113 KB smem, 128 TMEM columns -> 2 CTAs / SM.
#include "host_util.h"
#include "ptx.cuh"
#include "slot_attention.h"

namespace tocvp {

constexpr int ST_TILE = 128;                 // locations per tile
constexpr int ST_STAGES = 3;
constexpr int ST_X_BYTES = 2 * ST_TILE * 128;   // two channel halves of [128 loc][64 ch] f16
constexpr int ST_G_BYTES = 2 * 32 * 128;        // Gt: two channel halves of [32 rows][64 ch]
constexpr int ST_W_BYTES = 2 * 16 * 128;        // Wt: two location halves of [16 rows][64 loc]
constexpr int ST_OFF_G = ST_STAGES * ST_X_BYTES;
constexpr int ST_OFF_W = ST_OFF_G + ST_G_BYTES;
constexpr int ST_OFF_BAR = ST_OFF_W + 2 * ST_W_BYTES;
// barriers + reduction scratch; NO alignment slack: the dynamic shared array is declared __align__(1024) (checked at run
// time), which brings the CTA to 115712 B -- with the 1 KB the hardware reserves per CTA exactly half of the SM's 233472 B,
// so TWO CTAs are resident per SM.  (With the usual 1 KB of slack it was 116736 B and one CTA per SM: ncu r1 showed
// occupancy limited to 1 by shared memory and 42 % DRAM throughput.)
constexpr int ST_SMEM = ST_OFF_BAR + 1024;
constexpr int ST_TMEM_COLS = 128;               // D1: 2 x 32 columns, U: 16 columns at column 64

// byte offset of element (row r, k-element e) in a K-major 128B-swizzled operand stored as halves of 64 k-elements
__device__ __forceinline__ uint32_t sw128_kmajor_off(int r, int e, int rows_per_half) {
  const int half = e >> 6, ee = e & 63;
  return uint32_t(half * rows_per_half * 128 + r * 128 + ((((ee >> 3) ^ (r & 7)) << 4) | ((ee & 7) << 1)));
}

constexpr int ST_THREADS = 320;   // TMA warp, MMA warp, 2 softmax groups of 4 warps (even / odd tiles)

__global__ void __launch_bounds__(ST_THREADS, 2)
sa_stream_tc_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ gvec, float* __restrict__ partial,
                    int tiles_per_chunk, float ln_eps, float attn_eps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();     // the swizzled TMA / UMMA tiles need 1024-byte alignment
  uint8_t* sG = smem + ST_OFF_G;
  uint8_t* sW = smem + ST_OFF_W;
  uint64_t* x_full = reinterpret_cast<uint64_t*>(smem + ST_OFF_BAR);
  uint64_t* x_empty = x_full + ST_STAGES;
  uint64_t* d_full = x_empty + ST_STAGES;   // [2]
  uint64_t* w_full = d_full + 2;            // [2]
  uint64_t* w_empty = w_full + 2;           // [2]
  uint64_t* u_full = w_empty + 2;           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(u_full + 1);
  float* s_red = reinterpret_cast<float*>(tmem_slot + 2);   // [8 warps][16]

  const int b = blockIdx.y, chunk = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* gv = gvec + size_t(b) * SA_GVEC;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < ST_STAGES; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d_full[i], 1);
      mbar_init(&w_full[i], 128);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(u_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, ST_TMEM_COLS);
  // ---- Gt operand: rows 0-7 g_hi, 8-15 g_lo, 16 ones, 17-31 zero (K-major, swizzled exactly as TMA would write it);
  //      one 16-byte chunk (8 channels) per item
  for (int e = threadIdx.x; e < 32 * 16; e += blockDim.x) {
    const int r = e >> 4, c16 = e & 15;                 // row, 8-channel chunk
    const int half = c16 >> 3, c = c16 & 7;
    uint8_t* dst = sG + half * 32 * 128 + r * 128 + ((c ^ (r & 7)) << 4);
    if (r < 8) {
      const float4 g0 = *reinterpret_cast<const float4*>(gv + r * SA_D + c16 * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(gv + r * SA_D + c16 * 8 + 4);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __half h0 = __float2half_rn(g[2 * j]), h1 = __float2half_rn(g[2 * j + 1]);
        hi[j] = pack_half2(__half2float(h0), __half2float(h1));
        lo[j] = pack_half2(g[2 * j] - __half2float(h0), g[2 * j + 1] - __half2float(h1));
      }
      *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(dst + 8 * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);   // row r+8: same swizzle phase
    } else if (r >= 16) {
      const uint32_t one2 = (r == 16) ? 0x3C003C00u : 0u;   // half2(1, 1)
      *reinterpret_cast<uint4*>(dst) = make_uint4(one2, one2, one2, one2);
    }
  }
  fence_proxy_async();     // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int loc0 = chunk * tiles_per_chunk * ST_TILE;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < tiles_per_chunk; ++t) {
        mbar_wait(&x_empty[s], ph ^ 1);
        mbar_expect_tx(&x_full[s], ST_X_BYTES);
        uint8_t* dst = smem + s * ST_X_BYTES;
        tma_load_3d(&tmX, &x_full[s], dst, 0, loc0 + t * ST_TILE, b);
        tma_load_3d(&tmX, &x_full[s], dst + ST_TILE * 128, 64, loc0 + t * ST_TILE, b);
        if (++s == ST_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (converged warp, elected lane)
    constexpr uint32_t idesc1 = make_idesc_f16_ex(128, 32, 0, 0, 0);   // X (K-major) . Gt (K-major)
    constexpr uint32_t idesc2 = make_idesc_f16_ex(128, 16, 0, 1, 0);   // Xt (MN-major view) . Wt (K-major)
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t leader = elect_one_sync();
    const uint32_t g_base = smem_u32(sG);
    auto mma1 = [&](int t) {
      const int s = t % ST_STAGES;
      mbar_wait(&x_full[s], (t / ST_STAGES) & 1);
      tc_fence_after();
      const uint32_t xb = smem_u32(smem + s * ST_X_BYTES);
      const uint32_t d1 = tmem_u + uint32_t((t & 1) * 32);
#pragma unroll
      for (int k = 0; k < 8; ++k) {   // 8 x 16 channels; k-steps 0-3 in channel half 0, 4-7 in half 1
        const uint64_t da = make_desc_sw128(xb + uint32_t((k >> 2) * ST_TILE * 128 + (k & 3) * 32), 1024);
        const uint64_t db = make_desc_sw128(g_base + uint32_t((k >> 2) * 32 * 128 + (k & 3) * 32), 1024);
        umma_f16(d1, da, db, idesc1, k != 0, leader);
      }
      umma_commit(&d_full[t & 1], leader);
    };
    mma1(0);
    for (int t = 0; t < tiles_per_chunk; ++t) {
      if (t + 1 < tiles_per_chunk) mma1(t + 1);
      const int s = t % ST_STAGES;
      mbar_wait(&w_full[t & 1], (t >> 1) & 1);
      tc_fence_after();
      const uint32_t xb = smem_u32(smem + s * ST_X_BYTES);
      const uint32_t wb = smem_u32(sW + (t & 1) * ST_W_BYTES);
#pragma unroll
      for (int k = 0; k < 8; ++k) {   // 8 x 16 locations
        // A = Xt: MN-major, 64-channel blocks ST_TILE*128 B apart (LBO), 8-location groups 1024 B apart (SBO)
        const uint64_t da = make_desc_sw128_ex(xb + uint32_t(k * 2048), ST_TILE * 128, 1024);
        const uint64_t db = make_desc_sw128(wb + uint32_t((k >> 2) * 16 * 128 + (k & 3) * 32), 1024);
        umma_f16(tmem_u + 64, da, db, idesc2, (t | k) != 0, leader);
      }
      umma_commit(&x_empty[s], leader);
      umma_commit(&w_empty[t & 1], leader);
    }
    umma_commit(u_full, leader);
  } else {
    // ---------------------------------------------------------------- softmax warps: thread = location / channel
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;        // 0: even tiles, 1: odd tiles
    const int row = q * 32 + lane;          // TMEM lane
    float sg[SA_S], cb[SA_S], a_acc[SA_S], mw_acc[SA_S];
#pragma unroll
    for (int i = 0; i < SA_S; ++i) {
      sg[i] = gv[SA_S * SA_D + i];
      cb[i] = gv[SA_S * SA_D + SA_S + i];
      a_acc[i] = 0.f;
      mw_acc[i] = 0.f;
    }
    for (int t = grp; t < tiles_per_chunk; t += 2) {
      const int s = t % ST_STAGES;
      mbar_wait(&x_full[s], (t / ST_STAGES) & 1);      // acquire the TMA-written tile for generic loads
      const uint8_t* xr = smem + s * ST_X_BYTES + row * 128;
      float sq[4] = {0.f, 0.f, 0.f, 0.f};               // independent chains (ILP)
#pragma unroll
      for (int hc = 0; hc < 16; ++hc) {
        const int half = hc >> 3, c = hc & 7;
        const uint4 u = *reinterpret_cast<const uint4*>(xr + half * ST_TILE * 128 + ((c ^ (row & 7)) << 4));
        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w4[j]));
          sq[j] += f.x * f.x + f.y * f.y;
        }
      }
      const float ssq = (sq[0] + sq[1]) + (sq[2] + sq[3]);
      mbar_wait(&d_full[t & 1], (t >> 1) & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t((t & 1) * 32), v);
      tmem_ld_wait();
      const float mu = __uint_as_float(v[16]) * (1.f / SA_D);
      const float var = fmaxf(ssq * (1.f / SA_D) - mu * mu, 0.f);
      const float rstd = rsqrtf(var + ln_eps);
      float d[SA_S], mx = -1e30f;
#pragma unroll
      for (int i = 0; i < SA_S; ++i) {
        d[i] = rstd * ((__uint_as_float(v[i]) + __uint_as_float(v[8 + i])) - mu * sg[i]) + cb[i];
        mx = fmaxf(mx, d[i]);
      }
      float den = 0.f;
#pragma unroll
      for (int i = 0; i < SA_S; ++i) {
        d[i] = __expf(d[i] - mx);
        den += d[i];
      }
      const float inv = 1.f / den;
      // Wt buffer (t & 1) is free once MMA2 of tile t-2 has completed
      mbar_wait(&w_empty[t & 1], ((t >> 1) & 1) ^ 1);
      // element (slot row i, location `row`): half = row / 64, 16-byte chunk (row % 64) / 8 XOR i, halfword row % 8;
      // the lo row 8+i has the same swizzle phase, 1024 B further
      uint8_t* wdst = sW + (t & 1) * ST_W_BYTES + (row >> 6) * 16 * 128 + ((row & 7) << 1);
      const int wc = (row & 63) >> 3;
#pragma unroll
      for (int i = 0; i < SA_S; ++i) {
        const float a = d[i] * inv + attn_eps;         // softmax over SLOTS, + eps (attention.py:100)
        const float w = a * rstd;
        a_acc[i] += a;
        mw_acc[i] += w * mu;
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        *reinterpret_cast<__half*>(wdst + i * 128 + ((wc ^ i) << 4)) = hi;
        *reinterpret_cast<__half*>(wdst + i * 128 + ((wc ^ i) << 4) + 1024) = lo;
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&w_full[t & 1]);
    }
    // ---- results: U^T[ch = row][16] (hi + lo)  ->  partial[b][chunk][slot][ch]   (group 0 reads TMEM)
    float* out = partial + (size_t(b) * SA_CHUNKS + chunk) * SA_PART;
    if (grp == 0) {
      mbar_wait(u_full, 0);
      tc_fence_after();
      uint32_t u[16];
      tmem_ld16(tmem_base + (uint32_t(q * 32) << 16) + 64, u);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < SA_S; ++i) out[i * SA_D + row] = __uint_as_float(u[i]) + __uint_as_float(u[8 + i]);
    }
#pragma unroll
    for (int i = 0; i < SA_S; ++i) {
      a_acc[i] = warp_sum(a_acc[i]);
      mw_acc[i] = warp_sum(mw_acc[i]);
    }
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < SA_S; ++i) {
        s_red[(warp - 2) * 16 + i] = a_acc[i];
        s_red[(warp - 2) * 16 + 8 + i] = mw_acc[i];
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (threadIdx.x - 64 < 16) {
      const int i = threadIdx.x - 64;
      float t8 = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) t8 += s_red[w8 * 16 + i];
      out[SA_S * SA_D + i] = t8;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ST_TMEM_COLS);
  }
}

int sa_stream_tc(const __half* feats, size_t seq_stride, int B, int N, const float* gvec, float* partial, float ln_eps,
                 float attn_eps, cudaStream_t stream) {
  TOCVP_CHECK_ARG(N % (SA_CHUNKS * ST_TILE) == 0 && (reinterpret_cast<uintptr_t>(feats) & 15) == 0);
  static SmemAttrOnce attr_once;
  TOCVP_TRY(ensure_smem_attr(attr_once, sa_stream_tc_kernel, ST_SMEM));
  CUtensorMap tmX;
  const uint64_t dims[3] = {uint64_t(SA_D), uint64_t(N), uint64_t(B)};
  const uint64_t str[2] = {uint64_t(SA_D) * 2, uint64_t(seq_stride) * 2};
  const uint32_t box[3] = {64, uint32_t(ST_TILE), 1};
  TOCVP_TRY(encode_tmap(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, feats, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  const dim3 grid(SA_CHUNKS, B);
  sa_stream_tc_kernel<<<grid, ST_THREADS, ST_SMEM, stream>>>(tmX, gvec, partial, N / (SA_CHUNKS * ST_TILE), ln_eps, attn_eps);
  TOCVP_LAUNCHED();
  return TOCVP_OK;
}

}  // namespace tocvp
