// Multi-head attention for the predictor's short sequences (<= 128 keys, head dim 64), fp32 softmax.
// One CTA per (sequence, head) with one warp per 16-query tile (ceil(Tq/16) warps, so the 80-token window of the named
// config is ONE pass of 5 warps).  K and V of the head are staged once in shared memory, both row-major with the row
// stride padded to 4 (mod 32) words: K fragments are plain 32-bit loads, V fragments come from ldmatrix.trans (no
// transposing scalar stores).  Each warp runs
//     S = Q K^T  ->  softmax in registers (quad shuffles)  ->  O = P V
// on mma.sync.m16n8k16 (f16 operands, fp32 accumulate) with the S accumulator fragments re-used directly as the
// A operand of the second product (no smem round trip).  The problem per CTA (<= 80 x 80 x 64) is far too small to
// amortise a tcgen05/TMEM pipeline, and the kernel is a few % of the step.
// Serves the unmasked self-attention over the <= 10-frame slot window (reference
// src/models/Blocks/attention.py:245-265, 183-193) and the text cross-attention (attention.py:303-319): q, k, v are
// addressed as (base + row*ld + head*64), so the fused QKV GEMM output and the hoisted text K|V are used in place.
#include "host_util.h"
#include "ptx.cuh"

namespace tocvp {

constexpr int ATT_DH = 64;
constexpr int ATT_MAXK = 128;
constexpr int ATT_MAX_WARPS = 8;
constexpr int ATT_KS = 72;    // K / V row stride (halfs): 36 words = 144 B (16-byte aligned rows for ldmatrix)

// two transposed 8x8 b16 tiles x 2: B fragments (k = key, n = head-dim column) of two adjacent 8-column blocks
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}

// four 8x8 b16 tiles, not transposed: B fragments (k = head-dim, n = key) of two adjacent k-steps of one 8-key tile
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}

// 16-byte asynchronous global -> shared copy; `valid` false writes 16 zero bytes (src-size 0: nothing is read)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// NT = compile-time bound on the number of 8-key tiles (score registers s[NT][4]): fewer registers for short key
// sequences -> more resident CTAs per SM for a kernel that is pure latency.
// WARPS / MINB: launch bounds.  The kernel is a latency chain per CTA (load K/V -> sync -> compute -> store), so resident
// CTAs per SM are what hide it: the 5-warp / 80-key shape of the named config is capped at 96 registers for 4 CTAs per SM.
// FULL: Tk == NT * 8 and Tq == WARPS * 16 (the full 10-frame window of the named config): no row / key predicates, no
// masking, one pass per warp.  ncu r2 had the general kernel at ~970 instructions per warp against ~450 of tensor /
// softmax work -- 64-bit address arithmetic per 16-byte copy, run-time tile guards, generic -> shared conversions -- and
// issue-bound (IPC 2.0 with 4 schedulers); shared-memory addresses are 32-bit integers throughout and the global pointers
// of a thread advance by one precomputed stride.
__device__ __forceinline__ void cp_async16_u32(uint32_t smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async16_full(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void sts_u32(uint32_t saddr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 lds_u128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}

template <int NT, int WARPS, int MINB, bool FULL>
__global__ void __launch_bounds__(WARPS * 32, MINB)
mha_kernel(const __half* __restrict__ q, int ldq, int q_seq_rows, const __half* __restrict__ k,
           const __half* __restrict__ v, int ldkv, int Tq, int Tk, int heads, float scale_log2e, __half* __restrict__ out,
           int ldo, int rev) {
  extern __shared__ __align__(16) __half att_smem[];
  constexpr uint32_t ROWB = ATT_KS * 2;                              // bytes per shared row (144)
  const uint32_t sK = smem_u32(att_smem);                            // [NT * 8][ATT_KS]
  const uint32_t sV = sK + NT * 8 * ROWB;                            // [NT * 8][ATT_KS]
  const uint32_t sQ = sV + NT * 8 * ROWB;                            // [WARPS * 16][ATT_KS]: per-warp query / output tiles
  const int bid = rev ? int(gridDim.x) - 1 - int(blockIdx.x) : int(blockIdx.x);   // rev: last sequences first (L2 reuse)
  const int b = bid / heads, h = bid - b * heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int TkP = FULL ? NT * 8 : ((Tk + 15) & ~15);
  const __half* kb = k + size_t(b) * Tk * ldkv + h * ATT_DH;
  const __half* vb = v + size_t(b) * Tk * ldkv + h * ATT_DH;
  const __half* qb = q + size_t(b) * q_seq_rows * ldq + h * ATT_DH;   // q_seq_rows > Tq: a row subset per sequence
  __half* ob = out + size_t(b) * Tq * ldo + h * ATT_DH;
  pdl_wait();
  pdl_trigger();
  // The warp's 16 query rows go through a warp-private shared tile: 16-byte coalesced global loads (8 lanes per 128-byte
  // row) + ldmatrix, instead of 4-byte fragment loads that touch 8 half-used sectors per instruction; the same tile is
  // reused to write the output rows back in full 16-byte pieces (the LSU wavefront count was the busiest unit, ncu r1).
  const uint32_t sQw = sQ + uint32_t(warp) * 16 * ROWB;
  const int lr = lane >> 3, lc = lane & 7;
  const uint32_t sQl = sQw + uint32_t(lr) * ROWB + uint32_t(lc) * 16;   // this lane's 16-byte slot of rows lr, lr+4, ..
  // All of the CTA's inputs are requested up front with cp.async (16 bytes per request, zero-filled past the ends): no
  // staging registers, every load in flight at once, one global-latency phase per CTA.
  auto load_q = [&](int m0) {
    const __half* qp = qb + size_t(m0 + lr) * ldq + lc * 8;
    const size_t q4 = size_t(4) * ldq;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if constexpr (FULL) {
        cp_async16_full(sQl + uint32_t(4 * i) * ROWB, qp + i * q4);
      } else {
        const bool ok = m0 + lr + 4 * i < Tq;
        cp_async16_u32(sQl + uint32_t(4 * i) * ROWB, ok ? qp + i * q4 : qb, ok);
      }
    }
  };
  if (FULL || warp * 16 < Tq) load_q(warp * 16);
  {
    const int r0 = threadIdx.x >> 3, c = threadIdx.x & 7;
    const int rstep = WARPS * 4;                                  // rows per pass of the whole CTA (blockDim / 8)
    const __half* kp = kb + size_t(r0) * ldkv + c * 8;
    const __half* vp = vb + size_t(r0) * ldkv + c * 8;
    const size_t gstep = size_t(rstep) * ldkv;
    uint32_t so = uint32_t(r0) * ROWB + uint32_t(c) * 16;
    if constexpr (FULL) {
      // blockDim is WARPS * 32 here (FULL: Tq == WARPS * 16)
#pragma unroll
      for (int r = 0; r < NT * 8; r += rstep) {
        if (r + r0 < NT * 8) {
          cp_async16_full(sK + so, kp);
          cp_async16_full(sV + so, vp);
        }
        kp += gstep; vp += gstep; so += uint32_t(rstep) * ROWB;
      }
    } else {
      const int rs = int(blockDim.x >> 3);
      const size_t gs = size_t(rs) * ldkv;
      for (int r = r0; r < TkP; r += rs) {
        const bool ok = r < Tk;                                    // padded keys are zero rows (their P is exactly 0)
        cp_async16_u32(sK + so, ok ? kp : kb, ok);
        cp_async16_u32(sV + so, ok ? vp : vb, ok);
        kp += gs; vp += gs; so += uint32_t(rs) * ROWB;
      }
    }
  }
  cp_async_wait_all();
  __syncthreads();
  const int n_tiles = FULL ? NT : TkP / 8;    // key tiles of 8
  // fragment addresses of this lane (bytes, relative to the tile bases)
  const uint32_t q_frag = sQw + uint32_t((lane & 7) + ((lane >> 3) & 1) * 8) * ROWB + uint32_t(lane >> 4) * 16;
  const uint32_t k_frag = sK + uint32_t(lane & 7) * ROWB + uint32_t(lane >> 3) * 16;
  const uint32_t v_frag = sV + uint32_t(lane & 15) * ROWB + uint32_t(lane >> 4) * 16;
  for (int m0 = warp * 16; FULL ? (m0 == warp * 16) : (m0 < Tq); m0 += (blockDim.x >> 5) * 16) {
    if (!FULL && m0 != warp * 16) {
      __syncwarp();
      load_q(m0);
      cp_async_wait_all();
      __syncwarp();
    }
    uint32_t qa[4][4];   // A fragments for the 4 k-steps of the head dim
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(qa[ks], q_frag + ks * 32);
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      if (FULL || nt < n_tiles) {
        // lane l supplies the row address of key nt*8 + (l & 7), head-dim block p*32 + 8*(l >> 3)
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          uint32_t kb4[4];
          ldsm_x4(kb4, k_frag + uint32_t(nt * 8) * ROWB + p * 64);
          mma_16816(s[nt], qa[2 * p], kb4[0], kb4[1]);
          mma_16816(s[nt], qa[2 * p + 1], kb4[2], kb4[3]);
        }
      }
    }
    // softmax over keys (rows g and g+8 of this tile; a row lives in the 4 lanes of a quad)
    float mx0 = -1e30f, mx1 = -1e30f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (FULL || nt < n_tiles) {
        if constexpr (!FULL) {
          const int c = nt * 8 + 2 * t;
          if (c >= Tk) s[nt][0] = s[nt][2] = -1e30f;
          if (c + 1 >= Tk) s[nt][1] = s[nt][3] = -1e30f;
        }
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float nm0 = -mx0 * scale_log2e, nm1 = -mx1 * scale_log2e;    // exp2(s * scale - max * scale): one FFMA + MUFU
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (FULL || nt < n_tiles) {
        s[nt][0] = ex2_ftz(fmaf(s[nt][0], scale_log2e, nm0));
        s[nt][1] = ex2_ftz(fmaf(s[nt][1], scale_log2e, nm0));
        s[nt][2] = ex2_ftz(fmaf(s[nt][2], scale_log2e, nm1));
        s[nt][3] = ex2_ftz(fmaf(s[nt][3], scale_log2e, nm1));
        sum0 += s[nt][0] + s[nt][1];
        sum1 += s[nt][2] + s[nt][3];
      }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    // O = P V : the S fragments of key tiles (2kk, 2kk+1) are exactly the A fragment of k-step kk
    float o[ATT_DH / 8][4];
#pragma unroll
    for (int nd = 0; nd < ATT_DH / 8; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
      if (FULL || kk * 2 < n_tiles) {
        uint32_t pa[4];
        pa[0] = pack_half2(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack_half2(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack_half2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_half2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        // lane l supplies the row address of key kk*16 + (l & 15), column block nd2*16 + 8*(l >> 4)
#pragma unroll
        for (int nd2 = 0; nd2 < ATT_DH / 16; ++nd2) {
          uint32_t vb4[4];
          ldsm_x4_trans(vb4, v_frag + uint32_t(kk * 16) * ROWB + nd2 * 32);
          mma_16816(o[2 * nd2], pa, vb4[0], vb4[1]);
          mma_16816(o[2 * nd2 + 1], pa, vb4[2], vb4[3]);
        }
      }
    }
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
    __syncwarp();                                   // every lane has consumed its Q fragments
    const uint32_t o_lo = sQw + uint32_t(g) * ROWB + uint32_t(t) * 4, o_hi = o_lo + 8 * ROWB;
#pragma unroll
    for (int nd = 0; nd < ATT_DH / 8; ++nd) {
      sts_u32(o_lo + nd * 16, pack_half2(o[nd][0] * inv0, o[nd][1] * inv0));
      sts_u32(o_hi + nd * 16, pack_half2(o[nd][2] * inv1, o[nd][3] * inv1));
    }
    __syncwarp();
    __half* op = ob + size_t(m0 + lr) * ldo + lc * 8;
    const size_t o4 = size_t(4) * ldo;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (FULL || m0 + lr + 4 * i < Tq)
        *reinterpret_cast<uint4*>(op + i * o4) = lds_u128(sQl + uint32_t(4 * i) * ROWB);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Long sequences (> 128 keys: the ViT front-end at the reference JSON's 336 x 336 / 577-token geometry).  Same fragment
// scheme, but the keys stream through shared memory in blocks of 64 (cp.async, double-buffered) with a running
// (max, sum) per query row -- the standard online softmax -- so any Tk fits.  One CTA = 64 queries of one (sequence, head),
// 4 warps x 16 query rows.
// ------------------------------------------------------------------------------------------------------------------
constexpr int AL_BQ = 64, AL_BK = 64;
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(128, 3)
mha_long_kernel(const __half* __restrict__ q, int ldq, const __half* __restrict__ k, const __half* __restrict__ v, int ldkv,
                int Tq, int Tk, int heads, float scale_log2e, __half* __restrict__ out, int ldo) {
  __shared__ __align__(16) __half sQ[AL_BQ * ATT_KS];
  __shared__ __align__(16) __half sK[2][AL_BK * ATT_KS];
  __shared__ __align__(16) __half sV[2][AL_BK * ATT_KS];
  const int b = blockIdx.y / heads, h = blockIdx.y % heads;
  const int q0 = blockIdx.x * AL_BQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const __half* kb = k + size_t(b) * Tk * ldkv + h * ATT_DH;
  const __half* vb = v + size_t(b) * Tk * ldkv + h * ATT_DH;
  const __half* qb = q + size_t(b) * Tq * ldq + h * ATT_DH;
  __half* ob = out + size_t(b) * Tq * ldo + h * ATT_DH;
  pdl_wait();
  pdl_trigger();
  auto load_kv = [&](int blk, int buf) {
    for (int e = threadIdx.x; e < AL_BK * 8; e += 128) {
      const int r = e >> 3, c = e & 7;
      const int key = blk * AL_BK + r;
      const bool ok = key < Tk;                                  // keys past the end are zero rows, masked below
      cp_async16(&sK[buf][r * ATT_KS + c * 8], ok ? kb + size_t(key) * ldkv + c * 8 : kb, ok);
      cp_async16(&sV[buf][r * ATT_KS + c * 8], ok ? vb + size_t(key) * ldkv + c * 8 : vb, ok);
    }
  };
  for (int e = threadIdx.x; e < AL_BQ * 8; e += 128) {
    const int r = e >> 3, c = e & 7;
    const bool ok = q0 + r < Tq;
    cp_async16(&sQ[r * ATT_KS + c * 8], ok ? qb + size_t(q0 + r) * ldq + c * 8 : qb, ok);
  }
  load_kv(0, 0);
  cp_async_commit();
  const int n_blk = (Tk + AL_BK - 1) / AL_BK;
  __half* sQw = sQ + warp * 16 * ATT_KS;
  uint32_t qa[4][4];
  float o[ATT_DH / 8][4];
#pragma unroll
  for (int nd = 0; nd < ATT_DH / 8; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
  float m0 = -1e30f, m1 = -1e30f, l0 = 0.f, l1 = 0.f;           // running max (raw scores) and sum of rows g, g + 8
  for (int blk = 0; blk < n_blk; ++blk) {
    const int buf = blk & 1;
    if (blk + 1 < n_blk) load_kv(blk + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait_group<1>();                                    // block `blk` (and Q) have landed
    __syncthreads();
    if (blk == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        ldmatrix_x4(qa[ks], sQw + ((lane & 7) + ((lane >> 3) & 1) * 8) * ATT_KS + ks * 16 + (lane >> 4) * 8);
    }
    float s[AL_BK / 8][4];
#pragma unroll
    for (int nt = 0; nt < AL_BK / 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const __half* kr = &sK[buf][(nt * 8 + (lane & 7)) * ATT_KS + ((lane >> 3) << 3)];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        uint32_t kb4[4];
        ldmatrix_x4(kb4, kr + p * 32);
        mma_16816(s[nt], qa[2 * p], kb4[0], kb4[1]);
        mma_16816(s[nt], qa[2 * p + 1], kb4[2], kb4[3]);
      }
    }
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nt = 0; nt < AL_BK / 8; ++nt) {
      const int c = blk * AL_BK + nt * 8 + 2 * t;
      if (c >= Tk) s[nt][0] = s[nt][2] = -1e30f;
      if (c + 1 >= Tk) s[nt][1] = s[nt][3] = -1e30f;
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float c0 = ex2_ftz((m0 - mx0) * scale_log2e), c1 = ex2_ftz((m1 - mx1) * scale_log2e);   // rescale of the past
    m0 = mx0; m1 = mx1;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < AL_BK / 8; ++nt) {
      s[nt][0] = ex2_ftz((s[nt][0] - mx0) * scale_log2e);
      s[nt][1] = ex2_ftz((s[nt][1] - mx0) * scale_log2e);
      s[nt][2] = ex2_ftz((s[nt][2] - mx1) * scale_log2e);
      s[nt][3] = ex2_ftz((s[nt][3] - mx1) * scale_log2e);
      sum0 += s[nt][0] + s[nt][1];
      sum1 += s[nt][2] + s[nt][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    l0 = l0 * c0 + sum0;
    l1 = l1 * c1 + sum1;
#pragma unroll
    for (int nd = 0; nd < ATT_DH / 8; ++nd) {
      o[nd][0] *= c0; o[nd][1] *= c0; o[nd][2] *= c1; o[nd][3] *= c1;
    }
#pragma unroll
    for (int kk = 0; kk < AL_BK / 16; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_half2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_half2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_half2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_half2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      const __half* vrow = &sV[buf][(kk * 16 + (lane & 15)) * ATT_KS + ((lane >> 4) << 3)];
#pragma unroll
      for (int nd2 = 0; nd2 < ATT_DH / 16; ++nd2) {
        uint32_t vb4[4];
        ldmatrix_x4_trans(vb4, vrow + nd2 * 16);
        mma_16816(o[2 * nd2], pa, vb4[0], vb4[1]);
        mma_16816(o[2 * nd2 + 1], pa, vb4[2], vb4[3]);
      }
    }
    __syncthreads();                                             // every warp is done with this K / V buffer
  }
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
#pragma unroll
  for (int nd = 0; nd < ATT_DH / 8; ++nd) {
    const int c = nd * 8 + 2 * t;
    *reinterpret_cast<__half2*>(sQw + g * ATT_KS + c) = __floats2half2_rn(o[nd][0] * inv0, o[nd][1] * inv0);
    *reinterpret_cast<__half2*>(sQw + (g + 8) * ATT_KS + c) = __floats2half2_rn(o[nd][2] * inv1, o[nd][3] * inv1);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (lane >> 3) + 4 * i, c = lane & 7;
    const int row = q0 + warp * 16 + r;
    if (row < Tq) *reinterpret_cast<uint4*>(ob + size_t(row) * ldo + c * 8) = *reinterpret_cast<const uint4*>(sQw + r * ATT_KS + c * 8);
  }
}

// Self-attention over sequences of any length (head dim 64): the short kernel for Tk <= 128, the streaming one above beyond.
int mha_f16_any(const __half* q, int ldq, const __half* k, const __half* v, int ldkv, int B, int Tq, int Tk, int heads,
                __half* out, int ldo, cudaStream_t stream);

int mha_f16_sub(const __half* q, int ldq, int q_seq_rows, const __half* k, const __half* v, int ldkv, int B, int Tq, int Tk,
                int heads, __half* out, int ldo, cudaStream_t stream);

// q [B*Tq, ldq], k/v [B*Tk, ldkv] (all f16, head h at column h*64), out [B*Tq, ldo] f16.
int mha_f16(const __half* q, int ldq, const __half* k, const __half* v, int ldkv, int B, int Tq, int Tk, int heads,
            __half* out, int ldo, cudaStream_t stream) {
  return mha_f16_sub(q, ldq, Tq, k, v, ldkv, B, Tq, Tk, heads, out, ldo, stream);
}

// Same with the Tq queries of sequence b starting at row b*q_seq_rows of q (q_seq_rows >= Tq): attention for a subset of
// a sequence's rows (the newest frame's tokens in the predictor's last layer); out stays compact [B*Tq, ldo].
int mha_f16_sub(const __half* q, int ldq, int q_seq_rows, const __half* k, const __half* v, int ldkv, int B, int Tq, int Tk,
                int heads, __half* out, int ldo, cudaStream_t stream) {
  TOCVP_CHECK_ARG(q_seq_rows >= Tq);
  TOCVP_CHECK_ARG(q && k && v && out && B > 0 && Tq > 0 && Tk > 0 && Tk <= ATT_MAXK && heads > 0);
  TOCVP_CHECK_ARG(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0);
  const float scale_log2e = 0.125f * 1.4426950408889634f;   // head_dim^-0.5 (attention.py:187) * log2(e)
  int warps = (Tq + 15) / 16;
  warps = warps > ATT_MAX_WARPS ? ATT_MAX_WARPS : warps;
  const int nt = ((Tk + 15) & ~15) / 8;
  const int rev = tile_order_reversed();
#define MHA_LAUNCH(NTV, W, MB, FULLV)                                                                                 \
  do {                                                                                                                \
    constexpr int smem_bytes = (2 * NTV * 8 + W * 16) * ATT_KS * 2;                                                   \
    static SmemAttrOnce attr_once;                                                                                    \
    if (smem_bytes > 48 * 1024) TOCVP_TRY(ensure_smem_attr(attr_once, mha_kernel<NTV, W, MB, FULLV>, smem_bytes));    \
    TOCVP_CUDA(launch_pdl(mha_kernel<NTV, W, MB, FULLV>, dim3(B * heads), dim3(warps * 32), smem_bytes, stream, q, ldq, \
                          q_seq_rows, k, v, ldkv, Tq, Tk, heads, scale_log2e, out, ldo, rev));                        \
  } while (0)
  if (nt <= 4) {
    if (warps == 5 && Tq == 80 && Tk == 32) MHA_LAUNCH(4, 5, 5, true);       // the named config's text cross-attention
    else if (warps <= 5) MHA_LAUNCH(4, 5, 5, false);
    else MHA_LAUNCH(4, 8, 3, false);
  } else if (nt <= 10) {
    if (warps == 5 && Tq == 80 && Tk == 80) MHA_LAUNCH(10, 5, 4, true);      // ... and its full 10-frame self-attention
    else if (warps <= 5) MHA_LAUNCH(10, 5, 4, false);
    else MHA_LAUNCH(10, 8, 2, false);
  } else if (nt <= 14) {
    MHA_LAUNCH(14, 8, 2, false);
  } else {
    MHA_LAUNCH(16, 8, 2, false);
  }
#undef MHA_LAUNCH
  count_launch();
  return TOCVP_OK;
}

int mha_f16_any(const __half* q, int ldq, const __half* k, const __half* v, int ldkv, int B, int Tq, int Tk, int heads,
                __half* out, int ldo, cudaStream_t stream) {
  if (Tk <= ATT_MAXK) return mha_f16(q, ldq, k, v, ldkv, B, Tq, Tk, heads, out, ldo, stream);
  TOCVP_CHECK_ARG(q && k && v && out && B > 0 && Tq > 0 && heads > 0 && B * heads <= 65535);
  TOCVP_CHECK_ARG(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0);
  const float scale_log2e = 0.125f * 1.4426950408889634f;
  (void)tile_order_reversed();
  TOCVP_CUDA(launch_pdl(mha_long_kernel, dim3((Tq + AL_BQ - 1) / AL_BQ, B * heads), dim3(128), 0, stream, q, ldq, k, v, ldkv,
                        Tq, Tk, heads, scale_log2e, out, ldo));
  count_launch();
  return TOCVP_OK;
}

}  // namespace tocvp

extern "C" int tocvp_mha_f16(const void* q, int ldq, const void* k, const void* v, int ldkv, int B, int Tq, int Tk,
                             int heads, void* out, int ldo, void* stream) {
  return tocvp::mha_f16_any(static_cast<const __half*>(q), ldq, static_cast<const __half*>(k),
                            static_cast<const __half*>(v), ldkv, B, Tq, Tk, heads, static_cast<__half*>(out), ldo,
                            static_cast<cudaStream_t>(stream));
}
