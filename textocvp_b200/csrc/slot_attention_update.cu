// Per-slot update kernel of the corrector, second version (round 2).
//
// What it computes is unchanged (slot_attention.cu, sa_update_kernel): from the streaming pass' partial sums, the weighted
// mean, V projection, GRUCell, LayerNorm + residual MLP (reference src/models/Blocks/attention.py:103-110), optionally the
// post-norm transition TransformerBlock (attention.py:387-395) and the query / folded-key vectors of the next streaming pass,
// all at fp32-level accuracy (the slot state is recurrent over the whole video).
//
// What changed, in the order the measurements forced it:
//  1. Where the weights come from.  A CTA owns 16 slot rows and needs ALL 1.8 MB of weights for them; the first version read
//     every B fragment with __ldg straight from L2 inside the MMA loop (ncu r1: 42 M warp instructions, issue active 38 %,
//     91-120 us per launch against 60 us for the streaming pass it follows).  Here the weights stream through a 3-slot
//     shared-memory ring filled by bulk async copies (cp.async.bulk, completion on an mbarrier); the ring runs ahead ACROSS
//     layers, following a schedule built at kernel start, so the copies of layer i+1 are in flight while layer i computes.
//     One copy per CHUNK (24-32 KB): a first attempt with one 512-byte copy per weight row (2176 per CTA) was slower than
//     the kernel it replaces -- the copy engine's per-request cost, not bytes, set the pace.
//  2. The arithmetic.  With the weights in shared memory the kernel was bound by the legacy tensor path itself: 3xTF32 needs
//     six m16n8k8 MMAs per 16 x 8 x 16 block (66 us per launch).  Now every operand is split into TWO IEEE f16 values,
//     x = hi + lo / 2048 (hi = f16(x), lo = f16((x - hi) * 2048): 22 mantissa bits, the 2^11 pre-scale keeps lo out of the
//     f16 subnormals), and the product is hi.hi + (hi.lo + lo.hi) / 2048 in two fp32 accumulators: THREE m16n8k16 MMAs per
//     block, same accuracy class as 3xTF32 (measured against the fp32 SIMT version: tools/ab_corrector.py).
//  3. The weight layout.  The host packs the split weights in MMA-fragment order (tocvp_sa_weights.stream_*): per 16 x 8
//     block 32 lanes x 16 bytes = {b0_hi, b1_hi, b0_lo, b1_lo}, so a lane's whole B operand is ONE conflict-free LDS.128 and
//     no split arithmetic is spent on weights in the kernel.  Bytes per weight stay 4.
//  4. What did NOT help (measured, B = 256, C|T|A launch, CUPTI): the same stream fetched ONCE per cluster of 2 / 4 CTAs by
//     multicast bulk copies (UBLKCP.S.G.MULTICAST, remote `empty` arrives): 65 / 63 us against 59 us without clusters --
//     a quarter of the L2 -> SM bytes buys nothing, the cluster barriers cost a little; an evict_last L2 policy on the
//     weight stream and a fourth ring slot: +-1 us.  The ncu stall profile of this version is flat (wait 24 %, long
//     scoreboard 18 %, barrier 15 %, short scoreboard 13 %, L2 throughput 4 %): with 8 warps per SM the kernel is a
//     latency chain per chunk (LDS -> split -> dependent HMMAs -> barrier), ~1 us per 32 KB chunk, not a bandwidth problem.
//     The multicast path is kept behind U2_CLUSTER (1 = off, the default).
// Activations are row-major fp32 [16][K + 8] (conflict-free 64-bit A loads), split per k-step in registers.
#include "host_util.h"
#include "ptx.cuh"
#include "slot_attention.h"

namespace tocvp {

using SaWeights = tocvp_sa_weights;

constexpr int U2_R = 16;                 // slot rows per CTA (one mma M tile)
constexpr int U2_THREADS = 256;
constexpr int U2_WARPS = U2_THREADS / 32;
constexpr int U2_LDA = SA_D + 8;         // 136: row pitch of the [16][128] activation buffers (pitch = 8 mod 32 words: the
constexpr int U2_LDB = 512 + 8;          // 520: 64-bit A-fragment loads and C stores of a half-warp hit 32 distinct banks)
constexpr int U2_STAGES = 4;             // ring depth: three chunks in flight while one is consumed
constexpr int U2_CLUSTER = 1;            // > 1: CTAs of a cluster share one weight stream by multicast (measured slower, see 4.)
constexpr int U2_SLOT_FLOATS = 64 * 128;         // 8192 words = 32 KB: largest chunk (rows x N x 4 bytes)
constexpr int U2_MAX_SEGS = 16;
constexpr int U2_DO_C = 1, U2_DO_T = 2, U2_DO_A = 4;   // same flags as sa_update_kernel

struct U2Seg {
  const float* w;   // K x N weights as split f16 pairs in fragment order (see header), K * N 32-bit words
  int K, N;
};

// weight rows per ring slot: 32 KB (N = 128, 256, 512) or 24 KB (N = 384) of payload
__host__ __device__ constexpr int u2_chunk_rows(int N) { return N <= 128 ? 64 : (N <= 256 ? 32 : 16); }

// L2 eviction-priority policy for the weight stream: evict_last.  Between two update launches the streaming pass reads
// 268 MB of features through the 126 MB L2; without the hint every launch found its 1.8 MB of weights evicted and paid DRAM
// latency on the first touch of every chunk (2 us per 32 KB chunk with two chunks of look-ahead = the kernel's whole time).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// one copy, delivered to the same shared-memory offset of every CTA in `cta_mask`, completing `bytes` on the mbarrier at
// the same offset in each of them
__device__ __forceinline__ void bulk_g2s_multicast(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t cta_mask,
                                                   uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint [%0], [%1], %2, "
      "[%3], %4, %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

struct U2Pipe {
  const U2Seg* segs;   // shared memory
  int n_segs;
  float* ring;         // shared memory: U2_STAGES slots of U2_SLOT_FLOATS floats
  uint64_t* full;      // shared memory: [U2_STAGES]  data of the slot has landed (this CTA)
  uint64_t* empty;     // shared memory: [U2_STAGES]  every CTA of the cluster has released the slot (waited on by the issuer)
  int cidx;            // chunks consumed so far (uniform over the CTA)
  int pseg, pk, pidx;  // producer cursor: next chunk to request (used by warp 0 only, uniform there)
  int n_chunks;        // chunks in the whole schedule
  uint32_t rank;       // CTA rank in the cluster
  uint32_t ewaits;     // bit s: parity of the next wait on empty[s] (issuer side)
  uint64_t policy;     // L2 evict_last policy of the weight stream
};

// warp 0 (converged): request the next chunk of the schedule, if any, into ring slot pidx % STAGES
__device__ __forceinline__ void u2_issue(U2Pipe& p, int lane) {
  if (p.pseg >= p.n_segs) return;
  const U2Seg sg = p.segs[p.pseg];
  const int rows = u2_chunk_rows(sg.N);
  const int slot = p.pidx % U2_STAGES;
  float* dst = p.ring + slot * U2_SLOT_FLOATS;
  if (lane == 0) {
    const uint32_t bytes = uint32_t(rows) * uint32_t(sg.N) * 4u;
    mbar_expect_tx(&p.full[slot], bytes);                       // every CTA expects the chunk on its own barrier
    if constexpr (U2_CLUSTER == 1) {
      bulk_g2s(dst, sg.w + size_t(p.pk) * sg.N, bytes, &p.full[slot], p.policy);
    } else if (uint32_t(p.pidx % U2_CLUSTER) == p.rank) {       // ... one CTA fetches it for the whole cluster
      if (p.pidx >= U2_STAGES) {                                // refill: all CTAs must have released the slot
        mbar_wait_cluster(&p.empty[slot], (p.ewaits >> slot) & 1u);
        p.ewaits ^= 1u << slot;
      }
      bulk_g2s_multicast(dst, sg.w + size_t(p.pk) * sg.N, bytes, &p.full[slot], uint16_t((1u << U2_CLUSTER) - 1), p.policy);
    }
  }
  p.pk += rows;
  p.pidx += 1;
  if (p.pk >= sg.K) {
    p.pk = 0;
    p.pseg += 1;
  }
}

constexpr float U2_LO_SCALE = 2048.f;   // lo is stored pre-scaled by 2^11 (see header)

// two fp32 -> (hi, lo) f16x2 pairs, element 0 in the low half
__device__ __forceinline__ void u2_split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = pack_half2(x0, x1);
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = pack_half2((x0 - h.x) * U2_LO_SCALE, (x1 - h.y) * U2_LO_SCALE);
}
__device__ __forceinline__ void u2_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ys[r][n] = act(bias[n] + sum_k xs[r][k] W[k][n]) (+ add[r][n]) with W = the NEXT segment of the schedule (K x N, N = 64 NT).
// Warp w owns the 8-column tiles w, w + 8, ..., w + 8 (NT - 1).  Ends with a CTA barrier (ys complete, xs reusable).
template <int NT>
__device__ __noinline__ void u2_linear(U2Pipe& p, int K, const float* xs, int ldx, const float* __restrict__ bias,
                                          float* ys, int ldy, bool relu, const float* add, int ldadd) {
  constexpr int N = NT * 64;
  constexpr int rows = u2_chunk_rows(N);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  float acc[NT][4], cor[NT][4];           // hi.hi products | (hi.lo + lo.hi) products, scaled by 2048
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    cor[i][0] = cor[i][1] = cor[i][2] = cor[i][3] = 0.f;
  }
  for (int k0 = 0; k0 < K; k0 += rows) {
    const int slot = p.cidx % U2_STAGES;
    mbar_wait(&p.full[slot], uint32_t(p.cidx / U2_STAGES) & 1u);
    const uint4* wsl = reinterpret_cast<const uint4*>(p.ring + slot * U2_SLOT_FLOATS);
#pragma unroll
    for (int ks = 0; ks < rows / 16; ++ks) {
      uint32_t ah[4], al[4];
      const float* xa = xs + k0 + ks * 16 + 2 * t;
      const float2 v0 = *reinterpret_cast<const float2*>(xa + g * ldx);
      const float2 v1 = *reinterpret_cast<const float2*>(xa + (g + 8) * ldx);
      const float2 v2 = *reinterpret_cast<const float2*>(xa + g * ldx + 8);
      const float2 v3 = *reinterpret_cast<const float2*>(xa + (g + 8) * ldx + 8);
      u2_split2(v0.x, v0.y, ah[0], al[0]);
      u2_split2(v1.x, v1.y, ah[1], al[1]);
      u2_split2(v2.x, v2.y, ah[2], al[2]);
      u2_split2(v3.x, v3.y, ah[3], al[3]);
      uint4 bw[NT];
#pragma unroll
      for (int i = 0; i < NT; ++i) bw[i] = wsl[(ks * (N / 8) + warp + U2_WARPS * i) * 32 + lane];
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        u2_mma(acc[i], ah, bw[i].x, bw[i].y);      // hi . hi
        u2_mma(cor[i], ah, bw[i].z, bw[i].w);      // hi . lo
        u2_mma(cor[i], al, bw[i].x, bw[i].y);      // lo . hi
      }
    }
    __syncthreads();                         // every warp of this CTA is done with the slot
    if (U2_CLUSTER > 1 && warp == 0) {
      // release the slot towards the CTA that will refill it (chunk cidx + STAGES), if that chunk exists
      const int nxt = p.cidx + U2_STAGES;
      // relaxed: the arrive publishes no data; this CTA's reads of the slot have completed (their values fed the MMAs above
      // and every warp passed the barrier).  A release here compiles to MEMBAR + ERRBAR on the one lane the whole CTA then
      // waits for at the next barrier: 26 % of the kernel's stall samples in the first cluster version (ncu r2).
      if (lane == 0 && nxt < p.n_chunks)
        mbar_arrive_cluster_relaxed(mapa_rank(smem_u32(&p.empty[slot]), uint32_t(nxt % U2_CLUSTER)));
    }
    p.cidx += 1;
    if (warp == 0) u2_issue(p, lane);
  }
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int n = (warp + U2_WARPS * i) * 8 + 2 * t;
    const float b0 = bias ? __ldg(bias + n) : 0.f, b1 = bias ? __ldg(bias + n + 1) : 0.f;
    float v0 = fmaf(cor[i][0], 1.f / U2_LO_SCALE, acc[i][0]) + b0, v1 = fmaf(cor[i][1], 1.f / U2_LO_SCALE, acc[i][1]) + b1;
    float v2 = fmaf(cor[i][2], 1.f / U2_LO_SCALE, acc[i][2]) + b0, v3 = fmaf(cor[i][3], 1.f / U2_LO_SCALE, acc[i][3]) + b1;
    if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
    if (add != nullptr) {
      v0 += add[g * ldadd + n]; v1 += add[g * ldadd + n + 1];
      v2 += add[(g + 8) * ldadd + n]; v3 += add[(g + 8) * ldadd + n + 1];
    }
    *reinterpret_cast<float2*>(ys + g * ldy + n) = make_float2(v0, v1);
    *reinterpret_cast<float2*>(ys + (g + 8) * ldy + n) = make_float2(v2, v3);
  }
  __syncthreads();
}

__device__ __forceinline__ void u2_linear_n(U2Pipe& p, int K, int N, const float* xs, int ldx, const float* bias, float* ys,
                                            int ldy, bool relu, const float* add = nullptr, int ldadd = 0) {
  switch (N) {
    case 128: u2_linear<2>(p, K, xs, ldx, bias, ys, ldy, relu, add, ldadd); break;
    case 256: u2_linear<4>(p, K, xs, ldx, bias, ys, ldy, relu, add, ldadd); break;
    case 384: u2_linear<6>(p, K, xs, ldx, bias, ys, ldy, relu, add, ldadd); break;
    default: u2_linear<8>(p, K, xs, ldx, bias, ys, ldy, relu, add, ldadd); break;
  }
}

// LayerNorm over the 128 features of each row: warp per row (2 rows per warp), two-pass variance.  ys may alias xs.
__device__ __forceinline__ void u2_layernorm(const float* xs, int ldx, float* ys, int ldy, const float* __restrict__ gam,
                                             const float* __restrict__ bet, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < U2_R; r += U2_WARPS) {
    const float4 v = *reinterpret_cast<const float4*>(xs + r * ldx + lane * 4);
    const float mean = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.f / SA_D);
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    const float rstd = rsqrtf(warp_sum(a * a + b * b + c * c + d * d) * (1.f / SA_D) + eps);
    const float4 gg = __ldg(reinterpret_cast<const float4*>(gam) + lane);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bet) + lane);
    *reinterpret_cast<float4*>(ys + r * ldy + lane * 4) =
        make_float4(a * rstd * gg.x + bb.x, b * rstd * gg.y + bb.y, c * rstd * gg.z + bb.z, d * rstd * gg.w + bb.w);
  }
  __syncthreads();
}

constexpr int U2_SMEM_FLOATS = 3 * U2_R * U2_LDA + 2 * U2_R * U2_LDB + U2_STAGES * U2_SLOT_FLOATS + 64;
constexpr int U2_SMEM = U2_SMEM_FLOATS * 4 + U2_MAX_SEGS * 16 + 2 * U2_STAGES * 8 + 64;

__global__ void __launch_bounds__(U2_THREADS, 1)
sa_update2_kernel(SaWeights w, int S, int chunks, int n_rows /* B*S */, int flags, const float* __restrict__ slots_in,
                  const float* __restrict__ partial, float* __restrict__ slots_out, int slots_out_stride /* per seq */,
                  float* __restrict__ pred_out, float* __restrict__ gvec) {
  extern __shared__ __align__(16) float sm2[];
  float* cur = sm2;                          // [16][136] slots (previous, then new)
  float* t0 = cur + U2_R * U2_LDA;           // [16][132]
  float* t1 = t0 + U2_R * U2_LDA;            // [16][132]
  float* big0 = t1 + U2_R * U2_LDA;          // [16][516]
  float* big1 = big0 + U2_R * U2_LDB;        // [16][516]
  float* ring = big1 + U2_R * U2_LDB;        // 3 x 8704
  float* s_am = ring + U2_STAGES * U2_SLOT_FLOATS;   // [16][2] A, Mw per row (+ padding)
  U2Seg* segs = reinterpret_cast<U2Seg*>(s_am + 64);
  uint64_t* full = reinterpret_cast<uint64_t*>(segs + U2_MAX_SEGS);
  uint64_t* empty = full + U2_STAGES;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = SA_D;
  const int RPC = (U2_R / S) * S;            // a CTA owns whole sequences: 16 rows for 8 slots, 10 for 10 slots
  const int row0 = blockIdx.x * RPC;
  const int PART = sa_part(S);
  n_rows = min(n_rows, row0 + RPC);          // rows of the next CTA are not ours

  // ---- weight schedule (consumption order below must match) + barriers
  if (tid == 0) {
    int n = 0;
    // segments of a stream follow each other: K * N 32-bit words each
    auto push = [&](const float*& base, int K, int N) {
      segs[n].w = base; segs[n].K = K; segs[n].N = N; ++n;
      base += size_t(K) * N;
    };
    if (flags & U2_DO_C) {
      const float* b = w.stream_c;
      push(b, D, D);                 // wv_t
      push(b, D, 3 * D);             // w_ih_t
      push(b, D, 3 * D);             // w_hh_t
      push(b, D, w.mlp_hidden);      // w1_t
      push(b, w.mlp_hidden, D);      // w2_t
    }
    if (flags & U2_DO_T) {
      const float* b = w.stream_t;
      push(b, D, D);                 // t_wq_t
      push(b, D, D);                 // t_wk_t
      push(b, D, D);                 // t_wv_t
      push(b, D, D);                 // t_wo_t
      push(b, D, w.t_hidden);        // t_w1_t
      push(b, w.t_hidden, D);        // t_w2_t
    }
    if (flags & U2_DO_A) {
      const float* b = w.stream_a;
      push(b, D, D);                 // wq_t
      push(b, D, D);                 // wk (as stored: in-major for q -> qt)
    }
    int nch = 0;
    for (int i = 0; i < n; ++i) nch += segs[i].K / u2_chunk_rows(segs[i].N);
    s_am[62] = __int_as_float(n);
    s_am[63] = __int_as_float(nch);
    for (int s = 0; s < U2_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], U2_CLUSTER);      // one release per CTA of the cluster
    }
    fence_barrier_init();
  }
  if (U2_CLUSTER > 1) cluster_sync_all();    // barriers of every CTA are initialised before any peer copies / arrives
  else __syncthreads();
  U2Pipe p;
  p.segs = segs;
  p.n_segs = __float_as_int(s_am[62]);
  p.ring = ring;
  p.full = full;
  p.empty = empty;
  p.cidx = 0;
  p.pseg = 0; p.pk = 0; p.pidx = 0;
  p.n_chunks = __float_as_int(s_am[63]);
  p.rank = U2_CLUSTER > 1 ? cluster_ctarank() : 0;
  p.ewaits = 0;
  p.policy = l2_policy_evict_last();
  if (warp == 0) {
    for (int s = 0; s < U2_STAGES; ++s) u2_issue(p, lane);      // the weights do not depend on the previous kernel
  }

  // ---- slots_in [rows][D] -> cur
  for (int e = tid; e < U2_R * (D / 4); e += U2_THREADS) {
    const int r = e / (D / 4), c4 = e % (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < n_rows) v = __ldg(reinterpret_cast<const float4*>(slots_in + size_t(row0 + r) * D) + c4);
    *reinterpret_cast<float4*>(cur + r * U2_LDA + c4 * 4) = v;
  }
  __syncthreads();

  if (flags & U2_DO_C) {
    // ---- A_i = sum_j a_ij, Mw_i = sum_j w_ij mu_j summed over the location chunks (one thread per row and quantity)
    if (tid < 2 * U2_R) {
      const int r = tid >> 1, which = tid & 1, row = row0 + r;
      float acc = 0.f;
      if (row < n_rows) {
        const int b = row / S, i = row % S;
        const float* pp = partial + size_t(b) * chunks * PART + S * D + which * S + i;
        for (int c = 0; c < chunks; ++c) acc += pp[size_t(c) * PART];
      }
      s_am[r * 2 + which] = acc;
    }
    __syncthreads();
    // ---- weighted mean: uhat[r][f] = (gamma_f (U - Mw) + beta_f A) / A
    for (int e = tid; e < U2_R * (D / 4); e += U2_THREADS) {
      const int r = e / (D / 4), c4 = e % (D / 4), row = row0 + r;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < n_rows) {
        const int b = row / S, i = row % S;
        const float* pp = partial + size_t(b) * chunks * PART + i * D + c4 * 4;
        float4 U = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < chunks; ++c) {
          const float4 u = __ldg(reinterpret_cast<const float4*>(pp + size_t(c) * PART));
          U.x += u.x; U.y += u.y; U.z += u.z; U.w += u.w;
        }
        const float A = s_am[r * 2], Mw = s_am[r * 2 + 1];
        const float4 gg = __ldg(reinterpret_cast<const float4*>(w.ln_in_g) + c4);
        const float4 bb = __ldg(reinterpret_cast<const float4*>(w.ln_in_b) + c4);
        val = make_float4((gg.x * (U.x - Mw) + bb.x * A) / A, (gg.y * (U.y - Mw) + bb.y * A) / A,
                          (gg.z * (U.z - Mw) + bb.z * A) / A, (gg.w * (U.w - Mw) + bb.w * A) / A);
      }
      *reinterpret_cast<float4*>(t0 + r * U2_LDA + c4 * 4) = val;
    }
    __syncthreads();
    u2_linear_n(p, D, D, t0, U2_LDA, w.bv, t1, U2_LDA, false);                    // updates = Wv uhat + bv      -> t1
    u2_linear_n(p, D, 3 * D, t1, U2_LDA, w.b_ih, big0, U2_LDB, false);            // gi                          -> big0
    u2_linear_n(p, D, 3 * D, cur, U2_LDA, w.b_hh, big1, U2_LDB, false);           // gh (hidden = slots_prev)    -> big1
    for (int e = tid; e < U2_R * D; e += U2_THREADS) {                            // GRUCell, gate order r, z, n
      const int r = e / D, k = e % D;
      const float* gi = big0 + r * U2_LDB;
      const float* gh = big1 + r * U2_LDB;
      const float rg = 1.f / (1.f + __expf(-(gi[k] + gh[k])));
      const float zg = 1.f / (1.f + __expf(-(gi[D + k] + gh[D + k])));
      const float ng = tanhf(gi[2 * D + k] + rg * gh[2 * D + k]);
      const float h = cur[r * U2_LDA + k];
      cur[r * U2_LDA + k] = (1.f - zg) * ng + zg * h;
    }
    __syncthreads();
    u2_layernorm(cur, U2_LDA, t0, U2_LDA, w.ln_mlp_g, w.ln_mlp_b, w.ln_eps_sa);
    u2_linear_n(p, D, w.mlp_hidden, t0, U2_LDA, w.b1, big0, U2_LDB, true);
    u2_linear_n(p, w.mlp_hidden, D, big0, U2_LDB, w.b2, cur, U2_LDA, false, cur, U2_LDA);   // slots + MLP(LN(slots))
    if (slots_out) {
      for (int e = tid; e < U2_R * (D / 4); e += U2_THREADS) {
        const int r = e / (D / 4), c4 = e % (D / 4), row = row0 + r;
        if (row < n_rows)
          *reinterpret_cast<float4*>(slots_out + size_t(row / S) * slots_out_stride + size_t(row % S) * D + c4 * 4) =
              *reinterpret_cast<const float4*>(cur + r * U2_LDA + c4 * 4);
      }
    }
  }

  if (flags & U2_DO_T) {
    // ---- post-norm TransformerBlock: y = LN(MHSA(x) + x); z = LN(MLP(y) + y)   (attention.py:387-395)
    float* qb = t0;                 // q  [16][132]
    float* kb = t1;                 // k
    float* vb = big0;               // v  (pitch U2_LDB)
    float* att = big1;              // attention output: columns 0..127 of big1 (pitch U2_LDB)
    float* sc = big1 + 128;         // scores: columns 128 + h*16 + j of the same rows (H <= 8 heads x 16 key slots)
    u2_linear_n(p, D, D, cur, U2_LDA, nullptr, qb, U2_LDA, false);
    u2_linear_n(p, D, D, cur, U2_LDA, nullptr, kb, U2_LDA, false);
    u2_linear_n(p, D, D, cur, U2_LDA, nullptr, vb, U2_LDB, false);
    const int H = w.t_heads, dh = D / H;
    const float scl = rsqrtf(float(dh));
    // phase 1: (query row r, head h, key j) -> score
    for (int e = tid; e < U2_R * H * 16; e += U2_THREADS) {
      const int j = e & 15, h = (e >> 4) % H, r = e / (16 * H);
      if (r >= RPC || j >= S) continue;
      const int rs = (r / S) * S;
      const float4* qv = reinterpret_cast<const float4*>(qb + r * U2_LDA + h * dh);
      const float4* kv = reinterpret_cast<const float4*>(kb + (rs + j) * U2_LDA + h * dh);
      float d = 0.f;
      for (int c = 0; c < dh / 4; ++c) {
        const float4 a = qv[c], b = kv[c];
        d = fmaf(a.x, b.x, d); d = fmaf(a.y, b.y, d); d = fmaf(a.z, b.z, d); d = fmaf(a.w, b.w, d);
      }
      sc[r * U2_LDB + h * 16 + j] = d * scl;
    }
    __syncthreads();
    // phase 2: (row r, channel c): softmax over the S keys of the channel's head, weighted sum of V
    for (int e = tid; e < U2_R * D; e += U2_THREADS) {
      const int r = e / D, c = e % D, h = c / dh;
      float o = 0.f;
      if (r < RPC) {
        const int rs = (r / S) * S;
        const float* srow = sc + r * U2_LDB + h * 16;
        float mx = -1e30f;
        for (int j = 0; j < S; ++j) mx = fmaxf(mx, srow[j]);
        float den = 0.f;
        for (int j = 0; j < S; ++j) {
          const float pj = __expf(srow[j] - mx);
          den += pj;
          o = fmaf(pj, vb[(rs + j) * U2_LDB + c], o);
        }
        o /= den;
      }
      att[r * U2_LDB + c] = o;
    }
    __syncthreads();
    u2_linear_n(p, D, D, att, U2_LDB, nullptr, t0, U2_LDA, false, cur, U2_LDA);              // MHSA(x) + x   -> t0
    u2_layernorm(t0, U2_LDA, t1, U2_LDA, w.t_ln1_g, w.t_ln1_b, w.ln_eps_tf);                 // y             -> t1
    u2_linear_n(p, D, w.t_hidden, t1, U2_LDA, w.t_b1, big0, U2_LDB, true);
    u2_linear_n(p, w.t_hidden, D, big0, U2_LDB, w.t_b2, t0, U2_LDA, false, t1, U2_LDA);      // MLP(y) + y    -> t0
    u2_layernorm(t0, U2_LDA, cur, U2_LDA, w.t_ln2_g, w.t_ln2_b, w.ln_eps_tf);                // z             -> cur
    if (pred_out) {
      for (int e = tid; e < U2_R * (D / 4); e += U2_THREADS) {
        const int r = e / (D / 4), c4 = e % (D / 4);
        if (row0 + r < n_rows)
          *reinterpret_cast<float4*>(pred_out + size_t(row0 + r) * D + c4 * 4) =
              *reinterpret_cast<const float4*>(cur + r * U2_LDA + c4 * 4);
      }
    }
  }

  if (flags & U2_DO_A) {
    // ---- next pass: q = Wq LN(slots) + bq ; qt = Wk^T q ; g = scale*qt*gamma ; sg = sum g ; cb = scale*(qt.beta + q.bk)
    u2_layernorm(cur, U2_LDA, t0, U2_LDA, w.ln_slot_g, w.ln_slot_b, w.ln_eps_sa);
    u2_linear_n(p, D, D, t0, U2_LDA, w.bq, t1, U2_LDA, false);                    // q  -> t1
    u2_linear_n(p, D, D, t1, U2_LDA, nullptr, t0, U2_LDA, false);                 // qt -> t0 (Wk as stored [d][f] is in-major)
    for (int r = warp; r < U2_R; r += U2_WARPS) {                                 // warp per row
      const int row = row0 + r;
      const float4 qt = *reinterpret_cast<const float4*>(t0 + r * U2_LDA + lane * 4);
      const float4 qq = *reinterpret_cast<const float4*>(t1 + r * U2_LDA + lane * 4);
      const float4 gg = __ldg(reinterpret_cast<const float4*>(w.ln_in_g) + lane);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(w.ln_in_b) + lane);
      const float4 bk = __ldg(reinterpret_cast<const float4*>(w.bk) + lane);
      float sgv = qt.x * gg.x + qt.y * gg.y + qt.z * gg.z + qt.w * gg.w;
      float cbv = qt.x * bb.x + qt.y * bb.y + qt.z * bb.z + qt.w * bb.w + qq.x * bk.x + qq.y * bk.y + qq.z * bk.z + qq.w * bk.w;
      sgv = warp_sum(sgv);
      cbv = warp_sum(cbv);
      if (row < n_rows) {
        const int b = row / S, i = row % S;
        float* gv = gvec + size_t(b) * PART;
        *reinterpret_cast<float4*>(gv + i * D + lane * 4) =
            make_float4(w.scale * qt.x * gg.x, w.scale * qt.y * gg.y, w.scale * qt.z * gg.z, w.scale * qt.w * gg.w);
        if (lane == 0) {
          gv[S * D + i] = w.scale * sgv;
          gv[S * D + S + i] = w.scale * cbv;
        }
      }
    }
  }
  if (U2_CLUSTER > 1) cluster_sync_all();   // no CTA leaves while a peer may still multicast into its ring / arrive on its barriers
}

int launch_update2(const SaWeights& w, int S, int chunks, int B, int flags, const float* slots_in, const float* partial,
                   float* slots_out, int out_stride, float* pred_out, float* gvec, cudaStream_t stream) {
  static SmemAttrOnce attr_once;
  TOCVP_TRY(ensure_smem_attr(attr_once, sa_update2_kernel, U2_SMEM));
  const int rows = B * S;
  const int rpc = (U2_R / S) * S;
  int grid = (rows + rpc - 1) / rpc;
  grid = (grid + U2_CLUSTER - 1) / U2_CLUSTER * U2_CLUSTER;   // padding CTAs own no rows but take part in the weight pipeline
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(U2_THREADS);
  cfg.dynamicSmemBytes = U2_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = U2_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TOCVP_CUDA(cudaLaunchKernelEx(&cfg, sa_update2_kernel, w, S, chunks, rows, flags, slots_in, partial, slots_out, out_stride,
                                pred_out, gvec));
  count_launch();
  return TOCVP_OK;
}

// shapes the second-version kernel is instantiated for (everything else takes the first version)
bool update2_supported(const SaWeights& w, int flags) {
  auto ok_hidden = [](int h) { return h == 128 || h == 256 || h == 384 || h == 512; };
  if (((flags & U2_DO_C) && !w.stream_c) || ((flags & U2_DO_T) && !w.stream_t) || ((flags & U2_DO_A) && !w.stream_a)) return false;
  return ok_hidden(w.mlp_hidden) && (w.t_heads == 0 || (ok_hidden(w.t_hidden) && w.t_heads <= 8 && SA_D % w.t_heads == 0 &&
                                                      (SA_D / w.t_heads) % 4 == 0));
}

}  // namespace tocvp
