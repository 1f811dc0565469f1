// SAVi corrector: SlotAttention (reference src/models/Blocks/attention.py:67-112) and the post-norm transition
// TransformerBlock (attention.py:387-395) as bandwidth-oriented fused kernels.
//
// K/V are never materialised.  With x_j the input feature of location j, LN(x_j) = (x_j - mu_j) * rstd_j * gamma + beta:
//     q_i . k_j   = rstd_j * (g_i . x_j - mu_j * sum(g_i)) + (qt_i . beta + q_i . bk),     qt_i = Wk^T q_i,  g_i = qt_i * gamma
//     updates_i   = Wv * ( gamma * (sum_j w_ij x_j - sum_j w_ij mu_j) + beta * A_i ) / A_i + bv
// with a_ij = softmax_i(scale * q_i.k_j) + eps, A_i = sum_j a_ij, w_ij = a_ij * rstd_j  (sum_j a_ij / A_i = 1, so bv passes
// through unchanged).  One iteration = one streaming pass over the features (`sa_stream_kernel`, the HBM-bound part: the
// 8-way slot softmax lives in registers, location-axis sums use a warp reduce-scatter) + one small per-slot kernel
// (`sa_update_kernel`: weighted-mean finalisation, V projection, GRUCell, LayerNorm, residual MLP, and the next
// iteration's query / g vectors, or the transition block after the last iteration).  All of it fp32.
#include "host_util.h"
#include "ptx.cuh"
#include "slot_attention.h"

namespace tocvp {

int sa_stream_tc(const __half* feats, size_t seq_stride, int B, int N, const float* gvec, float* partial, float ln_eps,
                 float attn_eps, cudaStream_t stream);


template <typename T>
__device__ __forceinline__ float4 ldx4(const T* p);
template <>
__device__ __forceinline__ float4 ldx4<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <>
__device__ __forceinline__ float4 ldx4<__half>(const __half* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// Reduce-scatter of N per-lane values over the 32 lanes of a warp: afterwards v[0] of lane L is the warp-wide
// total of value index (L >> log2(32/N))  (N = 32: index L).  N-1 shuffles instead of 5N.
template <int N>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[N], int lane) {
#pragma unroll
  for (int n = N, off = 16; n > 1; n >>= 1, off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < n / 2; ++k) {
      const float send = upper ? v[k] : v[k + n / 2];
      const float keep = upper ? v[k + n / 2] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Streaming pass.  grid = (SA_CHUNKS, B), 256 threads.  Lane l of every warp owns channels 4l..4l+3;
// a warp handles 4 locations per step.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256, 2)
sa_stream_kernel(const T* __restrict__ feats, size_t seq_stride, int N, const float* __restrict__ gvec, float* __restrict__ partial,
                 float ln_eps, float attn_eps) {
  __shared__ __align__(16) float s_w[8][4][8];            // per warp: w[l][i]
  __shared__ __align__(16) float s_red[8][SA_S][SA_D];    // cross-warp reduction of the update accumulators (32 KB)
  __shared__ float s_am[8][2][SA_S];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int LC = N / SA_CHUNKS;
  const T* x = feats + size_t(b) * seq_stride + size_t(chunk) * LC * SA_D + lane * 4;
  const float* gv = gvec + size_t(b) * SA_GVEC;

  float g[SA_S][4];
#pragma unroll
  for (int i = 0; i < SA_S; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(gv + i * SA_D + lane * 4);
    g[i][0] = t.x; g[i][1] = t.y; g[i][2] = t.z; g[i][3] = t.w;
  }
  const int my_i = lane & 7, my_l = lane >> 3;
  const float sg = gv[SA_S * SA_D + my_i];
  const float cb = gv[SA_S * SA_D + SA_S + my_i];

  float acc[SA_S][4];
#pragma unroll
  for (int i = 0; i < SA_S; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  float a_sum = 0.f, mw_sum = 0.f;

  const int groups = LC / 4;
  float4 xn[4];
  if (warp < groups) {
#pragma unroll
    for (int l = 0; l < 4; ++l) xn[l] = ldx4<T>(x + size_t(warp * 4 + l) * SA_D);
  }
  for (int gi = warp; gi < groups; gi += 8) {
    float4 xv[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) xv[l] = xn[l];
    if (gi + 8 < groups) {  // software prefetch of the next group's features
#pragma unroll
      for (int l = 0; l < 4; ++l) xn[l] = ldx4<T>(x + size_t((gi + 8) * 4 + l) * SA_D);
    }
    float v[32], st[8];
#pragma unroll
    for (int l = 0; l < 4; ++l) {
#pragma unroll
      for (int i = 0; i < SA_S; ++i)
        v[l * 8 + i] = g[i][0] * xv[l].x + g[i][1] * xv[l].y + g[i][2] * xv[l].z + g[i][3] * xv[l].w;
      st[l] = (xv[l].x + xv[l].y) + (xv[l].z + xv[l].w);
      st[4 + l] = xv[l].x * xv[l].x + xv[l].y * xv[l].y + xv[l].z * xv[l].z + xv[l].w * xv[l].w;
    }
    warp_reduce_scatter<32>(v, lane);   // lane (l,i) = 8l+i now holds g_i . x_l
    warp_reduce_scatter<8>(st, lane);   // lane L holds stat index (L>>2)&7 summed over lanes with equal L&3 ...
    float sv = st[0];
    sv += __shfl_xor_sync(0xffffffffu, sv, 2);
    sv += __shfl_xor_sync(0xffffffffu, sv, 1);
    const float sum = __shfl_sync(0xffffffffu, sv, my_l * 4);
    const float ssq = __shfl_sync(0xffffffffu, sv, (4 + my_l) * 4);
    const float mu = sum * (1.f / SA_D);
    const float var = fmaxf(ssq * (1.f / SA_D) - mu * mu, 0.f);
    const float rstd = rsqrtf(var + ln_eps);
    const float d = rstd * (v[0] - mu * sg) + cb;            // already multiplied by the attention scale
    float m = d;
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
    const float e = __expf(d - m);
    float s = e;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    const float a = e / s + attn_eps;                        // softmax over SLOTS, + eps (attention.py:100)
    const float w = a * rstd;
    a_sum += a;
    mw_sum += w * mu;
    __syncwarp();
    s_w[warp][my_l][my_i] = w;
    __syncwarp();
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const float4 w0 = *reinterpret_cast<const float4*>(&s_w[warp][l][0]);
      const float4 w1 = *reinterpret_cast<const float4*>(&s_w[warp][l][4]);
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < SA_S; ++i) {
        acc[i][0] += ww[i] * xv[l].x;
        acc[i][1] += ww[i] * xv[l].y;
        acc[i][2] += ww[i] * xv[l].z;
        acc[i][3] += ww[i] * xv[l].w;
      }
    }
  }
  // ---- block reduction -> partial[b][chunk]
  a_sum += __shfl_xor_sync(0xffffffffu, a_sum, 8);
  a_sum += __shfl_xor_sync(0xffffffffu, a_sum, 16);
  mw_sum += __shfl_xor_sync(0xffffffffu, mw_sum, 8);
  mw_sum += __shfl_xor_sync(0xffffffffu, mw_sum, 16);
  if (lane < 8) {
    s_am[warp][0][lane] = a_sum;
    s_am[warp][1][lane] = mw_sum;
  }
#pragma unroll
  for (int i = 0; i < SA_S; ++i)
    *reinterpret_cast<float4*>(&s_red[warp][i][lane * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  __syncthreads();
  float* out = partial + (size_t(b) * SA_CHUNKS + chunk) * SA_PART;
  for (int e = threadIdx.x; e < SA_S * SA_D; e += 256) {
    float t = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) t += (&s_red[w8][0][0])[e];
    out[e] = t;
  }
  if (threadIdx.x < 16) {
    float t = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) t += (&s_am[w8][0][0])[threadIdx.x];
    out[SA_S * SA_D + threadIdx.x] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// Generic streaming pass (any slot count 4 <= S <= 11, any N, fp32 or f16 features): the CLIPort / ExtendedDINOSAUR shape
// (10 slots over 81..576 patch tokens, reference src/models/ExtendedDINOSAUR.py:188-194) and any shape the two
// specialised kernels do not cover.  grid = (chunks, B), 256 threads; one warp per location, lane l owns channels
// 4l..4l+3, every lane holds the full S-way softmax.  Same folded-K/V algebra and the same partial[] layout.
// ------------------------------------------------------------------------------------------------
template <typename T, int S>
__global__ void __launch_bounds__(256)
sa_stream_generic_kernel(const T* __restrict__ feats, size_t seq_stride, int N, int chunks, const float* __restrict__ gvec,
                         float* __restrict__ partial, float ln_eps, float attn_eps) {
  constexpr int PART = sa_part(S);
  __shared__ __align__(16) float s_red[8][S][SA_D];
  __shared__ float s_am[8][2][S];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int LC = (N + chunks - 1) / chunks;
  const int l0 = chunk * LC, l1 = min(N, l0 + LC);
  const T* x = feats + size_t(b) * seq_stride + lane * 4;
  const float* gv = gvec + size_t(b) * PART;
  float g[S][4], sg[S], cb[S], acc[S][4], a_acc[S], mw_acc[S];
#pragma unroll
  for (int i = 0; i < S; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(gv + i * SA_D + lane * 4);
    g[i][0] = t.x; g[i][1] = t.y; g[i][2] = t.z; g[i][3] = t.w;
    sg[i] = gv[S * SA_D + i];
    cb[i] = gv[S * SA_D + S + i];
    acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    a_acc[i] = 0.f;
    mw_acc[i] = 0.f;
  }
  for (int l = l0 + warp; l < l1; l += 8) {
    const float4 xv = ldx4<T>(x + size_t(l) * SA_D);
    const float sum = warp_sum((xv.x + xv.y) + (xv.z + xv.w));
    const float ssq = warp_sum(xv.x * xv.x + xv.y * xv.y + xv.z * xv.z + xv.w * xv.w);
    const float mu = sum * (1.f / SA_D);
    const float rstd = rsqrtf(fmaxf(ssq * (1.f / SA_D) - mu * mu, 0.f) + ln_eps);
    float d[S];
    float m = -1e30f;
#pragma unroll
    for (int i = 0; i < S; ++i) {
      const float dot = warp_sum(g[i][0] * xv.x + g[i][1] * xv.y + g[i][2] * xv.z + g[i][3] * xv.w);
      d[i] = rstd * (dot - mu * sg[i]) + cb[i];                // already multiplied by the attention scale
      m = fmaxf(m, d[i]);
    }
    float den = 0.f;
#pragma unroll
    for (int i = 0; i < S; ++i) {
      d[i] = __expf(d[i] - m);
      den += d[i];
    }
    const float inv = 1.f / den;
#pragma unroll
    for (int i = 0; i < S; ++i) {
      const float a = d[i] * inv + attn_eps;                   // softmax over SLOTS, + eps (attention.py:100)
      const float w = a * rstd;
      a_acc[i] += a;
      mw_acc[i] += w * mu;
      acc[i][0] += w * xv.x; acc[i][1] += w * xv.y; acc[i][2] += w * xv.z; acc[i][3] += w * xv.w;
    }
  }
#pragma unroll
  for (int i = 0; i < S; ++i) {
    *reinterpret_cast<float4*>(&s_red[warp][i][lane * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (lane == 0) {
      s_am[warp][0][i] = a_acc[i];
      s_am[warp][1][i] = mw_acc[i];
    }
  }
  __syncthreads();
  float* out = partial + (size_t(b) * chunks + chunk) * PART;
  for (int e = threadIdx.x; e < S * SA_D; e += 256) {
    float t = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) t += (&s_red[w8][0][0])[e];
    out[e] = t;
  }
  if (threadIdx.x < 2 * S) {
    float t = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) t += (&s_am[w8][0][0])[threadIdx.x];
    out[S * SA_D + threadIdx.x] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// Per-slot update kernel.  Each CTA owns R = 16 slot rows (2 sequences); activations live in smem
// transposed as [feature][row] so that a thread computing one output column reads the rows as float4.
// ------------------------------------------------------------------------------------------------
constexpr int UP_R = 16;
constexpr int UP_THREADS = 512;   // 16 warps: the update kernel is latency-bound (weight fragments from L2), not issue-bound
constexpr int UP_WARPS = UP_THREADS / 32;

using SaWeights = tocvp_sa_weights;  // include/tocvp.h

// 3xTF32 split: x = hi + lo with hi = x truncated to tf32 (low 13 mantissa bits cleared) and lo = x - hi (exact in fp32;
// the tensor core reads only the tf32 bits of it); hi*hi + lo*hi + hi*lo keeps ~20 mantissa bits, i.e. fp32-level accuracy
// for the recurrent slot state, on the (legacy-path) tensor cores.  Two instructions.  The first version used
// cvt.rna.tf32.f32 for both halves: sm_100 has no instruction for it, each expands to ~7 FSETP / SEL / LOP3 / IADD3, and they
// were HALF of the update kernel's 42 M warp instructions (ncu opcode histogram: FSETP 18 %, IADD3 13 %, LOP3 10 %, SEL 9 %,
// HMMA 6 %).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ys[o][r] = act(bias[o] + sum_k Wt[k][o] * xs[k][r])  for r < UP_R = 16 = one mma M tile.
// mma.sync m16n8k8 (3xTF32): warp w owns the 8-column tiles w, w+16, ...; the A fragment of a k-step (the 16 activation rows)
// is split once and reused for all of the warp's column tiles.  ~5x fewer instructions than the SIMT loop it replaces
// (kept below as linear_T_simt) at the same fp32-level accuracy.
template <int NT>
__device__ __forceinline__ void linear_T_mma(const float* __restrict__ Wt, const float* __restrict__ bias, int IN, int OUT,
                                             const float* xs, float* ys, bool relu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  float acc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll 4
  for (int k0 = 0; k0 < IN; k0 += 8) {
    uint32_t ah[4], al[4];
    split_tf32(xs[(k0 + t) * UP_R + g], ah[0], al[0]);
    split_tf32(xs[(k0 + t) * UP_R + g + 8], ah[1], al[1]);
    split_tf32(xs[(k0 + t + 4) * UP_R + g], ah[2], al[2]);
    split_tf32(xs[(k0 + t + 4) * UP_R + g + 8], ah[3], al[3]);
    float bw[NT][2];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const int n = (warp + UP_WARPS * i) * 8 + g;
      bw[i][0] = __ldg(Wt + size_t(k0 + t) * OUT + n);
      bw[i][1] = __ldg(Wt + size_t(k0 + t + 4) * OUT + n);
    }
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      uint32_t bh0, bl0, bh1, bl1;
      split_tf32(bw[i][0], bh0, bl0);
      split_tf32(bw[i][1], bh1, bl1);
      mma_tf32(acc[i], al, bh0, bh1);
      mma_tf32(acc[i], ah, bl0, bl1);
      mma_tf32(acc[i], ah, bh0, bh1);
    }
  }
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int n = (warp + UP_WARPS * i) * 8 + 2 * t;
    const float b0 = bias ? __ldg(bias + n) : 0.f, b1 = bias ? __ldg(bias + n + 1) : 0.f;
    float v0 = acc[i][0] + b0, v1 = acc[i][1] + b1, v2 = acc[i][2] + b0, v3 = acc[i][3] + b1;
    if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
    ys[n * UP_R + g] = v0;
    ys[(n + 1) * UP_R + g] = v1;
    ys[n * UP_R + g + 8] = v2;
    ys[(n + 1) * UP_R + g + 8] = v3;
  }
  __syncthreads();
}

__device__ void linear_T_simt(const float* __restrict__ Wt, const float* __restrict__ bias, int IN, int OUT,
                              const float* xs, float* ys, bool relu);

// simt: tocvp_tuning.corrector_mode = 1 (per call): first-version SIMT matvecs (A/B, tests)
__device__ __forceinline__ void linear_T(bool simt, const float* __restrict__ Wt, const float* __restrict__ bias, int IN, int OUT,
                                         const float* xs, float* ys, bool relu) {
  if (simt || (OUT != 128 && OUT != 256 && OUT != 384 && OUT != 512) || IN % 8 != 0) {
    linear_T_simt(Wt, bias, IN, OUT, xs, ys, relu);
    return;
  }
  switch (OUT) {   // column tiles per warp = OUT / (8 * UP_WARPS)
    case 128: linear_T_mma<128 / (8 * UP_WARPS)>(Wt, bias, IN, OUT, xs, ys, relu); break;
    case 256: linear_T_mma<256 / (8 * UP_WARPS)>(Wt, bias, IN, OUT, xs, ys, relu); break;
    case 384: linear_T_mma<384 / (8 * UP_WARPS)>(Wt, bias, IN, OUT, xs, ys, relu); break;
    default: linear_T_mma<512 / (8 * UP_WARPS)>(Wt, bias, IN, OUT, xs, ys, relu); break;
  }
}

// First version: ys[o][r] = act(bias[o] + sum_k Wt[k][o] * xs[k][r]) with one thread per output column.
__device__ void linear_T_simt(const float* __restrict__ Wt, const float* __restrict__ bias, int IN, int OUT,
                              const float* xs, float* ys, bool relu) {
  const int tid = threadIdx.x;
  if (OUT != 128 || UP_THREADS != 256) {
    for (int o = tid; o < OUT; o += UP_THREADS) {
      float acc[UP_R];
      const float bv = bias ? bias[o] : 0.f;
#pragma unroll
      for (int r = 0; r < UP_R; ++r) acc[r] = bv;
#pragma unroll 16
      for (int k = 0; k < IN; ++k) {
        const float w = __ldg(Wt + size_t(k) * OUT + o);
        const float4* xr = reinterpret_cast<const float4*>(xs + k * UP_R);
#pragma unroll
        for (int r4 = 0; r4 < UP_R / 4; ++r4) {
          const float4 xx = xr[r4];
          acc[r4 * 4 + 0] += w * xx.x; acc[r4 * 4 + 1] += w * xx.y;
          acc[r4 * 4 + 2] += w * xx.z; acc[r4 * 4 + 3] += w * xx.w;
        }
      }
#pragma unroll
      for (int r = 0; r < UP_R; ++r) ys[o * UP_R + r] = relu ? fmaxf(acc[r], 0.f) : acc[r];
    }
  } else {  // OUT == 128: two row-halves per column
    const int o = tid % OUT, half = tid / OUT;  // half in {0,1}
    constexpr int RP = UP_R / 2;
    float acc[RP];
    const float bv = bias ? bias[o] : 0.f;
#pragma unroll
    for (int r = 0; r < RP; ++r) acc[r] = bv;
#pragma unroll 16
    for (int k = 0; k < IN; ++k) {
      const float w = __ldg(Wt + size_t(k) * OUT + o);
      const float4* xr = reinterpret_cast<const float4*>(xs + k * UP_R + half * RP);
#pragma unroll
      for (int r4 = 0; r4 < RP / 4; ++r4) {
        const float4 xx = xr[r4];
        acc[r4 * 4 + 0] += w * xx.x; acc[r4 * 4 + 1] += w * xx.y;
        acc[r4 * 4 + 2] += w * xx.z; acc[r4 * 4 + 3] += w * xx.w;
      }
    }
#pragma unroll
    for (int r = 0; r < RP; ++r) ys[o * UP_R + half * RP + r] = relu ? fmaxf(acc[r], 0.f) : acc[r];
  }
  __syncthreads();
}

// LayerNorm over the feature axis of xs[D][R] -> ys[D][R]: one warp per row (2 rows per warp), lane l owns features
// l, l+32, ... (two-pass variance).  The first version used one THREAD per row (16 of 256 threads busy, 3 serial passes
// over 128 features) and was, with the other 16-thread sections, what the update kernel actually spent its time in once
// the matrix products moved to the tensor cores.
__device__ void layernorm_T(const float* xs, float* ys, const float* __restrict__ g, const float* __restrict__ b,
                            float eps, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < UP_R; r += UP_THREADS / 32) {
    float v[4];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = xs[(lane + 32 * j) * UP_R + r];
      s += v[j];
    }
    const float mean = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] -= mean;
      q += v[j] * v[j];
    }
    const float rstd = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
    for (int j = 0; j < 4; ++j) ys[(lane + 32 * j) * UP_R + r] = v[j] * rstd * g[lane + 32 * j] + b[lane + 32 * j];
  }
  __syncthreads();
}

// flags
constexpr int UP_DO_C = 1;   // finish the iteration from the streaming partials (weighted mean, V, GRU, MLP)
constexpr int UP_DO_T = 2;   // apply the transition block to the result -> pred_out
constexpr int UP_DO_A = 4;   // emit g / sg / cb for the next streaming pass (from the result, or from pred if DO_T)

__global__ void __launch_bounds__(UP_THREADS, 1)
sa_update_kernel(SaWeights w, int S, int chunks, int n_rows /* B*S */, int flags, const float* __restrict__ slots_in,
                 const float* __restrict__ partial, float* __restrict__ slots_out, int slots_out_stride /* per seq */,
                 float* __restrict__ pred_out, float* __restrict__ gvec, int simt_i) {
  const bool simt = simt_i != 0;
  extern __shared__ float sm[];
  float* cur = sm;                        // [128][R]  slots (prev, then new)
  float* t0 = cur + SA_D * UP_R;          // [128][R]
  float* t1 = t0 + SA_D * UP_R;           // [128][R]
  float* big0 = t1 + SA_D * UP_R;         // [512][R]
  float* big1 = big0 + 512 * UP_R;        // [512][R]
  // a CTA owns whole sequences: RPC = floor(16 / S) * S slot rows (16 for 8 slots, 10 for 10 slots); rows >= RPC are padding
  const int RPC = (UP_R / S) * S;
  const int row0 = blockIdx.x * RPC;
  const int tid = threadIdx.x;
  const int D = SA_D;
  const int PART = sa_part(S);
  n_rows = min(n_rows, row0 + RPC);   // rows of the next CTA are not ours

  // load slots_in [rows][D] -> cur[D][R]
  for (int e = tid; e < UP_R * D; e += UP_THREADS) {
    const int r = e / D, k = e % D;
    cur[k * UP_R + r] = (row0 + r < n_rows) ? slots_in[size_t(row0 + r) * D + k] : 0.f;
  }
  __syncthreads();

  if (flags & UP_DO_C) {
    // ---- weighted mean: uhat[f][r] = (gamma_f (U - Mw) + beta_f A) / A   (sum over chunks first)
    for (int e = tid; e < UP_R * D; e += UP_THREADS) {
      const int r = e / D, f = e % D;
      const int row = row0 + r;
      float val = 0.f;
      if (row < n_rows) {
        const int b = row / S, i = row % S;
        const float* p = partial + size_t(b) * chunks * PART;
        float U = 0.f, A = 0.f, Mw = 0.f;
        for (int c = 0; c < chunks; ++c) {
          U += p[c * PART + i * D + f];
          A += p[c * PART + S * D + i];
          Mw += p[c * PART + S * D + S + i];
        }
        val = (w.ln_in_g[f] * (U - Mw) + w.ln_in_b[f] * A) / A;
      }
      t0[f * UP_R + r] = val;
    }
    __syncthreads();
    linear_T(simt, w.wv_t, w.bv, D, D, t0, t1, false);                 // updates = Wv uhat + bv       -> t1
    linear_T(simt, w.w_ih_t, w.b_ih, D, 3 * D, t1, big0, false);       // gi                          -> big0 [384][R]
    linear_T(simt, w.w_hh_t, w.b_hh, D, 3 * D, cur, big1, false);      // gh (hidden = slots_prev)    -> big1
    for (int e = tid; e < UP_R * D; e += UP_THREADS) {           // GRUCell, gate order r,z,n
      const int k = e / UP_R, r = e % UP_R;
      const float ir = big0[k * UP_R + r], iz = big0[(D + k) * UP_R + r], in_ = big0[(2 * D + k) * UP_R + r];
      const float hr = big1[k * UP_R + r], hz = big1[(D + k) * UP_R + r], hn = big1[(2 * D + k) * UP_R + r];
      const float rg = 1.f / (1.f + __expf(-(ir + hr)));
      const float zg = 1.f / (1.f + __expf(-(iz + hz)));
      const float ng = tanhf(in_ + rg * hn);
      const float h = cur[k * UP_R + r];
      cur[k * UP_R + r] = (1.f - zg) * ng + zg * h;
    }
    __syncthreads();
    layernorm_T(cur, t0, w.ln_mlp_g, w.ln_mlp_b, w.ln_eps_sa, D);
    linear_T(simt, w.w1_t, w.b1, D, w.mlp_hidden, t0, big0, true);
    linear_T(simt, w.w2_t, w.b2, w.mlp_hidden, D, big0, t1, false);
    for (int e = tid; e < UP_R * D; e += UP_THREADS) cur[e] += t1[e];   // slots + MLP(LN(slots))
    __syncthreads();
    if (slots_out) {
      for (int e = tid; e < UP_R * D; e += UP_THREADS) {
        const int r = e / D, k = e % D;
        const int row = row0 + r;
        if (row < n_rows) slots_out[size_t(row / S) * slots_out_stride + size_t(row % S) * D + k] = cur[k * UP_R + r];
      }
    }
  }

  if (flags & UP_DO_T) {
    // ---- post-norm TransformerBlock: y = LN(MHSA(x) + x); z = LN(MLP(y) + y)   (attention.py:387-395)
    linear_T(simt, w.t_wq_t, nullptr, D, D, cur, t0, false);           // q -> t0
    linear_T(simt, w.t_wk_t, nullptr, D, D, cur, t1, false);           // k -> t1
    linear_T(simt, w.t_wv_t, nullptr, D, D, cur, big0, false);         // v -> big0[0..127]
    float* att = big1;                                           // attention output [128][R]
    float* sc = big1 + D * UP_R;                                 // scores [R][H][16]
    const int H = w.t_heads, dh = D / H;
    const float scl = rsqrtf(float(dh));
    // phase 1: thread = (query row r, head h, key pair jp): two of the <= 16 scores each
    for (int e = tid; e < UP_R * H * 8; e += UP_THREADS) {
      const int r = e % UP_R, h = (e / UP_R) % H, jp = e / (UP_R * H);
      if (r >= RPC) continue;                                    // padding rows
      const int rs = (r / S) * S;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = 2 * jp + jj;
        if (j < S) {
          float d = 0.f;
          for (int c = 0; c < dh; ++c) d = fmaf(t0[(h * dh + c) * UP_R + r], t1[(h * dh + c) * UP_R + rs + j], d);
          sc[(r * H + h) * 16 + j] = d * scl;
        }
      }
    }
    __syncthreads();
    // phase 2: thread = (row r, head h, 1/4 of the head's channels): softmax over the S keys, weighted sum of V
    for (int e = tid; e < UP_R * H * 4; e += UP_THREADS) {
      const int r = e % UP_R, h = (e / UP_R) % H, cg = e / (UP_R * H);
      if (r >= RPC) continue;
      const int rs = (r / S) * S;
      float pj[16];
      float mx = -1e30f;
      for (int j = 0; j < S; ++j) { pj[j] = sc[(r * H + h) * 16 + j]; mx = fmaxf(mx, pj[j]); }
      float den = 0.f;
      for (int j = 0; j < S; ++j) { pj[j] = __expf(pj[j] - mx); den += pj[j]; }
      const float inv = 1.f / den;
      const int cw = dh / 4;
      for (int c = cg * cw; c < (cg + 1) * cw; ++c) {
        float o = 0.f;
        for (int j = 0; j < S; ++j) o = fmaf(pj[j], big0[(h * dh + c) * UP_R + rs + j], o);
        att[(h * dh + c) * UP_R + r] = o * inv;
      }
    }
    __syncthreads();
    linear_T(simt, w.t_wo_t, nullptr, D, D, att, t0, false);
    for (int e = tid; e < UP_R * D; e += UP_THREADS) t0[e] += cur[e];
    __syncthreads();
    layernorm_T(t0, t1, w.t_ln1_g, w.t_ln1_b, w.ln_eps_tf, D);   // y -> t1
    linear_T(simt, w.t_w1_t, w.t_b1, D, w.t_hidden, t1, big0, true);
    linear_T(simt, w.t_w2_t, w.t_b2, w.t_hidden, D, big0, t0, false);
    for (int e = tid; e < UP_R * D; e += UP_THREADS) t0[e] += t1[e];
    __syncthreads();
    layernorm_T(t0, cur, w.t_ln2_g, w.t_ln2_b, w.ln_eps_tf, D);  // z -> cur
    if (pred_out) {
      for (int e = tid; e < UP_R * D; e += UP_THREADS) {
        const int r = e / D, k = e % D;
        if (row0 + r < n_rows) pred_out[size_t(row0 + r) * D + k] = cur[k * UP_R + r];
      }
    }
  }

  if (flags & UP_DO_A) {
    // ---- next pass: q = Wq LN(slots) + bq ; qt = Wk^T q ; g = scale*qt*gamma ; sg = sum g ; cb = scale*(qt.beta + q.bk)
    layernorm_T(cur, t0, w.ln_slot_g, w.ln_slot_b, w.ln_eps_sa, D);
    linear_T(simt, w.wq_t, w.bq, D, D, t0, t1, false);                 // q  -> t1
    linear_T(simt, w.wk, nullptr, D, D, t1, t0, false);                // qt -> t0  (Wk as stored [d][f] is "in-major" here)
    for (int e = tid; e < UP_R * D; e += UP_THREADS) {
      const int r = e / D, f = e % D;
      if (row0 + r < n_rows) {
        const int b = (row0 + r) / S, i = (row0 + r) % S;
        gvec[size_t(b) * PART + i * D + f] = w.scale * t0[f * UP_R + r] * w.ln_in_g[f];
      }
    }
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int r = warp; r < UP_R; r += UP_THREADS / 32) {          // warp per row: sg = sum qt*gamma, cb = qt.beta + q.bk
        float sgv = 0.f, cbv = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int f = lane + 32 * j;
          const float qt = t0[f * UP_R + r];
          sgv += qt * w.ln_in_g[f];
          cbv += qt * w.ln_in_b[f] + t1[f * UP_R + r] * w.bk[f];
        }
        sgv = warp_sum(sgv);
        cbv = warp_sum(cbv);
        if (lane == 0 && row0 + r < n_rows) {
          const int b = (row0 + r) / S, i = (row0 + r) % S;
          gvec[size_t(b) * PART + S * D + i] = w.scale * sgv;
          gvec[size_t(b) * PART + S * D + S + i] = w.scale * cbv;
        }
      }
    }
  }
}

constexpr int UP_SMEM = (3 * SA_D + 2 * 512) * UP_R * 4;   // 90112 B

int launch_update2(const SaWeights& w, int S, int chunks, int B, int flags, const float* slots_in, const float* partial,
                   float* slots_out, int out_stride, float* pred_out, float* gvec, cudaStream_t stream);
bool update2_supported(const SaWeights& w, int flags);

// tocvp_tuning.corrector_mode: 0 (default) = second-version update kernel (weights streamed through a shared-memory ring,
// slot_attention_update.cu); 2 = first tensor-core version (B fragments read from L2 inside the MMA loop); 1 = first-version
// fp32 SIMT loops.  All three compute the same thing at fp32-level accuracy.
static int launch_update(const SaWeights& w, int S, int chunks, int B, int flags, const float* slots_in, const float* partial,
                         float* slots_out, int out_stride, float* pred_out, float* gvec, cudaStream_t stream) {
  if (opts().corrector_mode == 0 && update2_supported(w, flags))
    return launch_update2(w, S, chunks, B, flags, slots_in, partial, slots_out, out_stride, pred_out, gvec, stream);
  static SmemAttrOnce attr_once;
  TOCVP_TRY(ensure_smem_attr(attr_once, sa_update_kernel, UP_SMEM));
  const int rows = B * S;
  const int rpc = (UP_R / S) * S;
  sa_update_kernel<<<(rows + rpc - 1) / rpc, UP_THREADS, UP_SMEM, stream>>>(w, S, chunks, rows, flags, slots_in, partial,
                                                                           slots_out, out_stride, pred_out, gvec,
                                                                           opts().corrector_mode == 1 ? 1 : 0);
  TOCVP_LAUNCHED();
  return TOCVP_OK;
}

constexpr int SA_MAX_S = 11;   // generic path: cross-warp reduction buffer [8][S][128] floats must fit static smem

size_t slot_attention_workspace_bytes(int B) {   // sized for the largest supported slot count
  constexpr size_t part = sa_part(SA_MAX_S);
  return (size_t(B) * part + size_t(B) * SA_CHUNKS * part + size_t(B) * SA_MAX_S * SA_D) * sizeof(float);
}

template <int S>
static int launch_generic(const void* feats, int feats_f16, size_t seq_stride, int B, int N, int chunks, const float* gvec,
                          float* partial, float ln_eps, float attn_eps, cudaStream_t stream) {
  const dim3 grid(chunks, B);
  if (feats_f16)
    sa_stream_generic_kernel<__half, S><<<grid, 256, 0, stream>>>(static_cast<const __half*>(feats), seq_stride, N, chunks,
                                                                  gvec, partial, ln_eps, attn_eps);
  else
    sa_stream_generic_kernel<float, S><<<grid, 256, 0, stream>>>(static_cast<const float*>(feats), seq_stride, N, chunks,
                                                                 gvec, partial, ln_eps, attn_eps);
  TOCVP_LAUNCHED();
  return TOCVP_OK;
}

// One streaming pass of the corrector over the features of one frame (g / sg / cb in gvec -> partial sums).
static int sa_stream_pass(const SaWeights& w, const void* feats, int feats_f16, size_t feats_seq_stride, int B, int N,
                          bool fast, int chunks, const float* gvec, float* partial, cudaStream_t stream) {
  const int S = w.num_slots;
  if (fast && feats_f16) {
    // pipeline format: tcgen05 streaming kernel (slot_attention_tc.cu)
    TOCVP_TRY(sa_stream_tc(static_cast<const __half*>(feats), feats_seq_stride, B, N, gvec, partial, w.ln_eps_sa,
                           w.attn_eps, stream));
  } else if (fast) {
    // fp32 features (the reference dtype at the module boundary): all-fp32 SIMT kernel
    const dim3 grid(SA_CHUNKS, B);
    sa_stream_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(feats), feats_seq_stride, N, gvec,
                                                      partial, w.ln_eps_sa, w.attn_eps);
    TOCVP_LAUNCHED();
  } else {
    int r = TOCVP_ERR_BAD_ARG;
    switch (S) {
#define SA_GENERIC_CASE(SS)                                                                                        \
  case SS:                                                                                                         \
  r = launch_generic<SS>(feats, feats_f16, feats_seq_stride, B, N, chunks, gvec, partial, w.ln_eps_sa, w.attn_eps, \
                         stream);                                                                                \
  break;
      SA_GENERIC_CASE(4) SA_GENERIC_CASE(5) SA_GENERIC_CASE(6) SA_GENERIC_CASE(7) SA_GENERIC_CASE(8)
      SA_GENERIC_CASE(9) SA_GENERIC_CASE(10) SA_GENERIC_CASE(11)
#undef SA_GENERIC_CASE
      default:
        set_last_error(__FILE__, __LINE__, "slot_attention: num_slots must be in 4..11");
        return TOCVP_ERR_BAD_ARG;
    }
    TOCVP_TRY(r);
  }
  return TOCVP_OK;
}

// feats [B,N,128] (fp32 or f16), slots_in [B,S,128] fp32 -> slots_out (row b at slots_out + b*out_stride), and,
// if pred_out != null, pred_out = transition(slots_out) [B,S,128].
int slot_attention(const SaWeights& w, const void* feats, int feats_f16, size_t feats_seq_stride, int B, int N, const float* slots_in, int iters,
                   float* slots_out, int out_stride, float* pred_out, void* workspace, size_t ws_bytes,
                   cudaStream_t stream) {
  const int S = w.num_slots;
  TOCVP_CHECK_ARG(feats && slots_in && slots_out && workspace && B > 0 && iters >= 1 && N >= 1);
  TOCVP_CHECK_ARG(S >= 1 && S <= SA_MAX_S);
  TOCVP_CHECK_ARG(feats_seq_stride >= size_t(N) * SA_D && feats_seq_stride % 8 == 0);
  TOCVP_CHECK_ARG(w.mlp_hidden <= 512 && w.t_hidden <= 512 && w.mlp_hidden % 256 == 0);
  TOCVP_CHECK_ARG(pred_out == nullptr || (w.t_heads > 0 && SA_D % w.t_heads == 0 && w.t_hidden % 256 == 0));
  if (ws_bytes < slot_attention_workspace_bytes(B)) {
    set_last_error(__FILE__, __LINE__, "slot_attention: workspace too small");
    return TOCVP_ERR_WORKSPACE;
  }
  // kernel choice: the two specialised streaming kernels cover 8 slots over N % 512 == 0 locations (SAVi / CATER);
  // everything else (10 slots over 81 / 576 patch tokens for ExtendedDINOSAUR) takes the generic kernel
  const bool fast = (S == SA_S) && (N % (SA_CHUNKS * 128) == 0);
  const int chunks = fast ? SA_CHUNKS : (N >= 1024 ? 4 : (N >= 256 ? 2 : 1));
  const int part = sa_part(S);
  float* gvec = static_cast<float*>(workspace);
  float* partial = gvec + size_t(B) * part;
  float* tmp_slots = partial + size_t(B) * chunks * part;   // [B,S,128] intermediate iterates
  TOCVP_TRY(launch_update(w, S, chunks, B, UP_DO_A, slots_in, nullptr, nullptr, 0, nullptr, gvec, stream));
  const float* cur = slots_in;
  for (int it = 0; it < iters; ++it) {
    TOCVP_TRY(sa_stream_pass(w, feats, feats_f16, feats_seq_stride, B, N, fast, chunks, gvec, partial, stream));
    const bool last = (it == iters - 1);
    if (!last) {
      TOCVP_TRY(launch_update(w, S, chunks, B, UP_DO_C | UP_DO_A, cur, partial, tmp_slots, S * SA_D, nullptr, gvec, stream));
      cur = tmp_slots;
    } else {
      TOCVP_TRY(launch_update(w, S, chunks, B, UP_DO_C | (pred_out ? UP_DO_T : 0), cur, partial, slots_out, out_stride,
                              pred_out, gvec, stream));
    }
  }
  return TOCVP_OK;
}

size_t slot_attention_seq_workspace_bytes(int B) {
  return slot_attention_workspace_bytes(B) + size_t(B) * SA_MAX_S * SA_D * sizeof(float);
}

// The corrector chain of SAVi.forward_decomp (SAVi.py:178-204) over n_frames consecutive frames in ONE call:
//   for t: slots_t = SlotAttention(feats_t, cur, iters_t) -> slot_history[:, t];  cur = transition(slots_t)
// The update launch that finishes frame t also applies the transition and emits the query vectors of frame t+1's
// streaming pass (UP_DO_C | UP_DO_T | UP_DO_A), so a frame costs two kernels (one streaming pass + one per-slot update)
// instead of three, and the host enqueues the whole chain without returning to Python.
int slot_attention_seq(const SaWeights& w, const void* feats, int feats_f16, size_t feats_seq_stride,
                       size_t feats_frame_stride, int B, int N, int n_frames, int iters_first, int iters,
                       const float* slots_in, float* slot_history, size_t hist_seq_stride, size_t hist_frame_stride,
                       float* carry_out, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  const int S = w.num_slots;
  TOCVP_CHECK_ARG(feats && slots_in && slot_history && carry_out && workspace && B > 0 && N >= 1 && n_frames >= 1);
  TOCVP_CHECK_ARG(iters_first >= 1 && iters >= 1 && S >= 1 && S <= SA_MAX_S);
  TOCVP_CHECK_ARG(feats_seq_stride >= size_t(N) * SA_D && feats_seq_stride % 8 == 0 && feats_frame_stride % 8 == 0);
  TOCVP_CHECK_ARG(w.mlp_hidden <= 512 && w.t_hidden <= 512 && w.mlp_hidden % 256 == 0);
  TOCVP_CHECK_ARG(w.t_heads > 0 && SA_D % w.t_heads == 0 && w.t_hidden % 256 == 0);   // the chain needs the transition
  TOCVP_CHECK_ARG(hist_seq_stride <= size_t(INT32_MAX));
  if (ws_bytes < slot_attention_seq_workspace_bytes(B)) {
    set_last_error(__FILE__, __LINE__, "slot_attention_seq: workspace too small");
    return TOCVP_ERR_WORKSPACE;
  }
  const bool fast = (S == SA_S) && (N % (SA_CHUNKS * 128) == 0);
  const int chunks = fast ? SA_CHUNKS : (N >= 1024 ? 4 : (N >= 256 ? 2 : 1));
  const int part = sa_part(S);
  float* gvec = static_cast<float*>(workspace);
  float* partial = gvec + size_t(B) * part;
  float* tmp_slots = partial + size_t(B) * chunks * part;
  float* pp = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + slot_attention_workspace_bytes(B));
  const size_t esz = feats_f16 ? 2 : 4;
  TOCVP_TRY(launch_update(w, S, chunks, B, UP_DO_A, slots_in, nullptr, nullptr, 0, nullptr, gvec, stream));
  const float* cur = slots_in;
  for (int t = 0; t < n_frames; ++t) {
    const void* ft = static_cast<const uint8_t*>(feats) + size_t(t) * feats_frame_stride * esz;
    float* out_t = slot_history + size_t(t) * hist_frame_stride;
    float* nxt = ((n_frames - 1 - t) & 1) ? pp : carry_out;      // the last frame's transition output lands in carry_out
    const int its = (t == 0) ? iters_first : iters;
    const float* c = cur;
    for (int it = 0; it < its; ++it) {
      TOCVP_TRY(sa_stream_pass(w, ft, feats_f16, feats_seq_stride, B, N, fast, chunks, gvec, partial, stream));
      if (it < its - 1) {
        TOCVP_TRY(launch_update(w, S, chunks, B, UP_DO_C | UP_DO_A, c, partial, tmp_slots, S * SA_D, nullptr, gvec, stream));
        c = tmp_slots;
      } else {
        const int flags = UP_DO_C | UP_DO_T | ((t + 1 < n_frames) ? UP_DO_A : 0);
        TOCVP_TRY(launch_update(w, S, chunks, B, flags, c, partial, out_t, int(hist_seq_stride), nxt, gvec, stream));
      }
    }
    cur = nxt;
  }
  return TOCVP_OK;
}

}  // namespace tocvp

extern "C" size_t tocvp_slot_attention_workspace_bytes(int B) { return tocvp::slot_attention_workspace_bytes(B); }

extern "C" int tocvp_slot_attention(const tocvp_sa_weights* w, const void* feats, int feats_f16,
                                    size_t feats_seq_stride, int B, int N,
                                    const float* slots_in, int iters, float* slots_out, int out_stride,
                                    float* pred_out, void* workspace, size_t ws_bytes, void* stream) {
  if (w == nullptr) {
    tocvp::set_last_error(__FILE__, __LINE__, "null weights");
    return TOCVP_ERR_BAD_ARG;
  }
  tocvp::OptsScope scope(w->tuning);
  return tocvp::slot_attention(*w, feats, feats_f16, feats_seq_stride, B, N, slots_in, iters, slots_out, out_stride, pred_out, workspace,
                               ws_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" size_t tocvp_slot_attention_seq_workspace_bytes(int B) { return tocvp::slot_attention_seq_workspace_bytes(B); }

extern "C" int tocvp_slot_attention_seq(const tocvp_sa_weights* w, const void* feats, int feats_f16,
                                        size_t feats_seq_stride, size_t feats_frame_stride, int B, int N, int n_frames,
                                        int iters_first, int iters, const float* slots_in, float* slot_history,
                                        size_t hist_seq_stride, size_t hist_frame_stride, float* carry_out,
                                        void* workspace, size_t ws_bytes, void* stream) {
  if (w == nullptr) {
    tocvp::set_last_error(__FILE__, __LINE__, "null weights");
    return TOCVP_ERR_BAD_ARG;
  }
  tocvp::OptsScope scope(w->tuning);
  return tocvp::slot_attention_seq(*w, feats, feats_f16, feats_seq_stride, feats_frame_stride, B, N, n_frames,
                                   iters_first, iters, slots_in, slot_history, hist_seq_stride, hist_frame_stride,
                                   carry_out, workspace, ws_bytes, static_cast<cudaStream_t>(stream));
}

// TransformerBlock.forward, post-norm flavour = the SAVi transition module applied on its own (reference
// src/models/Blocks/attention.py:387-395, called as `self.transition_module(slots)` at src/models/SAVi.py:193).
extern "C" int tocvp_transition(const tocvp_sa_weights* w, const float* slots, int B, float* out, void* stream) {
  using namespace tocvp;
  TOCVP_CHECK_ARG(w && slots && out && B > 0);
  const int S = w->num_slots;
  TOCVP_CHECK_ARG(S >= 1 && S <= SA_MAX_S);
  TOCVP_CHECK_ARG(w->t_heads > 0 && SA_D % w->t_heads == 0 && w->t_hidden % 256 == 0 && w->t_hidden <= 512);
  TOCVP_CHECK_ARG(w->t_wq_t && w->t_wk_t && w->t_wv_t && w->t_wo_t && w->t_w1_t && w->t_w2_t);
  OptsScope scope(w->tuning);
  return launch_update(*w, S, 1, B, UP_DO_T, slots, nullptr, nullptr, 0, out, nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" size_t tocvp_sizeof_sa_weights(void) { return sizeof(tocvp_sa_weights); }
