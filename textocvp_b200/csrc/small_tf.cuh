// Building blocks of the "whole small transformer in one CTA" kernels (text_encoder.cu, ocvp.cu): fp32 SIMT linear layers
// over shared-memory activations with transposed ([in][out]) weights read from L2, LayerNorm, grouped attention.
#pragma once
#include "ptx.cuh"

namespace tocvp {

constexpr int TE_THREADS = 256;
constexpr int TE_MAXL = 64;
constexpr int TE_RB = 16;   // rows per accumulator block

// y[r][n] = act(bias[n] + sum_k x[r][k] * Wt[k][n])   r < L;  x: smem [L][ldx], y: smem [L][ldy]
__device__ inline void te_linear(const float* __restrict__ Wt, const float* __restrict__ bias, int K, int N, const float* x,
                          int ldx, float* y, int ldy, int L, int act /*0 none, 1 gelu, 2 relu*/, const float* res, int ldr) {
  for (int n = threadIdx.x; n < N; n += TE_THREADS) {
    const float bv = bias ? __ldg(bias + n) : 0.f;
    for (int r0 = 0; r0 < L; r0 += TE_RB) {
      float acc[TE_RB];
#pragma unroll
      for (int r = 0; r < TE_RB; ++r) acc[r] = bv;
#pragma unroll 4
      for (int k = 0; k < K; ++k) {
        const float w = __ldg(Wt + size_t(k) * N + n);
#pragma unroll
        for (int r = 0; r < TE_RB; ++r) acc[r] = fmaf(w, x[(r0 + r) * ldx + k], acc[r]);   // smem broadcast reads
      }
#pragma unroll
      for (int r = 0; r < TE_RB; ++r) {
        if (r0 + r < L) {
          float v = acc[r];
          if (act == 1) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752f));            // exact GELU (torch default)
          else if (act == 2) v = fmaxf(v, 0.f);
          if (res) v += res[(r0 + r) * ldr + n];
          y[(r0 + r) * ldy + n] = v;
        }
      }
    }
  }
  __syncthreads();
}

// in-place LayerNorm of x[L][D] (one warp per row), optional row mask (rows with keep[r] == 0 are zeroed afterwards)
__device__ void te_layernorm_to(const float* x, float* y, int ld, int L, int D, const float* __restrict__ g,
                                const float* __restrict__ b, float eps, const unsigned char* keep);
__device__ inline void te_layernorm(float* x, int ld, int L, int D, const float* __restrict__ g, const float* __restrict__ b,
                                    float eps, const unsigned char* keep) {
  te_layernorm_to(x, x, ld, L, D, g, b, eps, keep);
}
// y may alias x (each lane rewrites only the elements it read)
__device__ inline void te_layernorm_to(const float* x, float* y, int ld, int L, int D, const float* __restrict__ g,
                                       const float* __restrict__ b, float eps, const unsigned char* keep) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < L; r += TE_THREADS / 32) {
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += x[r * ld + c];
    const float mean = warp_sum(s) / D;
    float q = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float d = x[r * ld + c] - mean;
      q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / D + eps);
    const float m = (keep && !keep[r]) ? 0.f : 1.f;
    for (int c = lane; c < D; c += 32) y[r * ld + c] = ((x[r * ld + c] - mean) * rstd * __ldg(g + c) + __ldg(b + c)) * m;
  }
  __syncthreads();
}


// Multi-head self-attention over packed q|k|v rows (row stride ldq, q at column 0, k at D, v at 2D): thread = (head,
// query row), online softmax over the keys j < n_keys that satisfy the group rule:
//   group 0: every key;  group 1: same frame (j / S == i / S);  group 2: same slot (j % S == i % S).
__device__ inline void te_attention(const float* qkv, int ldq, float* att, int ld_att, int L, int n_keys, int D, int H,
                                    int group, int S) {
  const int dh = D / H;
  const float scale = rsqrtf(float(dh));
  for (int e = threadIdx.x; e < H * L; e += TE_THREADS) {
    const int h = e / L, i = e - h * L;
    const float* q = qkv + i * ldq + h * dh;
    float m = -1e30f, den = 0.f;
    float o[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) o[c] = 0.f;
    int j0 = 0, j1 = n_keys, js = 1;
    if (group == 1) { j0 = (i / S) * S; j1 = j0 + S; }
    else if (group == 2) { j0 = i % S; js = S; }
    for (int j = j0; j < j1; j += js) {
      const float* kk = qkv + j * ldq + D + h * dh;
      const float* vv = qkv + j * ldq + 2 * D + h * dh;
      float s = 0.f;
      for (int c = 0; c < dh; ++c) s = fmaf(q[c], kk[c], s);
      s *= scale;
      const float mn = fmaxf(m, s);
      const float corr = __expf(m - mn), p = __expf(s - mn);
      den = den * corr + p;
#pragma unroll
      for (int c = 0; c < 64; ++c)
        if (c < dh) o[c] = o[c] * corr + p * vv[c];
      m = mn;
    }
    const float inv = 1.f / den;
#pragma unroll
    for (int c = 0; c < 64; ++c)
      if (c < dh) att[i * ld_att + h * dh + c] = o[c] * inv;
  }
  __syncthreads();
}

}  // namespace tocvp
