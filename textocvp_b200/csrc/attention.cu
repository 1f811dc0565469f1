// Multi-head attention for the predictor's short sequences (<= 128 keys, head dim 64), fp32 softmax.
// One CTA per (sequence, head); K/V of the head are staged once in shared memory (K rows padded to 33 words so that
// lanes reading different keys hit different banks); each warp then owns query rows: lanes split the keys for
// Q.K^T, warp-shuffle max/sum for the softmax, and split the 64 output dims for P.V.
// Serves both the unmasked self-attention over the <= 10-frame slot window (reference
// src/models/Blocks/attention.py:245-265, 183-193) and the text cross-attention (attention.py:303-319): q, k, v
// are addressed as (base + row*ld + head*64) so the fused QKV GEMM output and the hoisted text K|V are used in place.
#include "host_util.h"
#include "ptx.cuh"

namespace tocvp {

constexpr int ATT_DH = 64;
constexpr int ATT_MAXK = 128;
constexpr int ATT_THREADS = 128;

__global__ void __launch_bounds__(ATT_THREADS)
mha_kernel(const __half* __restrict__ q, int ldq, const __half* __restrict__ k, const __half* __restrict__ v, int ldkv,
           int Tq, int Tk, int heads, float scale, __half* __restrict__ out, int ldo) {
  __shared__ __align__(16) uint32_t sK[ATT_MAXK * 33];      // half2 words, row stride 33
  __shared__ __align__(16) uint32_t sV[ATT_MAXK * 32];      // half2 words, row stride 32
  __shared__ float sQ[ATT_THREADS / 32][ATT_DH];
  __shared__ float sP[ATT_THREADS / 32][ATT_MAXK];
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __half* kb = k + size_t(b) * Tk * ldkv + h * ATT_DH;
  const __half* vb = v + size_t(b) * Tk * ldkv + h * ATT_DH;
  // stage K, V: 8 x 16-byte chunks per row
  for (int e = threadIdx.x; e < Tk * 8; e += ATT_THREADS) {
    const int r = e >> 3, c = e & 7;
    const uint4 kk = *reinterpret_cast<const uint4*>(kb + size_t(r) * ldkv + c * 8);
    const uint4 vv = *reinterpret_cast<const uint4*>(vb + size_t(r) * ldkv + c * 8);
    uint32_t* dk = sK + r * 33 + c * 4;
    dk[0] = kk.x; dk[1] = kk.y; dk[2] = kk.z; dk[3] = kk.w;
    *reinterpret_cast<uint4*>(sV + r * 32 + c * 4) = vv;
  }
  __syncthreads();
  const __half* qb = q + size_t(b) * Tq * ldq + h * ATT_DH;
  __half* ob = out + size_t(b) * Tq * ldo + h * ATT_DH;
  for (int i = warp; i < Tq; i += ATT_THREADS / 32) {
    const __half2 q2 = *reinterpret_cast<const __half2*>(qb + size_t(i) * ldq + lane * 2);
    const float2 qf = __half22float2(q2);
    __syncwarp();
    sQ[warp][lane * 2] = qf.x * scale;
    sQ[warp][lane * 2 + 1] = qf.y * scale;
    __syncwarp();
    float sc[ATT_MAXK / 32];
    float mx = -1e30f;
#pragma unroll
    for (int t = 0; t < ATT_MAXK / 32; ++t) {
      const int j = lane + t * 32;
      float d = -1e30f;
      if (j < Tk) {
        d = 0.f;
        const uint32_t* kr = sK + j * 33;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float2 kf = __half22float2(*reinterpret_cast<const __half2*>(&kr[c]));
          d += sQ[warp][2 * c] * kf.x + sQ[warp][2 * c + 1] * kf.y;
        }
      }
      sc[t] = d;
      mx = fmaxf(mx, d);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < ATT_MAXK / 32; ++t) {
      const int j = lane + t * 32;
      const float p = (j < Tk) ? __expf(sc[t] - mx) : 0.f;
      sum += p;
      if (j < Tk) sP[warp][j] = p;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < Tk; ++j) {
      const float p = sP[warp][j];
      const float2 vf = __half22float2(*reinterpret_cast<const __half2*>(&sV[j * 32 + lane]));
      o0 += p * vf.x;
      o1 += p * vf.y;
    }
    const float inv = 1.f / sum;
    *reinterpret_cast<__half2*>(ob + size_t(i) * ldo + lane * 2) = __floats2half2_rn(o0 * inv, o1 * inv);
  }
}

// q [B*Tq, ldq], k/v [B*Tk, ldkv] (all f16, head h at column h*64), out [B*Tq, ldo] f16.
int mha_f16(const __half* q, int ldq, const __half* k, const __half* v, int ldkv, int B, int Tq, int Tk, int heads,
            __half* out, int ldo, cudaStream_t stream) {
  TOCVP_CHECK_ARG(q && k && v && out && B > 0 && Tq > 0 && Tk > 0 && Tk <= ATT_MAXK && heads > 0);
  TOCVP_CHECK_ARG(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 2 == 0);
  mha_kernel<<<B * heads, ATT_THREADS, 0, stream>>>(q, ldq, k, v, ldkv, Tq, Tk, heads, 0.125f, out, ldo);
  TOCVP_LAUNCHED();
  return TOCVP_OK;
}

}  // namespace tocvp

extern "C" int tocvp_mha_f16(const void* q, int ldq, const void* k, const void* v, int ldkv, int B, int Tq, int Tk,
                             int heads, void* out, int ldo, void* stream) {
  return tocvp::mha_f16(static_cast<const __half*>(q), ldq, static_cast<const __half*>(k),
                        static_cast<const __half*>(v), ldkv, B, Tq, Tk, heads, static_cast<__half*>(out), ldo,
                        static_cast<cudaStream_t>(stream));
}
