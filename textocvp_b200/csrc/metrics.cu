// Frame metrics of the evaluator on the device: clamp to [0,1] (reference src/05_evaluate_predictor.py:96-99), then per
// image MSE, PSNR and SSIM as fused reduction kernels feeding the metric accumulators that are all-reduced across ranks
// (src/lib/metrics.py:181-270; the reference calls piqa 1.2.2 -- psnr(value_range=1, epsilon=1e-8) and
// SSIM(window_size=11, sigma=1.5, k1=0.01, k2=0.03, valid convolution, mean over channels and positions) -- whose
// published formulas are restated here; piqa itself is not available offline, so metric parity is pinned against the
// oracle's restatement only).
#include "host_util.h"
#include "ptx.cuh"

namespace tocvp {

struct MetricArgs {
  const float* pred;      // [n_img, C, H, W]
  const float* target;    // image i at target + (i / fps) * seq_stride + (frame0 + i % fps) * C*H*W
  size_t seq_stride;
  int fps, frame0;
  int n_img, C, H, W;
  int clamp;
};

__device__ __forceinline__ const float* target_img(const MetricArgs& a, int i) {
  return a.target + size_t(i / a.fps) * a.seq_stride + size_t(a.frame0 + i % a.fps) * a.C * a.H * a.W;
}
__device__ __forceinline__ float clamp01(float v, int on) { return on ? fminf(fmaxf(v, 0.f), 1.f) : v; }

// one CTA per image: mse = mean((p - t)^2), psnr = 10 log10(1 / (mse + 1e-8))
__global__ void __launch_bounds__(256)
mse_psnr_kernel(MetricArgs a, float* __restrict__ mse, float* __restrict__ psnr) {
  __shared__ float s_red[8];
  const int i = blockIdx.x;
  const int n = a.C * a.H * a.W;
  const float4* p = reinterpret_cast<const float4*>(a.pred + size_t(i) * n);
  const float4* t = reinterpret_cast<const float4*>(target_img(a, i));
  float acc = 0.f;
  for (int e = threadIdx.x; e < n / 4; e += 256) {
    const float4 x = __ldg(p + e), y = __ldg(t + e);
    const float d0 = clamp01(x.x, a.clamp) - clamp01(y.x, a.clamp), d1 = clamp01(x.y, a.clamp) - clamp01(y.y, a.clamp);
    const float d2 = clamp01(x.z, a.clamp) - clamp01(y.z, a.clamp), d3 = clamp01(x.w, a.clamp) - clamp01(y.w, a.clamp);
    acc += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += s_red[w];
    const float m = s / float(n);
    if (mse) mse[i] = m;
    if (psnr) psnr[i] = 10.f * log10f(1.f / (m + 1e-8f));
  }
}

// SSIM with an 11-tap separable Gaussian window, valid convolution.  grid = n_img; a CTA walks the (channel, 32 x 32 tile)
// items of its image's (H-10) x (W-10) output map: loads the 42 x 42 inputs, filters x, y, x^2, y^2, xy horizontally
// into shared memory, then vertically per output pixel, and accumulates the mean in a fixed order.
constexpr int SS_T = 32, SS_K = 11, SS_IN = SS_T + SS_K - 1;

__global__ void __launch_bounds__(256)
ssim_kernel(MetricArgs a, float sigma, float c1, float c2, float* __restrict__ ssim) {
  __shared__ float sx[SS_IN][SS_IN + 1], sy[SS_IN][SS_IN + 1];
  __shared__ float sh[5][SS_IN][SS_T + 1];
  __shared__ float sg[SS_K];
  __shared__ float s_red[8];
  const int Ho = a.H - SS_K + 1, Wo = a.W - SS_K + 1;
  const int tiles_x = (Wo + SS_T - 1) / SS_T, tiles_y = (Ho + SS_T - 1) / SS_T;
  const int img = blockIdx.x;
  if (threadIdx.x < SS_K) {
    float g[SS_K], s = 0.f;
    for (int k = 0; k < SS_K; ++k) {
      const float d = float(k) - 0.5f * (SS_K - 1);
      g[k] = expf(-d * d / (2.f * sigma * sigma));
      s += g[k];
    }
    sg[threadIdx.x] = g[threadIdx.x] / s;
  }
  // One CTA per image walks its (channel, tile) work items in a fixed order and thread 0 carries the running sum: the
  // result does not depend on scheduling (no atomics), so a frame's SSIM is bit-identical wherever it sits in a batch.
  float total = 0.f;
  const int n_work = a.C * tiles_y * tiles_x;
  const bool vec = (a.W & 3) == 0;
  // rows of the 42 x 44 input window as aligned float4 pieces (tile origins are multiples of 32): every thread has its two
  // pieces of both tensors in flight before the first one is used -- the scalar version exposed one global latency per
  // element pair, 7 times per item, 60 % of the kernel's stall samples (ncu r2).  (Requesting item work + 1 before item
  // `work` is filtered costs 16 more registers, a resident CTA per SM, and measured slower: 0.83 vs 0.67 ms.)
  constexpr int VPR = (SS_IN + 3) / 4;                      // 11 float4 per row (44 columns; 42 are used)
  constexpr int NV = SS_IN * VPR;                           // 462
  constexpr int IT = (NV + 255) / 256;                      // 2
  float4 vp[IT], vt[IT];
  auto origin = [&](int work, int& ch, int& ty0, int& tx0) {
    ch = work / (tiles_y * tiles_x);
    const int tile = work % (tiles_y * tiles_x);
    ty0 = (tile / tiles_x) * SS_T;
    tx0 = (tile % tiles_x) * SS_T;
  };
  auto fetch = [&](int work) {
    int ch, ty0, tx0;
    origin(work, ch, ty0, tx0);
    const float* p = a.pred + (size_t(img) * a.C + ch) * a.H * a.W;
    const float* t = target_img(a, img) + size_t(ch) * a.H * a.W;
#pragma unroll
    for (int u = 0; u < IT; ++u) {
      const int e = threadIdx.x + u * 256;
      const int r = e / VPR, c4 = (e % VPR) * 4;
      const int y = ty0 + r, x = tx0 + c4;
      const bool ok = e < NV && y < a.H && x < a.W;
      vp[u] = ok ? __ldg(reinterpret_cast<const float4*>(p + size_t(y) * a.W + x)) : make_float4(0.f, 0.f, 0.f, 0.f);
      vt[u] = ok ? __ldg(reinterpret_cast<const float4*>(t + size_t(y) * a.W + x)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  for (int work = 0; work < n_work; ++work) {
    int ch, ty0, tx0;
    origin(work, ch, ty0, tx0);
    if (vec) {
      fetch(work);
#pragma unroll
      for (int u = 0; u < IT; ++u) {
        const int e = threadIdx.x + u * 256;
        if (e < NV) {
          const int r = e / VPR, c4 = (e % VPR) * 4;
          const float px4[4] = {vp[u].x, vp[u].y, vp[u].z, vp[u].w}, tx4[4] = {vt[u].x, vt[u].y, vt[u].z, vt[u].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (c4 + j < SS_IN) {
              sx[r][c4 + j] = clamp01(px4[j], a.clamp);       // outside the image: 0 (clamp01(0) = 0)
              sy[r][c4 + j] = clamp01(tx4[j], a.clamp);
            }
          }
        }
      }
    } else {
      const float* p = a.pred + (size_t(img) * a.C + ch) * a.H * a.W;
      const float* t = target_img(a, img) + size_t(ch) * a.H * a.W;
      for (int e = threadIdx.x; e < SS_IN * SS_IN; e += 256) {
        const int r = e / SS_IN, c = e % SS_IN;
        const int y = ty0 + r, x = tx0 + c;
        const bool ok = y < a.H && x < a.W;
        sx[r][c] = ok ? clamp01(__ldg(p + size_t(y) * a.W + x), a.clamp) : 0.f;
        sy[r][c] = ok ? clamp01(__ldg(t + size_t(y) * a.W + x), a.clamp) : 0.f;
      }
    }
    __syncthreads();
    // Both filter passes slide an 18-value register window over 8 consecutive outputs: 36 (horizontal) / 90 (vertical)
    // shared loads per 8 outputs instead of 176 / 440 -- the first version was bound by its scalar LDS rate (0.96 ms per
    // 4864 frames against 0.07 ms of HBM time).  Lane -> (row, segment) maps are bank-conflict free (43 / 33-float rows).
    float gk[SS_K];
#pragma unroll
    for (int k = 0; k < SS_K; ++k) gk[k] = sg[k];
    constexpr int SEG = 8, WIN = SEG + SS_K - 1;
    for (int e = threadIdx.x; e < SS_IN * (SS_T / SEG); e += 256) {
      const int r = e / (SS_T / SEG), c0 = (e % (SS_T / SEG)) * SEG;
      float vx[WIN], vy[WIN], gx[1], gy[1];
#pragma unroll
      for (int k = 0; k < WIN; ++k) { vx[k] = sx[r][c0 + k]; vy[k] = sy[r][c0 + k]; }
#pragma unroll
      for (int j = 0; j < SEG; ++j) {
        float mx = 0.f, my = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
#pragma unroll
        for (int k = 0; k < SS_K; ++k) {
          const float g = gk[k], ax = vx[j + k], ay = vy[j + k];
          gx[0] = g * ax; gy[0] = g * ay;                     // (g * x) * x etc.: the products of the first version, shared
          mx += gx[0]; my += gy[0]; xx += gx[0] * ax; yy += gy[0] * ay; xy += gx[0] * ay;
        }
        sh[0][r][c0 + j] = mx; sh[1][r][c0 + j] = my; sh[2][r][c0 + j] = xx; sh[3][r][c0 + j] = yy; sh[4][r][c0 + j] = xy;
      }
    }
    __syncthreads();
    float acc = 0.f;
    for (int e = threadIdx.x; e < (SS_T / SEG) * SS_T; e += 256) {
      const int r0 = (e / SS_T) * SEG, c = e % SS_T;
      if (tx0 + c < Wo && ty0 + r0 < Ho) {
        float o[5][SEG];
#pragma unroll
        for (int m = 0; m < 5; ++m) {
          float v[WIN];
#pragma unroll
          for (int k = 0; k < WIN; ++k) v[k] = sh[m][r0 + k][c];
#pragma unroll
          for (int j = 0; j < SEG; ++j) {
            float a_ = 0.f;
#pragma unroll
            for (int k = 0; k < SS_K; ++k) a_ += gk[k] * v[j + k];
            o[m][j] = a_;
          }
        }
#pragma unroll
        for (int j = 0; j < SEG; ++j) {
          if (ty0 + r0 + j < Ho) {
            const float mx = o[0][j], my = o[1][j], xx = o[2][j], yy = o[3][j], xy = o[4][j];
            const float mxx = mx * mx, myy = my * my, mxy = mx * my;
            const float cs = (2.f * (xy - mxy) + c2) / ((xx - mxx) + (yy - myy) + c2);
            acc += (2.f * mxy + c1) / (mxx + myy + c1) * cs;
          }
        }
      }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();                    // also: everyone is done with sx / sy / sh before the next item overwrites them
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += s_red[w];
      total += s;
    }
  }
  if (threadIdx.x == 0) ssim[img] = total / (float(a.C) * float(Ho) * float(Wo));
}

}  // namespace tocvp

using namespace tocvp;

extern "C" int tocvp_frame_metrics(const float* pred, const float* target, size_t target_seq_stride, int frames_per_seq,
                                   int target_frame0, int n_img, int C, int H, int W, int clamp, float* mse, float* psnr,
                                   float* ssim, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TOCVP_CHECK_ARG(pred && target && n_img > 0 && C > 0 && frames_per_seq > 0 && (mse || psnr || ssim));
  TOCVP_CHECK_ARG((C * H * W) % 4 == 0 && target_seq_stride % 4 == 0);
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(pred) & 15) == 0 && (reinterpret_cast<uintptr_t>(target) & 15) == 0);
  MetricArgs a{pred, target, target_seq_stride, frames_per_seq, target_frame0, n_img, C, H, W, clamp};
  if (mse || psnr) {
    mse_psnr_kernel<<<n_img, 256, 0, st>>>(a, mse, psnr);
    TOCVP_LAUNCHED();
  }
  if (ssim) {
    TOCVP_CHECK_ARG(H >= SS_K && W >= SS_K);
    ssim_kernel<<<n_img, 256, 0, st>>>(a, 1.5f, 0.01f * 0.01f, 0.03f * 0.03f, ssim);
    TOCVP_LAUNCHED();
  }
  return TOCVP_OK;
}
