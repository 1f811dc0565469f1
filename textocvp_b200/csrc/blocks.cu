// Stand-alone forwards of the reference's small sub-modules (the fused stage kernels never call these; they exist so that
// `savi.decoder(x)`, `savi.encoder(x)`, `pe(x)`, `pos_embedding(x)` and friends work on the CUDA path exactly like the
// reference's nn.Module calls do):
//   tocvp_cast_f16            fp32 -> f16 (saturating), the entry format of every tensor-core kernel
//   tocvp_add_table           x[row] + table[(row / div) % mod]: SoftPositionEmbed.forward (reference
//                             src/models/Blocks/model_blocks.py:215-226) and TemporalPositionalEncoding.forward (:358-379)
//   tocvp_clamp01             in-place clamp to [0,1] (src/05_evaluate_predictor.py:96)
//   tocvp_conv5x5_generic     conv5x5 + bias (+ReLU) on a MATERIALISED NCHW fp32 input with any channel counts, fp32 SIMT:
//                             first layer of ConvDecoder.forward (src/models/EncodersDecoders/decoders.py:111-125) when it
//                             is called on an arbitrary tensor instead of broadcast slots
#include "host_util.h"
#include "ptx.cuh"

namespace tocvp {

__global__ void __launch_bounds__(256) cast_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t n) {
  const size_t n4 = n / 4;
  for (size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n4; e += size_t(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(in) + e);
    uint2 p;
    p.x = pack_half2(v.x, v.y);
    p.y = pack_half2(v.z, v.w);
    reinterpret_cast<uint2*>(out)[e] = p;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const size_t e = n4 * 4 + threadIdx.x;
    reinterpret_cast<uint16_t*>(out)[e] = uint16_t(pack_half2(in[e], 0.f) & 0xFFFFu);
  }
}

__global__ void __launch_bounds__(256)
add_table_kernel(const float* __restrict__ x, const float* __restrict__ table, int div, int mod, int D, size_t rows,
                 float* __restrict__ out) {
  const size_t total4 = rows * size_t(D / 4);
  for (size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total4; e += size_t(gridDim.x) * blockDim.x) {
    const size_t row = e / size_t(D / 4);
    const int c4 = int(e % size_t(D / 4));
    const size_t tr = (row / size_t(div)) % size_t(mod);
    const float4 a = __ldg(reinterpret_cast<const float4*>(x) + e);
    const float4 t = __ldg(reinterpret_cast<const float4*>(table + tr * D) + c4);
    reinterpret_cast<float4*>(out)[e] = make_float4(a.x + t.x, a.y + t.y, a.z + t.z, a.w + t.w);
  }
}

__global__ void __launch_bounds__(256) clamp01_kernel(float* __restrict__ x, size_t n) {
  const size_t n4 = n / 4;
  for (size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n4; e += size_t(gridDim.x) * blockDim.x) {
    float4 v = reinterpret_cast<float4*>(x)[e];
    v.x = fminf(fmaxf(v.x, 0.f), 1.f); v.y = fminf(fmaxf(v.y, 0.f), 1.f);
    v.z = fminf(fmaxf(v.z, 0.f), 1.f); v.w = fminf(fmaxf(v.w, 0.f), 1.f);
    reinterpret_cast<float4*>(x)[e] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const size_t e = n4 * 4 + threadIdx.x;
    x[e] = fminf(fmaxf(x[e], 0.f), 1.f);
  }
}

// One CTA = one image row segment of 32 pixels x 16 output channels; the 5 x 36 x CIN_CHUNK input patch and the weight
// slice are staged through shared memory.  fp32 FMA throughout (this is the convenience path, not the hot one).
constexpr int GC_TW = 32, GC_CO = 16, GC_CI = 16;
__global__ void __launch_bounds__(128)
conv5x5_generic_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                       __half* __restrict__ out16, float* __restrict__ out32, int H, int W, int cin, int cout, int relu) {
  __shared__ float s_in[GC_CI][5][GC_TW + 4];
  __shared__ float s_w[GC_CO][GC_CI][25];
  const int tiles_x = (W + GC_TW - 1) / GC_TW;
  const int x0 = (blockIdx.x % tiles_x) * GC_TW, y = blockIdx.x / tiles_x;
  const int co0 = blockIdx.y * GC_CO, img = blockIdx.z;
  const int px = threadIdx.x % GC_TW, cg = threadIdx.x / GC_TW;     // 4 channel groups of 4 output channels
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const size_t plane = size_t(H) * W;
  for (int ci0 = 0; ci0 < cin; ci0 += GC_CI) {
    for (int e = threadIdx.x; e < GC_CI * 5 * (GC_TW + 4); e += 128) {
      const int c = e / (5 * (GC_TW + 4)), r = (e / (GC_TW + 4)) % 5, xx = e % (GC_TW + 4);
      const int yy = y + r - 2, xg = x0 + xx - 2, ci = ci0 + c;
      float v = 0.f;
      if (ci < cin && yy >= 0 && yy < H && xg >= 0 && xg < W) v = __ldg(x + (size_t(img) * cin + ci) * plane + size_t(yy) * W + xg);
      s_in[c][r][xx] = v;
    }
    for (int e = threadIdx.x; e < GC_CO * GC_CI * 25; e += 128) {
      const int o = e / (GC_CI * 25), c = (e / 25) % GC_CI, t = e % 25;
      const int co = co0 + o, ci = ci0 + c;
      s_w[o][c][t] = (co < cout && ci < cin) ? __ldg(w + (size_t(co) * cin + ci) * 25 + t) : 0.f;
    }
    __syncthreads();
    for (int c = 0; c < GC_CI; ++c) {
#pragma unroll
      for (int t = 0; t < 25; ++t) {
        const float v = s_in[c][t / 5][px + t % 5];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fmaf(v, s_w[cg * 4 + j][c][t], acc[j]);
      }
    }
    __syncthreads();
  }
  const int xg = x0 + px;
  if (xg >= W) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int co = co0 + cg * 4 + j;
    if (co >= cout) continue;
    float v = acc[j] + (bias ? __ldg(bias + co) : 0.f);
    if (relu) v = fmaxf(v, 0.f);
    const size_t pix = size_t(img) * plane + size_t(y) * W + xg;
    if (out16) reinterpret_cast<uint16_t*>(out16)[pix * cout + co] = uint16_t(pack_half2(v, 0.f) & 0xFFFFu);   // NHWC
    if (out32) out32[(size_t(img) * cout + co) * plane + size_t(y) * W + xg] = v;                              // NCHW
  }
}

static inline int ew_blocks(size_t n, int per_block) {
  size_t g = (n + per_block - 1) / per_block;
  return int(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace tocvp

using namespace tocvp;

extern "C" int tocvp_cast_f16(const float* in, void* out, size_t n, void* stream) {
  TOCVP_CHECK_ARG(in && out && n > 0);
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0);
  cast_f16_kernel<<<ew_blocks(n / 4 + 1, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, static_cast<__half*>(out), n);
  TOCVP_LAUNCHED();
  return TOCVP_OK;
}

extern "C" int tocvp_add_table(const float* x, const float* table, int div, int mod, int D, size_t rows, float* out,
                               void* stream) {
  TOCVP_CHECK_ARG(x && table && out && div >= 1 && mod >= 1 && D > 0 && D % 4 == 0 && rows > 0);
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(table) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  add_table_kernel<<<ew_blocks(rows * size_t(D / 4), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, table, div, mod,
                                                                                                       D, rows, out);
  TOCVP_LAUNCHED();
  return TOCVP_OK;
}

extern "C" int tocvp_clamp01(float* x, size_t n, void* stream) {
  TOCVP_CHECK_ARG(x && n > 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0);
  clamp01_kernel<<<ew_blocks(n / 4 + 1, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n);
  TOCVP_LAUNCHED();
  return TOCVP_OK;
}

extern "C" int tocvp_conv5x5_generic(const float* x, const float* weight, const float* bias, int relu, void* out_nhwc_f16,
                                     float* out_nchw_f32, int n_img, int H, int W, int cin, int cout, void* stream) {
  TOCVP_CHECK_ARG(x && weight && (out_nhwc_f16 || out_nchw_f32) && n_img > 0 && H > 0 && W > 0 && cin > 0 && cout > 0);
  TOCVP_CHECK_ARG(n_img <= 65535 && (cout + GC_CO - 1) / GC_CO <= 65535);
  const dim3 grid(((W + GC_TW - 1) / GC_TW) * H, (cout + GC_CO - 1) / GC_CO, n_img);
  conv5x5_generic_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(x, weight, bias, static_cast<__half*>(out_nhwc_f16),
                                                                              out_nchw_f32, H, W, cin, cout, relu);
  TOCVP_LAUNCHED();
  return TOCVP_OK;
}
