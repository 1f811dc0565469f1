// SAVi.encode (reference src/models/SAVi.py:226-238): SimpleConvEncoder (4 x conv5x5 + ReLU, encoders.py:136-159),
// SoftPositionEmbed add (model_blocks.py:215-226), LayerNorm(32) and the 32 -> 128 -> 128 pointwise MLP (SAVi.py:115-120).
//   conv 1 (3 -> 32, K = 75)  : direct fp32 SIMT convolution straight from the NCHW fp32 frame, writes NHWC f16
//   conv 2-4 (32 -> 32)       : tcgen05 implicit GEMM (conv5x5_tc.cu)
//   posemb + LN               : one fused bandwidth pass (norm.cu), the positional table is batch independent
//   MLP                       : two tcgen05 GEMMs (bias+ReLU / bias epilogues)
#include "host_util.h"
#include "ptx.cuh"

namespace tocvp {

int gemm_f16(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K, const float* bias, int relu,
             const float* residual, int ldr, int res_div, int res_mod, float* out32, int ld32, __half* out16,
             int ld16, cudaStream_t stream);
int conv5x5_f16(const __half* x, const __half* wpacked, const float* bias, __half* out, int n_img, int H, int W, int cin,
                int cout, int relu, cudaStream_t stream);
int layernorm(const void* x, int x_is_f16, int ldx, const float* add, int add_rows, const float* gamma,
              const float* beta, float eps, int rows, int D, __half* out16, int ld16, float* out32, int ld32,
              cudaStream_t stream);

int enc_mlp_f16(const __half* x, const __half* w1, const float* b1, const __half* w2, const float* b2, __half* y, int M,
                cudaStream_t stream);
int conv5x5_vp_f16(const __half* xp, const __half* wpacked, const float* bias, __half* out, int n_img, int H, int W,
                   cudaStream_t stream);
int conv5x5_xp_f16(const __half* x, const __half* wxp, const float* bias, const float* posemb, const float* ln_g,
                   const float* ln_b, float ln_eps, __half* out, int n_img, int H, int W, cudaStream_t stream);
int conv5x5_ln_f16(const __half* x, const __half* wpacked, const float* bias, const float* posemb, const float* ln_g,
                   const float* ln_b, float ln_eps, __half* out, int n_img, int H, int W, cudaStream_t stream);

// tocvp_tuning.encode_mode (per call): bit 0 = first-version SIMT fp32 conv1, bit 1 = separate posemb + LayerNorm pass (first version),
// bit 2 = the MLP as two separate GEMMs (first version), bit 3 = conv 1 over zero-padded channels (25 taps, second version)

// Frame -> tensor-core input of conv 1: NCHW fp32 [n,3,H,W] -> NHWC f16 [n,H,W,32] with channels 3..31 zero, so that conv 1
// (3 -> 32, K = 75) runs on the same tcgen05 implicit-GEMM kernel as conv 2-4 with zero-padded input channels (ten times
// the useful FLOPs, still 1.7x faster than the fp32 SIMT convolution it replaces: 3.5 -> 2.0 ms per 5120 frames).
__global__ void __launch_bounds__(256)
enc_pack_input_kernel(const float* __restrict__ x, size_t img_stride, __half* __restrict__ out, int plane, int n_img) {
  const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= size_t(n_img) * plane) return;
  const int img = int(idx / plane), pix = int(idx % plane);
  const float* xi = x + size_t(img) * img_stride + pix;
  uint4 p0 = make_uint4(pack_half2(xi[0], xi[plane]), pack_half2(xi[2 * size_t(plane)], 0.f), 0u, 0u);
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  uint4* o = reinterpret_cast<uint4*>(out + idx * 32);
  o[0] = p0; o[1] = z; o[2] = z; o[3] = z;
}

// Frame -> x-im2col input of conv 1 (third version): NCHW fp32 [n,3,H,W] -> f16 [n, H+1, W, 32].  Stored row r stands for
// image row y' = r-1; a pixel's 32 values are the 5 x 3 x-neighbourhood of image row y' (k = kx*3 + c, k = 15 zero) and of
// image row y'+1 (k = 16 + kx*3 + c, k = 31 zero), zeros outside the image.  Conv 1 then needs 3 vertical taps of K = 32
// (6 MMAs per 128 pixels) instead of 25 taps over zero-padded channels (50): the 5 x 3 useful inputs of a filter row fill
// half a k-block instead of 3 of its 32 lanes.  modules.im2col_x_row_pairs is the torch statement of this layout.
__global__ void __launch_bounds__(256)
enc_pack_vp_kernel(const float* __restrict__ x, size_t img_stride, __half* __restrict__ out, int H, int W, int n_img) {
  const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t per_img = size_t(H + 1) * W;
  if (idx >= size_t(n_img) * per_img) return;
  const int img = int(idx / per_img);
  const int rem = int(idx % per_img);
  const int r = rem / W, px = rem % W;
  const float* xi = x + size_t(img) * img_stride;
  const int plane = H * W;
  uint32_t w[16];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int y = r - 1 + half;
    const bool yok = y >= 0 && y < H;
    const float* row = xi + (yok ? y : 0) * W + px - 2;          // one pointer per row; taps / channels are constant offsets
    float v[16];
#pragma unroll
    for (int kx = 0; kx < 5; ++kx) {
      const int xx = px + kx - 2;
      const bool ok = yok && xx >= 0 && xx < W;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[kx * 3 + c] = ok ? __ldg(row + c * plane + kx) : 0.f;
    }
    v[15] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) w[half * 8 + j] = pack_half2(v[2 * j], v[2 * j + 1]);
  }
  __half* o = out + idx * 32;                    // 64 bytes per thread: two full 32-byte sectors (256-bit stores)
  stg_256(o, make_uint4(w[0], w[1], w[2], w[3]), make_uint4(w[4], w[5], w[6], w[7]));
  stg_256(o + 16, make_uint4(w[8], w[9], w[10], w[11]), make_uint4(w[12], w[13], w[14], w[15]));
}

constexpr int E1_TH = 8, E1_TW = 32, E1_CO = 32;

// x: fp32 [n, 3, H, W] with image i at x + i*img_stride; w: fp32 [75][32] ((ky*5+kx)*3+ci major); out: f16 NHWC [n,H,W,32]
__global__ void __launch_bounds__(256)
enc_conv1_kernel(const float* __restrict__ x, size_t img_stride, const float* __restrict__ w,
                 const float* __restrict__ bias, __half* __restrict__ out, int H, int W) {
  __shared__ float sIn[3][E1_TH + 4][E1_TW + 4 + 1];
  __shared__ __align__(16) float sW[75 * E1_CO];
  const int img = blockIdx.y;
  const int tiles_x = W / E1_TW;
  const int y0 = (blockIdx.x / tiles_x) * E1_TH, x0 = (blockIdx.x % tiles_x) * E1_TW;
  for (int e = threadIdx.x; e < 75 * E1_CO; e += 256) sW[e] = w[e];
  const float* xi = x + size_t(img) * img_stride;
  for (int e = threadIdx.x; e < 3 * (E1_TH + 4) * (E1_TW + 4); e += 256) {
    const int c = e / ((E1_TH + 4) * (E1_TW + 4)), r = e % ((E1_TH + 4) * (E1_TW + 4));
    const int yy = r / (E1_TW + 4), xx = r % (E1_TW + 4);
    const int gy = y0 - 2 + yy, gx = x0 - 2 + xx;
    sIn[c][yy][xx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? xi[(size_t(c) * H + gy) * W + gx] : 0.f;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[E1_CO];
#pragma unroll
  for (int o = 0; o < E1_CO; ++o) acc[o] = bias[o];
#pragma unroll 1
  for (int tap = 0; tap < 25; ++tap) {
    const int ky = tap / 5, kx = tap % 5;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = sIn[c][ty + ky][tx + kx];
      const float4* wr = reinterpret_cast<const float4*>(sW + (tap * 3 + c) * E1_CO);
#pragma unroll
      for (int o4 = 0; o4 < E1_CO / 4; ++o4) {
        const float4 ww = wr[o4];
        acc[o4 * 4 + 0] += v * ww.x; acc[o4 * 4 + 1] += v * ww.y;
        acc[o4 * 4 + 2] += v * ww.z; acc[o4 * 4 + 3] += v * ww.w;
      }
    }
  }
  __half* o = out + ((size_t(img) * H + (y0 + ty)) * W + (x0 + tx)) * E1_CO;
#pragma unroll
  for (int o8 = 0; o8 < E1_CO / 8; ++o8) {
    uint4 p;
    p.x = pack_half2(fmaxf(acc[o8 * 8 + 0], 0.f), fmaxf(acc[o8 * 8 + 1], 0.f));
    p.y = pack_half2(fmaxf(acc[o8 * 8 + 2], 0.f), fmaxf(acc[o8 * 8 + 3], 0.f));
    p.z = pack_half2(fmaxf(acc[o8 * 8 + 4], 0.f), fmaxf(acc[o8 * 8 + 5], 0.f));
    p.w = pack_half2(fmaxf(acc[o8 * 8 + 6], 0.f), fmaxf(acc[o8 * 8 + 7], 0.f));
    *reinterpret_cast<uint4*>(o + o8 * 8) = p;
  }
}

struct EncBuffers {
  __half *actA, *actB, *h16, *mid16;
};
static size_t align256e(size_t n) { return (n + 255) & ~size_t(255); }
static size_t enc_carve(const tocvp_enc_weights& w, int n, EncBuffers* eb, uint8_t* base) {
  const size_t px = size_t(n) * w.H * w.W;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* p = base ? base + off : nullptr;
    off += align256e(bytes);
    return p;
  };
  EncBuffers t;
  const size_t px1 = size_t(n) * (w.H + 1) * w.W;     // conv 1's packed input carries one extra row per image
  t.actA = reinterpret_cast<__half*>(take(px * w.hidden * 2));
  t.actB = reinterpret_cast<__half*>(take(px1 * w.hidden * 2));
  t.h16 = t.actA;   // LN output reuses actA (conv 4 lands in actB)
  t.mid16 = reinterpret_cast<__half*>(take(px * w.feat_dim * 2));
  if (eb) *eb = t;
  return off;
}

// conv 1 .. conv 3 of SimpleConvEncoder (each + bias + ReLU): frames -> eb.actA (NHWC f16); conv 4 follows in the caller
// (fused with the positional embedding + LayerNorm in SAVi.encode, plain in the stand-alone encoder forward).
// pixel-pair kernel for 32 -> 32 layer i (conv5x5_tc.cu, XP): full-width 16 x 64 tiles, an even tile count
static bool enc_xp_ok(const tocvp_enc_weights& w, int i, int n_img) {
  return !(opts().encode_mode & 16) && w.w_conv_xp[i] != nullptr && w.W % 64 == 0 && w.H % 16 == 0 &&
         ((n_img * (w.H / 16) * (w.W / 64)) % 2 == 0);
}

static int enc_convs_1_to_3(const tocvp_enc_weights& wr, EncBuffers& eb, const float* frames, size_t img_stride, int n_img,
                            int g_enc_mode, cudaStream_t st) {
  const tocvp_enc_weights* w = &wr;
  const int H = w->H, W = w->W, C = w->hidden;
  if ((g_enc_mode & 1) || w->w_conv1_tc == nullptr) {
    const dim3 g1((H / E1_TH) * (W / E1_TW), n_img);
    enc_conv1_kernel<<<g1, 256, 0, st>>>(frames, img_stride, w->w_conv1, w->b_conv1, eb.actA, H, W);
    TOCVP_LAUNCHED();
  } else if (!(g_enc_mode & 8) && w->w_conv1_vp != nullptr && ((n_img * (H / 16) * (W / 32)) % 2 == 0)) {
    const size_t npx = size_t(n_img) * (H + 1) * W;
    enc_pack_vp_kernel<<<int((npx + 255) / 256), 256, 0, st>>>(frames, img_stride, eb.actB, H, W, n_img);
    TOCVP_LAUNCHED();
    TOCVP_TRY(conv5x5_vp_f16(eb.actB, static_cast<const __half*>(w->w_conv1_vp), w->b_conv1, eb.actA, n_img, H, W, st));
  } else {
    const size_t npx = size_t(n_img) * H * W;
    enc_pack_input_kernel<<<int((npx + 255) / 256), 256, 0, st>>>(frames, img_stride, eb.actB, H * W, n_img);
    TOCVP_LAUNCHED();
    TOCVP_TRY(conv5x5_f16(eb.actB, static_cast<const __half*>(w->w_conv1_tc), w->b_conv1, eb.actA, n_img, H, W, C, C, 1, st));
  }
  __half* src = eb.actA;
  __half* dst = eb.actB;
  for (int i = 0; i < 2; ++i) {
    if (enc_xp_ok(wr, i, n_img))
      TOCVP_TRY(conv5x5_xp_f16(src, static_cast<const __half*>(w->w_conv_xp[i]), w->b_conv[i], nullptr, nullptr, nullptr, 0.f,
                               dst, n_img, H, W, st));
    else
      TOCVP_TRY(conv5x5_f16(src, static_cast<const __half*>(w->w_conv[i]), w->b_conv[i], dst, n_img, H, W, C, C, 1, st));
    __half* t = src; src = dst; dst = t;
  }
  return TOCVP_OK;      // conv 3's output is back in eb.actA
}

}  // namespace tocvp

using namespace tocvp;

extern "C" size_t tocvp_sizeof_enc_weights(void) { return sizeof(tocvp_enc_weights); }

extern "C" size_t tocvp_savi_encode_workspace_bytes(const tocvp_enc_weights* w, int n_img) {
  if (!w || n_img <= 0) return 0;
  return enc_carve(*w, n_img, nullptr, nullptr);
}

// frames: fp32, image i (3 x H x W, NCHW planes) at frames + i*img_stride floats; feats out: [n_img, H*W, feat_dim]
extern "C" int tocvp_savi_encode(const tocvp_enc_weights* w, const float* frames, size_t img_stride, int n_img,
                                 void* feats_f16, float* feats_f32, void* workspace, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TOCVP_CHECK_ARG(w && frames && (feats_f16 || feats_f32) && workspace && n_img > 0);
  OptsScope scope(w->tuning);
  const int g_enc_mode = opts().encode_mode & 15;
  TOCVP_CHECK_ARG(w->in_channels == 3 && w->hidden == 32 && w->feat_dim % 8 == 0);
  TOCVP_CHECK_ARG(w->H % 16 == 0 && w->W % 32 == 0);
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  if (ws_bytes < enc_carve(*w, n_img, nullptr, nullptr)) {
    set_last_error(__FILE__, __LINE__, "savi_encode: workspace too small");
    return TOCVP_ERR_WORKSPACE;
  }
  EncBuffers eb;
  enc_carve(*w, n_img, &eb, static_cast<uint8_t*>(workspace));
  const int H = w->H, W = w->W, C = w->hidden, F = w->feat_dim;
  const int M = n_img * H * W;
  TOCVP_TRY(enc_convs_1_to_3(*w, eb, frames, img_stride, n_img, g_enc_mode, st));
  const bool fuse_ln = !(g_enc_mode & 2) && ((n_img * (H / 16) * (W / 32)) % 2 == 0);
  if (fuse_ln && enc_xp_ok(*w, 2, n_img)) {
    TOCVP_TRY(conv5x5_xp_f16(eb.actA, static_cast<const __half*>(w->w_conv_xp[2]), w->b_conv[2], w->posemb, w->ln_g, w->ln_b,
                             1e-5f, eb.actB, n_img, H, W, st));
    eb.h16 = eb.actB;
  } else if (fuse_ln) {
    // conv 4 + positional embedding + LayerNorm(32) (eps 1e-5, SAVi.py:116) in one kernel: the LN input never leaves fp32
    TOCVP_TRY(conv5x5_ln_f16(eb.actA, static_cast<const __half*>(w->w_conv[2]), w->b_conv[2], w->posemb, w->ln_g, w->ln_b,
                             1e-5f, eb.actB, n_img, H, W, st));
    eb.h16 = eb.actB;
  } else {
    TOCVP_TRY(conv5x5_f16(eb.actA, static_cast<const __half*>(w->w_conv[2]), w->b_conv[2], eb.actB, n_img, H, W, C, C, 1, st));
    TOCVP_TRY(layernorm(eb.actB, 1, C, w->posemb, H * W, w->ln_g, w->ln_b, 1e-5f, M, C, eb.h16, C, nullptr, 0, st));
  }
  if (!(g_enc_mode & 4) && feats_f32 == nullptr && F == 128 && C == 32) {
    // both MLP layers in one kernel: the 128-wide hidden activation never leaves the SM (enc_mlp_tc.cu)
    return enc_mlp_f16(eb.h16, static_cast<const __half*>(w->w_mlp1), w->b_mlp1, static_cast<const __half*>(w->w_mlp2),
                       w->b_mlp2, static_cast<__half*>(feats_f16), M, st);
  }
  TOCVP_TRY(gemm_f16(eb.h16, C, static_cast<const __half*>(w->w_mlp1), C, M, F, C, w->b_mlp1, 1, nullptr, 0, 1, 0, nullptr,
                     0, eb.mid16, F, st));
  TOCVP_TRY(gemm_f16(eb.mid16, F, static_cast<const __half*>(w->w_mlp2), F, M, F, F, w->b_mlp2, 0, nullptr, 0, 1, 0,
                     feats_f32, F, static_cast<__half*>(feats_f16), F, st));
  return TOCVP_OK;
}

// SimpleConvEncoder.forward on its own (reference src/models/EncodersDecoders/encoders.py:156-159): 4 x (conv5x5 + bias +
// ReLU) -> NHWC f16 [n_img, H, W, 32] (no positional embedding, no LayerNorm, no MLP).  Same workspace as tocvp_savi_encode.
extern "C" int tocvp_savi_conv_stack(const tocvp_enc_weights* w, const float* frames, size_t img_stride, int n_img,
                                     void* out_nhwc_f16, void* workspace, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TOCVP_CHECK_ARG(w && frames && out_nhwc_f16 && workspace && n_img > 0);
  TOCVP_CHECK_ARG(w->in_channels == 3 && w->hidden == 32 && w->H % 16 == 0 && w->W % 32 == 0);
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && (reinterpret_cast<uintptr_t>(out_nhwc_f16) & 15) == 0);
  OptsScope scope(w->tuning);
  if (ws_bytes < enc_carve(*w, n_img, nullptr, nullptr)) {
    set_last_error(__FILE__, __LINE__, "savi_conv_stack: workspace too small");
    return TOCVP_ERR_WORKSPACE;
  }
  EncBuffers eb;
  enc_carve(*w, n_img, &eb, static_cast<uint8_t*>(workspace));
  TOCVP_TRY(enc_convs_1_to_3(*w, eb, frames, img_stride, n_img, opts().encode_mode & 15, st));
  if (enc_xp_ok(*w, 2, n_img))
    return conv5x5_xp_f16(eb.actA, static_cast<const __half*>(w->w_conv_xp[2]), w->b_conv[2], nullptr, nullptr, nullptr, 0.f,
                          static_cast<__half*>(out_nhwc_f16), n_img, w->H, w->W, st);
  return conv5x5_f16(eb.actA, static_cast<const __half*>(w->w_conv[2]), w->b_conv[2], static_cast<__half*>(out_nhwc_f16),
                     n_img, w->H, w->W, w->hidden, w->hidden, 1, st);
}
