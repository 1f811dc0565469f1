// ExtendedDINOSAUR-specific stages of the rollout path (CLIPort shape, BASELINE.json configs[3]):
//
//  * tocvp_dino_project: linear_feat_proj = LayerNorm(768) -> Linear(768,768) -> ReLU -> Linear(768,128)
//    (reference src/models/ExtendedDINOSAUR.py:97-102, applied at :186) on the (synthetic) ViT patch features.
//  * tocvp_patch_decode: MLPPatchDecoder.forward (src/models/EncodersDecoders/decoders.py:232-282):
//      broadcast slots + learned positional embedding + LayerNorm          -> one fused bandwidth kernel (norm.cu)
//      MLP 128 -> 1024 -> 1024 -> 1024 -> 769                              -> tcgen05 GEMMs (bias/ReLU epilogues)
//      softmax over slots of the alpha column, weighted sum of the features -> `patch_composite_kernel`
//      conv_patch_decoder (decoders.py:325-365): [conv3x3 + BN + ReLU, nearest x2]* + conv3x3 -> RGB
//                                                                          -> tcgen05 implicit GEMMs (gemm.h, ConvMap):
//         eval-mode BatchNorm is folded into the conv weights/bias at pack time, and every "Upsample(2) -> conv3x3" pair
//         runs as four 2x2 phase convolutions on the LOW-resolution activation (4/9 of the FLOPs, the upsampled tensor
//         is never materialised)
//      bilinear resize to img_size (align_corners = False, decoders.py:268-275) -> `resize_bilinear_kernel`
#include "gemm.h"
#include "host_util.h"
#include "ptx.cuh"

namespace tocvp {

int layernorm_bcast(const void* x, int x_is_f16, int ldx, int x_div, const float* add, int add_rows, const float* gamma,
                    const float* beta, float eps, int rows, int D, __half* out16, int ld16, float* out32, int ld32,
                    cudaStream_t stream);
int layernorm(const void* x, int x_is_f16, int ldx, const float* add, int add_rows, const float* gamma,
              const float* beta, float eps, int rows, int D, __half* out16, int ld16, float* out32, int ld32,
              cudaStream_t stream);

static size_t al256(size_t n) { return (n + 255) & ~size_t(255); }

// ------------------------------------------------------------------------------------------------ feature projection
constexpr int PROJ_CHUNK_ROWS = 65536;

static size_t proj_carve(const tocvp_proj_weights& w, int rows, __half** a16, __half** h16, uint8_t* base) {
  const int cr = rows < PROJ_CHUNK_ROWS ? rows : PROJ_CHUNK_ROWS;
  size_t off = 0;
  if (a16) *a16 = reinterpret_cast<__half*>(base + off);
  off += al256(size_t(cr) * w.feat_dim * 2);
  if (h16) *h16 = reinterpret_cast<__half*>(base + off);
  off += al256(size_t(cr) * w.hidden_dim * 2);
  return off;
}

// ------------------------------------------------------------------------------------------------ slot compositing
// y: fp32 [n_frames*S*N, ldy] (last MLP layer: F feature columns + 1 alpha logit at column F).
// feats32 [n_frames*N, F] fp32, act16: zero-bordered NHWC f16 [n_frames, g+2, g+2, F] (input of the CNN),
// masks fp32 [n_frames, S, N].  One CTA per (frame, patch), thread t owns channels 4t..4t+3.
__global__ void __launch_bounds__(256)
patch_composite_kernel(const float* __restrict__ y, int ldy, int S, int N, int F, int g, float* __restrict__ feats32,
                       __half* __restrict__ act16, float* __restrict__ masks) {
  __shared__ float s_alpha[16];
  const int fn = blockIdx.x;
  const int f = fn / N, n = fn - f * N;
  const float* y0 = y + (size_t(f) * S * N + n) * ldy;          // slot s at y0 + s*N*ldy
  if (threadIdx.x < 32) {
    const int s = threadIdx.x;
    const float logit = s < S ? y0[size_t(s) * N * ldy + F] : -1e30f;
    const float m = warp_max(logit);
    const float e = s < S ? __expf(logit - m) : 0.f;
    const float den = warp_sum(e);
    if (s < S) {
      const float a = e / den;
      s_alpha[s] = a;
      if (masks) masks[(size_t(f) * S + s) * N + n] = a;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x * 4; c < F; c += blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < S; ++s) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(y0 + size_t(s) * N * ldy + c));
      const float a = s_alpha[s];
      acc.x += a * v.x; acc.y += a * v.y; acc.z += a * v.z; acc.w += a * v.w;
    }
    if (feats32) *reinterpret_cast<float4*>(feats32 + size_t(fn) * F + c) = acc;
    if (act16) {
      const int py = n / g + 1, px = n % g + 1;
      uint2 p;
      p.x = pack_half2(acc.x, acc.y);
      p.y = pack_half2(acc.z, acc.w);
      *reinterpret_cast<uint2*>(act16 + ((size_t(f) * (g + 2) + py) * (g + 2) + px) * F + c) = p;
    }
  }
}

// zero the 1-pixel border of an NHWC f16 activation [n_img, Hp, Wp, C] (C % 8 == 0): the convolution's zero padding
__global__ void zero_border_kernel(__half* __restrict__ act, int n_img, int Hp, int Wp, int C) {
  const int per_img = 2 * Wp + 2 * (Hp - 2);
  const int c8 = C / 8;
  const size_t total = size_t(n_img) * per_img * c8;
  for (size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += size_t(gridDim.x) * blockDim.x) {
    const int c = int(e % c8);
    const size_t r = e / c8;
    const int b = int(r % per_img), img = int(r / per_img);
    int y, x;
    if (b < Wp) { y = 0; x = b; }
    else if (b < 2 * Wp) { y = Hp - 1; x = b - Wp; }
    else { const int k = b - 2 * Wp; y = 1 + (k >> 1); x = (k & 1) ? Wp - 1 : 0; }
    *reinterpret_cast<uint4*>(act + ((size_t(img) * Hp + y) * Wp + x) * C + c * 8) = make_uint4(0, 0, 0, 0);
  }
}

// in: fp32 NHWC [n_img, Hs, Ws, ld] (first 3 channels) -> out fp32 NCHW [n_img, 3, Ho, Wo], bilinear,
// align_corners = False (torch F.interpolate semantics: src = max(0, (dst + 0.5) * in/out - 0.5)); identity if equal.
__global__ void __launch_bounds__(256)
resize_bilinear_kernel(const float* __restrict__ in, int ld, int Hs, int Ws, float* __restrict__ out, int Ho, int Wo,
                       int n_img) {
  const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= size_t(n_img) * Ho * Wo) return;
  const int x = int(idx % Wo), y = int((idx / Wo) % Ho), img = int(idx / (size_t(Wo) * Ho));
  const float sy = fmaxf((y + 0.5f) * (float(Hs) / float(Ho)) - 0.5f, 0.f);
  const float sx = fmaxf((x + 0.5f) * (float(Ws) / float(Wo)) - 0.5f, 0.f);
  const int y0 = min(int(sy), Hs - 1), x0 = min(int(sx), Ws - 1);
  const int y1 = min(y0 + 1, Hs - 1), x1 = min(x0 + 1, Ws - 1);
  const float ly = sy - float(y0), lx = sx - float(x0);
  const float* b = in + size_t(img) * Hs * Ws * ld;
  const float4 v00 = *reinterpret_cast<const float4*>(b + (size_t(y0) * Ws + x0) * ld);
  const float4 v01 = *reinterpret_cast<const float4*>(b + (size_t(y0) * Ws + x1) * ld);
  const float4 v10 = *reinterpret_cast<const float4*>(b + (size_t(y1) * Ws + x0) * ld);
  const float4 v11 = *reinterpret_cast<const float4*>(b + (size_t(y1) * Ws + x1) * ld);
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  float* o = out + size_t(img) * 3 * Ho * Wo + size_t(y) * Wo + x;
  o[0] = w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x;
  o[size_t(Ho) * Wo] = w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y;
  o[2 * size_t(Ho) * Wo] = w00 * v00.z + w01 * v01.z + w10 * v10.z + w11 * v11.z;
}

// ------------------------------------------------------------------------------------------------ workspace layout
struct PatchBuffers {
  int cf;                       // frames per pass
  __half *x16, *h16[2];
  float* y32;
  __half* act[TOCVP_PATCH_MAX_CNN + 1];   // act[0] = CNN input, act[i+1] = output of conv block i (zero-bordered NHWC)
  int sp[TOCVP_PATCH_MAX_CNN + 2];        // unpadded spatial size of act[i]; sp[n_cnn+1] = final conv output size
  float* rgb32;
  int rgb_ld;
};

static int round8(int n) { return (n + 7) & ~7; }

static size_t patch_carve(const tocvp_patch_weights& w, int n_frames, PatchBuffers* pb, uint8_t* base) {
  const int S = w.num_slots, N = w.num_patches, D = w.slot_dim, F = w.feat_dim, g = w.grid;
  int hmax = 8;
  for (int i = 0; i + 1 < w.n_mlp; ++i) hmax = w.mlp_out[i] > hmax ? w.mlp_out[i] : hmax;
  const int ypad = round8(w.mlp_out[w.n_mlp - 1]);
  PatchBuffers t{};
  t.sp[0] = g;
  for (int i = 0; i < w.n_cnn; ++i) t.sp[i + 1] = t.sp[i] * (w.cnn_up[i] ? 2 : 1);
  t.sp[w.n_cnn + 1] = t.sp[w.n_cnn] * (w.out_up ? 2 : 1);
  t.rgb_ld = w.out_up ? 4 : 8;
  // bytes per frame -> frames per pass (bounds the workspace to ~2 GiB, at most 256 frames)
  size_t per = size_t(S) * N * (size_t(D) * 2 + size_t(hmax) * 4 + size_t(ypad) * 4);
  per += size_t(g + 2) * (g + 2) * F * 2;
  for (int i = 0; i < w.n_cnn; ++i) per += size_t(t.sp[i + 1] + 2) * (t.sp[i + 1] + 2) * w.cnn_cout[i] * 2;
  per += size_t(t.sp[w.n_cnn + 1]) * t.sp[w.n_cnn + 1] * t.rgb_ld * 4;
  long long cf = (long long)((size_t(2) << 30) / per);
  cf = cf < 1 ? 1 : (cf > 256 ? 256 : cf);
  if (cf > n_frames) cf = n_frames;
  t.cf = int(cf);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* p = base ? base + off : nullptr;
    off += al256(bytes);
    return p;
  };
  const size_t rows = size_t(cf) * S * N;
  t.x16 = reinterpret_cast<__half*>(take(rows * D * 2));
  t.h16[0] = reinterpret_cast<__half*>(take(rows * hmax * 2));
  t.h16[1] = reinterpret_cast<__half*>(take(rows * hmax * 2));
  t.y32 = reinterpret_cast<float*>(take(rows * ypad * 4));
  t.act[0] = reinterpret_cast<__half*>(take(size_t(cf) * (g + 2) * (g + 2) * F * 2));
  for (int i = 0; i < w.n_cnn; ++i)
    t.act[i + 1] = reinterpret_cast<__half*>(take(size_t(cf) * (t.sp[i + 1] + 2) * (t.sp[i + 1] + 2) * w.cnn_cout[i] * 2));
  t.rgb32 = reinterpret_cast<float*>(take(size_t(cf) * t.sp[w.n_cnn + 1] * t.sp[w.n_cnn + 1] * t.rgb_ld * 4));
  if (pb) *pb = t;
  return off;
}

// tap offsets of a conv layer over a zero-bordered input of padded width Wp
static void fill_offsets(ConvMap* cm, int taps, int Wp) {
  if (taps == 9) {            // full 3x3 neighbourhood, one tap set for every column
    for (int t = 0; t < 9; ++t) cm->off[0][t] = (t / 3 - 1) * Wp + (t % 3 - 1);
  } else {                    // 4 phase-specific 2x2 taps: (dy, dx) -> input offset (dy + py - 1, dx + px - 1)
    for (int ph = 0; ph < 4; ++ph)
      for (int t = 0; t < 4; ++t) cm->off[ph][t] = ((t >> 1) + (ph >> 1) - 1) * Wp + ((t & 1) + (ph & 1) - 1);
  }
}

static int check_patch(const tocvp_patch_weights& w) {
  TOCVP_CHECK_ARG(w.n_mlp >= 1 && w.n_mlp <= TOCVP_PATCH_MAX_MLP && w.n_cnn >= 0 && w.n_cnn <= TOCVP_PATCH_MAX_CNN);
  TOCVP_CHECK_ARG(w.num_slots >= 1 && w.num_slots <= 16 && w.slot_dim % 8 == 0 && w.feat_dim % 64 == 0);
  TOCVP_CHECK_ARG(w.grid * w.grid == w.num_patches && w.pos_embed != nullptr);
  TOCVP_CHECK_ARG(w.mlp_out[w.n_mlp - 1] == w.feat_dim + 1);
  for (int i = 0; i + 1 < w.n_mlp; ++i) TOCVP_CHECK_ARG(w.mlp_out[i] % 8 == 0);
  for (int i = 0; i < w.n_cnn; ++i) {
    TOCVP_CHECK_ARG(w.cnn_cin[i] % 64 == 0 && w.cnn_cout[i] % 64 == 0 && w.cnn_w[i] && w.cnn_b[i]);
    TOCVP_CHECK_ARG(w.cnn_cin[i] == (i == 0 ? w.feat_dim : w.cnn_cout[i - 1]));
  }
  TOCVP_CHECK_ARG(!w.reconstruct_images || (w.out_w && w.out_b && w.out_cin % 64 == 0));
  return TOCVP_OK;
}

}  // namespace tocvp

using namespace tocvp;

extern "C" size_t tocvp_sizeof_proj_weights(void) { return sizeof(tocvp_proj_weights); }
extern "C" size_t tocvp_sizeof_patch_weights(void) { return sizeof(tocvp_patch_weights); }

extern "C" size_t tocvp_dino_project_workspace_bytes(const tocvp_proj_weights* w, int rows) {
  if (!w || rows <= 0) return 0;
  return proj_carve(*w, rows, nullptr, nullptr, nullptr);
}

extern "C" int tocvp_dino_project(const tocvp_proj_weights* w, const float* feats, int rows, void* out_f16,
                                  float* out_f32, void* workspace, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TOCVP_CHECK_ARG(w && feats && rows > 0 && (out_f16 || out_f32) && workspace);
  OptsScope scope(w->tuning);
  TOCVP_CHECK_ARG(w->feat_dim % 8 == 0 && w->hidden_dim % 8 == 0 && w->slot_dim % 8 == 0 && w->feat_dim <= 1024);
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  if (ws_bytes < proj_carve(*w, rows, nullptr, nullptr, nullptr)) {
    set_last_error(__FILE__, __LINE__, "dino_project: workspace too small");
    return TOCVP_ERR_WORKSPACE;
  }
  __half *a16, *h16;
  proj_carve(*w, rows, &a16, &h16, static_cast<uint8_t*>(workspace));
  const int F = w->feat_dim, Hd = w->hidden_dim, D = w->slot_dim;
  for (int r0 = 0; r0 < rows; r0 += PROJ_CHUNK_ROWS) {
    const int nr = (rows - r0) < PROJ_CHUNK_ROWS ? (rows - r0) : PROJ_CHUNK_ROWS;
    TOCVP_TRY(layernorm(feats + size_t(r0) * F, 0, F, nullptr, 0, w->ln_g, w->ln_b, w->ln_eps, nr, F, a16, F, nullptr, 0, st));
    TOCVP_TRY(gemm_f16(a16, F, static_cast<const __half*>(w->w1), F, nr, Hd, F, w->b1, 1, nullptr, 0, 1, 0, nullptr, 0, h16,
                       Hd, st));
    TOCVP_TRY(gemm_f16(h16, Hd, static_cast<const __half*>(w->w2), Hd, nr, D, Hd, w->b2, 0, nullptr, 0, 1, 0,
                       out_f32 ? out_f32 + size_t(r0) * D : nullptr, D,
                       out_f16 ? static_cast<__half*>(out_f16) + size_t(r0) * D : nullptr, D, st));
  }
  return TOCVP_OK;
}

extern "C" size_t tocvp_patch_decode_workspace_bytes(const tocvp_patch_weights* w, int n_frames) {
  if (!w || n_frames <= 0 || check_patch(*w) != TOCVP_OK) return 0;
  return patch_carve(*w, n_frames, nullptr, nullptr);
}

extern "C" int tocvp_patch_decode(const tocvp_patch_weights* w, const float* slots, int n_frames, float* recons_imgs,
                                  float* recons_feats, float* masks, void* workspace, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TOCVP_CHECK_ARG(w && slots && workspace && n_frames > 0 && (recons_imgs || recons_feats || masks));
  OptsScope scope(w->tuning);
  TOCVP_TRY(check_patch(*w));
  TOCVP_CHECK_ARG(!recons_imgs || w->reconstruct_images);
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  if (ws_bytes < patch_carve(*w, n_frames, nullptr, nullptr)) {
    set_last_error(__FILE__, __LINE__, "patch_decode: workspace too small");
    return TOCVP_ERR_WORKSPACE;
  }
  PatchBuffers pb;
  patch_carve(*w, n_frames, &pb, static_cast<uint8_t*>(workspace));
  const int S = w->num_slots, N = w->num_patches, D = w->slot_dim, F = w->feat_dim, g = w->grid;
  const int ypad = round8(F + 1);
  // zero borders of the CNN activations (the workspace is caller-owned and may hold anything)
  if (recons_imgs) {
    for (int i = 0; i <= w->n_cnn; ++i) {
      const int Hp = pb.sp[i] + 2, C = (i == 0) ? F : w->cnn_cout[i - 1];
      const size_t items = size_t(pb.cf) * (4 * Hp - 4) * (C / 8);
      const int grid = int((items + 255) / 256 > 148 * 8 ? 148 * 8 : (items + 255) / 256);
      zero_border_kernel<<<grid, 256, 0, st>>>(pb.act[i], pb.cf, Hp, Hp, C);
      TOCVP_LAUNCHED();
    }
  }
  for (int f0 = 0; f0 < n_frames; f0 += pb.cf) {
    const int nf = (n_frames - f0) < pb.cf ? (n_frames - f0) : pb.cf;
    const int rows = nf * S * N;
    // ---- broadcast + pos_embed (+ LayerNorm) -> f16 rows (decoders.py:244-247)
    TOCVP_CHECK_ARG(w->ln_g != nullptr && w->ln_b != nullptr);   // initial_layer_norm = true in the named config
    TOCVP_TRY(layernorm_bcast(slots + size_t(f0) * S * D, 0, D, N, w->pos_embed, N, w->ln_g, w->ln_b, w->ln_eps, rows, D,
                              pb.x16, D, nullptr, 0, st));
    // ---- MLP
    const __half* a = pb.x16;
    int ka = D;
    for (int i = 0; i < w->n_mlp; ++i) {
      const bool last = (i == w->n_mlp - 1);
      const int n_out = last ? ypad : w->mlp_out[i];
      if (last) {
        TOCVP_TRY(gemm_f16(a, ka, static_cast<const __half*>(w->mlp_w[i]), ka, rows, n_out, ka, w->mlp_b[i], 0, nullptr, 0,
                           1, 0, pb.y32, ypad, nullptr, 0, st));
      } else {
        __half* o = pb.h16[i & 1];
        TOCVP_TRY(gemm_f16(a, ka, static_cast<const __half*>(w->mlp_w[i]), ka, rows, n_out, ka, w->mlp_b[i], 1, nullptr, 0,
                           1, 0, nullptr, 0, o, n_out, st));
        a = o;
        ka = n_out;
      }
    }
    // ---- alpha softmax over slots + weighted feature sum (decoders.py:252-256)
    patch_composite_kernel<<<nf * N, F / 4 < 256 ? F / 4 : 256, 0, st>>>(
        pb.y32, ypad, S, N, F, g, recons_feats ? recons_feats + size_t(f0) * N * F : nullptr,
        recons_imgs ? pb.act[0] : nullptr, masks ? masks + size_t(f0) * S * N : nullptr);
    TOCVP_LAUNCHED();
    if (!recons_imgs) continue;
    // ---- CNN (decoders.py:259-265): BN folded, upsampling folded into phase convolutions
    for (int i = 0; i < w->n_cnn; ++i) {
      ConvMap cm{};
      const int up = w->cnn_up[i];
      cm.taps = up ? 4 : 9;
      cm.cin = w->cnn_cin[i];
      cm.cpp = w->cnn_cout[i];
      cm.up = up;
      cm.tiles_per_phase = up ? 1 : 0;     // "phased": the GEMM derives the tile count from its tile width
      cm.Hp = cm.Wp = pb.sp[i] + 2;
      cm.Hop = cm.Wop = pb.sp[i + 1] + 2;
      cm.pad = 1;
      fill_offsets(&cm, cm.taps, cm.Wp);
      TOCVP_TRY(gemm_conv_f16(pb.act[i], static_cast<const __half*>(w->cnn_w[i]), nf, (up ? 4 : 1) * cm.cpp, cm,
                              w->cnn_b[i], 1, nullptr, pb.act[i + 1], cm.cpp, st));
    }
    {
      // final conv3x3 -> RGB (decoders.py:355-362); with a preceding upsample the 4 phases x (3+1 pad) channels form
      // N = 16 columns over the full 3x3 low-resolution neighbourhood (zero weights where a phase does not see a tap)
      ConvMap cm{};
      cm.taps = 9;
      cm.cin = w->out_cin;
      cm.up = w->out_up;
      cm.cpp = w->out_up ? 4 : 8;
      cm.tiles_per_phase = 0;
      cm.Hp = cm.Wp = pb.sp[w->n_cnn] + 2;
      cm.Hop = cm.Wop = pb.sp[w->n_cnn + 1];
      cm.pad = 0;
      fill_offsets(&cm, 9, cm.Wp);
      TOCVP_TRY(gemm_conv_f16(pb.act[w->n_cnn], static_cast<const __half*>(w->out_w), nf, w->out_up ? 16 : 8, cm, w->out_b,
                              0, pb.rgb32, nullptr, pb.rgb_ld, st));
    }
    const int sf = pb.sp[w->n_cnn + 1], I = w->img_size;
    const size_t npix = size_t(nf) * I * I;
    resize_bilinear_kernel<<<int((npix + 255) / 256), 256, 0, st>>>(pb.rgb32, pb.rgb_ld, sf, sf,
                                                                    recons_imgs + size_t(f0) * 3 * I * I, I, I, nf);
    TOCVP_LAUNCHED();
  }
  return TOCVP_OK;
}
