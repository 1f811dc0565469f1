// Frozen ViT front-end of ExtendedDINOSAUR: the reference wraps a timm VisionTransformer
// (vit_base_patch14_dinov2: patch 14, 768-d, 12 pre-norm blocks with LayerScale, 12 heads x 64, GELU MLP x4) in
// ViTEncoder.forward (reference src/models/EncodersDecoders/timm_encoders.py:58-69):
//     normalise -> patch_embed (conv p x p, stride p) -> cat(cls) + pos_embed -> blocks -> drop the class token
// (no final norm: the wrapper stops after `blocks`).  timm is not installed in this environment, so the block structure
// below restates timm's VisionTransformer / Block / LayerScale from their published definition -- PARITY UNPINNED (the
// tests compare against a plain-torch restatement with the same parameter names, oracle/vit_oracle.py).
//
// One host-side driver enqueues the whole forward on the caller's stream, built from the kernels the predictor uses:
//   patch embedding   = one im2col pass (normalisation fused, f16) + tcgen05 GEMM [n_img*N, 3pp] x [E, 3pp]^T + bias
//   block             = LayerNorm -> QKV GEMM + bias -> attention (short kernel <= 128 tokens, streaming kernel beyond)
//                       -> proj GEMM + bias -> x += ls1.gamma * (.)
//                       -> LayerNorm -> fc1 GEMM + bias + exact GELU -> fc2 GEMM + bias -> x += ls2.gamma * (.)
// Residual stream fp32, GEMM operands f16, fp32 accumulate -- the same precision recipe as the predictor.  LayerScale is
// applied in fp32 by a small kernel rather than folded into the f16 weights: at its initial value (1e-5) the folded
// weights would fall into the f16 subnormals.
#include "gemm.h"
#include "host_util.h"
#include "ptx.cuh"

namespace tocvp {

int layernorm(const void* x, int x_is_f16, int ldx, const float* add, int add_rows, const float* gamma,
              const float* beta, float eps, int rows, int D, __half* out16, int ld16, float* out32, int ld32,
              cudaStream_t stream);
int mha_f16_any(const __half* q, int ldq, const __half* k, const __half* v, int ldkv, int B, int Tq, int Tk, int heads,
                __half* out, int ldo, cudaStream_t stream);

// images fp32 NCHW (image i at x + i*img_stride) -> patches f16 [n_img * N, k_pad], k = c*p*p + py*p + px (the flattening of
// patch_embed.proj.weight [E, 3, p, p]), value = (x - mean_c) * inv_std_c, zeros for k >= 3*p*p.
__global__ void __launch_bounds__(256)
vit_patchify_kernel(const float* __restrict__ x, size_t img_stride, __half* __restrict__ out, int n_img, int H, int W, int p,
                    int gh, int gw, int k_pad, float m0, float m1, float m2, float s0, float s1, float s2) {
  const int kp2 = k_pad / 2;
  const size_t total = size_t(n_img) * gh * gw * kp2;
  const int pp = p * p;
  for (size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += size_t(gridDim.x) * blockDim.x) {
    const int k2 = int(e % kp2);
    const size_t patch = e / kp2;
    const int gx = int(patch % gw), gy = int((patch / gw) % gh);
    const size_t img = patch / (size_t(gw) * gh);
    float v[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = 2 * k2 + j;
      v[j] = 0.f;
      if (k < 3 * pp) {
        const int c = k / pp, r = k % pp, py = r / p, px = r % p;
        const float raw = __ldg(x + img * img_stride + (size_t(c) * H + size_t(gy * p + py)) * W + (gx * p + px));
        const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), inv = c == 0 ? s0 : (c == 1 ? s1 : s2);
        v[j] = (raw - mean) * inv;
      }
    }
    reinterpret_cast<uint32_t*>(out)[e] = pack_half2(v[0], v[1]);
  }
}

// tokens[img][0] = cls_token + pos_embed[0];  tokens[img][1 + p] = patch_embedding[img][p] + pos_embed[1 + p]
__global__ void __launch_bounds__(256)
vit_assemble_kernel(const float* __restrict__ emb, const float* __restrict__ cls_pos0, const float* __restrict__ pos,
                    float* __restrict__ tokens, int n_img, int N, int E) {
  const int e4 = E / 4;
  const size_t total = size_t(n_img) * (N + 1) * e4;
  for (size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += size_t(gridDim.x) * blockDim.x) {
    const int c4 = int(e % e4);
    const size_t row = e / e4;
    const int tk = int(row % (N + 1));
    const size_t img = row / (N + 1);
    float4 v;
    if (tk == 0) {
      v = __ldg(reinterpret_cast<const float4*>(cls_pos0) + c4);
    } else {
      const float4 a = __ldg(reinterpret_cast<const float4*>(emb + (img * N + (tk - 1)) * E) + c4);
      const float4 b = __ldg(reinterpret_cast<const float4*>(pos + size_t(tk - 1) * E) + c4);
      v = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
    reinterpret_cast<float4*>(tokens)[e] = v;
  }
}

// x[row][:] += gamma[:] * t[row][:]   (timm LayerScale + the block's residual add)
__global__ void __launch_bounds__(256)
vit_scale_residual_kernel(float* __restrict__ x, const float* __restrict__ t, const float* __restrict__ gamma, size_t rows,
                          int E) {
  const int e4 = E / 4;
  const size_t total = rows * e4;
  for (size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += size_t(gridDim.x) * blockDim.x) {
    const int c4 = int(e % e4);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    const float4 tv = __ldg(reinterpret_cast<const float4*>(t) + e);
    float4 xv = reinterpret_cast<float4*>(x)[e];
    xv.x = fmaf(g.x, tv.x, xv.x); xv.y = fmaf(g.y, tv.y, xv.y); xv.z = fmaf(g.z, tv.z, xv.z); xv.w = fmaf(g.w, tv.w, xv.w);
    reinterpret_cast<float4*>(x)[e] = xv;
  }
}

// out[img][p] = tokens[img][1 + p]   ("removing class patch", timm_encoders.py:68)
__global__ void __launch_bounds__(256)
vit_drop_cls_kernel(const float* __restrict__ tokens, float* __restrict__ out, int n_img, int N, int E) {
  const int e4 = E / 4;
  const size_t total = size_t(n_img) * N * e4;
  for (size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += size_t(gridDim.x) * blockDim.x) {
    const int c4 = int(e % e4);
    const size_t row = e / e4;
    const size_t img = row / N, p = row % N;
    reinterpret_cast<float4*>(out)[e] = __ldg(reinterpret_cast<const float4*>(tokens + (img * (N + 1) + p + 1) * E) + c4);
  }
}

struct VitBuffers {
  __half *patches, *h16, *qkv16, *att16, *mid16;
  float *emb32, *x32, *y32;
};

static size_t vit_align(size_t n) { return (n + 255) & ~size_t(255); }
constexpr int VIT_CHUNK_IMGS = 256;   // images per internal pass (bounds the workspace: ~0.6 GB at 577 tokens)

static size_t vit_carve(const tocvp_vit_weights& w, int n_img, VitBuffers* vb, uint8_t* base) {
  const int n = n_img < VIT_CHUNK_IMGS ? n_img : VIT_CHUNK_IMGS;
  const size_t N = size_t(w.grid_h) * w.grid_w, M = size_t(n) * (N + 1), E = w.embed_dim;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* p = base ? base + off : nullptr;
    off += vit_align(bytes);
    return p;
  };
  VitBuffers t;
  t.patches = reinterpret_cast<__half*>(take(size_t(n) * N * w.k_pad * 2));
  t.emb32 = reinterpret_cast<float*>(take(size_t(n) * N * E * 4));
  t.x32 = reinterpret_cast<float*>(take(M * E * 4));
  t.y32 = reinterpret_cast<float*>(take(M * E * 4));
  t.h16 = reinterpret_cast<__half*>(take(M * E * 2));
  t.qkv16 = reinterpret_cast<__half*>(take(M * 3 * E * 2));
  t.att16 = reinterpret_cast<__half*>(take(M * E * 2));
  t.mid16 = reinterpret_cast<__half*>(take(M * size_t(w.mlp_dim) * 2));
  if (vb) *vb = t;
  return off;
}

static int vit_check(const tocvp_vit_weights& w) {
  TOCVP_CHECK_ARG(w.blocks != nullptr && w.num_blocks >= 0 && w.num_blocks <= 64);
  TOCVP_CHECK_ARG(w.embed_dim % 64 == 0 && w.embed_dim / w.num_heads == 64 && w.mlp_dim % 8 == 0);
  TOCVP_CHECK_ARG(w.patch >= 1 && w.grid_h >= 1 && w.grid_w >= 1 && w.img_h >= w.grid_h * w.patch && w.img_w >= w.grid_w * w.patch);
  TOCVP_CHECK_ARG(w.k_pad % 8 == 0 && w.k_pad >= 3 * w.patch * w.patch);
  TOCVP_CHECK_ARG(w.w_patch && w.b_patch && w.cls_pos0 && w.pos);
  return TOCVP_OK;
}

static inline int vit_grid(size_t n) {
  size_t g = (n + 255) / 256;
  return int(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace tocvp

using namespace tocvp;

extern "C" size_t tocvp_sizeof_vit_weights(void) { return sizeof(tocvp_vit_weights); }
extern "C" size_t tocvp_sizeof_vit_block(void) { return sizeof(tocvp_vit_block); }

extern "C" size_t tocvp_vit_workspace_bytes(const tocvp_vit_weights* w, int n_img) {
  if (!w || n_img <= 0) return 0;
  return vit_carve(*w, n_img, nullptr, nullptr);
}

extern "C" int tocvp_vit_forward(const tocvp_vit_weights* w, const float* images, size_t img_stride, int n_img,
                                 float* feats, void* workspace, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TOCVP_CHECK_ARG(w && images && feats && workspace && n_img > 0);
  TOCVP_TRY(vit_check(*w));
  OptsScope scope(w->tuning);
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  if (ws_bytes < vit_carve(*w, n_img, nullptr, nullptr)) {
    set_last_error(__FILE__, __LINE__, "vit_forward: workspace too small");
    return TOCVP_ERR_WORKSPACE;
  }
  VitBuffers vb;
  vit_carve(*w, n_img, &vb, static_cast<uint8_t*>(workspace));
  const int E = w->embed_dim, Hm = w->mlp_dim, N = w->grid_h * w->grid_w, T = N + 1;
  for (int i0 = 0; i0 < n_img; i0 += VIT_CHUNK_IMGS) {
    const int n = (n_img - i0) < VIT_CHUNK_IMGS ? (n_img - i0) : VIT_CHUNK_IMGS;
    const int M = n * T;
    // ---- normalise + patch embedding + class token + positional embedding (timm_encoders.py:62-65)
    vit_patchify_kernel<<<vit_grid(size_t(n) * N * (w->k_pad / 2)), 256, 0, st>>>(
        images + size_t(i0) * img_stride, img_stride, vb.patches, n, w->img_h, w->img_w, w->patch, w->grid_h, w->grid_w,
        w->k_pad, w->mean[0], w->mean[1], w->mean[2], w->inv_std[0], w->inv_std[1], w->inv_std[2]);
    TOCVP_LAUNCHED();
    TOCVP_TRY(gemm_f16(vb.patches, w->k_pad, static_cast<const __half*>(w->w_patch), w->k_pad, n * N, E, w->k_pad, w->b_patch,
                       0, nullptr, 0, 1, 0, vb.emb32, E, nullptr, 0, st));
    vit_assemble_kernel<<<vit_grid(size_t(M) * (E / 4)), 256, 0, st>>>(vb.emb32, w->cls_pos0, w->pos, vb.x32, n, N, E);
    TOCVP_LAUNCHED();
    // ---- blocks: x = x + ls1(attn(norm1(x))); x = x + ls2(mlp(norm2(x)))
    for (int l = 0; l < w->num_blocks; ++l) {
      const tocvp_vit_block& bk = w->blocks[l];
      TOCVP_TRY(layernorm(vb.x32, 0, E, nullptr, 0, bk.ln1_g, bk.ln1_b, w->ln_eps, M, E, vb.h16, E, nullptr, 0, st));
      TOCVP_TRY(gemm_f16(vb.h16, E, static_cast<const __half*>(bk.w_qkv), E, M, 3 * E, E, bk.b_qkv, 0, nullptr, 0, 1, 0,
                         nullptr, 0, vb.qkv16, 3 * E, st));
      TOCVP_TRY(mha_f16_any(vb.qkv16, 3 * E, vb.qkv16 + E, vb.qkv16 + 2 * E, 3 * E, n, T, T, w->num_heads, vb.att16, E, st));
      TOCVP_TRY(gemm_f16(vb.att16, E, static_cast<const __half*>(bk.w_proj), E, M, E, E, bk.b_proj, 0, nullptr, 0, 1, 0,
                         vb.y32, E, nullptr, 0, st));
      vit_scale_residual_kernel<<<vit_grid(size_t(M) * (E / 4)), 256, 0, st>>>(vb.x32, vb.y32, bk.ls1, size_t(M), E);
      TOCVP_LAUNCHED();
      TOCVP_TRY(layernorm(vb.x32, 0, E, nullptr, 0, bk.ln2_g, bk.ln2_b, w->ln_eps, M, E, vb.h16, E, nullptr, 0, st));
      TOCVP_TRY(gemm_f16(vb.h16, E, static_cast<const __half*>(bk.w_fc1), E, M, Hm, E, bk.b_fc1, 2 /* GELU */, nullptr, 0, 1, 0,
                         nullptr, 0, vb.mid16, Hm, st));
      TOCVP_TRY(gemm_f16(vb.mid16, Hm, static_cast<const __half*>(bk.w_fc2), Hm, M, E, Hm, bk.b_fc2, 0, nullptr, 0, 1, 0,
                         vb.y32, E, nullptr, 0, st));
      vit_scale_residual_kernel<<<vit_grid(size_t(M) * (E / 4)), 256, 0, st>>>(vb.x32, vb.y32, bk.ls2, size_t(M), E);
      TOCVP_LAUNCHED();
    }
    vit_drop_cls_kernel<<<vit_grid(size_t(n) * N * (E / 4)), 256, 0, st>>>(vb.x32, feats + size_t(i0) * N * E, n, N, E);
    TOCVP_LAUNCHED();
  }
  return TOCVP_OK;
}
