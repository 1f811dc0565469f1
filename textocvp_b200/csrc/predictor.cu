// Text-conditioned transformer predictor: BaseTextOCVP.forward (reference src/models/Predictors/text_cond_OCVP.py:79-105),
// its AdaptedEncoderBlock layers (src/models/Blocks/attention.py:504-524, 445-463) and the autoregressive
// PredictorWrapper loop (src/models/Predictors/predictor_wrapper.py:50-87, 143-153) as ONE host-side driver that
// enqueues every kernel of the whole rollout on a stream: tcgen05 GEMMs with fused bias / ReLU / residual / positional
// epilogues, fp32 LayerNorm passes, and the small fp32-softmax attention kernel.
//   * the residual stream stays fp32; only GEMM operands are f16
//   * the text K|V projections are rollout-constant and hoisted out of the step loop (the reference recomputes them)
//   * self-attention is unmasked and the temporal PE is re-flipped every step -> full window recompute (no KV cache)
#include "gemm.h"
#include "host_util.h"
#include "ptx.cuh"

namespace tocvp {

int layernorm(const void* x, int x_is_f16, int ldx, const float* add, int add_rows, const float* gamma,
              const float* beta, float eps, int rows, int D, __half* out16, int ld16, float* out32, int ld32,
              cudaStream_t stream);
int mha_f16(const __half* q, int ldq, const __half* k, const __half* v, int ldkv, int B, int Tq, int Tk, int heads,
            __half* out, int ldo, cudaStream_t stream);
int mha_f16_sub(const __half* q, int ldq, int q_seq_rows, const __half* k, const __half* v, int ldkv, int B, int Tq, int Tk,
                int heads, __half* out, int ldo, cudaStream_t stream);

// rows of the newest frame, fp32: tokens [B, n*S, T] -> [B*S, T]
__global__ void last_frame_f32_kernel(const float* __restrict__ tokens, int n, int S, int T, int B, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const size_t total4 = size_t(B) * S * T / 4;
  for (size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total4; e += size_t(gridDim.x) * blockDim.x) {
    const size_t el = e * 4;
    const int row = int(el / T), c = int(el % T);
    const int b = row / S, s = row % S;
    *reinterpret_cast<float4*>(out + el) =
        *reinterpret_cast<const float4*>(tokens + (size_t(b) * n * S + size_t(n - 1) * S + s) * T + c);
  }
}

// frames [B, F_total, S*D] fp32 -> window tokens f16 [B, n, S*D] starting at frame f0
__global__ void window_to_f16_kernel(const float* __restrict__ frames, size_t seq_stride, int f0, int n, int SD, int B,
                                     __half* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const size_t total4 = size_t(B) * n * SD / 4;
  for (size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total4; e += size_t(gridDim.x) * blockDim.x) {
    const size_t el = e * 4;
    const int b = int(el / (size_t(n) * SD));
    const size_t r = el % (size_t(n) * SD);
    const float4 v = *reinterpret_cast<const float4*>(frames + size_t(b) * seq_stride + size_t(f0) * SD + r);
    uint2 p;
    p.x = pack_half2(v.x, v.y);
    p.y = pack_half2(v.z, v.w);
    *reinterpret_cast<uint2*>(out + el) = p;
  }
}

// rows of the newest frame: tokens fp32 [B, n*S, T] -> f16 [B*S, T]
__global__ void last_frame_to_f16_kernel(const float* __restrict__ tokens, int n, int S, int T, int B,
                                         __half* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const size_t total4 = size_t(B) * S * T / 4;
  for (size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total4; e += size_t(gridDim.x) * blockDim.x) {
    const size_t el = e * 4;
    const int row = int(el / T), c = int(el % T);
    const int b = row / S, s = row % S;
    const float4 v = *reinterpret_cast<const float4*>(tokens + (size_t(b) * n * S + size_t(n - 1) * S + s) * T + c);
    uint2 p;
    p.x = pack_half2(v.x, v.y);
    p.y = pack_half2(v.z, v.w);
    *reinterpret_cast<uint2*>(out + el) = p;
  }
}

// pred [B*S, D] (+ residual frames[b][f_last]) -> frames[b][f_new] and pred_out[b][t]
__global__ void commit_prediction_kernel(const float* __restrict__ pred, float* __restrict__ frames, size_t seq_stride,
                                         int f_last, int f_new, int residual, float* __restrict__ pred_out,
                                         size_t pred_seq_stride, int t, int SD, int B) {
  pdl_wait();
  pdl_trigger();
  const size_t total = size_t(B) * SD;
  for (size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += size_t(gridDim.x) * blockDim.x) {
    const int b = int(e / SD), r = int(e % SD);
    float v = pred[e];
    if (residual) v += frames[size_t(b) * seq_stride + size_t(f_last) * SD + r];
    frames[size_t(b) * seq_stride + size_t(f_new) * SD + r] = v;
    pred_out[size_t(b) * pred_seq_stride + size_t(t) * SD + r] = v;
  }
}

__global__ void copy_context_kernel(const float* __restrict__ src, size_t src_seq_stride, float* __restrict__ frames,
                                    size_t seq_stride, int nctx, int SD, int B) {
  pdl_wait();
  pdl_trigger();
  const size_t total = size_t(B) * nctx * SD;
  for (size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += size_t(gridDim.x) * blockDim.x) {
    const int b = int(e / (size_t(nctx) * SD));
    const size_t r = e % (size_t(nctx) * SD);
    frames[size_t(b) * seq_stride + r] = src[size_t(b) * src_seq_stride + r];
  }
}

static inline int ew_grid(size_t n, int threads = 256) {
  size_t g = (n + threads - 1) / threads;
  return int(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

struct PredBuffers {
  float *frames, *x32, *y32, *z32, *pred32, *stats, *xl32;
  __half *tok16, *h16, *qkv16, *att16, *q16, *mid16, *last16, *text16, *kv16;
};

static size_t align256(size_t n) { return (n + 255) & ~size_t(255); }

static size_t carve(const tocvp_pred_weights& w, int B, int L, int nctx, int npreds, PredBuffers* pb, uint8_t* base) {
  const int S = w.num_slots, D = w.slot_dim, T = w.token_dim, H = w.hidden_dim;
  const int nmax = w.buffer_size;
  const size_t Mmax = size_t(B) * S * nmax;
  const size_t ftot = size_t(nctx + npreds);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* p = base ? base + off : nullptr;
    off += align256(bytes);
    return p;
  };
  PredBuffers t{};
  t.frames = reinterpret_cast<float*>(take(size_t(B) * ftot * S * D * 4));
  t.x32 = reinterpret_cast<float*>(take(Mmax * T * 4));
  t.y32 = reinterpret_cast<float*>(take(Mmax * T * 4));
  t.z32 = reinterpret_cast<float*>(take(Mmax * T * 4));
  t.pred32 = reinterpret_cast<float*>(take(size_t(B) * S * D * 4));
  t.stats = reinterpret_cast<float*>(take(Mmax * size_t(T / 64) * 2 * 4));
  t.xl32 = reinterpret_cast<float*>(take(size_t(B) * S * T * 4));
  t.tok16 = reinterpret_cast<__half*>(take(Mmax * D * 2));
  t.h16 = reinterpret_cast<__half*>(take(Mmax * T * 2));
  t.qkv16 = reinterpret_cast<__half*>(take(Mmax * 3 * T * 2));
  t.att16 = reinterpret_cast<__half*>(take(Mmax * T * 2));
  t.q16 = reinterpret_cast<__half*>(take(Mmax * T * 2));
  t.mid16 = reinterpret_cast<__half*>(take(Mmax * size_t(H > w.cross_hidden ? H : w.cross_hidden) * 2));
  t.last16 = reinterpret_cast<__half*>(take(size_t(B) * S * T * 2));
  t.text16 = reinterpret_cast<__half*>(take(size_t(B) * L * T * 2));
  t.kv16 = reinterpret_cast<__half*>(take(size_t(w.num_layers) * B * L * 2 * T * 2));
  if (pb) *pb = t;
  return off;
}

size_t predictor_workspace_bytes(const tocvp_pred_weights& w, int B, int L, int nctx, int npreds) {
  return carve(w, B, L, nctx, npreds, nullptr, nullptr);
}

// One BaseTextOCVP.forward over the window [f0, f0+n) of pb.frames; result (without residual) in pb.pred32.
static int predictor_step(const tocvp_pred_weights& w, const PredBuffers& pb, int B, int L, int f0, int n, int ftot,
                          cudaStream_t st) {
  const int S = w.num_slots, D = w.slot_dim, T = w.token_dim;
  const int M = B * n * S;
  const size_t seq_stride = size_t(ftot) * S * D;
  TOCVP_CUDA(launch_pdl(window_to_f16_kernel, dim3(ew_grid(size_t(M) * D / 4)), dim3(256), 0, st, pb.frames, seq_stride,
                        f0, n, S * D, B, pb.tok16));
  count_launch();
  // tokens = mlp_in(slots) + flip(pe[:n])   (text_cond_OCVP.py:86-91, model_blocks.py:375-377): the pre-flipped
  // table for window length n is added in the GEMM epilogue, row -> frame index (row / S) % n.
  const float* pe_n = w.pe_flipped + size_t(n - 1) * w.buffer_size * T;
  TOCVP_TRY(gemm_f16(pb.tok16, D, static_cast<const __half*>(w.mlp_in_w), D, M, T, D, w.mlp_in_b, 0, pe_n, T, S, n,
                     pb.x32, T, nullptr, 0, st));
  // LayerNorm folding (>= 1024 rows): the GEMM that writes a residual-stream tensor also emits its f16 copy (h16) and
  // per-row partial [sum, sumsq] (stats); the projection that would consume LN(.) reads the raw copy with gamma-folded
  // weights and finishes the normalisation in its epilogue.  4 LayerNorm launches per layer and their fp32 re-reads
  // disappear.
  const float inv_t = 1.f / float(T);
  const int slots = T / 64;
  const GemmLn prod{nullptr, 0, nullptr, 0.f, 0.f, pb.stats};

  // Everything of a layer behind the self-attention, on Mr rows = B sequences x tq rows each (att16 holds their attention
  // output, xres their residual-stream rows): out-proj, cross-attention block, MLP.  Result in x32[0:Mr].
  auto layer_tail = [&](const tocvp_pred_layer& ly, int l, int Mr, int tq, const float* xres, bool emit_next) -> int {
    const bool fold = gemm_ln_supported(Mr, T);
    if (fold) {
      set_next_tile_order(1);   // out-proj: the attention wrote its last rows last
      TOCVP_TRY(gemm_f16_ln(pb.att16, T, static_cast<const __half*>(ly.w_o), T, Mr, T, T, nullptr, 0, xres, T, pb.y32, T,
                            pb.h16, T, prod, st));
      // ---- z = y + CrossAttn(LN(text), LN(y))                      (attention.py:445-463)
      const GemmLn c{pb.stats, slots, ly.c_cq, inv_t, w.ln_eps, nullptr};
      set_next_tile_order(0);
      TOCVP_TRY(gemm_f16_ln(pb.h16, T, static_cast<const __half*>(ly.wc_q_f), T, Mr, T, T, ly.d_cq, 0, nullptr, 0, nullptr,
                            0, pb.q16, T, c, st));
    } else {
      TOCVP_TRY(gemm_f16(pb.att16, T, static_cast<const __half*>(ly.w_o), T, Mr, T, T, nullptr, 0, xres, T, 1, 0, pb.y32, T,
                         nullptr, 0, st));
      TOCVP_TRY(layernorm(pb.y32, 0, T, nullptr, 0, ly.ln_cq_g, ly.ln_cq_b, w.ln_eps, Mr, T, pb.h16, T, nullptr, 0, st));
      TOCVP_TRY(gemm_f16(pb.h16, T, static_cast<const __half*>(ly.wc_q), T, Mr, T, T, nullptr, 0, nullptr, 0, 1, 0, nullptr,
                         0, pb.q16, T, st));
    }
    const __half* kv = pb.kv16 + size_t(l) * B * L * 2 * T;
    set_next_tile_order(1);
    TOCVP_TRY(mha_f16(pb.q16, T, kv, kv + T, 2 * T, B, tq, L, w.cross_heads, pb.att16, T, st));
    if (fold) {
      set_next_tile_order(0);
      TOCVP_TRY(gemm_f16_ln(pb.att16, T, static_cast<const __half*>(ly.wc_o), T, Mr, T, T, ly.bc_o, 0, pb.y32, T, pb.z32, T,
                            pb.h16, T, prod, st));
      // ---- z = z + MLP_c(LN(z))
      const GemmLn c{pb.stats, slots, ly.c_c1, inv_t, w.ln_eps, nullptr};
      set_next_tile_order(1);
      TOCVP_TRY(gemm_f16_ln(pb.h16, T, static_cast<const __half*>(ly.wc_1_f), T, Mr, w.cross_hidden, T, ly.d_c1, 1, nullptr,
                            0, nullptr, 0, pb.mid16, w.cross_hidden, c, st));
      set_next_tile_order(0);
      TOCVP_TRY(gemm_f16_ln(pb.mid16, w.cross_hidden, static_cast<const __half*>(ly.wc_2), w.cross_hidden, Mr, T,
                            w.cross_hidden, ly.bc_2, 0, pb.z32, T, nullptr, 0, pb.h16, T, prod, st));   // fp32 z is dead: the skip below is y
      // ---- out = y + MLP(LN(z))   (skip is y, attention.py:521-523)
      const GemmLn c2{pb.stats, slots, ly.c_1, inv_t, w.ln_eps, nullptr};
      set_next_tile_order(1);
      TOCVP_TRY(gemm_f16_ln(pb.h16, T, static_cast<const __half*>(ly.w_1_f), T, Mr, w.hidden_dim, T, ly.d_1, 1, nullptr, 0,
                            nullptr, 0, pb.mid16, w.hidden_dim, c2, st));
      set_next_tile_order(0);
      if (!emit_next) {
        TOCVP_TRY(gemm_f16(pb.mid16, w.hidden_dim, static_cast<const __half*>(ly.w_2), w.hidden_dim, Mr, T, w.hidden_dim,
                           ly.b_2, 0, pb.y32, T, 1, 0, pb.x32, T, nullptr, 0, st));
      } else {   // the next layer's QKV projection consumes the f16 copy + statistics of x
        TOCVP_TRY(gemm_f16_ln(pb.mid16, w.hidden_dim, static_cast<const __half*>(ly.w_2), w.hidden_dim, Mr, T, w.hidden_dim,
                              ly.b_2, 0, pb.y32, T, pb.x32, T, pb.h16, T, prod, st));
      }
    } else {
      TOCVP_TRY(gemm_f16(pb.att16, T, static_cast<const __half*>(ly.wc_o), T, Mr, T, T, ly.bc_o, 0, pb.y32, T, 1, 0, pb.z32,
                         T, nullptr, 0, st));
      TOCVP_TRY(layernorm(pb.z32, 0, T, nullptr, 0, ly.ln_cm_g, ly.ln_cm_b, w.ln_eps, Mr, T, pb.h16, T, nullptr, 0, st));
      TOCVP_TRY(gemm_f16(pb.h16, T, static_cast<const __half*>(ly.wc_1), T, Mr, w.cross_hidden, T, ly.bc_1, 1, nullptr, 0, 1,
                         0, nullptr, 0, pb.mid16, w.cross_hidden, st));
      TOCVP_TRY(gemm_f16(pb.mid16, w.cross_hidden, static_cast<const __half*>(ly.wc_2), w.cross_hidden, Mr, T,
                         w.cross_hidden, ly.bc_2, 0, pb.z32, T, 1, 0, pb.z32, T, nullptr, 0, st));
      TOCVP_TRY(layernorm(pb.z32, 0, T, nullptr, 0, ly.ln_m_g, ly.ln_m_b, w.ln_eps, Mr, T, pb.h16, T, nullptr, 0, st));
      TOCVP_TRY(gemm_f16(pb.h16, T, static_cast<const __half*>(ly.w_1), T, Mr, w.hidden_dim, T, ly.b_1, 1, nullptr, 0, 1, 0,
                         nullptr, 0, pb.mid16, w.hidden_dim, st));
      TOCVP_TRY(gemm_f16(pb.mid16, w.hidden_dim, static_cast<const __half*>(ly.w_2), w.hidden_dim, Mr, T, w.hidden_dim,
                         ly.b_2, 0, pb.y32, T, 1, 0, pb.x32, T, nullptr, 0, st));
    }
    return TOCVP_OK;
  };

  // Tile order: consecutive kernels of a layer walk the rows in opposite directions (set_next_tile_order, host_util.h) --
  // QKV down, attention up, out-proj down, cross-q up, cross-attention down, cross-out up, MLP_c up-proj down, down-proj up,
  // MLP up-proj down, down-proj up -- so each starts with what its producer wrote last.
  const bool fold_all = gemm_ln_supported(M, T);
  int n_out = n;                            // frames held by x32 after the last layer (1 when it was pruned)
  for (int l = 0; l < w.num_layers; ++l) {
    const tocvp_pred_layer& ly = w.layers[l];
    const bool last_layer = (l == w.num_layers - 1);
    // ---- y = x + MHSA(LN(x))                                       (attention.py:512-514)
    if (fold_all && l > 0) {
      const GemmLn c{pb.stats, slots, ly.c_qkv, inv_t, w.ln_eps, nullptr};
      set_next_tile_order(1);   // the previous layer's last GEMM walked ascending
      TOCVP_TRY(gemm_f16_ln(pb.h16, T, static_cast<const __half*>(ly.w_qkv_f), T, M, 3 * T, T, ly.d_qkv, 0, nullptr, 0,
                            nullptr, 0, pb.qkv16, 3 * T, c, st));
    } else {
      TOCVP_TRY(layernorm(pb.x32, 0, T, nullptr, 0, ly.ln_q_g, ly.ln_q_b, w.ln_eps, M, T, pb.h16, T, nullptr, 0, st));
      TOCVP_TRY(gemm_f16(pb.h16, T, static_cast<const __half*>(ly.w_qkv), T, M, 3 * T, T, nullptr, 0, nullptr, 0, 1, 0,
                         nullptr, 0, pb.qkv16, 3 * T, st));
    }
    if (last_layer && n > 1) {
      // mlp_out reads only the newest frame's tokens (text_cond_OCVP.py:103) and nothing else consumes this layer's
      // output: behind the K/V projection only the S newest rows of every sequence are computed (exact, not an
      // approximation) -- queries, out-proj, cross-attention and both MLPs on B*S rows instead of B*n*S.
      const int Mc = B * S;
      TOCVP_CUDA(launch_pdl(last_frame_f32_kernel, dim3(ew_grid(size_t(Mc) * T / 4)), dim3(256), 0, st, pb.x32, n, S, T,
                            B, pb.xl32));
      count_launch();
      TOCVP_TRY(mha_f16_sub(pb.qkv16 + size_t(n - 1) * S * 3 * T, 3 * T, n * S, pb.qkv16 + T, pb.qkv16 + 2 * T, 3 * T, B, S,
                            n * S, w.num_heads, pb.att16, T, st));
      TOCVP_TRY(layer_tail(ly, l, Mc, S, pb.xl32, false));
      n_out = 1;
    } else {
      set_next_tile_order(0);
      TOCVP_TRY(mha_f16(pb.qkv16, 3 * T, pb.qkv16 + T, pb.qkv16 + 2 * T, 3 * T, B, n * S, n * S, w.num_heads, pb.att16, T,
                        st));
      TOCVP_TRY(layer_tail(ly, l, M, n * S, pb.x32, fold_all && !last_layer));
    }
  }
  set_next_tile_order(0);
  // ---- mlp_out on the newest frame's tokens (text_cond_OCVP.py:103)
  TOCVP_CUDA(launch_pdl(last_frame_to_f16_kernel, dim3(ew_grid(size_t(B) * S * T / 4)), dim3(256), 0, st, pb.x32, n_out,
                        S, T, B, pb.last16));
  count_launch();
  TOCVP_TRY(gemm_f16(pb.last16, T, static_cast<const __half*>(w.mlp_out_w), T, B * S, D, T, w.mlp_out_b, 0, nullptr, 0,
                     1, 0, pb.pred32, D, nullptr, 0, st));
  return TOCVP_OK;
}

static int check_weights(const tocvp_pred_weights& w) {
  TOCVP_CHECK_ARG(w.num_layers > 0 && w.layers != nullptr);
  TOCVP_CHECK_ARG(w.token_dim % 64 == 0 && w.token_dim / w.num_heads == 64 && w.token_dim / w.cross_heads == 64);
  TOCVP_CHECK_ARG(w.slot_dim % 8 == 0 && w.hidden_dim % 8 == 0 && w.cross_hidden % 8 == 0);
  TOCVP_CHECK_ARG(w.buffer_size >= 1 && w.num_slots * w.buffer_size <= 128);
  TOCVP_CHECK_ARG(w.mlp_in_w && w.mlp_in_b && w.mlp_out_w && w.mlp_out_b && w.pe_flipped);
  return TOCVP_OK;
}

static int hoist_text_kv(const tocvp_pred_weights& w, const PredBuffers& pb, const float* text, int B, int L,
                         cudaStream_t st) {
  const int T = w.token_dim;
  TOCVP_CHECK_ARG(L >= 1 && L <= 128);
  for (int l = 0; l < w.num_layers; ++l) {
    const tocvp_pred_layer& ly = w.layers[l];
    // ln_cross_att_kv is applied to the RAW text embeddings in every layer (attention.py:454)
    TOCVP_TRY(layernorm(text, 0, T, nullptr, 0, ly.ln_ckv_g, ly.ln_ckv_b, w.ln_eps, B * L, T, pb.text16, T, nullptr, 0, st));
    TOCVP_TRY(gemm_f16(pb.text16, T, static_cast<const __half*>(ly.wc_kv), T, B * L, 2 * T, T, nullptr, 0, nullptr, 0, 1,
                       0, nullptr, 0, pb.kv16 + size_t(l) * B * L * 2 * T, 2 * T, st));
  }
  return TOCVP_OK;
}

}  // namespace tocvp

using namespace tocvp;

extern "C" size_t tocvp_sizeof_pred_weights(void) { return sizeof(tocvp_pred_weights); }
extern "C" size_t tocvp_sizeof_pred_layer(void) { return sizeof(tocvp_pred_layer); }

extern "C" size_t tocvp_predictor_workspace_bytes(const tocvp_pred_weights* w, int B, int L, int num_context,
                                                  int num_preds) {
  if (!w) return 0;
  return predictor_workspace_bytes(*w, B, L, num_context, num_preds);
}

extern "C" int tocvp_predictor_rollout(const tocvp_pred_weights* w, const float* slot_history, size_t hist_seq_stride,
                                       const float* text, int B, int L, int num_context, int num_preds,
                                       float* pred_slots, void* workspace, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TOCVP_CHECK_ARG(w && slot_history && text && pred_slots && workspace && B > 0 && num_context >= 1 && num_preds >= 1);
  OptsScope scope(w->tuning);
  TOCVP_TRY(check_weights(*w));
  if (ws_bytes < predictor_workspace_bytes(*w, B, L, num_context, num_preds)) {
    set_last_error(__FILE__, __LINE__, "predictor_rollout: workspace too small");
    return TOCVP_ERR_WORKSPACE;
  }
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  PredBuffers pb;
  carve(*w, B, L, num_context, num_preds, &pb, static_cast<uint8_t*>(workspace));
  const int S = w->num_slots, D = w->slot_dim, SD = S * D;
  const int ftot = num_context + num_preds;
  const size_t seq_stride = size_t(ftot) * SD;
  TOCVP_CUDA(launch_pdl(copy_context_kernel, dim3(ew_grid(size_t(B) * num_context * SD)), dim3(256), 0, st,
                        slot_history, hist_seq_stride, pb.frames, seq_stride, num_context, SD, B));
  count_launch();
  TOCVP_TRY(hoist_text_kv(*w, pb, text, B, L, st));
  for (int t = 0; t < num_preds; ++t) {
    const int have = num_context + t;                                   // frames available
    const int n = have < w->buffer_size ? have : w->buffer_size;        // predictor_wrapper.py:143-153
    const int f0 = have - n;
    TOCVP_TRY(predictor_step(*w, pb, B, L, f0, n, ftot, st));
    TOCVP_CUDA(launch_pdl(commit_prediction_kernel, dim3(ew_grid(size_t(B) * SD)), dim3(256), 0, st, pb.pred32,
                          pb.frames, seq_stride, have - 1, have, w->residual, pred_slots, size_t(num_preds) * SD, t, SD,
                          B));
    count_launch();
  }
  return TOCVP_OK;
}

// Single BaseTextOCVP.forward: slots [B,n,S,D] fp32, text [B,L,T] fp32 -> out [B,S,D] fp32.
extern "C" int tocvp_predictor_forward(const tocvp_pred_weights* w, const float* slots, const float* text, int B, int n,
                                       int L, float* out, void* workspace, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TOCVP_CHECK_ARG(w && slots && text && out && workspace && B > 0 && n >= 1);
  OptsScope scope(w->tuning);
  TOCVP_TRY(check_weights(*w));
  TOCVP_CHECK_ARG(n <= w->buffer_size);
  if (ws_bytes < predictor_workspace_bytes(*w, B, L, n, 1)) {
    set_last_error(__FILE__, __LINE__, "predictor_forward: workspace too small");
    return TOCVP_ERR_WORKSPACE;
  }
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  PredBuffers pb;
  carve(*w, B, L, n, 1, &pb, static_cast<uint8_t*>(workspace));
  const int SD = w->num_slots * w->slot_dim;
  const size_t seq_stride = size_t(n + 1) * SD;
  TOCVP_CUDA(launch_pdl(copy_context_kernel, dim3(ew_grid(size_t(B) * n * SD)), dim3(256), 0, st, slots, size_t(n) * SD,
                        pb.frames, seq_stride, n, SD, B));
  count_launch();
  TOCVP_TRY(hoist_text_kv(*w, pb, text, B, L, st));
  TOCVP_TRY(predictor_step(*w, pb, B, L, 0, n, n + 1, st));
  TOCVP_CUDA(launch_pdl(commit_prediction_kernel, dim3(ew_grid(size_t(B) * SD)), dim3(256), 0, st, pb.pred32, pb.frames,
                        seq_stride, n - 1, n, w->residual, out, size_t(SD), 0, SD, B));
  count_launch();
  return TOCVP_OK;
}
