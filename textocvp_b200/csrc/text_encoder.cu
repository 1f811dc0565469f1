// TransformerTextEncoder.forward (reference src/models/EncodersDecoders/text_encoders.py:84-124): token + position
// embeddings -> LayerNorm(eps 1e-8) -> zero the padding tokens -> num_layers x torch.nn.TransformerEncoderLayer
// (post-norm, GELU, key-padding mask from the caption lengths; built at :42-50) -> LayerNorm -> Linear(input_dim,
// output_dim).  It feeds PredictorWrapper.encode_text_caption (src/models/Predictors/predictor_wrapper.py:90-127) once
// per rollout, immediately before the hot path.
//
// The whole encoder of one caption runs in ONE CTA, fp32 end to end (it is <= 64 tokens x 128 features: a latency
// problem, not a throughput one): the residual stream, the packed QKV / FFN activations live in shared memory, weights
// are read transposed ([in][out], coalesced across the threads that own output columns) from L2, attention keeps an
// online softmax per (head, query) thread.
#include "host_util.h"
#include "ptx.cuh"
#include "small_tf.cuh"

namespace tocvp {

__global__ void __launch_bounds__(TE_THREADS, 1)
text_encoder_kernel(tocvp_text_weights w, const long long* __restrict__ tokens, const long long* __restrict__ lengths,
                    int L, float* __restrict__ out) {
  extern __shared__ float sm[];
  const int D = w.input_dim, F = w.ffn_dim, H = w.num_heads;
  const int ldq = 3 * D + 1;                                   // +1: query rows land in different banks
  float* x = sm;                                               // [L][D] residual stream
  float* big = x + TE_MAXL * D;                                // [L][3D+1] qkv, later [L][F] FFN hidden
  float* att = big + TE_MAXL * (F > ldq ? F : ldq);            // [L][D]
  __shared__ unsigned char keep[TE_MAXL];
  const int b = blockIdx.x;
  const long long* tok = tokens + size_t(b) * L;
  const int len = int(lengths[b]);
  // ---- embeddings (text_encoders.py:100-103), padding-token mask (:107-108)
  for (int e = threadIdx.x; e < L * D; e += TE_THREADS) {
    const int r = e / D, c = e - r * D;
    long long t = tok[r];
    t = t < 0 ? 0 : (t >= w.vocab_size ? w.vocab_size - 1 : t);
    x[r * D + c] = __ldg(w.tok_emb + size_t(t) * D + c) + __ldg(w.pos_emb + size_t(r) * D + c);
  }
  if (threadIdx.x < L) keep[threadIdx.x] = tok[threadIdx.x] != 0;
  __syncthreads();
  te_layernorm(x, D, L, D, w.ln0_g, w.ln0_b, 1e-8f, keep);
  // ---- encoder layers (post-norm): x = LN1(x + SA(x)); x = LN2(x + FFN(x))
  for (int l = 0; l < w.num_layers; ++l) {
    const tocvp_text_layer& ly = w.layers[l];
    te_linear(ly.in_w_t, ly.in_b, D, 3 * D, x, D, big, ldq, L, 0, nullptr, 0);
    // attention over the keys j < len (src_key_padding_mask, text_encoders.py:110)
    te_attention(big, ldq, att, D, L, len < L ? len : L, D, H, 0, 1);
    te_linear(ly.out_w_t, ly.out_b, D, D, att, D, x, D, L, 0, x, D);          // x += out_proj(att)   (residual in place)
    te_layernorm(x, D, L, D, ly.ln1_g, ly.ln1_b, 1e-5f, nullptr);
    te_linear(ly.ff1_w_t, ly.ff1_b, D, F, x, D, big, F, L, 1, nullptr, 0);
    te_linear(ly.ff2_w_t, ly.ff2_b, F, D, big, F, x, D, L, 0, x, D);
    te_layernorm(x, D, L, D, ly.ln2_g, ly.ln2_b, 1e-5f, nullptr);
  }
  // ---- text_out_projection: LayerNorm -> Linear(D, out_dim)  (text_encoders.py:66-69, 122)
  te_layernorm(x, D, L, D, w.lnf_g, w.lnf_b, 1e-5f, nullptr);
  float* ob = out + size_t(b) * L * w.output_dim;
  for (int n = threadIdx.x; n < w.output_dim; n += TE_THREADS) {
    const float bv = __ldg(w.proj_b + n);
    for (int r0 = 0; r0 < L; r0 += TE_RB) {
      float acc[TE_RB];
#pragma unroll
      for (int r = 0; r < TE_RB; ++r) acc[r] = bv;
      for (int k = 0; k < D; ++k) {
        const float wv = __ldg(w.proj_w_t + size_t(k) * w.output_dim + n);
#pragma unroll
        for (int r = 0; r < TE_RB; ++r) acc[r] = fmaf(wv, x[(r0 + r) * D + k], acc[r]);
      }
#pragma unroll
      for (int r = 0; r < TE_RB; ++r)
        if (r0 + r < L) ob[size_t(r0 + r) * w.output_dim + n] = acc[r];
    }
  }
}

}  // namespace tocvp

using namespace tocvp;

extern "C" size_t tocvp_sizeof_text_weights(void) { return sizeof(tocvp_text_weights); }

extern "C" int tocvp_text_encode(const tocvp_text_weights* w, const long long* tokens, const long long* lengths, int B,
                                 int L, float* out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TOCVP_CHECK_ARG(w && tokens && lengths && out && B > 0 && L >= 1 && L <= TE_MAXL && L <= w->context_length);
  TOCVP_CHECK_ARG(w->num_layers >= 1 && w->num_layers <= TOCVP_TEXT_MAX_LAYERS && w->input_dim % w->num_heads == 0);
  TOCVP_CHECK_ARG(w->input_dim / w->num_heads <= 64 && w->input_dim <= 256 && w->ffn_dim <= 1024);
  const int D = w->input_dim, F = w->ffn_dim;
  const int ldq = 3 * D + 1;
  const size_t smem = size_t(TE_MAXL) * (D + (F > ldq ? F : ldq) + D) * sizeof(float);
  TOCVP_CHECK_ARG(smem <= 220 * 1024);
  static SmemAttrOnce attr_once;   // opt in to the device maximum once per device; the launch passes the actual size
  TOCVP_TRY(ensure_smem_attr(attr_once, text_encoder_kernel, -1));
  text_encoder_kernel<<<B, TE_THREADS, smem, st>>>(*w, tokens, lengths, L, out);
  TOCVP_LAUNCHED();
  return TOCVP_OK;
}
