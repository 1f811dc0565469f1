// Sibling predictors that share PredictorWrapper's autoregressive loop: VanillaTransformerPredictor
// (reference src/models/Predictors/OCVP.py:24-141) and OCVPSeq / OCVPSeqLayer (OCVP.py:145-319): mlp_in -> sinusoidal
// time encoding shared by the slots of a frame (SlotPositionalEncoding, src/models/Blocks/model_blocks.py:230-290) ->
// pre-norm torch.nn.TransformerEncoderLayer blocks (ReLU, eps 1e-5) -> mlp_out on the newest frame (+ residual).
// Vanilla attends over all n*S tokens; an OCVP-Seq layer is an OBJECT block (attention inside each frame) followed by a
// TIME block (attention along each slot's history): the same token array with two different key groups.  OCVP-Par
// (OCVP.py:324-548) applies both attentions to the same normed input and adds them (block flags).
//
// One prediction step of one sequence runs in ONE CTA, fp32 end to end: <= 80 tokens x 128 features is a latency
// problem.  Token t = frame * S + slot.
#include "host_util.h"
#include "ptx.cuh"
#include "small_tf.cuh"

namespace tocvp {

constexpr int OC_MAXT = 80;

__global__ void __launch_bounds__(TE_THREADS, 1)
ocvp_step_kernel(tocvp_ocvp_weights w, const float* __restrict__ slots, size_t seq_stride, int n, float* __restrict__ out) {
  extern __shared__ float sm[];
  const int S = w.num_slots, Ds = w.slot_dim, D = w.token_dim, F = w.ffn_dim, H = w.num_heads;
  const int T = n * S;
  const int ldq = 3 * D + 1;
  float* x = sm;                        // [T][D] residual stream
  float* h = x + OC_MAXT * D;           // [T][D] normed input / attention output
  float* big = h + OC_MAXT * D;         // [T][3D+1] qkv, [T][F] FFN hidden, [T][Ds] input slots
  const float* in = slots + size_t(blockIdx.x) * seq_stride;
  for (int e = threadIdx.x; e < T * Ds; e += TE_THREADS) big[e] = in[e];
  __syncthreads();
  // tokens = mlp_in(slots) + pe[frame]   (OCVP.py:117-123; model_blocks.py:286-288: frame f of the window gets pe[f])
  te_linear(w.mlp_in_w_t, w.mlp_in_b, Ds, D, big, Ds, x, D, T, 0, nullptr, 0);
  for (int e = threadIdx.x; e < T * D; e += TE_THREADS) x[e] += __ldg(w.pe + size_t(e / (S * D)) * D + e % D);
  __syncthreads();
  for (int l = 0; l < w.num_blocks; ++l) {
    const tocvp_text_layer& ly = w.blocks[l];
    // pre-norm encoder layer: x = x + SA(LN1(x)); x = x + FFN(LN2(x)).  The attention output overwrites the q columns of the
    // packed qkv rows (a thread rewrites only the q it has read), so h keeps LN1(x): an OCVP-Par layer (OCVP.py:499-546,
    // x + SA_obj(LN1 x) + SA_time(LN1 x)) is two blocks -- the second reuses h (flag bit 1), the first skips the FFN (bit 0).
    const int flags = w.block_flags[l];
    if (!(flags & 2)) te_layernorm_to(x, h, D, T, D, ly.ln1_g, ly.ln1_b, 1e-5f, nullptr);
    te_linear(ly.in_w_t, ly.in_b, D, 3 * D, h, D, big, ldq, T, 0, nullptr, 0);
    te_attention(big, ldq, big, ldq, T, T, D, H, w.block_group[l], S);
    te_linear(ly.out_w_t, ly.out_b, D, D, big, ldq, x, D, T, 0, x, D);
    if (flags & 1) continue;
    te_layernorm_to(x, h, D, T, D, ly.ln2_g, ly.ln2_b, 1e-5f, nullptr);
    te_linear(ly.ff1_w_t, ly.ff1_b, D, F, h, D, big, F, T, 2, nullptr, 0);
    te_linear(ly.ff2_w_t, ly.ff2_b, F, D, big, F, x, D, T, 0, x, D);
  }
  // output = mlp_out(tokens of the newest frame) (+ slots[:, -1])   (OCVP.py:132-134)
  const float* xl = x + size_t(n - 1) * S * D;
  const float* sl = in + size_t(n - 1) * S * Ds;
  float* ob = out + size_t(blockIdx.x) * S * Ds;
  for (int e = threadIdx.x; e < S * Ds; e += TE_THREADS) {
    const int r = e / Ds, c = e - r * Ds;
    float acc = __ldg(w.mlp_out_b + c);
    for (int k = 0; k < D; ++k) acc = fmaf(__ldg(w.mlp_out_w_t + size_t(k) * Ds + c), xl[r * D + k], acc);
    ob[e] = w.residual ? acc + sl[e] : acc;
  }
}

}  // namespace tocvp

using namespace tocvp;

extern "C" size_t tocvp_sizeof_ocvp_weights(void) { return sizeof(tocvp_ocvp_weights); }

// slots fp32: sequence b's window [n, S, slot_dim] at slots + b*seq_stride (floats); out fp32 [B, S, slot_dim].
extern "C" int tocvp_ocvp_forward(const tocvp_ocvp_weights* w, const float* slots, size_t seq_stride, int B, int n,
                                  float* out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TOCVP_CHECK_ARG(w && slots && out && B > 0 && n >= 1 && n <= w->max_len);
  TOCVP_CHECK_ARG(w->num_blocks >= 1 && w->num_blocks <= TOCVP_OCVP_MAX_BLOCKS && n * w->num_slots <= OC_MAXT);
  TOCVP_CHECK_ARG(w->token_dim <= 128 && w->slot_dim <= 3 * w->token_dim && w->ffn_dim <= 3 * w->token_dim &&
                  w->token_dim % w->num_heads == 0 && w->token_dim / w->num_heads <= 64);
  const size_t smem = size_t(OC_MAXT) * (2 * w->token_dim + 3 * w->token_dim + 1) * sizeof(float);
  TOCVP_CHECK_ARG(smem <= 220 * 1024);
  static SmemAttrOnce attr_once;   // opt in to the device maximum once per device; the launch passes the actual size
  TOCVP_TRY(ensure_smem_attr(attr_once, ocvp_step_kernel, -1));
  ocvp_step_kernel<<<B, TE_THREADS, smem, st>>>(*w, slots, seq_stride, n, out);
  TOCVP_LAUNCHED();
  return TOCVP_OK;
}
