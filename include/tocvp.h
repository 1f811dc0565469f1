/*
 * tocvp.h -- C ABI of libtocvp.so: hand-written sm_100a kernels for the TextOCVP rollout path.
 *
 * The reference (angelvillar96/TextOCVP) is pure Python/PyTorch and has no FFI layer; its
 * "operator interface" for this path is the set of nn.Module.forward signatures listed in
 * SURVEY.md 8(b).  The entry points below sit directly under those signatures -- the Python
 * modules in textocvp_b200/ (same class names, constructor arguments, forward signatures
 * and state_dict keys as the reference) call them through ctypes.  Each declaration cites
 * the reference code it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - plain C types only: device pointers, sizes, a cudaStream_t passed as void*;
 *   - every call is asynchronous on the given stream and never allocates: outputs and
 *     workspaces are caller-owned (tocvp_*_workspace_bytes reports the size needed);
 *   - return value: TOCVP_OK (0) or a negative error code; tocvp_last_error() returns a
 *     thread-local message.  Nothing throws, nothing calls exit();
 *   - no process-wide tuning state: kernel-selection options travel with each call (tocvp_tuning below, caller-owned,
 *     NULL = defaults).  What the library keeps across calls: per-device caches of immutable attributes (SM count, the
 *     dynamic-shared-memory attribute of each kernel), the per-device internal side stream + events of tocvp_savi_decode
 *     (created on first use, mutex-guarded), a thread-local error string and a launch counter (statistics).  Calls are
 *     re-entrant across threads and streams and run on the caller's CURRENT device (the library never calls
 *     cudaSetDevice);
 *   - fp32 tensors are 16-byte aligned; fp16 tensors ("f16" below) are IEEE binary16, 16-byte
 *     aligned, row-major with the channel / feature dimension innermost (NHWC for images);
 *   - there is NO CPU fallback: on anything but an sm_100 device tocvp_init() fails.
 */
#ifndef TOCVP_H_
#define TOCVP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOCVP_ABI_VERSION 3

#define TOCVP_OK 0
#define TOCVP_ERR_BAD_ARG (-1)   /* bad shape / null or misaligned pointer / unsupported size */
#define TOCVP_ERR_CUDA (-2)      /* a CUDA runtime call or kernel launch failed                */
#define TOCVP_ERR_ARCH (-3)      /* device is not sm_100                                       */
#define TOCVP_ERR_WORKSPACE (-4) /* workspace too small                                         */

int tocvp_abi_version(void);
/* Checks that `device` is an sm_100 part (does not change the current device). */
int tocvp_init(int device);
const char* tocvp_last_error(void);
/* Statistics: number of kernels this library has launched so far in this process (monotonic). */
unsigned long long tocvp_kernel_launches(void);
/* For callers that replay captured library calls from a CUDA graph: adds the graph's kernel count to the statistic. */
void tocvp_note_graph_replay(unsigned long long n_kernels);

/* ------------------------------------------------------------------------------------------
 * Per-call kernel-selection options (tests, A/B measurements).  All-zero = the defaults a product caller gets with
 * tuning == NULL.  The struct is caller-owned and read during the call only.
 * ------------------------------------------------------------------------------------------ */
typedef struct tocvp_tuning {
  int gemm_mode;      /* 0 = automatic choice between the single-CTA and the CTA-pair (cta_group::2) GEMM kernels, 1 =
                         single-CTA kernel only, 128 / 256 = pair kernel with that tile width wherever applicable */
  int gemm_no_wres;   /* 1 = W-resident variant of the 256-wide pair kernel (K <= 512) off */
  int conv_mode;      /* 0 = CTA-pair conv5x5 kernel whenever the tile count is even, 1 = single-CTA kernel only */
  int encode_mode;    /* bit mask: bit 0 = first-version fp32 SIMT conv 1, bit 1 = separate posemb + LayerNorm pass
                         (default: fused into conv 4's epilogue), bit 2 = the 32 -> 128 -> 128 MLP as two GEMMs (default:
                         one kernel with two chained tcgen05 GEMMs when only f16 features are requested), bit 3 = conv 1 as
                         25 taps over zero-padded channels (default: x-taps folded into K, 3 taps), bit 4 = the 32 -> 32
                         layers on the 25-tap N = 32 kernel (default: pixel-pair kernel, N = 64, when w_conv_xp is set) */
  int decode_mode;    /* bit mask: bit 0 = decoder layer 1 generated inside the layer-2 convolution kernel (default:
                         separate bandwidth kernel), bit 1 = first-version head conv3x3 (shifted windows, N = 16; default:
                         nine taps in the GEMM's N dimension), bit 2 = first-version (image-stationary) layer-1 kernel,
                         bit 3 = serial chunks (default: chunk-pipelined, layer 1 of chunk i+1 on an internal side stream
                         under the convolutions of chunk i), bit 4 = separate compositing kernel (default: fused into the
                         head convolution's epilogue) */
  int corrector_mode; /* per-slot update kernel: 0 = 3xTF32 mma.sync with the weights streamed through a shared-memory ring
                         (default), 2 = first tensor-core version (weight fragments read from L2), 1 = fp32 SIMT loops */
  int no_pdl;         /* 1 = plain stream order (default: programmatic dependent launch, bit-identical results) */
  int no_tile_alternation; /* 1 = every predictor kernel walks its row blocks ascending (default: alternating, so a
                         consumer starts with the rows its producer wrote last; bit-identical results) */
  int reserved[8];
} tocvp_tuning;
size_t tocvp_sizeof_tuning(void);

/* ------------------------------------------------------------------------------------------
 * Dense projection: C[M,N] = A[M,K] . W[N,K]^T (+bias) (activation) (+residual), tcgen05 / TMEM / TMA.
 * A, W: f16, K innermost (W is torch.nn.Linear.weight as stored).  bias fp32[N] or NULL.  relu: activation code, 0 = none,
 * 1 = ReLU, 2 = exact (erf) GELU.  residual fp32 [M, ldr] or NULL (added after the activation).  Writes out_f32 and/or out_f16 (either may
 * be NULL, not both).  N, K, lda, ldw multiples of 8.
 * Replaces nn.Linear -> cuBLAS at src/models/Blocks/attention.py:167-175, 296-300, 352-356,
 * src/models/Predictors/text_cond_OCVP.py:47-48, src/models/SAVi.py:117-119.
 * ------------------------------------------------------------------------------------------ */
int tocvp_gemm_f16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                   int relu, const float* residual, int ldr, float* out_f32, int ld32, void* out_f16, int ld16,
                   const tocvp_tuning* tuning, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row LayerNorm (fp32 statistics): y = LN(x (+ add[row % add_rows])) * gamma + beta.
 * x: fp32 or f16 [rows, ldx]; add: optional fp32 table [add_rows, D] (the batch-independent
 * SoftPositionEmbed projection, src/models/Blocks/model_blocks.py:215-226); D % 4 == 0, D <= 1024.
 * Replaces nn.LayerNorm at src/models/Blocks/attention.py:49-51, 361-362, 427, 435-436 and
 * src/models/SAVi.py:116.
 * ------------------------------------------------------------------------------------------ */
int tocvp_layernorm(const void* x, int x_is_f16, int ldx, const float* add, int add_rows, const float* gamma,
                    const float* beta, float eps, int rows, int D, void* out_f16, int ld16, float* out_f32,
                    int ld32, void* stream);

/* ------------------------------------------------------------------------------------------
 * 5x5 stride-1 zero-padded convolution + bias (+ReLU) as a tcgen05 implicit GEMM.
 * x: f16 NHWC [n_img,H,W,cin]; w_packed: f16 [25,cout,cin] (tap-major, tap = ky*5+kx, i.e.
 * torch weight[co,ci,ky,kx] permuted); bias fp32[cout]; out: f16 NHWC [n_img,H,W,cout].
 * H % 16 == 0, W % 32 == 0; (cin,cout) in {(64,64),(32,32)}; x 16-byte and out 32-byte aligned (the CTA-pair
 * kernel stores with 256-bit accesses).
 * Replaces nn.Conv2d -> cuDNN at src/models/Blocks/model_blocks.py:83-91 as used by
 * src/models/EncodersDecoders/decoders.py:96-108 and encoders.py:141-153.
 * ------------------------------------------------------------------------------------------ */
int tocvp_conv5x5_f16(const void* x, const void* w_packed, const float* bias, void* out, int n_img, int H, int W,
                      int cin, int cout, int relu, const tocvp_tuning* tuning, void* stream);

/* ------------------------------------------------------------------------------------------
 * SAVi / ExtendedDINOSAUR corrector: SlotAttention.forward (src/models/Blocks/attention.py:67-112) for
 * num_slots x 128-d slots over N locations of 128-d features (tcgen05 streaming kernel for 8 slots and
 * N % 512 == 0 with f16 features, generic SIMT kernel otherwise), plus (optionally) the post-norm transition TransformerBlock
 * (attention.py:387-395, built at src/models/Blocks/transition_models.py:26-31).
 * All pointers are fp32 device arrays.  "_t" = transposed to [in][out]; wk is torch's [out][in].
 * ------------------------------------------------------------------------------------------ */
typedef struct tocvp_sa_weights {
  const float *ln_in_g, *ln_in_b, *ln_slot_g, *ln_slot_b, *ln_mlp_g, *ln_mlp_b; /* norm_input / norm_slot / norm_mlp */
  const float *wq_t, *bq, *wk, *bk, *wv_t, *bv;                                 /* to_q / to_k / to_v               */
  const float *w_ih_t, *w_hh_t, *b_ih, *b_hh;                                   /* gru (gate order r,z,n)           */
  const float *w1_t, *b1, *w2_t, *b2;                                           /* mlp.0 / mlp.2                    */
  const float *t_wq_t, *t_wk_t, *t_wv_t, *t_wo_t;                               /* transition attn q,k,v,out (no bias) */
  const float *t_ln1_g, *t_ln1_b, *t_ln2_g, *t_ln2_b;                           /* layernorm_query / layernorm_mlp  */
  const float *t_w1_t, *t_b1, *t_w2_t, *t_b2;                                   /* transition mlp.0 / mlp.2         */
  int mlp_hidden, t_heads, t_hidden;
  float attn_eps, ln_eps_sa, ln_eps_tf, scale; /* 1e-8, 1e-3, 1e-6, dim_feats^-0.5 */
  int num_slots;                               /* 4..11: 8 (SAVi.json) and 10 (ExtendedDINOSAUR.json) are the named ones */
  const tocvp_tuning* tuning;                  /* NULL = defaults */
  /* Second copies of the matrices above for the streaming update kernel (csrc/slot_attention_update.cu).  Each in-major
   * matrix W[K][N] is split into two IEEE f16 planes, hi = f16(W), lo = f16((W - hi) * 2048), and stored in mma.m16n8k16
   * B-fragment order: for k-step ks (16 rows), column tile nt (8 columns), lane l = g*4 + t: four 32-bit words
   * {b0_hi, b1_hi, b0_lo, b1_lo} with b0 = (W[ks*16 + 2t][nt*8 + g], W[ks*16 + 2t + 1][..]) (element 0 in the low half)
   * and b1 the same 8 rows further down: K * N words per matrix (modules._stream states it in torch).  Matrices are
   * concatenated in consumption order: stream_c = wv_t | w_ih_t | w_hh_t | w1_t | w2_t;  stream_t = t_wq_t | t_wk_t |
   * t_wv_t | t_wo_t | t_w1_t | t_w2_t;  stream_a = wq_t | wk.  NULL = not provided (the first-version kernel is used). */
  const float *stream_c, *stream_t, *stream_a;
} tocvp_sa_weights;

size_t tocvp_sizeof_sa_weights(void);
size_t tocvp_slot_attention_workspace_bytes(int B);
/* feats fp32 or f16: sequence b's [N,128] block starts at feats + b*feats_seq_stride (elements), so the features of
 * frame t inside a [B,T,N,128] encode batch are used in place; slots_in [B,S,128]; slots_out row b at
 * slots_out + b*out_stride (floats), so results land straight in slot_history[:, t]; pred_out (optional) =
 * transition(slots_out) [B,S,128].  iters = num_iterations_first for step 0, else num_iterations (attention.py:90). */
int tocvp_slot_attention(const tocvp_sa_weights* w, const void* feats, int feats_f16, size_t feats_seq_stride, int B,
                         int N,
                         const float* slots_in, int iters, float* slots_out, int out_stride, float* pred_out,
                         void* workspace, size_t ws_bytes, void* stream);

/* The corrector chain of SAVi.forward_decomp (reference src/models/SAVi.py:178-204) over n_frames consecutive frames in
 * one call: slots_t = SlotAttention(feats_t, cur, t == 0 ? iters_first : iters) -> slot_history + t*hist_frame_stride
 * (row b at + b*hist_seq_stride, floats); cur = transition(slots_t) (requires the TransformerBlock transition weights).
 * Frame t's features start at feats + t*feats_frame_stride (elements), sequence b's at + b*feats_seq_stride.
 * carry_out [B,S,128] receives transition(slots of the last frame) = the slots_in of a following call.  Pass
 * iters_first = iters when the first frame of this call is not step 0 of the video. */
size_t tocvp_slot_attention_seq_workspace_bytes(int B);
int tocvp_slot_attention_seq(const tocvp_sa_weights* w, const void* feats, int feats_f16, size_t feats_seq_stride,
                             size_t feats_frame_stride, int B, int N, int n_frames, int iters_first, int iters,
                             const float* slots_in, float* slot_history, size_t hist_seq_stride,
                             size_t hist_frame_stride, float* carry_out, void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Short-sequence multi-head attention (head dim 64, <= 128 keys, fp32 softmax, no mask).
 * q [B*Tq, ldq], k / v [B*Tk, ldkv] f16 with head h at column offset h*64; out [B*Tq, ldo] f16.
 * Replaces MetaAttention.attention + split/merge_heads, src/models/Blocks/attention.py:183-215.
 * ------------------------------------------------------------------------------------------ */
int tocvp_mha_f16(const void* q, int ldq, const void* k, const void* v, int ldkv, int B, int Tq, int Tk, int heads,
                  void* out, int ldo, void* stream);

/* ------------------------------------------------------------------------------------------
 * Text-conditioned predictor (TextOCVP body).  f16 weight matrices are torch Linear weights
 * [out,in] cast to f16; w_qkv = cat(attn.q, attn.k, attn.v), wc_kv = cat(cross_attn.k, cross_attn.v).
 * Layer = AdaptedEncoderBlock, src/models/Blocks/attention.py:504-524 (+ TransformerDecoderBlock :445-463).
 * ------------------------------------------------------------------------------------------ */
typedef struct tocvp_pred_layer {
  const float *ln_q_g, *ln_q_b;     /* layernorm_query                     */
  const void* w_qkv;                /* f16 [3T, T]                          */
  const void* w_o;                  /* f16 [T, T]   attn.out_projection.0   */
  const float *ln_cq_g, *ln_cq_b;   /* cross_attention.ln_cross_att_q       */
  const float *ln_ckv_g, *ln_ckv_b; /* cross_attention.ln_cross_att_kv      */
  const void* wc_q;                 /* f16 [T, T]                           */
  const void* wc_kv;                /* f16 [2T, T]                          */
  const void* wc_o;                 /* f16 [T, T]   cross_attn.out_projection (with bias) */
  const float* bc_o;
  const float *ln_cm_g, *ln_cm_b;   /* cross_attention.ln_mlp               */
  const void *wc_1, *wc_2;          /* f16 [Hc, T], [T, Hc]  cross_attention.mlp */
  const float *bc_1, *bc_2;
  const float *ln_m_g, *ln_m_b;     /* layernorm_mlp                        */
  const void *w_1, *w_2;            /* f16 [H, T], [T, H]    mlp            */
  const float *b_1, *b_2;
  /* LayerNorm folded into the projection that consumes it (used for M >= 1024 rows, where the CTA-pair GEMM applies
   * LN(x).W^T = rstd*(x.(W*gamma)^T - mu*c) + d in its epilogue from row statistics emitted by the producing GEMM):
   * *_f = f16(W * gamma[None,:]) with W the matrix above, c = row sums of *_f (fp32), d = W.beta (+ bias). */
  const void *w_qkv_f, *wc_q_f, *wc_1_f, *w_1_f;
  const float *c_qkv, *d_qkv, *c_cq, *d_cq, *c_c1, *d_c1, *c_1, *d_1;
} tocvp_pred_layer;

typedef struct tocvp_pred_weights {
  const tocvp_pred_layer* layers;   /* HOST array of num_layers entries     */
  int num_layers, num_slots, slot_dim, token_dim, hidden_dim, cross_hidden, num_heads, cross_heads;
  int buffer_size;                  /* input_buffer_size (max window, frames) */
  int residual;
  float ln_eps;                     /* 1e-6 */
  const void* mlp_in_w;             /* f16 [T, D] */
  const float* mlp_in_b;
  const void* mlp_out_w;            /* f16 [D, T] */
  const float* mlp_out_b;
  const float* pe_flipped;          /* fp32 [buffer_size][buffer_size][T]: table n-1 row f = pe[n-1-f] (f < n) */
  const tocvp_tuning* tuning;       /* NULL = defaults */
} tocvp_pred_weights;

size_t tocvp_sizeof_pred_weights(void);
size_t tocvp_sizeof_pred_layer(void);
size_t tocvp_predictor_workspace_bytes(const tocvp_pred_weights* w, int B, int L, int num_context, int num_preds);
/* PredictorWrapper.forward (src/models/Predictors/predictor_wrapper.py:50-87), teacher_force = False:
 * slot_history fp32 (sequence b at slot_history + b*hist_seq_stride, first num_context frames are read),
 * text fp32 [B, L, T] (the text-encoder output), pred_slots fp32 [B, num_preds, S, D].  256B-aligned workspace. */
int tocvp_predictor_rollout(const tocvp_pred_weights* w, const float* slot_history, size_t hist_seq_stride,
                            const float* text, int B, int L, int num_context, int num_preds, float* pred_slots,
                            void* workspace, size_t ws_bytes, void* stream);
/* BaseTextOCVP.forward (src/models/Predictors/text_cond_OCVP.py:79-105): slots [B,n,S,D] -> out [B,S,D].
 * Workspace: tocvp_predictor_workspace_bytes(w, B, L, n, 1). */
int tocvp_predictor_forward(const tocvp_pred_weights* w, const float* slots, const float* text, int B, int n, int L,
                            float* out, void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * SAVi.encode (src/models/SAVi.py:226-238; SimpleConvEncoder src/models/EncodersDecoders/encoders.py:136-159;
 * SoftPositionEmbed src/models/Blocks/model_blocks.py:215-226; encoder_mlp SAVi.py:115-120).
 * ------------------------------------------------------------------------------------------ */
typedef struct tocvp_enc_weights {
  const float* w_conv1;      /* fp32 [25*3][32]: row (ky*5+kx)*3+ci, col co  (encoder.encoder.0.block.0.weight permuted) */
  const float* b_conv1;
  const void* w_conv[3];     /* f16 [25,32,32] tap-major (encoder.encoder.{1,2,3}) */
  const float* b_conv[3];
  const float* posemb;       /* fp32 [H*W, 32] = Conv1x1(build_grid) + bias, precomputed (batch independent) */
  const float *ln_g, *ln_b;  /* encoder_mlp.0 */
  const void* w_mlp1;        /* f16 [F, 32]   encoder_mlp.1 */
  const float* b_mlp1;
  const void* w_mlp2;        /* f16 [F, F]    encoder_mlp.3 */
  const float* b_mlp2;
  int H, W, in_channels, hidden, feat_dim;
  const void* w_conv1_tc;    /* f16 [25,32,32] tap-major, input channels zero-padded 3 -> 32: conv 1 on the tensor cores */
  const void* w_conv1_vp;    /* f16 [3,32,32]: conv 1 with the 5 x-taps x 3 channels of two filter rows folded into K = 32
                                (k = half*16 + kx*3 + c), three vertical taps two rows apart; null = use w_conv1_tc */
  const tocvp_tuning* tuning; /* NULL = defaults */
  const void* w_conv_xp[3];  /* f16 [30,64,32] = [(ky, u)][parity*32 + co][ci]: the 32 -> 32 layers for the pixel-pair kernel
                                (one MMA row = pixels 2j, 2j+1; u = 0..5 horizontal input shifts; block (ky, u) holds
                                w[co][ci][ky][u - parity] or zeros).  NULL = the 25-tap kernel on w_conv.  ABI 3. */
} tocvp_enc_weights;

size_t tocvp_sizeof_enc_weights(void);
size_t tocvp_savi_encode_workspace_bytes(const tocvp_enc_weights* w, int n_img);
/* frames fp32: image i = 3 planes of H x W at frames + i*img_stride (floats), so x[:, t] of a [B,T,3,H,W] video is
 * addressed in place; feats [n_img, H*W, F] written as f16 and/or fp32 (either may be NULL, not both). */
int tocvp_savi_encode(const tocvp_enc_weights* w, const float* frames, size_t img_stride, int n_img, void* feats_f16,
                      float* feats_f32, void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * SAVi.decode + broadcast + ConvDecoder + compositing (src/models/SAVi.py:241-275,
 * src/models/EncodersDecoders/decoders.py:96-119).
 * ------------------------------------------------------------------------------------------ */
typedef struct tocvp_dec_weights {
  const void* w1_taps;       /* f16 [25*64, D]: row tap*64+co, col ci  (decoder.decoder.0 weight, per-tap matrices) */
  const float* p1;           /* fp32 [H*W, 64] = conv1(posemb map) + b1, precomputed (batch independent)            */
  const void* w_conv[3];     /* f16 [25,64,64] tap-major (decoder.decoder.{1,2,3}) */
  const float* b_conv[3];
  const void* w_out;         /* f16 [9][16][64]: (ky*3+kx, co zero-padded 4 -> 16, ci)  decoder.decoder.4 */
  const float* b_out;        /* [4] */
  int H, W, slot_dim, num_slots, hidden;
  const void* w_out_taps;    /* f16 [48][64]: row (ky*3+kx)*4 + co (rows 36..47 zero), col ci: the 9 taps in the GEMM N dim */
  const tocvp_tuning* tuning; /* NULL = defaults */
} tocvp_dec_weights;

size_t tocvp_sizeof_dec_weights(void);
size_t tocvp_savi_decode_workspace_bytes(const tocvp_dec_weights* w, int n_frames);
/* slots fp32 [n_frames, S, D] -> recons_imgs fp32 [n_frames,3,H,W]; optional recons [n_frames,S,3,H,W] and
 * masks [n_frames,S,1,H,W] (NULL to skip: the evaluator only consumes recons_imgs). */
#define TOCVP_DECODE_CHUNK_FRAMES 256 /* frames decoded per internal pass (bounds the activation workspace) */
/* conv_events (optional, may be NULL): cudaEvent_t handles; pair (2i, 2i+1) is recorded around the i-th conv5x5
 * launch (3 per chunk of TOCVP_DECODE_CHUNK_FRAMES frames) so a caller can time the dominant kernel inside a step. */
int tocvp_savi_decode(const tocvp_dec_weights* w, const float* slots, int n_frames, float* recons_imgs, float* recons,
                      float* masks, void* workspace, size_t ws_bytes, void* stream, void* const* conv_events,
                      int n_conv_events);

/* ------------------------------------------------------------------------------------------
 * ExtendedDINOSAUR.linear_feat_proj (src/models/ExtendedDINOSAUR.py:97-102, applied at :186):
 * LayerNorm(F) -> Linear(F,Hd) -> ReLU -> Linear(Hd,D) over rows of ViT patch features.
 * ------------------------------------------------------------------------------------------ */
typedef struct tocvp_proj_weights {
  const float *ln_g, *ln_b; /* linear_feat_proj.0 [F]                */
  const void* w1;           /* f16 [Hd, F]   linear_feat_proj.1      */
  const float* b1;
  const void* w2;           /* f16 [D, Hd]   linear_feat_proj.3      */
  const float* b2;
  int feat_dim, hidden_dim, slot_dim;
  float ln_eps;             /* 1e-5 (nn.LayerNorm default)           */
  const tocvp_tuning* tuning; /* NULL = defaults */
} tocvp_proj_weights;

size_t tocvp_sizeof_proj_weights(void);
size_t tocvp_dino_project_workspace_bytes(const tocvp_proj_weights* w, int rows);
/* feats fp32 [rows, F] -> projected features [rows, D] as f16 and/or fp32 (either may be NULL, not both). */
int tocvp_dino_project(const tocvp_proj_weights* w, const float* feats, int rows, void* out_f16, float* out_f32,
                       void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * MLPPatchDecoder.forward (src/models/EncodersDecoders/decoders.py:232-282; MLP built at :309-322,
 * CNN at :325-365; BasePatchDecoder.broadcast_slots / add_positional_encoding :152-199).
 * Weight packing (done once per load by the host module):
 *   mlp_w[i]  f16 [out_i, in_i] = torch Linear weight; the last layer (out = feat_dim + 1, alpha last) is zero-padded
 *             to a multiple of 8 rows, its bias likewise.
 *   cnn_w[i]  ConvBlock i with eval-mode BatchNorm folded in (w' = w * gamma/sqrt(var+eps), b' = (b-mean)*scale+beta):
 *             cnn_up[i] == 0: f16 [cout, 9*cin], K index = (ky*3+kx)*cin + ci;
 *             cnn_up[i] == 1 (an Upsample(2) precedes the conv: it is folded into 4 phase convolutions on the
 *             low-resolution input): f16 [4*cout, 4*cin], row = (py*2+px)*cout + co, K index = (dy*2+dx)*cin + ci,
 *             weight = sum of the 3x3 taps that land on low-res offset (dy+py-1, dx+px-1); bias repeated per phase.
 *   out_w     final conv3x3 -> 3 channels, padded to 4: out_up == 1: f16 [64 rows allocated, 16 used; 9*cin], row =
 *             phase*4 + c, K index = (oy*3+ox)*cin + ci over the low-res 3x3 neighbourhood (zeros where a phase does not
 *             see a tap); out_up == 0: rows = c (8 used).  out_b fp32 [16] / [8].
 * ------------------------------------------------------------------------------------------ */
#define TOCVP_PATCH_MAX_MLP 6
#define TOCVP_PATCH_MAX_CNN 6
typedef struct tocvp_patch_weights {
  const float* pos_embed;   /* fp32 [N, D]  decoder.pos_embed                    */
  const float *ln_g, *ln_b; /* decoder.mlp.0 (initial_layer_norm)                */
  const void* mlp_w[TOCVP_PATCH_MAX_MLP];
  const float* mlp_b[TOCVP_PATCH_MAX_MLP];
  const void* cnn_w[TOCVP_PATCH_MAX_CNN];
  const float* cnn_b[TOCVP_PATCH_MAX_CNN];
  const void* out_w;
  const float* out_b;
  int mlp_out[TOCVP_PATCH_MAX_MLP];
  int cnn_cin[TOCVP_PATCH_MAX_CNN], cnn_cout[TOCVP_PATCH_MAX_CNN], cnn_up[TOCVP_PATCH_MAX_CNN];
  int n_mlp, n_cnn, out_cin, out_up, reconstruct_images;
  int num_slots, slot_dim, num_patches, grid, feat_dim, img_size;
  float ln_eps;             /* 1e-5 */
  const tocvp_tuning* tuning; /* NULL = defaults */
} tocvp_patch_weights;

size_t tocvp_sizeof_patch_weights(void);
size_t tocvp_patch_decode_workspace_bytes(const tocvp_patch_weights* w, int n_frames);
/* slots fp32 [n_frames, S, D] -> recons_imgs fp32 [n_frames,3,I,I], recons_feats fp32 [n_frames,N,F],
 * masks fp32 [n_frames,S,1,g,g]; any output may be NULL (recons_imgs == NULL skips the CNN). */
int tocvp_patch_decode(const tocvp_patch_weights* w, const float* slots, int n_frames, float* recons_imgs,
                       float* recons_feats, float* masks, void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Evaluator metrics on the device (src/05_evaluate_predictor.py:96-103 -> src/lib/metrics.py:181-270, piqa 1.2.2):
 * optional clamp of both images to [0,1], then per image MSE, PSNR = 10 log10(1/(mse + 1e-8)) and SSIM (11-tap Gaussian
 * window, sigma 1.5, k1 = 0.01, k2 = 0.03, value range 1, valid convolution, mean over channels and positions).
 * pred fp32 [n_img, C, H, W]; the target of image i is read IN PLACE from a video tensor:
 *   target + (i / frames_per_seq) * target_seq_stride + (target_frame0 + i % frames_per_seq) * C*H*W   (floats),
 * i.e. videos[:, num_context : num_context + num_preds] needs no copy.  Outputs fp32 [n_img], any may be NULL.
 * n_img, C <= 65535.
 * ------------------------------------------------------------------------------------------ */
int tocvp_frame_metrics(const float* pred, const float* target, size_t target_seq_stride, int frames_per_seq,
                        int target_frame0, int n_img, int C, int H, int W, int clamp, float* mse, float* psnr, float* ssim,
                        void* stream);

/* ------------------------------------------------------------------------------------------
 * TransformerTextEncoder.forward (src/models/EncodersDecoders/text_encoders.py:84-124), the text encoder of
 * TextOCVP_CustomTF (src/models/Predictors/predictor_wrapper.py:106-110): embeddings -> LayerNorm(1e-8) -> padding mask ->
 * num_layers x nn.TransformerEncoderLayer (post-norm, exact GELU, key-padding mask j >= length) -> LayerNorm -> Linear.
 * All weights fp32; "_t" = transposed to [in][out].  One CTA per caption, fp32 end to end.
 * ------------------------------------------------------------------------------------------ */
#define TOCVP_TEXT_MAX_LAYERS 4
typedef struct tocvp_text_layer {
  const float *in_w_t, *in_b;   /* self_attn.in_proj_{weight,bias}: [D][3D], [3D] (q | k | v)  */
  const float *out_w_t, *out_b; /* self_attn.out_proj                                          */
  const float *ln1_g, *ln1_b;   /* norm1                                                       */
  const float *ff1_w_t, *ff1_b; /* linear1 [D][F]                                              */
  const float *ff2_w_t, *ff2_b; /* linear2 [F][D]                                              */
  const float *ln2_g, *ln2_b;   /* norm2                                                       */
} tocvp_text_layer;
typedef struct tocvp_text_weights {
  const float *tok_emb, *pos_emb; /* [vocab][D], [context_length][D]     */
  const float *ln0_g, *ln0_b;     /* layer_norm (eps 1e-8)               */
  tocvp_text_layer layers[TOCVP_TEXT_MAX_LAYERS];
  const float *lnf_g, *lnf_b;     /* text_out_projection.0               */
  const float *proj_w_t, *proj_b; /* text_out_projection.1: [D][out_dim] */
  int num_layers, input_dim, ffn_dim, num_heads, output_dim, vocab_size, context_length;
} tocvp_text_weights;

size_t tocvp_sizeof_text_weights(void);
/* tokens int64 [B, L], lengths int64 [B] (device) -> text embeddings fp32 [B, L, output_dim].  L <= 64. */
int tocvp_text_encode(const tocvp_text_weights* w, const long long* tokens, const long long* lengths, int B, int L,
                      float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Sibling predictors behind the same PredictorWrapper loop: VanillaTransformerPredictor (src/models/Predictors/OCVP.py:24-141)
 * OCVPSeq (OCVP.py:145-319) and OCVPPar (OCVP.py:324-548: object- and time-attention in PARALLEL on the same normed input,
 * expressed as two blocks: [attention over the frame's slots, no feed-forward] + [attention along the slot's history on the
 * SAME LayerNorm-1 output, then the layer's feed-forward]).  `blocks` are pre-norm nn.TransformerEncoderLayer parameter sets (same fields as
 * tocvp_text_layer); block_group[i] selects the keys a query attends to: 0 = all tokens of the window (Vanilla),
 * 1 = tokens of the same frame (OCVPSeqLayer.object_encoder_block), 2 = the same slot over time (time_encoder_block).
 * ------------------------------------------------------------------------------------------ */
#define TOCVP_OCVP_MAX_BLOCKS 8
typedef struct tocvp_ocvp_weights {
  const float *mlp_in_w_t, *mlp_in_b;   /* [slot_dim][token_dim]                                  */
  const float *mlp_out_w_t, *mlp_out_b; /* [token_dim][slot_dim]                                  */
  const float* pe;                      /* sinusoidal table [max_len][token_dim] (model_blocks.py:261-266) */
  tocvp_text_layer blocks[TOCVP_OCVP_MAX_BLOCKS];
  int block_group[TOCVP_OCVP_MAX_BLOCKS];
  int block_flags[TOCVP_OCVP_MAX_BLOCKS]; /* bit 0: no feed-forward half (first branch of an OCVP-Par layer); bit 1: reuse the
                                             previous block's LayerNorm-1 output (second branch of an OCVP-Par layer) */
  int num_blocks, num_slots, slot_dim, token_dim, ffn_dim, num_heads, max_len, residual;
} tocvp_ocvp_weights;

size_t tocvp_sizeof_ocvp_weights(void);
/* One prediction step (VanillaTransformerPredictor.forward / OCVPSeq.forward): sequence b's window [n, S, slot_dim] at
 * slots + b*seq_stride (floats) -> out [B, S, slot_dim].  n * S <= 80 tokens, token_dim <= 128. */
int tocvp_ocvp_forward(const tocvp_ocvp_weights* w, const float* slots, size_t seq_stride, int B, int n, float* out,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * Frozen ViT front-end of ExtendedDINOSAUR: ViTEncoder.forward (src/models/EncodersDecoders/timm_encoders.py:58-69) around a
 * timm VisionTransformer (vit_base_patch14_dinov2, timm_encoders.py:232-254: patch 14, 768-d, 12 pre-norm blocks with
 * LayerScale, 12 heads, GELU MLP): normalise -> patch embedding -> cat(cls) + pos_embed -> blocks -> drop the class token.
 * timm is a third-party dependency that is not vendored: the block structure restates its published definition (PARITY
 * UNPINNED against timm itself; tests compare with a plain-torch restatement).  f16 matrices are torch Linear weights
 * [out, in]; w_patch is patch_embed.proj.weight [E, 3, p, p] flattened to [E, 3*p*p] and zero-padded to k_pad columns.
 * ------------------------------------------------------------------------------------------ */
typedef struct tocvp_vit_block {
  const float *ln1_g, *ln1_b;  /* norm1 */
  const void* w_qkv;           /* f16 [3E, E]  attn.qkv */
  const float* b_qkv;
  const void* w_proj;          /* f16 [E, E]   attn.proj */
  const float* b_proj;
  const float* ls1;            /* [E] ls1.gamma */
  const float *ln2_g, *ln2_b;  /* norm2 */
  const void* w_fc1;           /* f16 [Hm, E]  mlp.fc1 */
  const float* b_fc1;
  const void* w_fc2;           /* f16 [E, Hm]  mlp.fc2 */
  const float* b_fc2;
  const float* ls2;            /* [E] ls2.gamma */
} tocvp_vit_block;

typedef struct tocvp_vit_weights {
  const tocvp_vit_block* blocks; /* HOST array of num_blocks entries */
  int num_blocks, embed_dim, num_heads, mlp_dim, patch, img_h, img_w, grid_h, grid_w, k_pad;
  const void* w_patch;           /* f16 [E, k_pad] */
  const float* b_patch;          /* [E] */
  const float* cls_pos0;         /* [E] = cls_token + pos_embed[0] */
  const float* pos;              /* [N, E] = pos_embed[1:]  (N = grid_h * grid_w) */
  float mean[3], inv_std[3];     /* input normalisation (timm_encoders.py:82-96) */
  float ln_eps;                  /* 1e-6 */
  const tocvp_tuning* tuning;
} tocvp_vit_weights;

size_t tocvp_sizeof_vit_weights(void);
size_t tocvp_sizeof_vit_block(void);
size_t tocvp_vit_workspace_bytes(const tocvp_vit_weights* w, int n_img);
/* images fp32: image i = 3 planes of img_h x img_w at images + i*img_stride (floats) -> patch features fp32 [n_img, N, E]. */
int tocvp_vit_forward(const tocvp_vit_weights* w, const float* images, size_t img_stride, int n_img, float* feats,
                      void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stand-alone forwards of the reference's sub-modules.  The fused stage entry points above never call these; they exist
 * so that calls a reference user can make on a sub-module (`savi.transition_module(slots)`, `savi.decoder(x)`,
 * `savi.encoder(x)`, `predictor.pe(x)`, `savi.encoder_pos_embedding(x)`) run on this library too.
 * ------------------------------------------------------------------------------------------ */
/* TransformerBlock.forward, post-norm flavour (src/models/Blocks/attention.py:387-395; SAVi transition, SAVi.py:193):
 * slots fp32 [B, S, 128] -> out fp32 [B, S, 128].  Only the t_* fields, num_slots, t_heads, t_hidden, ln_eps_tf of w are read. */
int tocvp_transition(const tocvp_sa_weights* w, const float* slots, int B, float* out, void* stream);
/* SimpleConvEncoder.forward (src/models/EncodersDecoders/encoders.py:156-159): frames as in tocvp_savi_encode ->
 * NHWC f16 [n_img, H, W, 32] after the fourth conv + ReLU.  Workspace: tocvp_savi_encode_workspace_bytes(w, n_img). */
int tocvp_savi_conv_stack(const tocvp_enc_weights* w, const float* frames, size_t img_stride, int n_img,
                          void* out_nhwc_f16, void* workspace, size_t ws_bytes, void* stream);
/* conv5x5 (stride 1, zero padding 2) + bias (+ReLU) on a materialised NCHW fp32 tensor with ANY channel counts, fp32 SIMT:
 * x [n_img, cin, H, W], weight [cout, cin, 5, 5] (torch layout), out as NHWC f16 and/or NCHW fp32.  First layer of
 * ConvDecoder.forward (decoders.py:111-125) when it is called on an arbitrary tensor rather than on broadcast slots. */
int tocvp_conv5x5_generic(const float* x, const float* weight, const float* bias, int relu, void* out_nhwc_f16,
                          float* out_nchw_f32, int n_img, int H, int W, int cin, int cout, void* stream);
/* Last layer of ConvDecoder.forward (decoders.py:104-108): conv3x3 64 -> 4, no activation; x NHWC f16 [n_img,H,W,64] ->
 * NHWC fp32 [n_img, H, W, 4]. */
int tocvp_conv3x3_head(const tocvp_dec_weights* w, const void* x_nhwc_f16, int n_img, float* out_nhwc4, void* stream);
/* out[row] = x[row] + table[(row / div) % mod], fp32 rows of D floats: SoftPositionEmbed.forward (model_blocks.py:215-226:
 * div = 1, mod = H*W, table = projection(grid)) and TemporalPositionalEncoding.forward (model_blocks.py:358-379: div = S,
 * mod = n, table = flip(pe[:n])). */
int tocvp_add_table(const float* x, const float* table, int div, int mod, int D, size_t rows, float* out, void* stream);
/* fp32 -> f16, saturating at +-65504 (the entry format of the tensor-core kernels). */
int tocvp_cast_f16(const float* in, void* out, size_t n, void* stream);
/* in-place clamp to [0, 1] (src/05_evaluate_predictor.py:96). */
int tocvp_clamp01(float* x, size_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TOCVP_H_ */
