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
 *   - no global mutable state besides caches of immutable device attributes;
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

#define TOCVP_ABI_VERSION 1

#define TOCVP_OK 0
#define TOCVP_ERR_BAD_ARG (-1)   /* bad shape / null or misaligned pointer / unsupported size */
#define TOCVP_ERR_CUDA (-2)      /* a CUDA runtime call or kernel launch failed                */
#define TOCVP_ERR_ARCH (-3)      /* device is not sm_100                                       */
#define TOCVP_ERR_WORKSPACE (-4) /* workspace too small                                         */

int tocvp_abi_version(void);
/* Checks that `device` is an sm_100 part and makes it current. */
int tocvp_init(int device);
const char* tocvp_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Dense projection: C[M,N] = A[M,K] . W[N,K]^T (+bias) (ReLU) (+residual), tcgen05 / TMEM / TMA.
 * A, W: f16, K innermost (W is torch.nn.Linear.weight as stored).  bias fp32[N] or NULL.
 * residual fp32 [M, ldr] or NULL (added after ReLU).  Writes out_f32 and/or out_f16 (either may
 * be NULL, not both).  N, K, lda, ldw multiples of 8.
 * Replaces nn.Linear -> cuBLAS at src/models/Blocks/attention.py:167-175, 296-300, 352-356,
 * src/models/Predictors/text_cond_OCVP.py:47-48, src/models/SAVi.py:117-119.
 * ------------------------------------------------------------------------------------------ */
int tocvp_gemm_f16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                   int relu, const float* residual, int ldr, float* out_f32, int ld32, void* out_f16, int ld16,
                   void* stream);

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
 * H % 16 == 0, W % 32 == 0; (cin,cout) in {(64,64),(32,32)}.
 * Replaces nn.Conv2d -> cuDNN at src/models/Blocks/model_blocks.py:83-91 as used by
 * src/models/EncodersDecoders/decoders.py:96-108 and encoders.py:141-153.
 * ------------------------------------------------------------------------------------------ */
int tocvp_conv5x5_f16(const void* x, const void* w_packed, const float* bias, void* out, int n_img, int H, int W,
                      int cin, int cout, int relu, void* stream);

/* Test-only hardware probe (not on the product path): D[128,64] = X[shift:shift+128, :64] . W^T with the
 * A operand descriptor started `shift` 128-byte rows into a swizzled TMA tile.  See csrc/probe.cu. */
int tocvp_probe_shifted_operand(const void* X, const void* W, float* out, int shift, int base_offset_mode,
                                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TOCVP_H_ */
