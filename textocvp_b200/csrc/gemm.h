// Internal interface of the tcgen05 GEMM (gemm_tc.cu): the plain dense projection and its implicit-convolution mode.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace tocvp {

int gemm_f16(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K, const float* bias, int relu,
             const float* residual, int ldr, int res_div, int res_mod, float* out32, int ld32, __half* out16,
             int ld16, cudaStream_t stream);

// LayerNorm folded into a pair of GEMMs (CTA-pair kernel only, M >= 1024):
//   producer  (stats_out != null): C = A.W^T (+bias) + residual in fp32, PLUS an f16 copy of C (out16) and per-row
//             partial [sum, sum of squares] pairs of C: stats_out[M][N/64][2], one pair per 64 output columns (each
//             written by exactly one warp: no atomics, nothing to zero);
//   consumer  (stats != null): A is that f16 copy, W holds W*gamma, and the epilogue applies
//             y = rstd*(acc - mu*c_n) + d_n with mu, rstd from the `slots` partial pairs, c_n = sum_k W'[n,k], d_n passed
//             as `bias`.
struct GemmLn {
  const float* stats;
  int slots;
  const float* c;
  float inv_k, eps;
  float* stats_out;
};
bool gemm_ln_supported(int M, int N);   // true when gemm_f16_ln will take the CTA-pair kernel for this shape
int gemm_f16_ln(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K, const float* bias, int relu,
                const float* residual, int ldr, float* out32, int ld32, __half* out16, int ld16, const GemmLn& ln,
                cudaStream_t stream);

// 3x3 convolution over a zero-bordered NHWC f16 activation [n_img, Hp, Wp, Cin] (Hp = H+2, Wp = W+2) run as a GEMM over
// the FLATTENED padded pixel index p: a filter tap is a constant row offset dy*Wp + dx of the A operand, so the A tile
// of tap t is one 2-D TMA box at row m0 + off[t] (rows outside the tensor are zero-filled by the TMA unit) and no
// im2col buffer exists.  Rows that are border pixels compute garbage and are not stored.
//
// With `up` = 1 the convolution consumes the nearest-neighbour x2 upsampling of the input (Upsample + Conv2d 3x3,
// reference src/models/EncodersDecoders/decoders.py:336-352) WITHOUT materialising it: output pixel (2y+py, 2x+px) only
// sees the 2x2 low-resolution neighbourhood {y+py-1, y+py} x {x+px-1, x+px}, so each of the 4 output phases is a 2x2
// convolution whose weights are sums of the original 3x3 taps (4/9 of the FLOPs).  N is then phase-major:
// N = 4 * cpp, W rows [phase][cout], K = taps * Cin with per-phase tap offsets.
struct ConvMap {
  int taps;             // K-blocks are grouped by tap: K = taps * cin
  int cin;              // multiple of 64
  int tiles_per_phase;  // != 0: each phase (cpp columns) has its own tap set (the GEMM fills in the tile count); 0 = one tap set
  int off[4][9];        // A row offset of (phase, tap)
  int Hp, Wp;           // padded input geometry; M = n_img * Hp * Wp
  int up;               // output is the 2x-upsampled grid, column n -> phase n / cpp
  int cpp;              // output channels per phase (multiple of 4)
  int Hop, Wop, pad;    // output geometry: [n_img, Hop, Wop, ldo] with `pad` border pixels (1: next layer's input, 0: final)
};

// out16 / out32: NHWC with `ldo` channels per pixel (f16: next layer's zero-bordered input; fp32: final image planes).
int gemm_conv_f16(const __half* X, const __half* W, int n_img, int N, const ConvMap& cm, const float* bias, int relu,
                  float* out32, __half* out16, int ldo, cudaStream_t stream);

}  // namespace tocvp
