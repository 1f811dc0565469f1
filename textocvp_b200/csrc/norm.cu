// LayerNorm over the innermost dimension (fp32 statistics), one warp per row, vectorised 16-byte loads.
//   y = (x (+ add[row % add_rows]) - mean) * rsqrt(var + eps) * gamma + beta      -> f16 and/or fp32
// Bandwidth-bound; feeds the f16 A-operand of the tcgen05 GEMMs.  Replaces nn.LayerNorm at reference
// src/models/Blocks/attention.py:49-51,361-362,427,435-436 and src/models/SAVi.py:116 (with the
// SoftPositionEmbed add of src/models/Blocks/model_blocks.py:215-226 fused in front of it).
#include "host_util.h"
#include "ptx.cuh"

namespace tocvp {

constexpr int LN_MAX_CHUNKS = 8;  // D <= 8 * 128 = 1024

template <typename TIN>
__device__ __forceinline__ float4 load4(const TIN* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 load4<__half>(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

template <typename TIN>
__global__ void __launch_bounds__(256)
layernorm_kernel(const TIN* __restrict__ x, int ldx, int x_div, const float* __restrict__ add, int add_rows,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int rows, int D,
                 __half* __restrict__ out16, int ld16, float* __restrict__ out32, int ld32) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  pdl_wait();
  pdl_trigger();
  if (warp >= rows) return;
  const TIN* xr = x + size_t(warp / x_div) * ldx;   // x_div > 1: each input row is broadcast to x_div output rows
  const float* ar = add ? add + size_t(warp % add_rows) * D : nullptr;
  float4 v[LN_MAX_CHUNKS];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < LN_MAX_CHUNKS; ++c) {
    const int i = c * 128 + lane * 4;
    if (i < D) {
      v[c] = load4<TIN>(xr + i);
      if (ar) {
        const float4 a = *reinterpret_cast<const float4*>(ar + i);
        v[c].x += a.x; v[c].y += a.y; v[c].z += a.z; v[c].w += a.w;
      }
      s += (v[c].x + v[c].y) + (v[c].z + v[c].w);
    }
  }
  const float mean = warp_sum(s) / float(D);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < LN_MAX_CHUNKS; ++c) {
    const int i = c * 128 + lane * 4;
    if (i < D) {
      const float a = v[c].x - mean, b = v[c].y - mean, cc = v[c].z - mean, d = v[c].w - mean;
      q += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / float(D) + eps);
#pragma unroll
  for (int c = 0; c < LN_MAX_CHUNKS; ++c) {
    const int i = c * 128 + lane * 4;
    if (i < D) {
      const float4 g = *reinterpret_cast<const float4*>(gamma + i);
      const float4 b = *reinterpret_cast<const float4*>(beta + i);
      const float y0 = (v[c].x - mean) * rstd * g.x + b.x;
      const float y1 = (v[c].y - mean) * rstd * g.y + b.y;
      const float y2 = (v[c].z - mean) * rstd * g.z + b.z;
      const float y3 = (v[c].w - mean) * rstd * g.w + b.w;
      if (out16) {
        uint2 p;
        p.x = pack_half2(y0, y1);
        p.y = pack_half2(y2, y3);
        *reinterpret_cast<uint2*>(out16 + size_t(warp) * ld16 + i) = p;
      }
      if (out32) *reinterpret_cast<float4*>(out32 + size_t(warp) * ld32 + i) = make_float4(y0, y1, y2, y3);
    }
  }
}


// Narrow rows (D = 4*LPR <= 64): LPR lanes per row, 32/LPR rows per warp, one float4 per lane -- keeps all 32 lanes busy
// for the encoder's LayerNorm(32) over B*T*4096 pixels (SAVi.py:116).
template <typename TIN, int LPR>
__global__ void __launch_bounds__(256)
layernorm_narrow_kernel(const TIN* __restrict__ x, int ldx, const float* __restrict__ add, int add_rows,
                        const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int rows,
                        __half* __restrict__ out16, int ld16, float* __restrict__ out32, int ld32) {
  constexpr int D = 4 * LPR;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const float4 g = *reinterpret_cast<const float4*>(gamma + sub * 4);
  const float4 bt = *reinterpret_cast<const float4*>(beta + sub * 4);
  const long long warp0 = (long long)(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)(gridDim.x) * (blockDim.x >> 5);
  for (long long base = warp0 * RPW; base < rows; base += nwarps * RPW) {   // warp-uniform trip count (shuffles inside)
    const long long row = base + lane / LPR;
    const bool valid = row < rows;
    float4 v = valid ? load4<TIN>(x + size_t(row) * ldx + sub * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (add && valid) {
      const float4 a = *reinterpret_cast<const float4*>(add + size_t(row % add_rows) * D + sub * 4);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / D);
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    float q = (a * a + b * b) + (c * c + d * d);
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / D) + eps);
    const float y0 = a * rstd * g.x + bt.x, y1 = b * rstd * g.y + bt.y, y2 = c * rstd * g.z + bt.z,
                y3 = d * rstd * g.w + bt.w;
    if (out16 && valid) {
      uint2 p;
      p.x = pack_half2(y0, y1);
      p.y = pack_half2(y2, y3);
      *reinterpret_cast<uint2*>(out16 + size_t(row) * ld16 + sub * 4) = p;
    }
    if (out32 && valid) *reinterpret_cast<float4*>(out32 + size_t(row) * ld32 + sub * 4) = make_float4(y0, y1, y2, y3);
  }
}

int layernorm_bcast(const void* x, int x_is_f16, int ldx, int x_div, const float* add, int add_rows, const float* gamma,
                    const float* beta, float eps, int rows, int D, __half* out16, int ld16, float* out32, int ld32,
                    cudaStream_t stream);

int layernorm(const void* x, int x_is_f16, int ldx, const float* add, int add_rows, const float* gamma,
              const float* beta, float eps, int rows, int D, __half* out16, int ld16, float* out32, int ld32,
              cudaStream_t stream) {
  return layernorm_bcast(x, x_is_f16, ldx, 1, add, add_rows, gamma, beta, eps, rows, D, out16, ld16, out32, ld32, stream);
}

// y[row] = LN(x[row / x_div] + add[row % add_rows]): with x_div = add_rows = N this is the broadcast-slots + positional
// embedding + LayerNorm front of MLPPatchDecoder (reference src/models/EncodersDecoders/decoders.py:152-199, 314).
int layernorm_bcast(const void* x, int x_is_f16, int ldx, int x_div, const float* add, int add_rows, const float* gamma,
                    const float* beta, float eps, int rows, int D, __half* out16, int ld16, float* out32, int ld32,
                    cudaStream_t stream) {
  TOCVP_CHECK_ARG(x && gamma && beta && rows > 0 && D > 0 && D % 4 == 0 && D <= LN_MAX_CHUNKS * 128 && x_div >= 1);
  TOCVP_CHECK_ARG(ldx % 4 == 0 && (out16 || out32));
  TOCVP_CHECK_ARG(!add || add_rows > 0);
  if ((D == 32 || D == 64) && x_div == 1) {
    const int rpw = 128 / D;
    long long warps = (rows + rpw - 1) / rpw;
    long long blocks = (warps + 7) / 8;
    const int grid = int(blocks > 148 * 32 ? 148 * 32 : blocks);
#define LN_NARROW(TIN, LPR)                                                                                         \
  layernorm_narrow_kernel<TIN, LPR><<<grid, 256, 0, stream>>>(static_cast<const TIN*>(x), ldx, add, add_rows, gamma, \
                                                              beta, eps, rows, out16, ld16, out32, ld32)
    if (x_is_f16) { if (D == 32) LN_NARROW(__half, 8); else LN_NARROW(__half, 16); }
    else          { if (D == 32) LN_NARROW(float, 8);  else LN_NARROW(float, 16); }
#undef LN_NARROW
    TOCVP_LAUNCHED();
    return TOCVP_OK;
  }
  const int wpb = 8;
  const int grid = (rows + wpb - 1) / wpb;
  if (x_is_f16)
    TOCVP_CUDA(launch_pdl(layernorm_kernel<__half>, dim3(grid), dim3(wpb * 32), 0, stream, static_cast<const __half*>(x),
                          ldx, x_div, add, add_rows, gamma, beta, eps, rows, D, out16, ld16, out32, ld32));
  else
    TOCVP_CUDA(launch_pdl(layernorm_kernel<float>, dim3(grid), dim3(wpb * 32), 0, stream, static_cast<const float*>(x),
                          ldx, x_div, add, add_rows, gamma, beta, eps, rows, D, out16, ld16, out32, ld32));
  count_launch();
  return TOCVP_OK;
}

}  // namespace tocvp

extern "C" int tocvp_layernorm(const void* x, int x_is_f16, int ldx, const float* add, int add_rows,
                               const float* gamma, const float* beta, float eps, int rows, int D, void* out_f16,
                               int ld16, float* out_f32, int ld32, void* stream) {
  return tocvp::layernorm(x, x_is_f16, ldx, add, add_rows, gamma, beta, eps, rows, D,
                          static_cast<__half*>(out_f16), ld16, out_f32, ld32, static_cast<cudaStream_t>(stream));
}
