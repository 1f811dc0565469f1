// Spatial-broadcast decoder + alpha compositing: SAVi.decode / SAVi.broadcast (reference src/models/SAVi.py:241-275)
// over ConvDecoder (src/models/EncodersDecoders/decoders.py:96-119).
//
//  layer 1 (conv5x5 128->64 on the broadcast slot + positional embedding) is never run as a convolution: its input is
//      slot (spatially constant) + posemb (batch independent), so
//          conv1(x)[y,x,:] = P[y,x,:] + sum_{taps valid at (y,x)} W_tap . slot,      P = conv1(posemb) + b1 (precomputed)
//      and the set of valid taps takes only 5 x 5 border patterns.  One tcgen05 GEMM [n_slots,128] x [128, 25*64] gives the
//      per-tap vectors; `dec_l1_kernel` forms the 25 pattern sums in smem and streams out relu(P + S[pattern]) as the f16
//      NHWC activation -- a pure bandwidth kernel replacing 40% of the decoder FLOPs (the broadcast tensor of
//      SAVi.py:264-275 is never materialised).
//  layers 2-4: tcgen05 implicit-GEMM conv5x5 64->64 (conv5x5_tc.cu).
//  final conv3x3 64->4 + softmax over slots + weighted sum (SAVi.py:251-255): `conv3x3_composite_kernel`, all 8 slots
//      of a pixel tile handled by one CTA with an online softmax, so the [B',8,4,H,W] maps never hit HBM unless the
//      caller asks for `recons` / `masks`.
#include "host_util.h"
#include "ptx.cuh"

namespace tocvp {

int gemm_f16(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K, const float* bias, int relu,
             const float* residual, int ldr, int res_div, int res_mod, float* out32, int ld32, __half* out16,
             int ld16, cudaStream_t stream);
int conv5x5_f16(const __half* x, const __half* wpacked, const float* bias, __half* out, int n_img, int H, int W, int cin,
                int cout, int relu, cudaStream_t stream);

__global__ void f32_to_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t n4) {
  for (size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < n4; e += size_t(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(in)[e];
    uint2 p;
    p.x = pack_half2(v.x, v.y);
    p.y = pack_half2(v.z, v.w);
    reinterpret_cast<uint2*>(out)[e] = p;
  }
}

// ------------------------------------------------------------------------------------------------ layer 1
// taps: fp32 [n, 25*C] (tap-major, tap = ky*5+kx), P: fp32 [H*W, C], out: f16 NHWC [n, H, W, C].  C = 64.
__device__ __forceinline__ int border_pattern(int v, int n) { return v < 2 ? v : (v >= n - 2 ? v - (n - 5) : 2); }

__global__ void __launch_bounds__(256)
dec_l1_kernel(const float* __restrict__ taps, const float* __restrict__ P, __half* __restrict__ out, int H, int W) {
  constexpr int C = 64;
  __shared__ float sT[25 * C];
  __shared__ __align__(16) float sS[25 * C];
  const int img = blockIdx.x;
  const float* t = taps + size_t(img) * 25 * C;
  for (int e = threadIdx.x; e < 25 * C; e += 256) sT[e] = t[e];
  __syncthreads();
  for (int e = threadIdx.x; e < 25 * C; e += 256) {
    const int pat = e / C, c = e % C;
    const int py = pat / 5, px = pat % 5;
    // pattern 0: first row (ky >= 2), 1: second row (ky >= 1), 2: interior, 3: ky <= 3, 4: ky <= 2
    const int ky0 = py == 0 ? 2 : (py == 1 ? 1 : 0), ky1 = py == 4 ? 2 : (py == 3 ? 3 : 4);
    const int kx0 = px == 0 ? 2 : (px == 1 ? 1 : 0), kx1 = px == 4 ? 2 : (px == 3 ? 3 : 4);
    float s = 0.f;
    for (int ky = ky0; ky <= ky1; ++ky)
      for (int kx = kx0; kx <= kx1; ++kx) s += sT[(ky * 5 + kx) * C + c];
    sS[e] = s;
  }
  __syncthreads();
  __half* o = out + size_t(img) * H * W * C;
  const int items = H * W * (C / 8);
  for (int e = threadIdx.x; e < items; e += 256) {
    const int pix = e >> 3, c8 = (e & 7) * 8;
    const int y = pix / W, x = pix % W;
    const float* s = sS + (border_pattern(y, H) * 5 + border_pattern(x, W)) * C + c8;
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(P + size_t(pix) * C + c8));
    const float4 p1 = __ldg(reinterpret_cast<const float4*>(P + size_t(pix) * C + c8 + 4));
    const float4 s0 = *reinterpret_cast<const float4*>(s);
    const float4 s1 = *reinterpret_cast<const float4*>(s + 4);
    uint4 r;
    r.x = pack_half2(fmaxf(p0.x + s0.x, 0.f), fmaxf(p0.y + s0.y, 0.f));
    r.y = pack_half2(fmaxf(p0.z + s0.z, 0.f), fmaxf(p0.w + s0.w, 0.f));
    r.z = pack_half2(fmaxf(p1.x + s1.x, 0.f), fmaxf(p1.y + s1.y, 0.f));
    r.w = pack_half2(fmaxf(p1.z + s1.z, 0.f), fmaxf(p1.w + s1.w, 0.f));
    *reinterpret_cast<uint4*>(o + size_t(pix) * C + c8) = r;
  }
}

// ------------------------------------------------------------------------------------------------ conv3x3 + compositing
constexpr int C3_TH = 16, C3_TW = 32, C3_C = 64, C3_CP = 72;   // 72 halfs = 144 B per pixel: conflict-free LDS.128
constexpr int C3_HALO = (C3_TH + 2) * (C3_TW + 2);
constexpr int C3_SMEM = C3_HALO * C3_CP * 2 + 9 * C3_C * 4 * 4;

// act: f16 NHWC [n_frames*S, H, W, 64]; w: fp32 [9][64][4]; imgs: fp32 [n_frames,3,H,W];
// recons (opt) fp32 [n_frames,S,3,H,W]; masks (opt) fp32 [n_frames,S,1,H,W]
__global__ void __launch_bounds__(256, 2)
conv3x3_composite_kernel(const __half* __restrict__ act, const float* __restrict__ w, const float* __restrict__ bias,
                         float* __restrict__ imgs, float* __restrict__ recons, float* __restrict__ masks, int S, int H,
                         int W) {
  extern __shared__ __align__(16) uint8_t c3_smem[];
  __half* sIn = reinterpret_cast<__half*>(c3_smem);
  float4* sW = reinterpret_cast<float4*>(c3_smem + C3_HALO * C3_CP * 2);
  const int frame = blockIdx.y;
  const int tiles_x = W / C3_TW;
  const int y0 = (blockIdx.x / tiles_x) * C3_TH, x0 = (blockIdx.x % tiles_x) * C3_TW;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // pixels (ty, tx) and (ty + 8, tx)
  for (int e = threadIdx.x; e < 9 * C3_C; e += 256) sW[e] = reinterpret_cast<const float4*>(w)[e];
  const float4 b4 = *reinterpret_cast<const float4*>(bias);
  float mx[2] = {-1e30f, -1e30f}, den[2] = {0.f, 0.f}, rgb[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  const size_t plane = size_t(H) * W;
  for (int s = 0; s < S; ++s) {
    const __half* a = act + (size_t(frame) * S + s) * plane * C3_C;
    __syncthreads();   // previous slot's tile fully consumed (also orders the sW fill on the first pass)
    for (int e = threadIdx.x; e < C3_HALO * 8; e += 256) {
      const int p = e >> 3, c = e & 7;
      const int gy = y0 - 1 + p / (C3_TW + 2), gx = x0 - 1 + p % (C3_TW + 2);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W)
        v = __ldg(reinterpret_cast<const uint4*>(a + (size_t(gy) * W + gx) * C3_C + c * 8));
      *reinterpret_cast<uint4*>(sIn + p * C3_CP + c * 8) = v;
    }
    __syncthreads();
    float acc[2][4] = {{b4.x, b4.y, b4.z, b4.w}, {b4.x, b4.y, b4.z, b4.w}};
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap % 3;
      const __half* pa = sIn + ((ty + ky) * (C3_TW + 2) + tx + kx) * C3_CP;
      const __half* pb = pa + 8 * (C3_TW + 2) * C3_CP;
      const float4* wt = sW + tap * C3_C;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        const uint4 ua = *reinterpret_cast<const uint4*>(pa + c8 * 8);
        const uint4 ub = *reinterpret_cast<const uint4*>(pb + c8 * 8);
        const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w}, wb[4] = {ub.x, ub.y, ub.z, ub.w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&wa[h]));
          const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&wb[h]));
          const float4 w0 = wt[c8 * 8 + h * 2], w1 = wt[c8 * 8 + h * 2 + 1];
          acc[0][0] += fa.x * w0.x + fa.y * w1.x; acc[0][1] += fa.x * w0.y + fa.y * w1.y;
          acc[0][2] += fa.x * w0.z + fa.y * w1.z; acc[0][3] += fa.x * w0.w + fa.y * w1.w;
          acc[1][0] += fb.x * w0.x + fb.y * w1.x; acc[1][1] += fb.x * w0.y + fb.y * w1.y;
          acc[1][2] += fb.x * w0.z + fb.y * w1.z; acc[1][3] += fb.x * w0.w + fb.y * w1.w;
        }
      }
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int gy = y0 + ty + 8 * p, gx = x0 + tx;
      const float logit = acc[p][3];                          // channel order RGB then mask (SAVi.py:251-252)
      const float nm = fmaxf(mx[p], logit);
      const float corr = __expf(mx[p] - nm), e = __expf(logit - nm);
      den[p] = den[p] * corr + e;
      rgb[p][0] = rgb[p][0] * corr + e * acc[p][0];
      rgb[p][1] = rgb[p][1] * corr + e * acc[p][1];
      rgb[p][2] = rgb[p][2] * corr + e * acc[p][2];
      mx[p] = nm;
      if (recons) {
        float* r = recons + ((size_t(frame) * S + s) * 3) * plane + size_t(gy) * W + gx;
        r[0] = acc[p][0]; r[plane] = acc[p][1]; r[2 * plane] = acc[p][2];
      }
      if (masks) masks[(size_t(frame) * S + s) * plane + size_t(gy) * W + gx] = logit;
    }
  }
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int gy = y0 + ty + 8 * p, gx = x0 + tx;
    const float inv = 1.f / den[p];
    float* o = imgs + size_t(frame) * 3 * plane + size_t(gy) * W + gx;
    o[0] = rgb[p][0] * inv; o[plane] = rgb[p][1] * inv; o[2 * plane] = rgb[p][2] * inv;
    if (masks) {
      for (int s = 0; s < S; ++s) {
        float* m = masks + (size_t(frame) * S + s) * plane + size_t(gy) * W + gx;
        *m = __expf(*m - mx[p]) * inv;
      }
    }
  }
}

constexpr int DEC_CHUNK_FRAMES = 256;

struct DecBuffers {
  __half* slots16;
  float* taps32;
  __half *actA, *actB;
};

static size_t align256d(size_t n) { return (n + 255) & ~size_t(255); }

static size_t dec_carve(const tocvp_dec_weights& w, int n_frames, DecBuffers* db, uint8_t* base) {
  const int chunk = n_frames < DEC_CHUNK_FRAMES ? n_frames : DEC_CHUNK_FRAMES;
  const size_t nsi = size_t(chunk) * w.num_slots;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* p = base ? base + off : nullptr;
    off += align256d(bytes);
    return p;
  };
  DecBuffers t;
  t.slots16 = reinterpret_cast<__half*>(take(nsi * w.slot_dim * 2));
  t.taps32 = reinterpret_cast<float*>(take(nsi * 25 * w.hidden * 4));
  t.actA = reinterpret_cast<__half*>(take(nsi * w.H * w.W * w.hidden * 2));
  t.actB = reinterpret_cast<__half*>(take(nsi * w.H * w.W * w.hidden * 2));
  if (db) *db = t;
  return off;
}

}  // namespace tocvp

using namespace tocvp;

extern "C" size_t tocvp_sizeof_dec_weights(void) { return sizeof(tocvp_dec_weights); }

extern "C" size_t tocvp_savi_decode_workspace_bytes(const tocvp_dec_weights* w, int n_frames) {
  if (!w || n_frames <= 0) return 0;
  return dec_carve(*w, n_frames, nullptr, nullptr);
}

extern "C" int tocvp_savi_decode(const tocvp_dec_weights* w, const float* slots, int n_frames, float* recons_imgs,
                                 float* recons, float* masks, void* workspace, size_t ws_bytes, void* stream,
                                 void* const* conv_events, int n_conv_events) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TOCVP_CHECK_ARG(w && slots && recons_imgs && workspace && n_frames > 0);
  TOCVP_CHECK_ARG(w->hidden == 64 && w->slot_dim % 8 == 0 && w->H % C3_TH == 0 && w->W % C3_TW == 0 && w->H >= 5 && w->W >= 5);
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  if (ws_bytes < dec_carve(*w, n_frames, nullptr, nullptr)) {
    set_last_error(__FILE__, __LINE__, "savi_decode: workspace too small");
    return TOCVP_ERR_WORKSPACE;
  }
  static bool attr_set = false;
  if (!attr_set) {
    TOCVP_CUDA(cudaFuncSetAttribute(conv3x3_composite_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C3_SMEM));
    attr_set = true;
  }
  DecBuffers db;
  dec_carve(*w, n_frames, &db, static_cast<uint8_t*>(workspace));
  const int S = w->num_slots, D = w->slot_dim, H = w->H, W = w->W, C = w->hidden;
  const size_t plane = size_t(H) * W;
  for (int f0 = 0; f0 < n_frames; f0 += DEC_CHUNK_FRAMES) {
    const int nf = (n_frames - f0) < DEC_CHUNK_FRAMES ? (n_frames - f0) : DEC_CHUNK_FRAMES;
    const int nsi = nf * S;
    const size_t n4 = size_t(nsi) * D / 4;
    f32_to_f16_kernel<<<int((n4 + 255) / 256), 256, 0, st>>>(slots + size_t(f0) * S * D, db.slots16, n4);
    TOCVP_LAUNCHED();
    TOCVP_TRY(gemm_f16(db.slots16, D, static_cast<const __half*>(w->w1_taps), D, nsi, 25 * C, D, nullptr, 0, nullptr, 0,
                       1, 0, db.taps32, 25 * C, nullptr, 0, st));
    dec_l1_kernel<<<nsi, 256, 0, st>>>(db.taps32, w->p1, db.actA, H, W);
    TOCVP_LAUNCHED();
    __half* bufs[2] = {db.actA, db.actB};
    for (int l = 0; l < 3; ++l) {
      // optional CUDA-event pair around each conv launch (bench.py measures the dominant kernel live, in the step)
      const int ev = 2 * ((f0 / DEC_CHUNK_FRAMES) * 3 + l);
      const bool prof = conv_events != nullptr && ev + 1 < n_conv_events;
      if (prof) TOCVP_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(conv_events[ev]), st));
      TOCVP_TRY(conv5x5_f16(bufs[l & 1], static_cast<const __half*>(w->w_conv[l]), w->b_conv[l], bufs[(l + 1) & 1], nsi, H,
                            W, C, C, 1, st));
      if (prof) TOCVP_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(conv_events[ev + 1]), st));
    }
    const dim3 grid((H / C3_TH) * (W / C3_TW), nf);
    conv3x3_composite_kernel<<<grid, 256, C3_SMEM, st>>>(
        db.actB, w->w_out, w->b_out, recons_imgs + size_t(f0) * 3 * plane,
        recons ? recons + size_t(f0) * S * 3 * plane : nullptr, masks ? masks + size_t(f0) * S * plane : nullptr, S, H, W);
    TOCVP_LAUNCHED();
  }
  return TOCVP_OK;
}
