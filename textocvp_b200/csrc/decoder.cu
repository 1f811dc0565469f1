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
//  final conv3x3 64->4 (decoders.py:110-116): the same tcgen05 implicit-GEMM kernel with a 3x3 tap set and the output
//      channels zero-padded to N = 16; it writes one float4 (RGB + mask logit) per pixel and slot.
//  compositing (SAVi.py:251-255): `composite_kernel`, softmax over the slot axis + weighted sum, one thread per pixel.
#include <mutex>

#include "host_util.h"
#include "ptx.cuh"

namespace tocvp {

int gemm_f16(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K, const float* bias, int relu,
             const float* residual, int ldr, int res_div, int res_mod, float* out32, int ld32, __half* out16,
             int ld16, cudaStream_t stream);
int conv5x5_f16(const __half* x, const __half* wpacked, const float* bias, __half* out, int n_img, int H, int W, int cin,
                int cout, int relu, cudaStream_t stream);
int conv3x3_head_f16(const __half* x, const __half* wpacked, const float* bias, float* out4, int n_img, int H, int W,
                     cudaStream_t stream);
int conv3x3_head_composite_f16(const __half* x, const __half* w_taps, const float* bias, int n_frames, int S, int H, int W,
                               float* imgs, float* recons, float* masks, cudaStream_t stream);
int conv3x3_head_taps_f16(const __half* x, const __half* w_taps, const float* bias, float* out4, int n_img, int H, int W,
                          cudaStream_t stream);
int conv5x5_gen_f16(const float* P, const float* S, const __half* x_dummy, const __half* wpacked, const float* bias,
                    __half* out, int n_img, int H, int W, cudaStream_t stream);

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

// Pattern sums only (fused path): S[img][pattern][c] = sum of the taps valid for that 5x5 border pattern; layer 1's
// activation is then generated inside the layer-2 convolution kernel (conv5x5_tc.cu, ConvGen) and never stored.
__global__ void __launch_bounds__(256)
dec_l1_patterns_kernel(const float* __restrict__ taps, float* __restrict__ S, int n_img) {
  constexpr int C = 64;
  const int sub = threadIdx.x >> 6, c = threadIdx.x & 63;       // 4 slot-images per CTA, one thread per channel
  const int img = blockIdx.x * 4 + sub;
  if (img >= n_img) return;
  const float* t = taps + size_t(img) * 25 * C + c;
  float tv[25];                                                 // this channel's 25 tap values (coalesced over c)
#pragma unroll
  for (int k = 0; k < 25; ++k) tv[k] = __ldg(t + k * C);
  float* o = S + size_t(img) * 25 * C + c;
  // all bounds are compile-time constants after unrolling: no branches, same summation order (ky outer, kx inner) as the
  // first version (which looped with run-time bounds over a shared-memory copy and was issue-bound: 91 % issue-active)
#pragma unroll
  for (int pat = 0; pat < 25; ++pat) {
    const int py = pat / 5, px = pat % 5;
    const int ky0 = py == 0 ? 2 : (py == 1 ? 1 : 0), ky1 = py == 4 ? 2 : (py == 3 ? 3 : 4);
    const int kx0 = px == 0 ? 2 : (px == 1 ? 1 : 0), kx1 = px == 4 ? 2 : (px == 3 ? 3 : 4);
    float s = 0.f;
#pragma unroll
    for (int ky = 0; ky < 5; ++ky)
#pragma unroll
      for (int kx = 0; kx < 5; ++kx)
        if (ky >= ky0 && ky <= ky1 && kx >= kx0 && kx <= kx1) s += tv[ky * 5 + kx];
    o[pat * C] = s;
  }
}

// Layer 1, second version (default): pixel-stationary.  A CTA owns 8 consecutive pixels of the image plane for ALL
// slot-images: its 8 x 64 P values stay in registers, and per slot-image it reads only the pattern sums those 8 pixels
// need (mostly the one interior pattern: 256 B instead of 8 x 256 B of P per image in dec_l1_kernel), so the kernel is
// bound by its 1 GiB of output writes rather than by 2 GiB of L2 reads.  thread = (slot-image lane 0..31, 8-channel chunk).
// S: fp32 [n_img, 25, 64] from dec_l1_patterns_kernel.  Requires W % 8 == 0.
__global__ void __launch_bounds__(256)
dec_l1_pixel_kernel(const float* __restrict__ S, const float* __restrict__ P, __half* __restrict__ out, int n_img, int H,
                    int W) {
  constexpr int C = 64;
  const int chunk = threadIdx.x & 7, il = threadIdx.x >> 3;         // 8-channel chunk, slot-image lane
  // Persistent over the 8-pixel groups (grid <= number of SMs): every CTA is resident from the first dispatch, so a
  // kernel submitted on another stream right behind this one (the decoder's convolutions, chunk pipeline) is not held
  // back by a queue of pending CTAs.
  for (int pg = blockIdx.x; pg < (H * W) / 8; pg += gridDim.x) {
    const int pix0 = pg * 8;
    const int y = pix0 / W, x0 = pix0 % W;
    float p[8][8];
    int pat[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(P + size_t(pix0 + k) * C + chunk * 8));
      const float4 b = __ldg(reinterpret_cast<const float4*>(P + size_t(pix0 + k) * C + chunk * 8 + 4));
      p[k][0] = a.x; p[k][1] = a.y; p[k][2] = a.z; p[k][3] = a.w;
      p[k][4] = b.x; p[k][5] = b.y; p[k][6] = b.z; p[k][7] = b.w;
      pat[k] = border_pattern(y, H) * 5 + border_pattern(x0 + k, W);
    }
    const size_t plane = size_t(H) * W;
    for (int img = il; img < n_img; img += 32) {
      const float* s = S + size_t(img) * 25 * C + chunk * 8;
      __half* o = out + (size_t(img) * plane + pix0) * C + chunk * 8;
      int cur = -1;
      float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (pat[k] != cur) {                                           // at most 3 distinct patterns in 8 pixels of a row
          cur = pat[k];
          s0 = __ldg(reinterpret_cast<const float4*>(s + cur * C));
          s1 = __ldg(reinterpret_cast<const float4*>(s + cur * C + 4));
        }
        uint4 r;
        r.x = pack_half2_relu(p[k][0] + s0.x, p[k][1] + s0.y);
        r.y = pack_half2_relu(p[k][2] + s0.z, p[k][3] + s0.w);
        r.z = pack_half2_relu(p[k][4] + s1.x, p[k][5] + s1.y);
        r.w = pack_half2_relu(p[k][6] + s1.z, p[k][7] + s1.w);
        *reinterpret_cast<uint4*>(o + size_t(k) * C) = r;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ compositing
// maps: fp32 [n_frames*S, H*W, 4] (RGB + mask logit, SAVi.py:251-252) from the tcgen05 conv3x3 head.
// imgs fp32 [n_frames,3,H,W]; recons (opt) [n_frames,S,3,H,W]; masks (opt) [n_frames,S,1,H,W].
// One thread per pixel: softmax over the slot axis + weighted sum in registers, fully coalesced 16-byte reads.
__global__ void __launch_bounds__(256)
composite_kernel(const float4* __restrict__ maps, float* __restrict__ imgs, float* __restrict__ recons,
                 float* __restrict__ masks, int S, int plane, int n_frames) {
  const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= size_t(n_frames) * plane) return;
  const int frame = int(idx / plane), pix = int(idx % plane);
  const float4* m = maps + size_t(frame) * S * plane + pix;
  float mx = -1e30f;
  for (int s = 0; s < S; ++s) mx = fmaxf(mx, __ldg(m + size_t(s) * plane).w);
  float den = 0.f, r = 0.f, g = 0.f, b = 0.f;
  for (int s = 0; s < S; ++s) {
    const float4 v = __ldg(m + size_t(s) * plane);
    const float e = __expf(v.w - mx);
    den += e; r += e * v.x; g += e * v.y; b += e * v.z;
    if (recons) {
      float* o = recons + (size_t(frame) * S + s) * 3 * plane + pix;
      o[0] = v.x; o[plane] = v.y; o[2 * size_t(plane)] = v.z;
    }
  }
  const float inv = 1.f / den;
  float* o = imgs + size_t(frame) * 3 * plane + pix;
  o[0] = r * inv; o[plane] = g * inv; o[2 * size_t(plane)] = b * inv;
  if (masks) {
    for (int s = 0; s < S; ++s)
      masks[(size_t(frame) * S + s) * plane + pix] = __expf(__ldg(m + size_t(s) * plane).w - mx) * inv;
  }
}

constexpr int DEC_CHUNK_FRAMES = 256;

// tocvp_tuning.decode_mode (per call), bit 0: 1 = generate layer 1 inside the layer-2 conv, 0 = separate bandwidth kernel (default).
// Measured r1 (tools/ab_decode.py, same process): the fused layer-2 conv takes 1.86 ms instead of 1.44 ms per chunk -- the
// generator warps share the shared-memory port and the issue slots the implicit GEMM is bound by -- which is more than the
// 0.36 ms of dec_l1_kernel it removes, so the fused path is kept (tested, parity-green) but not the default.
// bit 1: 0 = head conv3x3 with the 9 taps in the GEMM's N dimension (default), 1 = shifted-window
// kernel with N = 16 (first version, kept for A/B)
// bit 2: 0 = pixel-stationary layer-1 kernel (default), 1 = image-stationary first version

// bit 3: 0 = chunk-pipelined decode (default): layer 1 of chunk i+1 is written on a side stream while the convolutions of
// chunk i run (the pixel-stationary kernel needs no shared memory / TMEM and 1 CTA per SM of registers, so it is
// co-resident with the persistent conv CTAs and its 1 GiB write stream hides under tensor-bound time); 1 = serial.

struct DecBuffers {
  __half* slots16;    // [n_frames*S, D]   (all chunks: the layer-1 front end runs once for the whole call)
  float* taps32;      // [n_frames*S, 25*C]
  float* pat32;       // [n_frames*S, 25*C]
  __half *actA, *actB, *actC;
  float* maps4;
};

// Side stream + events of the pipelined decode, one set per device, created on first use.  The side stream only ever runs
// work that is forked from / joined back into the caller's stream by events inside one call, so calls stay ordered on
// the caller's stream (and capturable).  The set is guarded by a mutex held for the enqueue phase of a call.
struct DecSide {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, l1_done[2] = {nullptr, nullptr}, conv_done[2] = {nullptr, nullptr};
};
static std::mutex g_side_mutex;
static DecSide g_side[16];

static int dec_side(DecSide** out) {
  int dev = 0;
  TOCVP_CUDA(cudaGetDevice(&dev));
  TOCVP_CHECK_ARG(dev >= 0 && dev < 16);
  DecSide& sd = g_side[dev];
  if (!sd.stream) {
    TOCVP_CUDA(cudaStreamCreateWithFlags(&sd.stream, cudaStreamNonBlocking));
    TOCVP_CUDA(cudaEventCreateWithFlags(&sd.fork, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
      TOCVP_CUDA(cudaEventCreateWithFlags(&sd.l1_done[i], cudaEventDisableTiming));
      TOCVP_CUDA(cudaEventCreateWithFlags(&sd.conv_done[i], cudaEventDisableTiming));
    }
  }
  *out = &sd;
  return TOCVP_OK;
}

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
  const size_t nsi_all = size_t(n_frames) * w.num_slots;
  DecBuffers t;
  t.slots16 = reinterpret_cast<__half*>(take(nsi_all * w.slot_dim * 2));
  t.taps32 = reinterpret_cast<float*>(take(nsi_all * 25 * w.hidden * 4));
  t.pat32 = reinterpret_cast<float*>(take(nsi_all * 25 * w.hidden * 4));
  t.actA = reinterpret_cast<__half*>(take(nsi * w.H * w.W * w.hidden * 2));
  t.actB = reinterpret_cast<__half*>(take(nsi * w.H * w.W * w.hidden * 2));
  t.actC = n_frames > chunk ? reinterpret_cast<__half*>(take(nsi * w.H * w.W * w.hidden * 2)) : nullptr;
  t.maps4 = reinterpret_cast<float*>(take(nsi * w.H * w.W * 4 * 4));
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
  OptsScope scope(w->tuning);
  const int dmode = opts().decode_mode;
  const bool g_dec_fuse_l1 = (dmode & 1) != 0, g_dec_head_taps = (dmode & 2) == 0, g_dec_l1_pixel = (dmode & 4) == 0,
             g_dec_overlap = (dmode & 8) == 0;
  TOCVP_CHECK_ARG(w->hidden == 64 && w->slot_dim % 8 == 0 && w->H % 16 == 0 && w->W % 32 == 0);
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  if (ws_bytes < dec_carve(*w, n_frames, nullptr, nullptr)) {
    set_last_error(__FILE__, __LINE__, "savi_decode: workspace too small");
    return TOCVP_ERR_WORKSPACE;
  }
  DecBuffers db;
  dec_carve(*w, n_frames, &db, static_cast<uint8_t*>(workspace));
  const int S = w->num_slots, D = w->slot_dim, H = w->H, W = w->W, C = w->hidden;
  const size_t plane = size_t(H) * W;
  const int nsi_all = n_frames * S;
  // ---- layer-1 front end for the whole call: f16 slots, per-tap vectors (one GEMM), 5x5 border-pattern sums
  {
    const size_t n4 = size_t(nsi_all) * D / 4;
    f32_to_f16_kernel<<<int((n4 + 255) / 256), 256, 0, st>>>(slots, db.slots16, n4);
    TOCVP_LAUNCHED();
    TOCVP_TRY(gemm_f16(db.slots16, D, static_cast<const __half*>(w->w1_taps), D, nsi_all, 25 * C, D, nullptr, 0, nullptr,
                       0, 1, 0, db.taps32, 25 * C, nullptr, 0, st));
  }
  const int chunk0 = n_frames < DEC_CHUNK_FRAMES ? n_frames : DEC_CHUNK_FRAMES;
  // layer 1 is generated inside the layer-2 convolution only on request (tocvp_set_decode_mode bit 0) and when the pair
  // kernel applies to every chunk (even tile count)
  const bool even_tiles = ((chunk0 * S * (H / 16) * (W / 32)) % 2 == 0) && (((n_frames % chunk0) * S * (H / 16) * (W / 32)) % 2 == 0);
  const bool fused_l1 = even_tiles && g_dec_fuse_l1;
  const bool pixel_l1 = !fused_l1 && (W % 8 == 0) && g_dec_l1_pixel;
  if (fused_l1 || pixel_l1) {
    dec_l1_patterns_kernel<<<(nsi_all + 3) / 4, 256, 0, st>>>(db.taps32, db.pat32, nsi_all);
    TOCVP_LAUNCHED();
  }
  const int n_chunks = (n_frames + DEC_CHUNK_FRAMES - 1) / DEC_CHUNK_FRAMES;
  const bool overlap = g_dec_overlap && pixel_l1 && n_chunks > 1;
  std::unique_lock<std::mutex> side_lock(g_side_mutex, std::defer_lock);
  DecSide* sd = nullptr;
  if (overlap) {
    side_lock.lock();
    TOCVP_TRY(dec_side(&sd));
    TOCVP_CUDA(cudaEventRecord(sd->fork, st));
    TOCVP_CUDA(cudaStreamWaitEvent(sd->stream, sd->fork, 0));
  }
  // chunk i: layer 1 -> X[i&1];  L2: X -> actB;  L3: actB -> X;  L4: X -> actB;  head reads actB.  (serial mode: X = actA)
  __half* xbuf[2] = {db.actA, overlap ? db.actC : db.actA};
  auto layer1 = [&](int ci, cudaStream_t s_l1) -> int {
    const int f0 = ci * DEC_CHUNK_FRAMES;
    const int nsi = ((n_frames - f0) < DEC_CHUNK_FRAMES ? (n_frames - f0) : DEC_CHUNK_FRAMES) * S;
    const size_t so = size_t(f0) * S * 25 * C;
    if (pixel_l1) {
      const int groups = (H * W) / 8;
      dec_l1_pixel_kernel<<<groups < num_sms() ? groups : num_sms(), 256, 0, s_l1>>>(db.pat32 + so, w->p1, xbuf[ci & 1], nsi,
                                                                                    H, W);
      TOCVP_LAUNCHED();
    } else if (!fused_l1) {
      dec_l1_kernel<<<nsi, 256, 0, s_l1>>>(db.taps32 + so, w->p1, xbuf[ci & 1], H, W);
      TOCVP_LAUNCHED();
    }
    return TOCVP_OK;
  };
  if (overlap) {
    TOCVP_TRY(layer1(0, sd->stream));
    TOCVP_CUDA(cudaEventRecord(sd->l1_done[0], sd->stream));
  }
  for (int ci = 0; ci < n_chunks; ++ci) {
    const int f0 = ci * DEC_CHUNK_FRAMES;
    const int nf = (n_frames - f0) < DEC_CHUNK_FRAMES ? (n_frames - f0) : DEC_CHUNK_FRAMES;
    const int nsi = nf * S;
    if (overlap) {
      if (ci + 1 < n_chunks) {
        // Layer 1 of chunk ci+1 starts when layer 4 of chunk ci-1 has finished (the last reader of its buffer X[(ci+1)&1]):
        // it shares the HBM with the bandwidth-bound head conv + compositing of chunk ci-1 (neither is power-hungry) and
        // whatever is left of it runs co-resident with the first convolution of chunk ci.  Starting it under the
        // convolutions only (r1 A/B) gains on a box with power headroom but nothing on a power-capped one: the conv slows
        // down by exactly the layer-1 time.
        if (ci >= 1) TOCVP_CUDA(cudaStreamWaitEvent(sd->stream, sd->conv_done[(ci - 1) & 1], 0));
        TOCVP_TRY(layer1(ci + 1, sd->stream));
        TOCVP_CUDA(cudaEventRecord(sd->l1_done[(ci + 1) & 1], sd->stream));
      }
      TOCVP_CUDA(cudaStreamWaitEvent(st, sd->l1_done[ci & 1], 0));
    } else {
      TOCVP_TRY(layer1(ci, st));
    }
    __half* X = xbuf[ci & 1];
    const __half* src[3] = {X, db.actB, X};
    __half* dst[3] = {db.actB, X, db.actB};
    for (int l = 0; l < 3; ++l) {
      // optional CUDA-event pair around each conv launch (bench.py measures the dominant kernel live, in the step)
      const int ev = 2 * (ci * 3 + l);
      const bool prof = conv_events != nullptr && ev + 1 < n_conv_events;
      if (prof) TOCVP_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(conv_events[ev]), st));
      if (l == 0 && fused_l1) {
        TOCVP_TRY(conv5x5_gen_f16(w->p1, db.pat32 + size_t(f0) * S * 25 * C, X, static_cast<const __half*>(w->w_conv[0]),
                                  w->b_conv[0], db.actB, nsi, H, W, st));
      } else {
        TOCVP_TRY(conv5x5_f16(src[l], static_cast<const __half*>(w->w_conv[l]), w->b_conv[l], dst[l], nsi, H, W, C, C, 1,
                              st));
      }
      if (prof) TOCVP_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(conv_events[ev + 1]), st));
    }
    if (overlap) TOCVP_CUDA(cudaEventRecord(sd->conv_done[ci & 1], st));
    // conv3x3 64 -> 4 head on the tensor cores (N padded to 16), then softmax-over-slots compositing
    const size_t npix = size_t(nf) * plane;
    float* imgs_c = recons_imgs + size_t(f0) * 3 * plane;
    float* recons_c = recons ? recons + size_t(f0) * S * 3 * plane : nullptr;
    float* masks_c = masks ? masks + size_t(f0) * S * plane : nullptr;
    if (g_dec_head_taps && w->w_out_taps != nullptr && !(dmode & 16) && S <= 11) {
      // head conv with the slot-softmax compositing in its epilogue (SAVi.py:251-255): no fp32 map round trip, no extra launch
      TOCVP_TRY(conv3x3_head_composite_f16(db.actB, static_cast<const __half*>(w->w_out_taps), w->b_out, nf, S, H, W, imgs_c,
                                           recons_c, masks_c, st));
      continue;
    }
    if (g_dec_head_taps && w->w_out_taps != nullptr) {
      TOCVP_TRY(conv3x3_head_taps_f16(db.actB, static_cast<const __half*>(w->w_out_taps), w->b_out, db.maps4, nsi, H, W, st));
    } else {
      TOCVP_TRY(conv3x3_head_f16(db.actB, static_cast<const __half*>(w->w_out), w->b_out, db.maps4, nsi, H, W, st));
    }
    composite_kernel<<<int((npix + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(db.maps4), imgs_c, recons_c,
                                                              masks_c, S, int(plane), nf);
    TOCVP_LAUNCHED();
  }
  return TOCVP_OK;
}

// Last layer of ConvDecoder.forward on its own (reference src/models/EncodersDecoders/decoders.py:104-108): conv3x3 64 -> 4,
// no activation, on an NHWC f16 activation -> NHWC fp32 [n_img, H, W, 4] (RGB + mask logit per pixel).
extern "C" int tocvp_conv3x3_head(const tocvp_dec_weights* w, const void* x_nhwc_f16, int n_img, float* out_nhwc4,
                                  void* stream) {
  TOCVP_CHECK_ARG(w && x_nhwc_f16 && out_nhwc4 && n_img > 0 && w->hidden == 64);
  TOCVP_CHECK_ARG(w->H % 16 == 0 && w->W % 32 == 0);
  OptsScope scope(w->tuning);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!(opts().decode_mode & 2) && w->w_out_taps != nullptr)
    return conv3x3_head_taps_f16(static_cast<const __half*>(x_nhwc_f16), static_cast<const __half*>(w->w_out_taps), w->b_out,
                                 out_nhwc4, n_img, w->H, w->W, st);
  return conv3x3_head_f16(static_cast<const __half*>(x_nhwc_f16), static_cast<const __half*>(w->w_out), w->b_out, out_nhwc4,
                          n_img, w->H, w->W, st);
}
