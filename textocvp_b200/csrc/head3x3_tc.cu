// Decoder head: conv3x3 64 -> 4 (RGB + mask logit, reference src/models/EncodersDecoders/decoders.py:110-116) with the
// nine filter taps in the N dimension of ONE tcgen05 GEMM per halo tile.
//
// The shifted-window implicit GEMM of conv5x5_tc.cu re-reads the 4 KB A operand from shared memory once per tap and
// k-step; with only N = 16 output columns that operand read (not the tensor pipe, not HBM) sets the pace: 0.52 ms per
// 2048 slot-images where the 1 GiB input takes 0.17 ms to stream (r1 ncu).  Here the UNSHIFTED halo tile is multiplied
// by all nine taps at once,
//     D[r][tap*4 + co] = sum_c X[r][c] * W[co][c][tap]        r = halo pixel (row-major 10 x 34), N = 36 (padded to 48),
// 12 MMAs per 8 x 32 tile instead of 72, and the shift moves to the epilogue: D goes TMEM -> registers -> shared memory and
//     out[y][x][co] = b[co] + sum_tap D[(y+ty)*34 + (x+tx)][tap*4 + co]
// is gathered from there (36 shared loads per pixel, bank-conflict free with an odd row stride).
// Two halo buffers and two accumulator sets: load of tile i+1, MMAs of tile i and epilogue of tile i-1 overlap.
//   warp 0: TMA producer   warp 1: TMEM alloc + MMA issue   warps 2-5: epilogue
//
// COMPOSITE = true (round 2, default in tocvp_savi_decode): the slot-softmax compositing of SAVi.decode (reference
// src/models/SAVi.py:251-255) runs in this kernel's epilogue.  A CTA then walks a pixel tile through ALL S slot-images of a
// frame before it moves on; every epilogue thread keeps the four head outputs of its two pixels for each slot in a
// thread-private shared-memory column and, after the last slot, does softmax over the slot axis + the weighted RGB sum
// with exactly the arithmetic of composite_kernel (decoder.cu), so both routes are bit-identical.  The fp32
// [n, H, W, 4] map (134 MB per 256-frame chunk, written and read back) and the separate launch disappear.
#include "host_util.h"
#include "ptx.cuh"

namespace tocvp {

constexpr int HD_TH = 8, HD_TW = 32;             // output tile
constexpr int HD_HR = HD_TH + 2, HD_WB = HD_TW + 2;   // halo rows, halo row pitch (pixels)
constexpr int HD_ROWS = HD_HR * HD_WB;           // 340 halo pixels
constexpr int HD_MT = (HD_ROWS + 127) / 128;     // 3 M-tiles (the last one reads 44 rows past the tile: never used)
constexpr int HD_N = 48;                         // 9 taps x 4 channels = 36, padded to a multiple of 16
constexpr int HD_A_BYTES = HD_ROWS * 128;        // 43520 bytes landed by the TMA
constexpr int HD_A_STRIDE = HD_MT * 128 * 128;   // 49152: buffer pitch, covers the over-read, 1024-aligned
constexpr int HD_W_BYTES = HD_N * 128;           // 6144
// D staging row stride (floats).  36 = 9 taps x 4 channels, rows of 144 bytes: every access is a 128-bit LDS / STS of one
// tap's 4 channels, and consecutive rows land in different 16-byte bank groups (144 / 16 = 9, odd), so a quarter-warp -- the
// unit a 128-bit shared access is served in -- is conflict free.  (Round 1 used 37 and scalar accesses: 4x the LSU
// instructions in an epilogue that ncu showed to be the kernel's pace: 4 epilogue warps, one per scheduler, 9 % warps active.)
constexpr int HD_DS = 36;
constexpr int HD_D_BYTES = HD_MT * 128 * HD_DS * 4;
constexpr int HD_SMEM = 2 * HD_A_STRIDE + HD_W_BYTES + HD_D_BYTES + 256 + 1024;
constexpr int HD_MAX_SLOTS = 11;
constexpr int HD_COMP_BYTES = HD_TH * HD_TW * 16;   // per slot: float4 per pixel of the tile
constexpr int HD_THREADS = 320;                 // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter)
constexpr int HD_ETH = HD_THREADS - 64;         // epilogue threads = pixels of the 8 x 32 tile
static_assert(HD_ETH == HD_TH * HD_TW, "one epilogue thread per output pixel");
constexpr int HD_TMEM_COLS = 512;                // 2 x 3 x 48 = 288 columns used (two accumulator sets)

__device__ __forceinline__ uint64_t hd_desc(uint32_t saddr) {   // K-major, SWIZZLE_128B, dense 128-byte rows
  return uint64_t((saddr >> 4) & 0x3FFF) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
         (uint64_t(2) << 61);
}

// Pipeline (r1, second version): TWO halo buffers and TWO accumulator sets, so the TMA load of tile i+1, the MMAs of
// tile i and the shift-and-add epilogue of tile i-1 all overlap and the kernel runs at the rate its input streams in.
// (First version: one 18 x 40 halo buffer of a 16 x 32 tile; load and MMAs of consecutive tiles serialised, 303 us per
// 2048 slot-images.)  The 8 x 32 tile with a 34-pixel pitch re-reads 1.33x its pixels (L2 hits) instead of 1.41x.
struct HeadComposite {
  int S;            // slots per frame (n_img = n_frames * S, slot-images of a frame are consecutive)
  float* imgs;      // [n_frames, 3, H, W]
  float* recons;    // [n_frames, S, 3, H, W] or null
  float* masks;     // [n_frames, S, 1, H, W] or null
};

// S_CT: compile-time slot count (0 = run-time hc.S).  With S known the slot loops of the final compositing pass unroll and
// their shared loads / exponentials issue back to back; as run-time loops they were a ~1000-clock dependent chain per pixel
// on the ONE epilogue warp each scheduler has, every S-th tile -- more than the two-deep accumulator pipeline can hide
// (first fused version: 343 us per 2048 slot-images against 293 + 30 us for the separate kernels).
// Register cap: the decoder's next-chunk layer-1 kernel (256 threads x 128 registers) is meant to run CO-RESIDENT with
// this kernel (decoder.cu chunk pipeline).  An SM sub-partition has 16 K registers and gets up to 3 of this kernel's 10
// warps + 2 of layer 1's 8: 3 x 32 x R + 2 x 4096 <= 16384 needs R <= 85.  At 86 registers (one build of round 2) layer 1
// silently stopped overlapping and ran under the next convolution instead: +0.18 ms per chunk (same-box A/B of two builds).
template <bool COMPOSITE, int S_CT>
__global__ void __maxnreg__(80)
head3x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, int n_img, int H, int W,
               const float* __restrict__ bias, float* __restrict__ out4, HeadComposite hc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sW = smem + 2 * HD_A_STRIDE;
  float* sD = reinterpret_cast<float*>(sW + HD_W_BYTES);
  uint64_t* w_full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sD) + HD_D_BYTES);
  uint64_t* a_full = w_full + 1;     // [2]
  uint64_t* a_empty = a_full + 2;    // [2]
  uint64_t* t_full = a_empty + 2;    // [2]
  uint64_t* t_empty = t_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  float4* sM = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(sD) + HD_D_BYTES + 256);   // COMPOSITE: [S][256 pixels]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_x = W / HD_TW, tiles_y = H / HD_TH;
  const int tiles_per_img = tiles_x * tiles_y;
  const int num_tiles = n_img * tiles_per_img;
  // Work item t (all three roles walk the same sequence).  Plain: tile t of the flattened (image, tile) grid.  COMPOSITE:
  // slot-major inside a (frame, tile) group, so the S tiles a pixel tile needs are consecutive items of ONE CTA:
  // t = group * S + slot with groups handed out round-robin (gridDim.x groups per round).
  const int S = !COMPOSITE ? 1 : (S_CT > 0 ? S_CT : hc.S);
  const int num_groups = num_tiles / S;
  // every thread of every role decodes every item: with run-time divisors these four divisions were 16 M of the kernel's
  // 60 M instructions (ncu r2); tile counts are powers of two for every frame size the decoder is built for
  const bool pow2 = (tiles_per_img & (tiles_per_img - 1)) == 0 && (tiles_x & (tiles_x - 1)) == 0;
  const int sh_img = 31 - __clz(tiles_per_img), sh_x = 31 - __clz(tiles_x);
  auto item = [&](int it, int& img, int& y0, int& x0, int& slot) -> bool {
    const int g = int(blockIdx.x) + (it / S) * int(gridDim.x);
    if (g >= num_groups) return false;
    const int gr = num_groups - 1 - g;      // last groups first: layer 4 wrote them last, ~100 MB of them are still in L2
    slot = it % S;
    int q, r, ry;
    if (pow2) {
      q = gr >> sh_img;
      r = gr & (tiles_per_img - 1);
      ry = r >> sh_x;
      x0 = (r & (tiles_x - 1)) * HD_TW;
    } else {
      q = gr / tiles_per_img;
      r = gr - q * tiles_per_img;
      ry = r / tiles_x;
      x0 = (r - ry * tiles_x) * HD_TW;
    }
    img = COMPOSITE ? q * S + slot : q;
    y0 = ry * HD_TH;
    return true;
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    mbar_init(w_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&a_full[b], 1);
      mbar_init(&a_empty[b], 1);
      mbar_init(&t_full[b], 1);
      mbar_init(&t_empty[b], HD_ETH / 32);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, HD_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_full, HD_W_BYTES);
      tma_load_2d(&tmW, w_full, sW, 0, 0);
      int img, y0, x0, slot;
      for (int it = 0; item(it, img, y0, x0, slot); ++it) {
        const int buf = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(&a_empty[buf], ph ^ 1);                    // the MMAs that read this halo buffer have completed
        mbar_expect_tx(&a_full[buf], HD_A_BYTES);
        tma_load_4d(&tmX, &a_full[buf], sA + buf * HD_A_STRIDE, 0, x0 - 1, y0 - 1, img);   // out-of-image: zero-filled
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_f16(128, HD_N, 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t leader = elect_one_sync();
    mbar_wait(w_full, 0);
    const uint64_t db = hd_desc(smem_u32(sW));
    int img, y0, x0, slot;
    for (int it = 0; item(it, img, y0, x0, slot); ++it) {
      const int buf = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(&t_empty[buf], ph ^ 1);                      // the epilogue has copied this accumulator set out
      mbar_wait(&a_full[buf], ph);
      tc_fence_after();
      const uint32_t a_base = smem_u32(sA + buf * HD_A_STRIDE);
      const uint32_t d_base = tmem_u + uint32_t(buf * HD_MT * HD_N);
#pragma unroll
      for (int mt = 0; mt < HD_MT; ++mt) {
        const uint64_t da = hd_desc(a_base + uint32_t(mt * 128 * 128));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(d_base + uint32_t(mt * HD_N), da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, k != 0, leader);
      }
      umma_commit(&a_empty[buf], leader);
      umma_commit(&t_full[buf], leader);
    }
  } else {
    // 8 epilogue warps: with 4 (one per scheduler) the epilogue was a dependent-issue chain and the kernel's pace (ncu r1 /
    // r2); warps q and q + 4 share a TMEM lane quarter and split its M-tiles, then every thread owns one pixel of the tile
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;                         // 0..255
    const float4 b4 = *reinterpret_cast<const float4*>(bias);
    int img, y0, x0, slot;
    const size_t plane = size_t(H) * W;
    for (int it = 0; item(it, img, y0, x0, slot); ++it) {
      const int buf = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(&t_full[buf], ph);
      tc_fence_after();
      // ---- phase 1: accumulators -> shared memory, D[halo pixel][tap*4 + co]
#pragma unroll 1
      for (int mt = part; mt < HD_MT; mt += 2) {
        uint32_t v[32], v4[4];
        const uint32_t ta = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(buf * HD_MT * HD_N + mt * HD_N);
        tmem_ld32(ta, v);
        tmem_ld4(ta + 32, v4);
        tmem_ld_wait();
        float4* d = reinterpret_cast<float4*>(sD + (mt * 128 + q * 32 + lane) * HD_DS);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          d[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                             __uint_as_float(v[4 * j + 3]));
        d[8] = make_float4(__uint_as_float(v4[0]), __uint_as_float(v4[1]), __uint_as_float(v4[2]), __uint_as_float(v4[3]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[buf]);             // this accumulator set is free for the MMAs of tile it+2
      asm volatile("bar.sync 1, %0;" ::"n"(HD_ETH) : "memory");   // D complete
      // ---- phase 2: shift-and-add over the 9 taps; thread -> pixel (row et >> 5, column et & 31)
      const int px = et & 31, py = et >> 5;
      {
        float a0 = b4.x, a1 = b4.y, a2 = b4.z, a3 = b4.w;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const float4 d = *reinterpret_cast<const float4*>(sD + ((py + tap / 3) * HD_WB + px + tap % 3) * HD_DS + tap * 4);
          a0 += d.x; a1 += d.y; a2 += d.z; a3 += d.w;
        }
        if constexpr (!COMPOSITE) {
          *reinterpret_cast<float4*>(out4 + ((size_t(img) * H + (y0 + py)) * W + (x0 + px)) * 4) = make_float4(a0, a1, a2, a3);
        } else {
          sM[slot * (HD_TH * HD_TW) + py * HD_TW + px] = make_float4(a0, a1, a2, a3);   // thread-private column
          if (hc.recons != nullptr) {
            float* o = hc.recons + size_t(img) * 3 * plane + uint32_t((y0 + py) * W + (x0 + px));
            o[0] = a0; o[plane] = a1; o[2 * plane] = a2;
          }
        }
      }
      if constexpr (COMPOSITE) {
        if (slot == S - 1) {
          // softmax over the slot axis + weighted RGB sum (SAVi.py:252-255); same arithmetic, same order as composite_kernel
          const int frame = img / S;
          {
            const float4* m = sM + py * HD_TW + px;
            float mx = -1e30f;
#pragma unroll
            for (int s2 = 0; s2 < S; ++s2) mx = fmaxf(mx, m[s2 * (HD_TH * HD_TW)].w);
            float den = 0.f, r = 0.f, g = 0.f, b = 0.f;
#pragma unroll
            for (int s2 = 0; s2 < S; ++s2) {
              const float4 v = m[s2 * (HD_TH * HD_TW)];
              const float e = __expf(v.w - mx);
              den += e; r += e * v.x; g += e * v.y; b += e * v.z;
            }
            const float inv = 1.f / den;
            const size_t pix = size_t(y0 + py) * W + (x0 + px);
            float* o = hc.imgs + size_t(frame) * 3 * plane + pix;
            o[0] = r * inv; o[plane] = g * inv; o[2 * plane] = b * inv;
            if (hc.masks != nullptr) {
#pragma unroll
              for (int s2 = 0; s2 < S; ++s2)
                hc.masks[(size_t(frame) * S + s2) * plane + pix] = __expf(m[s2 * (HD_TH * HD_TW)].w - mx) * inv;
            }
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(HD_ETH) : "memory");   // everyone is done reading D
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, HD_TMEM_COLS);
  }
}

// x: f16 NHWC [n_img, H, W, 64]; w_taps: f16 [48, 64] (row tap*4 + co, rows 36..47 zero); bias fp32 [4];
// plain: out4 fp32 NHWC [n_img, H, W, 4];  composite (hc != null): n_img = n_frames * hc->S, outputs as in HeadComposite.
static int launch_head(const __half* x, const __half* w_taps, const float* bias, float* out4, const HeadComposite* hc,
                       int n_img, int H, int W, cudaStream_t stream) {
  TOCVP_CHECK_ARG(x && w_taps && bias && n_img > 0 && H % HD_TH == 0 && W % HD_TW == 0);
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  CUtensorMap tmX, tmW;
  {
    const uint64_t dims[4] = {64, uint64_t(W), uint64_t(H), uint64_t(n_img)};
    const uint64_t str[3] = {128, uint64_t(W) * 128, uint64_t(H) * W * 128};
    const uint32_t box[4] = {64, uint32_t(HD_WB), uint32_t(HD_HR), 1};
    TOCVP_TRY(encode_tmap(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  TOCVP_TRY(encode_tmap_2d_f16(&tmW, w_taps, HD_N, 64, 64, HD_N, 64));
  const int num_tiles = n_img * (H / HD_TH) * (W / HD_TW);
  if (hc != nullptr) {
    TOCVP_CHECK_ARG(hc->S >= 1 && hc->S <= HD_MAX_SLOTS && n_img % hc->S == 0 && hc->imgs != nullptr);
    const int smem = HD_SMEM + hc->S * HD_COMP_BYTES;
    const int groups = num_tiles / hc->S;
    const int grid = groups < num_sms() ? groups : num_sms();
    if (hc->S == 8) {                                    // SAVi.json: 8 slots
      static SmemAttrOnce attr_once;
      TOCVP_TRY(ensure_smem_attr(attr_once, head3x3_kernel<true, 8>, HD_SMEM + 8 * HD_COMP_BYTES));
      head3x3_kernel<true, 8><<<grid, HD_THREADS, smem, stream>>>(tmX, tmW, n_img, H, W, bias, nullptr, *hc);
    } else {
      static SmemAttrOnce attr_once;
      TOCVP_TRY(ensure_smem_attr(attr_once, head3x3_kernel<true, 0>, HD_SMEM + HD_MAX_SLOTS * HD_COMP_BYTES));
      head3x3_kernel<true, 0><<<grid, HD_THREADS, smem, stream>>>(tmX, tmW, n_img, H, W, bias, nullptr, *hc);
    }
  } else {
    TOCVP_CHECK_ARG(out4 != nullptr && (reinterpret_cast<uintptr_t>(out4) & 15) == 0);
    static SmemAttrOnce attr_once;
    TOCVP_TRY(ensure_smem_attr(attr_once, head3x3_kernel<false, 1>, HD_SMEM));
    const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
    head3x3_kernel<false, 1><<<grid, HD_THREADS, HD_SMEM, stream>>>(tmX, tmW, n_img, H, W, bias, out4, HeadComposite{1, nullptr, nullptr, nullptr});
  }
  TOCVP_LAUNCHED();
  return TOCVP_OK;
}

int conv3x3_head_taps_f16(const __half* x, const __half* w_taps, const float* bias, float* out4, int n_img, int H, int W,
                          cudaStream_t stream) {
  return launch_head(x, w_taps, bias, out4, nullptr, n_img, H, W, stream);
}

// Head conv + slot-softmax compositing in one kernel: x holds n_frames * S slot-images (a frame's S slot-images consecutive).
int conv3x3_head_composite_f16(const __half* x, const __half* w_taps, const float* bias, int n_frames, int S, int H, int W,
                               float* imgs, float* recons, float* masks, cudaStream_t stream) {
  const HeadComposite hc{S, imgs, recons, masks};
  return launch_head(x, w_taps, bias, nullptr, &hc, n_frames * S, H, W, stream);
}

}  // namespace tocvp
