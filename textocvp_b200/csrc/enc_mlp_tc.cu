// SAVi encoder MLP (reference src/models/SAVi.py:115-120, applied at :237): y = W2 . relu(W1 . x + b1) + b2 over the
// 21 M LayerNorm-ed pixel rows of a 5120-frame batch, 32 -> 128 -> 128, as ONE kernel with two chained tcgen05 GEMMs.
//
// As two separate GEMMs this stage is pure HBM traffic: the 128-wide hidden activation (5.4 GB f16 per 5120 frames) is
// written and read back (4.7 ms).  Here the hidden tile never leaves the SM:
//   TMA  : x tile [128 rows x 32] (64-byte rows, SWIZZLE_64B) through a 4-stage ring; W1 / W2 resident in shared memory
//   MMA 1: D1[128x128] = x . W1^T (2 k-steps)                                   -> TMEM (double buffered)
//   epi 1: warps 2-5: D1 + b1 -> ReLU -> f16 -> shared memory, written directly in the K-major SWIZZLE_128B layout of an
//          A operand (two 64-wide k-blocks)                                      (double buffered)
//   MMA 2: D2[128x128] = h . W2^T (8 k-steps), issued one tile behind MMA 1     -> TMEM (double buffered)
//   epi 2: warps 6-9: D2 + b2 -> f16 -> swizzled staging tiles -> TMA stores
// HBM traffic: 64 B in + 256 B out per row = the algorithmic minimum.
#include "host_util.h"
#include "ptx.cuh"

namespace tocvp {

constexpr int EM_K1 = 32, EM_N = 128;
constexpr int EM_STAGES = 4;
constexpr int EM_A1_BYTES = 128 * EM_K1 * 2;        // 8 KB
constexpr int EM_W1_BYTES = EM_N * EM_K1 * 2;       // 8 KB
constexpr int EM_W2_BYTES = EM_N * EM_N * 2;        // 32 KB (two k-blocks of [128 x 64])
constexpr int EM_A2_BYTES = 128 * EM_N * 2;         // 32 KB (two k-blocks of [128 x 64])
constexpr int EM_STG_BYTES = 4 * 2 * 4096;          // epilogue-2 warps: two [32 rows x 64 cols] f16 tiles each
constexpr int EM_OFF_W1 = EM_STAGES * EM_A1_BYTES;
constexpr int EM_OFF_W2 = EM_OFF_W1 + EM_W1_BYTES;
constexpr int EM_OFF_A2 = EM_OFF_W2 + EM_W2_BYTES;
constexpr int EM_OFF_STG = EM_OFF_A2 + 2 * EM_A2_BYTES;
constexpr int EM_OFF_BAR = EM_OFF_STG + EM_STG_BYTES;
constexpr int EM_SMEM = EM_OFF_BAR + 512 + 1024;

__device__ __forceinline__ uint64_t em_desc(uint32_t saddr, uint32_t sbo, int layout) {
  return uint64_t((saddr >> 4) & 0x3FFF) | (uint64_t(1) << 16) | (uint64_t((sbo >> 4) & 0x3FFF) << 32) | (uint64_t(1) << 46) |
         (uint64_t(layout) << 61);
}
__device__ __forceinline__ void em_tma_store(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}

__global__ void __launch_bounds__(320, 1)
enc_mlp_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmY, int M,
               const float* __restrict__ b1, const float* __restrict__ b2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + EM_OFF_BAR);
  uint64_t* a1_full = bars;                 // [4]
  uint64_t* a1_empty = a1_full + EM_STAGES; // [4]
  uint64_t* d1_full = a1_empty + EM_STAGES; // [2]
  uint64_t* d1_empty = d1_full + 2;         // [2]  4 arrivals (epilogue-1 warps)
  uint64_t* a2_full = d1_empty + 2;         // [2]  4 arrivals
  uint64_t* a2_empty = a2_full + 2;         // [2]
  uint64_t* d2_full = a2_empty + 2;         // [2]
  uint64_t* d2_empty = d2_full + 2;         // [2]  4 arrivals (epilogue-2 warps)
  uint64_t* w_full = d2_empty + 2;          // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (M + 127) / 128;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
    for (int s = 0; s < EM_STAGES; ++s) {
      mbar_init(&a1_full[s], 1);
      mbar_init(&a1_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&d1_full[b], 1);
      mbar_init(&d1_empty[b], 4);
      mbar_init(&a2_full[b], 4);
      mbar_init(&a2_empty[b], 1);
      mbar_init(&d2_full[b], 1);
      mbar_init(&d2_empty[b], 4);
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_full, EM_W1_BYTES + EM_W2_BYTES);
      tma_load_2d(&tmW1, w_full, smem + EM_OFF_W1, 0, 0);
      tma_load_2d(&tmW2, w_full, smem + EM_OFF_W2, 0, 0);
      tma_load_2d(&tmW2, w_full, smem + EM_OFF_W2 + 16384, 64, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&a1_empty[s], ph ^ 1);
        mbar_expect_tx(&a1_full[s], EM_A1_BYTES);
        tma_load_2d(&tmX, &a1_full[s], smem + s * EM_A1_BYTES, 0, t * 128);
        if (++s == EM_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_f16(128, EM_N, 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t leader = elect_one_sync();
    mbar_wait(w_full, 0);
    const uint64_t dw1 = em_desc(smem_u32(smem + EM_OFF_W1), 8 * 64, 4);          // SWIZZLE_64B, 64-byte rows
    auto mma2 = [&](int j) {                                                       // second GEMM of the j-th local tile
      const int b = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&a2_full[b], ph);
      mbar_wait(&d2_empty[b], ph ^ 1);
      tc_fence_after();
      const uint32_t a2 = smem_u32(smem + EM_OFF_A2 + b * EM_A2_BYTES);
      const uint32_t w2 = smem_u32(smem + EM_OFF_W2);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const uint64_t da = em_desc(a2 + kb * 16384, 1024, 2), db = em_desc(w2 + kb * 16384, 1024, 2);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_u + 256 + uint32_t(b * EM_N), da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0, leader);
      }
      umma_commit(&a2_empty[b], leader);
      umma_commit(&d2_full[b], leader);
    };
    int s = 0, j = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++j) {
      const int b = j & 1;
      const uint32_t bph = (j >> 1) & 1;
      mbar_wait(&a1_full[s], ph);
      mbar_wait(&d1_empty[b], bph ^ 1);
      tc_fence_after();
      const uint64_t da = em_desc(smem_u32(smem + s * EM_A1_BYTES), 8 * 64, 4);
#pragma unroll
      for (int k = 0; k < EM_K1 / 16; ++k)
        umma_f16(tmem_u + uint32_t(b * EM_N), da + uint64_t(2 * k), dw1 + uint64_t(2 * k), idesc, k != 0, leader);
      umma_commit(&a1_empty[s], leader);
      umma_commit(&d1_full[b], leader);
      if (++s == EM_STAGES) { s = 0; ph ^= 1; }
      if (j > 0) mma2(j - 1);                               // one tile behind: its hidden tile is ready by now
    }
    if (j > 0) mma2(j - 1);
  } else if (warp < 6) {
    // ---------------------------------------------------------------- epilogue 1: D1 -> relu -> f16 A operand in smem
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t sw = uint32_t(row & 7);
    int j = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++j) {
      const int b = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&d1_full[b], ph);
      tc_fence_after();
      mbar_wait(&a2_empty[b], ph ^ 1);                       // MMA 2 of the tile that used this buffer has completed
      uint8_t* dst = smem + EM_OFF_A2 + b * EM_A2_BYTES + row * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(b * EM_N + c * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          const int n = c * 32 + j8 * 8;
          const float4 ba = __ldg(reinterpret_cast<const float4*>(b1 + n));
          const float4 bb = __ldg(reinterpret_cast<const float4*>(b1 + n + 4));
          uint4 p;
          p.x = pack_half2_relu(__uint_as_float(v[j8 * 8 + 0]) + ba.x, __uint_as_float(v[j8 * 8 + 1]) + ba.y);
          p.y = pack_half2_relu(__uint_as_float(v[j8 * 8 + 2]) + ba.z, __uint_as_float(v[j8 * 8 + 3]) + ba.w);
          p.z = pack_half2_relu(__uint_as_float(v[j8 * 8 + 4]) + bb.x, __uint_as_float(v[j8 * 8 + 5]) + bb.y);
          p.w = pack_half2_relu(__uint_as_float(v[j8 * 8 + 6]) + bb.z, __uint_as_float(v[j8 * 8 + 7]) + bb.w);
          const int chunk = n >> 3;                           // 16-byte chunk of the 128-wide hidden row
          *reinterpret_cast<uint4*>(dst + (chunk >> 3) * 16384 + ((uint32_t(chunk & 7) ^ sw) << 4)) = p;
        }
      }
      tc_fence_before();
      fence_proxy_async();                                    // generic smem writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&d1_empty[b]);
        mbar_arrive(&a2_full[b]);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue 2: D2 + b2 -> f16 -> TMA store
    const int q = warp & 3;
    const uint32_t sw = uint32_t(lane & 7);
    uint8_t* stg = smem + EM_OFF_STG + (warp - 6) * 8192;
    int j = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++j) {
      const int b = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&d2_full[b], ph);
      tc_fence_after();
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // last tile's stores have drained
      __syncwarp();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(256 + b * EM_N + c * 32), v);
        tmem_ld_wait();
        uint8_t* dst = stg + (c >> 1) * 4096 + lane * 128;
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          const int n = c * 32 + j8 * 8;
          const float4 ba = __ldg(reinterpret_cast<const float4*>(b2 + n));
          const float4 bb = __ldg(reinterpret_cast<const float4*>(b2 + n + 4));
          uint4 p;
          p.x = pack_half2(__uint_as_float(v[j8 * 8 + 0]) + ba.x, __uint_as_float(v[j8 * 8 + 1]) + ba.y);
          p.y = pack_half2(__uint_as_float(v[j8 * 8 + 2]) + ba.z, __uint_as_float(v[j8 * 8 + 3]) + ba.w);
          p.z = pack_half2(__uint_as_float(v[j8 * 8 + 4]) + bb.x, __uint_as_float(v[j8 * 8 + 5]) + bb.y);
          p.w = pack_half2(__uint_as_float(v[j8 * 8 + 6]) + bb.z, __uint_as_float(v[j8 * 8 + 7]) + bb.w);
          *reinterpret_cast<uint4*>(dst + ((uint32_t((c & 1) * 4 + j8) ^ sw) << 4)) = p;
        }
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&d2_empty[b]);
        em_tma_store(&tmY, stg, 0, t * 128 + q * 32);
        em_tma_store(&tmY, stg + 4096, 64, t * 128 + q * 32);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// x f16 [M, 32] (LayerNorm output), w1 f16 [128, 32], w2 f16 [128, 128], biases fp32 -> y f16 [M, 128]
int enc_mlp_f16(const __half* x, const __half* w1, const float* b1, const __half* w2, const float* b2, __half* y, int M,
                cudaStream_t stream) {
  TOCVP_CHECK_ARG(x && w1 && b1 && w2 && b2 && y && M > 0);
  TOCVP_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0);
  static SmemAttrOnce attr_once;
  TOCVP_TRY(ensure_smem_attr(attr_once, enc_mlp_kernel, EM_SMEM));
  CUtensorMap tmX, tmW1, tmW2, tmY;
  {
    const uint64_t dims[2] = {uint64_t(EM_K1), uint64_t(M)};
    const uint64_t str[1] = {uint64_t(EM_K1) * 2};
    const uint32_t box[2] = {uint32_t(EM_K1), 128};
    TOCVP_TRY(encode_tmap(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, x, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B));
  }
  {
    const uint64_t dims[2] = {uint64_t(EM_K1), uint64_t(EM_N)};
    const uint64_t str[1] = {uint64_t(EM_K1) * 2};
    const uint32_t box[2] = {uint32_t(EM_K1), uint32_t(EM_N)};
    TOCVP_TRY(encode_tmap(&tmW1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, w1, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B));
  }
  TOCVP_TRY(encode_tmap_2d_f16(&tmW2, w2, EM_N, EM_N, EM_N, EM_N, 64));
  TOCVP_TRY(encode_tmap_2d_f16(&tmY, y, M, EM_N, EM_N, 32, 64));
  const int tiles = (M + 127) / 128;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  enc_mlp_kernel<<<grid, 320, EM_SMEM, stream>>>(tmX, tmW1, tmW2, tmY, M, b1, b2);
  TOCVP_LAUNCHED();
  return TOCVP_OK;
}

}  // namespace tocvp
