// Hardware probe (test-only library tests/native/libtocvp_probe.so; not part of libtocvp.so or include/tocvp.h): does a K-major SWIZZLE_128B UMMA
// operand descriptor accept a start address that is offset by an arbitrary number of 128-byte rows
// inside a 1024B-aligned TMA-written buffer?  The implicit-GEMM convolution relies on exactly that
// (every filter tap is the same smem halo tile read at a shifted start row), so the property is
// checked on the real part by tests/test_kernels_gpu.py before anything is built on it.
//   D[128,64] = X[shift : shift+128, 0:64] . W[64,64]^T
#include "host_util.h"
#include "ptx.cuh"

namespace tocvp {

__global__ void __launch_bounds__(128, 1)
probe_shift_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, float* out,
                   int shift, int base_offset_mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sX = smem;               // 256 rows x 128 B
  uint8_t* sW = smem + 256 * 128;   // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sW + 64 * 128);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 256 * 128 + 64 * 128);
    tma_load_2d(&tmX, bar, sX, 0, 0);
    tma_load_2d(&tmW, bar, sW, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a = smem_u32(sX) + uint32_t(shift) * 128u;
    uint64_t da = make_desc_sw128(a, 1024);
    if (base_offset_mode == 1) da |= uint64_t((a >> 7) & 7) << 49;
    const uint64_t db = make_desc_sw128(smem_u32(sW), 1024);
    constexpr uint32_t idesc = make_idesc_f16(128, 64, 0);
    for (int k = 0; k < 4; ++k) umma_f16(tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, k != 0);
    umma_commit(done);
  }
  mbar_wait(done, 0);
  __syncwarp();
  tc_fence_after();
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + uint32_t(c * 32), v);
    tmem_ld_wait();
    float* o = out + size_t(warp * 32 + lane) * 64 + c * 32;
    for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

}  // namespace tocvp

// X: f16 [256, 64] row-major, W: f16 [64, 64] row-major, out: fp32 [128, 64].
extern "C" int tocvp_probe_shifted_operand(const void* X, const void* W, float* out, int shift,
                                           int base_offset_mode, void* stream) {
  using namespace tocvp;
  TOCVP_CHECK_ARG(X && W && out && shift >= 0 && shift <= 128);
  CUtensorMap tmX, tmW;
  TOCVP_TRY(encode_tmap_2d_f16(&tmX, X, 256, 64, 64, 256, 64));
  TOCVP_TRY(encode_tmap_2d_f16(&tmW, W, 64, 64, 64, 64, 64));
  const int smem = 256 * 128 + 64 * 128 + 64 + 1024;
  TOCVP_CUDA(cudaFuncSetAttribute(probe_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_shift_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(tmX, tmW, out, shift, base_offset_mode);
  TOCVP_LAUNCHED();
  return TOCVP_OK;
}
