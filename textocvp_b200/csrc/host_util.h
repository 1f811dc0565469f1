// Host-side helpers shared by the C-ABI entry points: error codes, TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tocvp.h"

namespace tocvp {

#define TOCVP_CHECK_ARG(cond)                                                      \
  do {                                                                             \
    if (!(cond)) {                                                                 \
      tocvp::set_last_error(__FILE__, __LINE__, "bad argument: " #cond);           \
      return TOCVP_ERR_BAD_ARG;                                                    \
    }                                                                              \
  } while (0)

#define TOCVP_CUDA(call)                                                           \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      tocvp::set_last_error(__FILE__, __LINE__, cudaGetErrorString(e__));          \
      return TOCVP_ERR_CUDA;                                                       \
    }                                                                              \
  } while (0)

#define TOCVP_TRY(call)                                                            \
  do {                                                                             \
    int r__ = (call);                                                              \
    if (r__ != TOCVP_OK) return r__;                                               \
  } while (0)

void set_last_error(const char* file, int line, const char* msg);
void count_launch();
unsigned long long launch_count();

// after every kernel launch: surface launch errors and count the launch (tocvp_kernel_launches)
#define TOCVP_LAUNCHED()                 \
  do {                                   \
    TOCVP_CUDA(cudaGetLastError());      \
    tocvp::count_launch();               \
  } while (0)


// cuTensorMapEncodeTiled fetched through the runtime (no link-time dependency on libcuda).
int encode_tmap(CUtensorMap* map, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes /* rank-1 entries, dims 1.. */, const uint32_t* box,
                CUtensorMapSwizzle swizzle);

// 2D row-major fp16 matrix [rows, cols] (cols contiguous) with a {box_cols, box_rows} box, 128B swizzle.
int encode_tmap_2d_f16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                       uint32_t box_rows, uint32_t box_cols);

int num_sms();   // SM count of the CURRENT device (cached per device)

// Per-call options.  The library keeps no process-wide tuning state: an entry point that accepts a `const tocvp_tuning*`
// (directly or through its weights struct, include/tocvp.h) installs it for the duration of the call on the calling
// thread; the kernel-selection code deeper in the library reads it through opts().  Without a scope opts() returns the
// defaults (all zeros).
const tocvp_tuning& opts();
struct OptsScope {
  explicit OptsScope(const tocvp_tuning* t);
  ~OptsScope();
  OptsScope(const OptsScope&) = delete;
  OptsScope& operator=(const OptsScope&) = delete;
 private:
  const tocvp_tuning* saved_;
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: one flag per (kernel instantiation, device).  Two
// threads racing on the same flag both set the attribute (idempotent).
struct SmemAttrOnce {
  unsigned long long done_mask = 0;   // bit d: set for device d
};
template <typename Fn>
static inline int ensure_smem_attr(SmemAttrOnce& once, Fn* kernel, int bytes) {
  int dev = 0;
  TOCVP_CUDA(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if (__atomic_load_n(&once.done_mask, __ATOMIC_ACQUIRE) & bit) return TOCVP_OK;
  if (bytes < 0) {   // "as much as the device allows": the opt-in maximum minus the kernel's static shared memory
    cudaFuncAttributes fa;
    TOCVP_CUDA(cudaFuncGetAttributes(&fa, kernel));
    int optin = 0;
    TOCVP_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    bytes = optin - int(fa.sharedSizeBytes);
  }
  TOCVP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  __atomic_fetch_or(&once.done_mask, bit, __ATOMIC_RELEASE);
  return TOCVP_OK;
}

// Programmatic dependent launch (tocvp_tuning.no_pdl, default on): the kernel may become resident while the previous kernel of
// the stream drains; it runs its prologue (barrier init, TMEM allocation, descriptor prefetch, constant-weight loads) and
// blocks in griddepcontrol.wait (pdl_wait(), ptx.cuh) until the previous grid has completed and flushed.  ONLY for kernels
// that execute pdl_wait() before their first access to memory another kernel produces or consumes.
bool pdl_enabled();

// Tile traversal order of the NEXT pair-GEMM / attention launch of this thread (consumed by the launch, then reset to
// ascending).  A driver that chains producer -> consumer kernels over tensors larger than the L2 (predictor.cu) alternates
// the direction, so that a consumer starts with the rows its producer wrote LAST -- still L2 resident -- instead of
// streaming in the same order and missing on everything (LRU).  Results do not depend on the order.
void set_next_tile_order(int reversed);
int tile_order_reversed();          // reads and resets
bool tile_order_alternation();      // !tocvp_tuning.no_tile_alternation (default on)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace tocvp
