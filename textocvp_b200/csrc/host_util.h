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

int num_sms();

// Programmatic dependent launch (tocvp_set_pdl, default on): the kernel may become resident while the previous kernel of
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
bool tile_order_alternation();      // tocvp_set_tile_order knob (default on)

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
