#include "host_util.h"

#include <string.h>

#include <atomic>
#include <mutex>

namespace tocvp {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(); }

void set_last_error(const char* file, int line, const char* msg) {
  const char* base = strrchr(file, '/');
  snprintf(g_err, sizeof(g_err), "%s:%d: %s", base ? base + 1 : file, line, msg);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

int encode_tmap(CUtensorMap* map, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return TOCVP_ERR_CUDA;
  }
  cuuint64_t gdims[5];
  cuuint64_t gstr[5];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(map, dtype, cuuint32_t(rank), const_cast<void*>(base), gdims, gstr, gbox, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (CUresult %d, rank %d, dim0 %llu box0 %u)", int(r), rank,
             (unsigned long long)dims[0], box[0]);
    set_last_error(__FILE__, __LINE__, msg);
    return TOCVP_ERR_CUDA;
  }
  return TOCVP_OK;
}

int encode_tmap_2d_f16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                       uint32_t box_rows, uint32_t box_cols) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[1] = {ld_elems * 2};
  const uint32_t box[2] = {box_cols, box_rows};
  return encode_tmap(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// Options of the library call running on this thread (installed by OptsScope at the C-ABI boundary; caller-owned memory).
static const tocvp_tuning g_default_tuning = {};
static thread_local const tocvp_tuning* g_opts = nullptr;
const tocvp_tuning& opts() { return g_opts ? *g_opts : g_default_tuning; }
OptsScope::OptsScope(const tocvp_tuning* t) : saved_(g_opts) {
  if (t != nullptr) g_opts = t;   // a nested call without options of its own inherits the caller's
}
OptsScope::~OptsScope() { g_opts = saved_; }

bool pdl_enabled() { return opts().no_pdl == 0; }

static thread_local int g_next_rev = 0;   // consumed by the next pair-GEMM / attention launch of this thread
bool tile_order_alternation() { return opts().no_tile_alternation == 0; }
void set_next_tile_order(int reversed) { g_next_rev = (reversed && tile_order_alternation()) ? 1 : 0; }
int tile_order_reversed() {
  const int r = g_next_rev;
  g_next_rev = 0;
  return r;
}

int num_sms() {
  static std::atomic<int> cache[64];   // immutable device attribute, cached per device
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

}  // namespace tocvp

extern "C" const char* tocvp_last_error(void) { return tocvp::g_err; }

extern "C" int tocvp_init(int device) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    tocvp::set_last_error(__FILE__, __LINE__, "cudaGetDeviceProperties failed (no CUDA device)");
    return TOCVP_ERR_CUDA;
  }
  if (prop.major != 10) {
    char msg[128];
    snprintf(msg, sizeof(msg), "device is sm_%d%d; this library is built for sm_100a only (no fallback)", prop.major,
             prop.minor);
    tocvp::set_last_error(__FILE__, __LINE__, msg);
    return TOCVP_ERR_ARCH;
  }
  return TOCVP_OK;   // the current device is NOT changed: every call runs on the caller's current device
}

extern "C" int tocvp_abi_version(void) { return TOCVP_ABI_VERSION; }

// Statistics only (monotonic counter of kernels this library has launched in this process).
extern "C" unsigned long long tocvp_kernel_launches(void) { return tocvp::launch_count(); }

// A caller that captured library calls into a CUDA graph reports each replay here (n = kernels in the graph), so that the
// statistic keeps counting kernels that actually ran on the device.
extern "C" void tocvp_note_graph_replay(unsigned long long n_kernels) {
  for (unsigned long long i = 0; i < n_kernels; ++i) tocvp::count_launch();
}
