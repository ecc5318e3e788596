// Shared host-side plumbing for the C-ABI library: error reporting, launch accounting and the
// driver entry point for TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace ldm {

// thread-local last error (returned by ldmseg_last_error_string)
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launch_count;

inline int check_launch(const char* what) {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e) ? static_cast<int>(e) : -1;
  }
  return 0;
}

#define LDM_REQUIRE(cond, ...)    \
  do {                            \
    if (!(cond)) {                \
      ::ldm::set_error(__VA_ARGS__); \
      return -2;                  \
    }                             \
  } while (0)

#define LDM_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) {                                                 \
      ::ldm::set_error("%s failed: %s", #call, cudaGetErrorString(e__));      \
      return static_cast<int>(e__);                                           \
    }                                                                         \
  } while (0)

// Programmatic dependent launch (PDL): when enabled (ldmseg_set_pdl), every kernel is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization so that its prologue overlaps the tail of the
// previous kernel on the stream; all kernels execute `griddepcontrol.wait` before touching memory.
extern int g_pdl;

template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                          cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Encode a tiled, 128B-swizzled bf16 tensor map of the given rank (dims innermost first).
// strides_bytes has rank-1 entries (stride of dims 1..rank-1).  Returns 0 on success.
// elem_strides (optional, rank entries): traversal stride per dimension -- with stride s a box of s * n
// traversed elements lands as n elements in shared memory (strided convolutions without im2col).
int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides = nullptr);

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace ldm
