// Host-side shared declarations for libaqualora_b200.so (error handling, TMA descriptor factory).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/aqualora_b200.h"

namespace aq {

// thread-local error string returned by aq_last_error()
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define AQ_CHECK_CUDA(expr)                                                                     \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return ::aq::fail(AQ_ERR_LAUNCH, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                    \
  } while (0)

// after every kernel launch: bump the process-wide launch counter (aq_launch_count) and surface launch errors
#define AQ_LAUNCHED()                      \
  do {                                     \
    ::aq::count_launch();                  \
    AQ_CHECK_CUDA(cudaGetLastError());     \
  } while (0)

#define AQ_REQUIRE(cond, code, ...)                  \
  do {                                               \
    if (!(cond)) return ::aq::fail(code, __VA_ARGS__); \
  } while (0)

// Opt a kernel into more than 48 KiB of dynamic shared memory ONCE PER DEVICE (the attribute is per device: a process that
// touches a second GPU must opt in again there).  `static` inside the macro gives every call site (= every kernel instantiation)
// its own flag array; the race between threads is benign (the call is idempotent).
#define AQ_OPT_IN_SMEM(kernel, bytes)                                                                        \
  do {                                                                                                       \
    static bool _aq_done[64] = {};                                                                           \
    int _aq_dev = 0;                                                                                         \
    AQ_CHECK_CUDA(cudaGetDevice(&_aq_dev));                                                                  \
    if (_aq_dev < 0 || _aq_dev >= 64 || !_aq_done[_aq_dev]) {                                                \
      AQ_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      if (_aq_dev >= 0 && _aq_dev < 64) _aq_done[_aq_dev] = true;                                            \
    }                                                                                                        \
  } while (0)

enum TmaSwizzle { kSwzNone = 0, kSwz32 = 1, kSwz64 = 2, kSwz128 = 3 };

// Encode a tiled bf16 / fp32 tensor map (rank 2 or 3).  dims/box are innermost-first; strides are in
// BYTES for dims 1..rank-1.  Returns AQ_OK or an error code (aq_last_error() explains).
// l2_promotion: bytes an L2 miss of the map's loads is widened to (0 = none, 128, 256)
int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, TmaSwizzle swz, int l2_promotion = 256);

// Launch configuration with the programmatic-dependent-launch attribute (opt-in: AQ_PDL=1 in the environment): the kernel's
// prologue (barrier init, TMEM allocation, descriptor prefetch) and its launch latency overlap the tail of the previous kernel in
// the stream; every kernel launched through this calls griddep_wait() before touching global memory.
bool pdl_enabled();
struct PdlLaunch {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  PdlLaunch(dim3 grid, dim3 block, size_t smem, cudaStream_t stream) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
  }
};

void count_launch();
int sm_count();  // SMs of the current device (cached)
int check_arch();  // AQ_OK iff the current device is sm_100

}  // namespace aq
