// Error reporting, device queries and the TMA descriptor factory shared by all kernels.
#include <stdarg.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "aq_common.h"

namespace aq {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    // resolved through the runtime so the library carries no link-time dependency on libcuda.so
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, TmaSwizzle swz, int l2_promotion) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return fail(AQ_ERR_LAUNCH, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0)
    return fail(AQ_ERR_BAD_ALIGN, "tensor base %p is not 16-byte aligned", base);
  for (int i = 0; i + 1 < rank; ++i)
    if (strides_bytes[i] % 16 != 0)
      return fail(AQ_ERR_BAD_ALIGN, "tensor stride %llu bytes is not a multiple of 16", (unsigned long long)strides_bytes[i]);
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUtensorMapSwizzle sw = swz == kSwz128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swz == kSwz64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swz == kSwz32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                          : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  l2_promotion >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                  : l2_promotion >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                        : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(AQ_ERR_LAUNCH, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu x %llu box %u x %u)", (int)r, rank,
                (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 1), box[0], rank > 1 ? box[1] : 1u);
  return AQ_OK;
}

bool pdl_enabled() {
  static int cached = -1;
  if (cached < 0) {
    // opt-in: measured on B200 (profiles/r02_gemm_sweep_pdl{0,1}.log, r02_bench_pdl_ab.txt) it trims 1 - 12 % off isolated launches but
    // leaves the PPFT step unchanged (the next kernel's CTAs cannot co-reside with ours: 227 KiB of shared memory each), and it makes
    // CUPTI / ncu kernel durations include the time a dependent spends waiting in griddepcontrol.wait
    const char* e = getenv("AQ_PDL");
    cached = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return cached == 1;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached[dev] = n;
  }
  return cached[dev];
}

int check_arch() {
  static int cached[64] = {0};  // 0 unknown, 1 ok, -1 bad
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
    return fail(AQ_ERR_ARCH, "no CUDA device is current (this library has no CPU path)");
  if (cached[dev] == 0) {
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    cached[dev] = (major == 10 && minor == 0) ? 1 : -1;
  }
  if (cached[dev] < 0) return fail(AQ_ERR_ARCH, "device %d is not sm_100 (B200); libaqualora_b200 is built for sm_100a only", dev);
  return AQ_OK;
}

}  // namespace aq

extern "C" {
int aq_version(void) { return 1; }
int aq_arch(void) { return 100; }
const char* aq_last_error(void) { return aq::g_err; }
int aq_sm_count(void) { return aq::sm_count(); }
long long aq_launch_count(void) { return aq::launch_count(); }
}
