// Small HBM-bound kernels around the LoRA contraction: MapperNet, parameter casts/transposes, and the fused
// global-norm clip + AdamW update over the flat fp32 parameter / gradient buffers.
#include "aq_common.h"
#include "aq_ptx.cuh"

namespace aq {

// ---------------------------------------------------------------- MapperNet (utils/models.py:98-115)
__global__ void mapper_fwd_kernel(const float* __restrict__ msg, const float* __restrict__ emb, float* __restrict__ scale,
                                  int B, int bits, int r, int round_bf16) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * r) return;
  const int b = idx / r, j = idx % r;
  float acc = 0.f;
  for (int i = 0; i < bits; ++i) acc += emb[i * r + j] * msg[b * bits + i];   // same summation order as .sum(dim=1)
  float out = acc / sqrtf((float)bits) + 1.f;
  if (round_bf16) out = bf16_round(out);
  scale[idx] = out;
}

__global__ void mapper_bwd_kernel(const float* __restrict__ msg, const float* __restrict__ g_scale, float* __restrict__ g_emb,
                                  int B, int bits, int r) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= bits * r) return;
  const int i = idx / r, j = idx % r;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) acc += msg[b * bits + i] * g_scale[b * r + j];
  g_emb[idx] += acc / sqrtf((float)bits);
}

// ---------------------------------------------------------------- casts / transposes
template <typename TIn>
__device__ __forceinline__ float to_f32(TIn v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// 32x32 tile through shared memory: coalesced reads of src rows, coalesced writes of dst_t rows.
template <typename TIn>
__global__ void cast_transpose_kernel(const TIn* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                      __nv_bfloat16* __restrict__ dst_t, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int rr = r0 + dy, cc = c0 + threadIdx.x;
    float v = 0.f;
    if (rr < rows && cc < cols) {
      v = to_f32<TIn>(src[(size_t)rr * cols + cc]);
      if (dst != nullptr) dst[(size_t)rr * cols + cc] = __float2bfloat16_rn(v);
    }
    tile[dy][threadIdx.x] = v;
  }
  __syncthreads();
  if (dst_t != nullptr) {
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
      const int cc = c0 + dy, rr = r0 + threadIdx.x;
      if (rr < rows && cc < cols) dst_t[(size_t)cc * rows + rr] = __float2bfloat16_rn(tile[threadIdx.x][dy]);
    }
  }
}

// All 384 LoRA matrices in one launch (one per matrix cost 384 x 3 us per step): a device table of jobs, one 32x32 tile per
// block; the block finds its job by bisection over the tile prefix.
struct CastJob {   // 7 x int64, built by the host as a plain int64 tensor (include/aqualora_b200.h: aq_cast_job)
  long long src, dst, dst_t, rows, cols, tile_begin, tiles_x;
};

__global__ void cast_transpose_batched_kernel(const CastJob* __restrict__ jobs, int njobs) {
  __shared__ float tile[32][33];
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if ((long long)blockIdx.x >= jobs[mid].tile_begin) lo = mid; else hi = mid - 1;
  }
  const CastJob j = jobs[lo];
  const int t = (int)(blockIdx.x - j.tile_begin);
  const int c0 = (t % (int)j.tiles_x) * 32, r0 = (t / (int)j.tiles_x) * 32;
  const int rows = (int)j.rows, cols = (int)j.cols;
  const float* src = reinterpret_cast<const float*>(j.src);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(j.dst);
  __nv_bfloat16* dst_t = reinterpret_cast<__nv_bfloat16*>(j.dst_t);
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int rr = r0 + dy, cc = c0 + threadIdx.x;
    float v = 0.f;
    if (rr < rows && cc < cols) {
      v = src[(size_t)rr * cols + cc];
      if (dst != nullptr) dst[(size_t)rr * cols + cc] = __float2bfloat16_rn(v);
    }
    tile[dy][threadIdx.x] = v;
  }
  __syncthreads();
  if (dst_t != nullptr) {
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
      const int cc = c0 + dy, rr = r0 + threadIdx.x;
      if (rr < rows && cc < cols) dst_t[(size_t)cc * rows + rr] = __float2bfloat16_rn(tile[threadIdx.x][dy]);
    }
  }
}

// ---------------------------------------------------------------- flat clip + AdamW
__global__ void flat_sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  float acc = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(g4 + i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[(n4 << 2) + threadIdx.x];
    acc += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float warp_sums[32];
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? warp_sums[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

struct AdamArgs {
  float grad_scale, max_norm, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt;
};

__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, float gmul, const AdamArgs& a) {
  const float gg = g * gmul;
  p *= (1.f - a.lr * a.weight_decay);                      // decoupled weight decay (torch.optim.AdamW)
  m = a.beta1 * m + (1.f - a.beta1) * gg;
  v = a.beta2 * v + (1.f - a.beta2) * gg * gg;
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p -= (a.lr / a.bc1) * (m / denom);
  g = 0.f;                                                 // optimizer.zero_grad()
}

__global__ void flat_clip_adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                       float* __restrict__ v, long long n, const float* __restrict__ norm_sq, AdamArgs a) {
  // clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
  const float total_norm = sqrtf(*norm_sq) * fabsf(a.grad_scale);
  float coef = a.max_norm > 0.f ? a.max_norm / (total_norm + 1e-6f) : 1.f;
  coef = fminf(coef, 1.f);
  const float gmul = coef * a.grad_scale;
  const long long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* g4 = reinterpret_cast<float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    adam_one(pp.x, gg.x, mm.x, vv.x, gmul, a);
    adam_one(pp.y, gg.y, mm.y, vv.y, gmul, a);
    adam_one(pp.z, gg.z, mm.z, vv.z, gmul, a);
    adam_one(pp.w, gg.w, mm.w, vv.w, gmul, a);
    p4[i] = pp; g4[i] = gg; m4[i] = mm; v4[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    adam_one(p[i], g[i], m[i], v[i], gmul, a);
  }
}

static int grid_for(long long work_items, int block, int per_sm) {
  const int sms = sm_count() > 0 ? sm_count() : 148;
  long long blocks = (work_items + block - 1) / block;
  const long long cap = (long long)sms * per_sm;   // whole waves of resident CTAs
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace aq

using namespace aq;

extern "C" {

int aq_mapper_fwd(const float* msg, const float* emb, float* scale, int B, int bits, int r, int out_bf16_rounded, void* stream) {
  AQ_REQUIRE(B > 0 && bits > 0 && r > 0, AQ_ERR_BAD_SHAPE, "mapper_fwd: bad shape B=%d bits=%d r=%d", B, bits, r);
  int rc = check_arch();
  if (rc) return rc;
  const int n = B * r;
  mapper_fwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(msg, emb, scale, B, bits, r, out_bf16_rounded);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_mapper_bwd(const float* msg, const float* g_scale, float* g_emb, int B, int bits, int r, void* stream) {
  AQ_REQUIRE(B > 0 && bits > 0 && r > 0, AQ_ERR_BAD_SHAPE, "mapper_bwd: bad shape B=%d bits=%d r=%d", B, bits, r);
  int rc = check_arch();
  if (rc) return rc;
  const int n = bits * r;
  mapper_bwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(msg, g_scale, g_emb, B, bits, r);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_cast_transpose_bf16(const float* src, void* dst, void* dst_t, int rows, int cols, void* stream) {
  AQ_REQUIRE(rows > 0 && cols > 0, AQ_ERR_BAD_SHAPE, "cast_transpose: bad shape %d x %d", rows, cols);
  int rc = check_arch();
  if (rc) return rc;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  cast_transpose_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, (__nv_bfloat16*)dst_t, rows, cols);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_cast_transpose_bf16_batched(const aq_cast_job* jobs, int njobs, int64_t total_tiles, void* stream) {
  static_assert(sizeof(CastJob) == sizeof(aq_cast_job), "job table layout");
  AQ_REQUIRE(jobs != nullptr && njobs > 0 && total_tiles > 0 && total_tiles < (1ll << 31), AQ_ERR_BAD_SHAPE,
             "cast_transpose_batched: bad table (njobs=%d, tiles=%lld)", njobs, (long long)total_tiles);
  int rc = check_arch();
  if (rc) return rc;
  cast_transpose_batched_kernel<<<(unsigned)total_tiles, dim3(32, 8), 0, (cudaStream_t)stream>>>(reinterpret_cast<const CastJob*>(jobs), njobs);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_transpose_bf16(const void* src, void* dst, int rows, int cols, void* stream) {
  AQ_REQUIRE(rows > 0 && cols > 0, AQ_ERR_BAD_SHAPE, "transpose: bad shape %d x %d", rows, cols);
  int rc = check_arch();
  if (rc) return rc;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  cast_transpose_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, nullptr, (__nv_bfloat16*)dst, rows, cols);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_flat_sumsq(const float* g, int64_t n, float* norm_sq, void* stream) {
  AQ_REQUIRE(n > 0, AQ_ERR_BAD_SHAPE, "flat_sumsq: n=%lld", (long long)n);
  AQ_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15u) == 0, AQ_ERR_BAD_ALIGN, "flat_sumsq: buffer must be 16-byte aligned");
  int rc = check_arch();
  if (rc) return rc;
  flat_sumsq_kernel<<<grid_for(n / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(g, n, norm_sq);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_flat_clip_adamw(float* p, float* g, float* m, float* v, int64_t n, const float* norm_sq, float grad_scale, float max_norm,
                       float lr, float beta1, float beta2, float eps, float weight_decay, int step, void* stream) {
  AQ_REQUIRE(n > 0 && step >= 1, AQ_ERR_BAD_SHAPE, "flat_clip_adamw: n=%lld step=%d", (long long)n, step);
  AQ_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
               reinterpret_cast<uintptr_t>(v)) & 15u) == 0,
             AQ_ERR_BAD_ALIGN, "flat_clip_adamw: buffers must be 16-byte aligned");
  int rc = check_arch();
  if (rc) return rc;
  AdamArgs a;
  a.grad_scale = grad_scale; a.max_norm = max_norm; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  a.weight_decay = weight_decay;
  a.bc1 = 1.f - powf(beta1, (float)step);
  a.bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  flat_clip_adamw_kernel<<<grid_for(n / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, norm_sq, a);
  AQ_LAUNCHED();
  return AQ_OK;
}

}  // extern "C"
