// Training-mode BatchNorm2d (+ SiLU) over NHWC fp32 activations [M = B * H * W, C] for the message decoder's train() path
// (train/latent_wm_pretrain.py:160-217, rob_enhance_finetune.py:980: `sec_decoder.train()` -> torchvision's Conv2dNormActivation uses
// batch statistics).  HBM-bound; one statistics pass + one apply pass forward (PyTorch eager: statistics, normalise, SiLU = 3 passes
// + the saved SiLU input), one reduction pass + one apply pass backward (eager: SiLU backward, BN backward reduce, BN backward apply).
//
//   forward   mean_c, var_c over the M rows (biased, fp64 accumulation of per-thread fp32 partials) ; y = act(gamma (z - mean) rstd + beta)
//             running_mean / running_var updated with `momentum` (running_var takes the unbiased variance, as nn.BatchNorm2d does)
//   backward  g_u = gy * act'(u), u = gamma zhat + beta (recomputed from z: nothing but z and (mean, rstd) is saved)
//             g_gamma += sum g_u zhat ; g_beta += sum g_u ; gz = gamma rstd (g_u - mean(g_u) - zhat mean(g_u zhat))
#include "aq_common.h"

namespace aq {

constexpr int kBnThreads = 256;
constexpr int kBnMaxC = 2048;

__device__ __forceinline__ float bn_silu(float u) { return u / (1.f + expf(-u)); }
__device__ __forceinline__ float bn_silu_grad(float u) {
  const float s = 1.f / (1.f + expf(-u));
  return s * (1.f + u * (1.f - s));
}

// Thread -> (column slot, row phase): consecutive threads read consecutive float4 of a row (coalesced); when a row has fewer
// than 256 float4 the block covers 256 / C4 rows per pass, when it has more (C up to 2048) a thread owns two column slots.
struct BnThreadMap {
  int cols, rp, cslot, r0, nslots;
  bool active;
  __device__ BnThreadMap(int C4) {
    cols = C4 < kBnThreads ? C4 : kBnThreads;
    rp = kBnThreads / cols;
    cslot = (int)threadIdx.x % cols;
    r0 = (int)threadIdx.x / cols;
    active = r0 < rp;
    nslots = (C4 + kBnThreads - 1) / kBnThreads;
  }
};

// BWD = false: sums[c] = (sum z, sum z^2).   BWD = true: sums[c] = (sum g_u, sum g_u zhat).
template <bool BWD, int ACT>
__global__ void __launch_bounds__(kBnThreads) bn_reduce_kernel(const float* __restrict__ z, const float* __restrict__ gy,
                                                                const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, double* __restrict__ sums, long long M,
                                                                int C, int rows_per_block) {
  __shared__ float sh[2][kBnMaxC];
  const int C4 = C >> 2;
  const BnThreadMap tm(C4);
  for (int c = threadIdx.x; c < C; c += kBnThreads) { sh[0][c] = 0.f; sh[1][c] = 0.f; }
  __syncthreads();
  const long long rb = (long long)blockIdx.x * rows_per_block;
  const long long re = rb + rows_per_block < M ? rb + rows_per_block : M;
  const float4* z4 = reinterpret_cast<const float4*>(z);
  const float4* g4 = reinterpret_cast<const float4*>(gy);
  if (tm.active) {
    for (int j = 0; j < tm.nslots; ++j) {
      const int c4 = tm.cslot + j * kBnThreads;
      if (c4 >= C4) break;
      float mu[4] = {0, 0, 0, 0}, rs[4] = {1, 1, 1, 1}, ga[4] = {1, 1, 1, 1}, be[4] = {0, 0, 0, 0};
      if (BWD) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          mu[k] = mean_rstd[2 * (c4 * 4 + k)]; rs[k] = mean_rstd[2 * (c4 * 4 + k) + 1];
          ga[k] = gamma[c4 * 4 + k]; be[k] = beta[c4 * 4 + k];
        }
      }
      float a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
      for (long long r = rb + tm.r0; r < re; r += tm.rp) {
        const float4 v4 = __ldg(z4 + r * C4 + c4);
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
        if (!BWD) {
#pragma unroll
          for (int k = 0; k < 4; ++k) { a[k] += v[k]; b[k] = fmaf(v[k], v[k], b[k]); }
        } else {
          const float4 g4v = __ldg(g4 + r * C4 + c4);
          const float g[4] = {g4v.x, g4v.y, g4v.z, g4v.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float zh = (v[k] - mu[k]) * rs[k];
            const float gu = ACT ? g[k] * bn_silu_grad(fmaf(zh, ga[k], be[k])) : g[k];
            a[k] += gu; b[k] = fmaf(gu, zh, b[k]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        atomicAdd(&sh[0][c4 * 4 + k], a[k]);
        atomicAdd(&sh[1][c4 * 4 + k], b[k]);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kBnThreads) {
    atomicAdd(sums + 2 * c, (double)sh[0][c]);
    atomicAdd(sums + 2 * c + 1, (double)sh[1][c]);
  }
}

// per channel: (mean, rstd) from the fp64 sums; running statistics (nn.BatchNorm2d: running = (1 - m) running + m batch, unbiased var)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, float* __restrict__ mean_rstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, long long M, int C, float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = sums[2 * c] / (double)M;
  double var = sums[2 * c + 1] / (double)M - mean * mean;
  var = var < 0.0 ? 0.0 : var;
  mean_rstd[2 * c] = (float)mean;
  mean_rstd[2 * c + 1] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean != nullptr) {
    const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
    running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
    running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
  }
}

template <int ACT>
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(const float* __restrict__ z, const float* __restrict__ mean_rstd,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               float* __restrict__ y, long long n4, int C4) {
  const float4* z4 = reinterpret_cast<const float4*>(z);
  float4* y4 = reinterpret_cast<float4*>(y);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const float4 v4 = __ldg(z4 + i);
    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float sc = mean_rstd[2 * (c + k) + 1] * gamma[c + k];
      const float u = fmaf(v[k] - mean_rstd[2 * (c + k)], sc, beta[c + k]);
      o[k] = ACT ? bn_silu(u) : u;
    }
    y4[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

template <int ACT>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(const float* __restrict__ gy, const float* __restrict__ z,
                                                                   const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, const double* __restrict__ sums,
                                                                   float* __restrict__ gz, long long n4, int C4, double inv_m) {
  const float4* z4 = reinterpret_cast<const float4*>(z);
  const float4* g4 = reinterpret_cast<const float4*>(gy);
  float4* o4 = reinterpret_cast<float4*>(gz);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const float4 v4 = __ldg(z4 + i), gv4 = __ldg(g4 + i);
    const float v[4] = {v4.x, v4.y, v4.z, v4.w}, g[4] = {gv4.x, gv4.y, gv4.z, gv4.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float mu = mean_rstd[2 * (c + k)], rs = mean_rstd[2 * (c + k) + 1], ga = gamma[c + k];
      const float zh = (v[k] - mu) * rs;
      const float gu = ACT ? g[k] * bn_silu_grad(fmaf(zh, ga, beta[c + k])) : g[k];
      const float m1 = (float)(sums[2 * (c + k)] * inv_m), m2 = (float)(sums[2 * (c + k) + 1] * inv_m);
      o[k] = ga * rs * (gu - m1 - zh * m2);
    }
    o4[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void bn_param_grad_kernel(const double* __restrict__ sums, float* __restrict__ g_gamma, float* __restrict__ g_beta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (g_beta != nullptr) g_beta[c] += (float)sums[2 * c];
  if (g_gamma != nullptr) g_gamma[c] += (float)sums[2 * c + 1];
}

static int bn_rows_per_block(long long M, int C) {
  const int sms = sm_count() > 0 ? sm_count() : 148;
  const int C4 = C / 4;
  const int cols = C4 < kBnThreads ? C4 : kBnThreads;
  const int rp = kBnThreads / cols;
  long long rows = (M + 4LL * sms - 1) / (4LL * sms);      // ~4 blocks per SM
  const long long min_rows = 16LL * rp;                     // >= 16 passes per block: the shared / global atomics stay a small share
  if (rows < min_rows) rows = min_rows;
  return (int)rows;
}

static int bn_grid(long long n4) {
  const int sms = sm_count() > 0 ? sm_count() : 148;
  long long blocks = (n4 + kBnThreads - 1) / kBnThreads;
  const long long cap = (long long)sms * 16;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace aq

using namespace aq;

extern "C" {

size_t aq_bn_train_workspace_bytes(int C) { return (size_t)C * 2 * sizeof(double); }

int aq_bn_train_fwd(const float* z, const float* gamma, const float* beta, float* running_mean, float* running_var, float* mean_rstd,
                    float* y, int64_t M, int C, float eps, float momentum, int act, void* ws, size_t ws_bytes, void* stream) {
  AQ_REQUIRE(z && gamma && beta && mean_rstd && y && M > 0 && C > 0, AQ_ERR_BAD_SHAPE, "bn_train_fwd: bad arguments");
  AQ_REQUIRE(C % 4 == 0 && C <= kBnMaxC, AQ_ERR_BAD_SHAPE, "bn_train_fwd: C=%d must be a multiple of 4, <= %d", C, kBnMaxC);
  AQ_REQUIRE((running_mean == nullptr) == (running_var == nullptr), AQ_ERR_BAD_SHAPE, "bn_train_fwd: pass both running statistics or neither");
  AQ_REQUIRE(ws != nullptr && ws_bytes >= aq_bn_train_workspace_bytes(C), AQ_ERR_WORKSPACE, "bn_train_fwd: workspace too small");
  AQ_REQUIRE(((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0 && (reinterpret_cast<uintptr_t>(ws) & 7u) == 0,
             AQ_ERR_BAD_ALIGN, "bn_train_fwd: z / y must be 16-byte aligned, ws 8-byte aligned");
  int rc = check_arch();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  double* sums = reinterpret_cast<double*>(ws);
  AQ_CHECK_CUDA(cudaMemsetAsync(sums, 0, aq_bn_train_workspace_bytes(C), st));
  const int rows = bn_rows_per_block(M, C);
  const int grid = (int)((M + rows - 1) / rows);
  bn_reduce_kernel<false, 0><<<grid, kBnThreads, 0, st>>>(z, nullptr, nullptr, nullptr, nullptr, sums, M, C, rows);
  AQ_LAUNCHED();
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, mean_rstd, running_mean, running_var, M, C, eps, momentum);
  AQ_LAUNCHED();
  const long long n4 = M * (C / 4);
  if (act) bn_apply_kernel<1><<<bn_grid(n4), kBnThreads, 0, st>>>(z, mean_rstd, gamma, beta, y, n4, C / 4);
  else bn_apply_kernel<0><<<bn_grid(n4), kBnThreads, 0, st>>>(z, mean_rstd, gamma, beta, y, n4, C / 4);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_bn_train_bwd(const float* gy, const float* z, const float* gamma, const float* beta, const float* mean_rstd, float* gz,
                    float* g_gamma, float* g_beta, int64_t M, int C, int act, void* ws, size_t ws_bytes, void* stream) {
  AQ_REQUIRE(gy && z && gamma && beta && mean_rstd && gz && M > 0 && C > 0, AQ_ERR_BAD_SHAPE, "bn_train_bwd: bad arguments");
  AQ_REQUIRE(C % 4 == 0 && C <= kBnMaxC, AQ_ERR_BAD_SHAPE, "bn_train_bwd: C=%d must be a multiple of 4, <= %d", C, kBnMaxC);
  AQ_REQUIRE(ws != nullptr && ws_bytes >= aq_bn_train_workspace_bytes(C), AQ_ERR_WORKSPACE, "bn_train_bwd: workspace too small");
  AQ_REQUIRE(((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(gz)) & 15u) == 0 &&
                 (reinterpret_cast<uintptr_t>(ws) & 7u) == 0,
             AQ_ERR_BAD_ALIGN, "bn_train_bwd: gy / z / gz must be 16-byte aligned, ws 8-byte aligned");
  int rc = check_arch();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  double* sums = reinterpret_cast<double*>(ws);
  AQ_CHECK_CUDA(cudaMemsetAsync(sums, 0, aq_bn_train_workspace_bytes(C), st));
  const int rows = bn_rows_per_block(M, C);
  const int grid = (int)((M + rows - 1) / rows);
  if (act) bn_reduce_kernel<true, 1><<<grid, kBnThreads, 0, st>>>(z, gy, mean_rstd, gamma, beta, sums, M, C, rows);
  else bn_reduce_kernel<true, 0><<<grid, kBnThreads, 0, st>>>(z, gy, mean_rstd, gamma, beta, sums, M, C, rows);
  AQ_LAUNCHED();
  const long long n4 = M * (C / 4);
  const double inv_m = 1.0 / (double)M;
  if (act) bn_bwd_apply_kernel<1><<<bn_grid(n4), kBnThreads, 0, st>>>(gy, z, mean_rstd, gamma, beta, sums, gz, n4, C / 4, inv_m);
  else bn_bwd_apply_kernel<0><<<bn_grid(n4), kBnThreads, 0, st>>>(gy, z, mean_rstd, gamma, beta, sums, gz, n4, C / 4, inv_m);
  AQ_LAUNCHED();
  if (g_gamma != nullptr || g_beta != nullptr) {
    bn_param_grad_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, g_gamma, g_beta, C);
    AQ_LAUNCHED();
  }
  return AQ_OK;
}

}  // extern "C"
