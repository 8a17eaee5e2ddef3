// HBM-bound glue of the U-Net around the LoRA projections: GroupNorm (+ broadcast add, + SiLU) over
// channels-last bf16 activations, forward and backward, and GEGLU forward and backward.
//
// Replaces, in scripts/lib/original_unet.py: ResnetBlock2D's `norm1 -> silu`, `+ temb[:, :, None, None] -> norm2 -> silu`
// (:440-453), Transformer2DModel.norm (:826), conv_norm_out -> silu (:1416), and GEGLU.forward's chunk / gelu / mul (:708-729).
// PyTorch's native_group_norm only knows NCHW: on the channels-last tensors the tensor-core convolutions want it copies the
// input to NCHW, normalises, and the next convolution copies back (17 ms of layout copies per PPFT step at B = 16).  Here
// the rows stay [B, H*W, C]: a thread owns 8 consecutive channels (one 16-byte load) and walks down the rows.
//
// Two passes per direction (statistics, then apply); the second read of the activations comes from the 126 MB L2 at the
// BASELINE sizes ([16, 64, 64, 320] bf16 = 42 MB).  Statistics leave each CTA as fp32 partials and are combined in fp64.
#include "aq_common.h"
#include "aq_ptx.cuh"

namespace aq {

constexpr int kGnMaxGroups = 128;
// Rows per thread and round.  Every loop below is software-pipelined: the loads of round i + 1 are issued before round i is
// consumed, so a CTA streams continuously instead of alternating between a load burst and arithmetic (first version: all CTAs of
// a wave in lock-step, 2.2 TB/s; ncu in profiles/r01_glue_launches_v13.txt).
constexpr int kGnBatch = 2;      // backward: two tensors, 2 x 2 x 2 loads in flight
constexpr int kGnBatchFwd = 4;   // forward: one tensor, 2 x 4 loads in flight

struct GnParams {
  const uint4* x;        // [B, HW, C] bf16
  const uint4* dy;       // backward: [B, HW, C] bf16
  uint4* out;            // forward: y, backward: dx
  const __nv_bfloat16* gamma;   // [C]
  const __nv_bfloat16* beta;    // [C]
  const __nv_bfloat16* add_bc;  // [B, C] or null: x' = x + add_bc[b, c] is what gets normalised
  double* sums;          // [B, G, 2] zeroed by the caller of the statistics pass
  float* mean_rstd;      // [B, G, 2] forward: written; backward: read
  int HW, C, G, cpg, V, R, rows_per_cta;
  float eps;
};

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = bf16_lo(w[i]);
    f[2 * i + 1] = bf16_hi(w[i]);
  }
}

__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  return o;
}

__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&f)[8]) {
  unpack8(__ldg(reinterpret_cast<const uint4*>(p)), f);
}

// sigmoid(u) = 0.5 tanh(0.5 u) + 0.5 with the single-instruction tanh.approx (max rel. error 2^-11, below the bf16 rounding of the
// outputs): ONE special-function op per element instead of two (ex2 + rcp) -- at 16 MUFU results per clock and SM the exp / rcp
// pair alone cost 9 of the 21.7 us of gn_fwd_apply on [16, 320, 64, 64] (profiles/r01_ncu_glue_v14_full_summary.txt)
__device__ __forceinline__ float sigmoidf_fast(float u) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * u));
  return fmaf(0.5f, t, 0.5f);
}

// per-thread partials of 8 channels -> per-group partials of the CTA (shared fp32 atomics) -> fp64 global atomics
__device__ __forceinline__ void gn_reduce_to_groups(const float (&a)[8], const float (&b)[8], int ch0, const GnParams& p, int batch) {
  __shared__ float sh[2 * kGnMaxGroups];
  for (int i = threadIdx.x; i < 2 * p.G; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  // 8 consecutive channels touch one group, or two when the run crosses a boundary, or more when cpg < 8
  float ra = 0.f, rb = 0.f;
  int g_run = ch0 / p.cpg;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (ch0 + i) / p.cpg;
    if (g != g_run) {
      atomicAdd(&sh[2 * g_run], ra);
      atomicAdd(&sh[2 * g_run + 1], rb);
      ra = rb = 0.f;
      g_run = g;
    }
    ra += a[i];
    rb += b[i];
  }
  atomicAdd(&sh[2 * g_run], ra);
  atomicAdd(&sh[2 * g_run + 1], rb);
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * p.G; i += blockDim.x) atomicAdd(p.sums + (size_t)batch * 2 * p.G + i, (double)sh[i]);
}

// ---------------------------------------------------------------- forward
__global__ void __launch_bounds__(1024) gn_fwd_stats_kernel(const GnParams p) {
  const int v = threadIdx.x % p.V, rr = threadIdx.x / p.V;
  const int b = blockIdx.y;
  const int r_begin = blockIdx.x * p.rows_per_cta;
  const int r_end = min(r_begin + p.rows_per_cta, p.HW);
  const uint4* xb = p.x + (size_t)b * p.HW * p.V + v;
  float add[8] = {};
  if (p.add_bc != nullptr) load8_bf16(p.add_bc + (size_t)b * p.C + v * 8, add);
  float s[8] = {}, ss[8] = {};
  auto load = [&](uint4 (&q)[kGnBatchFwd], int r) {
#pragma unroll
    for (int j = 0; j < kGnBatchFwd; ++j) {
      const int rj = r + j * p.R;
      q[j] = rj < r_end ? __ldg(xb + (size_t)rj * p.V) : make_uint4(0, 0, 0, 0);
    }
  };
  uint4 q[kGnBatchFwd], nq[kGnBatchFwd];
  load(q, r_begin + rr);
  for (int r = r_begin + rr; r < r_end; r += kGnBatchFwd * p.R) {
    load(nq, r + kGnBatchFwd * p.R);      // rows past r_end come back as zeros without touching memory
#pragma unroll
    for (int j = 0; j < kGnBatchFwd; ++j) {
      if (r + j * p.R >= r_end) continue;
      float f[8];
      unpack8(q[j], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float t = f[i] + add[i];
        s[i] += t;
        ss[i] = fmaf(t, t, ss[i]);
      }
    }
#pragma unroll
    for (int j = 0; j < kGnBatchFwd; ++j) q[j] = nq[j];
  }
  gn_reduce_to_groups(s, ss, v * 8, p, b);
}

// (mean, rstd) of every group of sample b into shared memory, once per CTA.  Forward: from the fp64 sums -- only the cancelling
// subtraction E[x^2] - mean^2 is done in double (B200 issues fp64 at 1/64 of the fp32 rate: the first version did a double
// division and square root per channel per thread and spent more time there than on the data).  Backward: from mean_rstd.
__device__ __forceinline__ void gn_group_stats(const GnParams& p, int b, float* sh_mr) {
  for (int g = threadIdx.x; g < p.G; g += blockDim.x) {
    float m, rs;
    if (p.sums != nullptr) {
      const double inv_n = 1.0 / ((double)p.cpg * (double)p.HW);
      const double mu = p.sums[((size_t)b * p.G + g) * 2] * inv_n;
      double var = p.sums[((size_t)b * p.G + g) * 2 + 1] * inv_n - mu * mu;
      var = var < 0.0 ? 0.0 : var;
      m = (float)mu;
      rs = 1.0f / sqrtf((float)var + p.eps);
    } else {
      m = p.mean_rstd[((size_t)b * p.G + g) * 2];
      rs = p.mean_rstd[((size_t)b * p.G + g) * 2 + 1];
    }
    sh_mr[2 * g] = m;
    sh_mr[2 * g + 1] = rs;
  }
  __syncthreads();
}

// per-channel affine of this thread's 8 channels: y = x * a + c
__device__ __forceinline__ void gn_channel_affine(const GnParams& p, const float* sh_mr, int ch0, float (&a)[8], float (&c)[8],
                                                  float (&mean)[8], float (&rstd)[8]) {
  float gam[8], bet[8];
  load8_bf16(p.gamma + ch0, gam);
  load8_bf16(p.beta + ch0, bet);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (ch0 + i) / p.cpg;
    mean[i] = sh_mr[2 * g];
    rstd[i] = sh_mr[2 * g + 1];
    a[i] = rstd[i] * gam[i];
    c[i] = bet[i] - mean[i] * a[i];
  }
}

template <bool SILU>
__global__ void __launch_bounds__(1024) gn_fwd_apply_kernel(const GnParams p) {
  const int v = threadIdx.x % p.V, rr = threadIdx.x / p.V;
  const int b = blockIdx.y;
  const int r_begin = blockIdx.x * p.rows_per_cta;
  const int r_end = min(r_begin + p.rows_per_cta, p.HW);
  __shared__ float sh_mr[2 * kGnMaxGroups];
  gn_group_stats(p, b, sh_mr);
  float a[8], c[8], mean[8], rstd[8];
  gn_channel_affine(p, sh_mr, v * 8, a, c, mean, rstd);
  if (p.add_bc != nullptr) {
    float add[8];
    load8_bf16(p.add_bc + (size_t)b * p.C + v * 8, add);
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = fmaf(add[i], a[i], c[i]);
  }
  if (blockIdx.x == 0 && rr == 0 && p.mean_rstd != nullptr) {
    // one writer per group: the thread that owns the group's first channel
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ch = v * 8 + i;
      if (ch % p.cpg == 0) {
        const int g = ch / p.cpg;
        p.mean_rstd[((size_t)b * p.G + g) * 2] = mean[i];
        p.mean_rstd[((size_t)b * p.G + g) * 2 + 1] = rstd[i];
      }
    }
  }
  const uint4* xb = p.x + (size_t)b * p.HW * p.V + v;
  uint4* yb = p.out + (size_t)b * p.HW * p.V + v;
  auto load = [&](uint4 (&q)[kGnBatchFwd], int r) {
#pragma unroll
    for (int j = 0; j < kGnBatchFwd; ++j) {
      const int rj = r + j * p.R;
      q[j] = rj < r_end ? __ldg(xb + (size_t)rj * p.V) : make_uint4(0, 0, 0, 0);
    }
  };
  uint4 q[kGnBatchFwd], nq[kGnBatchFwd];
  load(q, r_begin + rr);
  for (int r = r_begin + rr; r < r_end; r += kGnBatchFwd * p.R) {
    load(nq, r + kGnBatchFwd * p.R);
#pragma unroll
    for (int j = 0; j < kGnBatchFwd; ++j) {
      if (r + j * p.R >= r_end) continue;
      float f[8];
      unpack8(q[j], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float u = fmaf(f[i], a[i], c[i]);
        f[i] = SILU ? u * sigmoidf_fast(u) : u;
      }
      yb[(size_t)(r + j * p.R) * p.V] = pack8(f);
    }
#pragma unroll
    for (int j = 0; j < kGnBatchFwd; ++j) q[j] = nq[j];
  }
}

// ---------------------------------------------------------------- backward (gamma / beta are frozen: only dx)
// u = xhat * gamma + beta, y = silu(u) or u;  du = dy * silu'(u);  t = gamma * du
// dx = rstd * (t - mean_g(t) - xhat * mean_g(t * xhat))
template <bool SILU>
__device__ __forceinline__ void gn_bwd_terms(const float (&x)[8], const float (&dy)[8], const float (&a)[8], const float (&c)[8],
                                             const float (&mean)[8], const float (&rstd)[8], const float (&gam)[8], float (&t)[8],
                                             float (&xhat)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    xhat[i] = (x[i] - mean[i]) * rstd[i];
    float du = dy[i];
    if (SILU) {
      const float u = fmaf(x[i], a[i], c[i]);
      const float sg = sigmoidf_fast(u);
      du *= sg * fmaf(u, 1.f - sg, 1.f);
    }
    t[i] = gam[i] * du;
  }
}

template <bool SILU>
__global__ void gn_bwd_stats_kernel(const GnParams p) {
  const int v = threadIdx.x % p.V, rr = threadIdx.x / p.V;
  const int b = blockIdx.y;
  const int r_begin = blockIdx.x * p.rows_per_cta;
  const int r_end = min(r_begin + p.rows_per_cta, p.HW);
  GnParams ps = p;
  ps.sums = nullptr;   // statistics of the forward come from mean_rstd
  __shared__ float sh_mr[2 * kGnMaxGroups];
  gn_group_stats(ps, b, sh_mr);
  float a[8], c[8], mean[8], rstd[8], gam[8];
  gn_channel_affine(ps, sh_mr, v * 8, a, c, mean, rstd);
  load8_bf16(p.gamma + v * 8, gam);
  if (p.add_bc != nullptr) {
    float add[8];
    load8_bf16(p.add_bc + (size_t)b * p.C + v * 8, add);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      c[i] = fmaf(add[i], a[i], c[i]);
      mean[i] -= add[i];   // xhat = (x + add - mean) * rstd
    }
  }
  const uint4* xb = p.x + (size_t)b * p.HW * p.V + v;
  const uint4* gb = p.dy + (size_t)b * p.HW * p.V + v;
  float s1[8] = {}, s2[8] = {};
  auto load = [&](uint4 (&q)[kGnBatch], uint4 (&d)[kGnBatch], int r) {
#pragma unroll
    for (int j = 0; j < kGnBatch; ++j) {
      const int rj = r + j * p.R;
      q[j] = rj < r_end ? __ldg(xb + (size_t)rj * p.V) : make_uint4(0, 0, 0, 0);
      d[j] = rj < r_end ? __ldg(gb + (size_t)rj * p.V) : make_uint4(0, 0, 0, 0);
    }
  };
  uint4 q[kGnBatch], d[kGnBatch], nq[kGnBatch], nd[kGnBatch];
  load(q, d, r_begin + rr);
  for (int r = r_begin + rr; r < r_end; r += kGnBatch * p.R) {
    load(nq, nd, r + kGnBatch * p.R);
#pragma unroll
    for (int j = 0; j < kGnBatch; ++j) {
      if (r + j * p.R >= r_end) continue;
      float x[8], dy[8], t[8], xhat[8];
      unpack8(q[j], x);
      unpack8(d[j], dy);
      gn_bwd_terms<SILU>(x, dy, a, c, mean, rstd, gam, t, xhat);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s1[i] += t[i];
        s2[i] = fmaf(t[i], xhat[i], s2[i]);
      }
    }
#pragma unroll
    for (int j = 0; j < kGnBatch; ++j) {
      q[j] = nq[j];
      d[j] = nd[j];
    }
  }
  gn_reduce_to_groups(s1, s2, v * 8, p, b);
}

template <bool SILU>
__global__ void gn_bwd_apply_kernel(const GnParams p) {
  const int v = threadIdx.x % p.V, rr = threadIdx.x / p.V;
  const int b = blockIdx.y;
  const int r_begin = blockIdx.x * p.rows_per_cta;
  const int r_end = min(r_begin + p.rows_per_cta, p.HW);
  GnParams ps = p;
  ps.sums = nullptr;
  __shared__ float sh_mr[2 * kGnMaxGroups];
  __shared__ float sh_m12[2 * kGnMaxGroups];
  {
    const float inv_n = 1.f / ((float)p.cpg * (float)p.HW);
    for (int i = threadIdx.x; i < 2 * p.G; i += blockDim.x) sh_m12[i] = (float)p.sums[(size_t)b * 2 * p.G + i] * inv_n;
  }
  gn_group_stats(ps, b, sh_mr);   // (its barrier also publishes sh_m12)
  float a[8], c[8], mean[8], rstd[8], gam[8], m1[8], m2[8];
  gn_channel_affine(ps, sh_mr, v * 8, a, c, mean, rstd);
  load8_bf16(p.gamma + v * 8, gam);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (v * 8 + i) / p.cpg;
    m1[i] = sh_m12[2 * g];
    m2[i] = sh_m12[2 * g + 1];
  }
  if (p.add_bc != nullptr) {
    float add[8];
    load8_bf16(p.add_bc + (size_t)b * p.C + v * 8, add);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      c[i] = fmaf(add[i], a[i], c[i]);
      mean[i] -= add[i];
    }
  }
  const uint4* xb = p.x + (size_t)b * p.HW * p.V + v;
  const uint4* gb = p.dy + (size_t)b * p.HW * p.V + v;
  uint4* ob = p.out + (size_t)b * p.HW * p.V + v;
  auto load = [&](uint4 (&q)[kGnBatch], uint4 (&d)[kGnBatch], int r) {
#pragma unroll
    for (int j = 0; j < kGnBatch; ++j) {
      const int rj = r + j * p.R;
      q[j] = rj < r_end ? __ldg(xb + (size_t)rj * p.V) : make_uint4(0, 0, 0, 0);
      d[j] = rj < r_end ? __ldg(gb + (size_t)rj * p.V) : make_uint4(0, 0, 0, 0);
    }
  };
  uint4 q[kGnBatch], d[kGnBatch], nq[kGnBatch], nd[kGnBatch];
  load(q, d, r_begin + rr);
  for (int r = r_begin + rr; r < r_end; r += kGnBatch * p.R) {
    load(nq, nd, r + kGnBatch * p.R);
#pragma unroll
    for (int j = 0; j < kGnBatch; ++j) {
      if (r + j * p.R >= r_end) continue;
      float x[8], dy[8], t[8], xhat[8];
      unpack8(q[j], x);
      unpack8(d[j], dy);
      gn_bwd_terms<SILU>(x, dy, a, c, mean, rstd, gam, t, xhat);
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = rstd[i] * (t[i] - m1[i] - xhat[i] * m2[i]);
      ob[(size_t)(r + j * p.R) * p.V] = pack8(t);
    }
#pragma unroll
    for (int j = 0; j < kGnBatch; ++j) {
      q[j] = nq[j];
      d[j] = nd[j];
    }
  }
}

template <typename K1, typename K2>
static int gn_ctas_per_sm(K1 k1, K2 k2, int threads) {
  static int cache[1025] = {0};   // per kernel pair (template instance); benign race: idempotent
  if (threads >= 0 && threads <= 1024 && cache[threads] > 0) return cache[threads];
  int a = 0, b = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k1, threads, 0) != cudaSuccess || a < 1) a = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k2, threads, 0) != cudaSuccess || b < 1) b = 1;
  if (threads >= 0 && threads <= 1024) cache[threads] = a < b ? a : b;
  return a < b ? a : b;
}

static int gn_setup(GnParams& p, dim3& grid, int& threads, const void* x, const void* gamma, const void* beta, const void* add_bc,
                    double* sums, float* mean_rstd, int B, int HW, int C, int G, float eps, const char* who, int (*occupancy)(int)) {
  AQ_REQUIRE(B > 0 && HW > 0 && C > 0 && G > 0, AQ_ERR_BAD_SHAPE, "%s: empty problem B=%d HW=%d C=%d G=%d", who, B, HW, C, G);
  AQ_REQUIRE(C % 8 == 0 && C % G == 0, AQ_ERR_BAD_SHAPE, "%s: C=%d must be a multiple of 8 and of G=%d", who, C, G);
  AQ_REQUIRE(G <= kGnMaxGroups, AQ_ERR_BAD_SHAPE, "%s: at most %d groups, got %d", who, kGnMaxGroups, G);
  AQ_REQUIRE(C / 8 <= 1024, AQ_ERR_BAD_SHAPE, "%s: C=%d exceeds 8192 channels", who, C);
  AQ_REQUIRE(B <= 65535, AQ_ERR_BAD_SHAPE, "%s: B=%d exceeds 65535 samples per call", who, B);
  AQ_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(gamma) & 15u) == 0 &&
                 (reinterpret_cast<uintptr_t>(beta) & 15u) == 0 && (reinterpret_cast<uintptr_t>(add_bc) & 15u) == 0,
             AQ_ERR_BAD_ALIGN, "%s: pointers must be 16-byte aligned", who);
  AQ_REQUIRE(sums != nullptr && mean_rstd != nullptr, AQ_ERR_WORKSPACE, "%s: workspace / statistics buffer is NULL", who);
  int rc = check_arch();
  if (rc) return rc;
  p.x = reinterpret_cast<const uint4*>(x);
  p.gamma = reinterpret_cast<const __nv_bfloat16*>(gamma);
  p.beta = reinterpret_cast<const __nv_bfloat16*>(beta);
  p.add_bc = reinterpret_cast<const __nv_bfloat16*>(add_bc);
  p.sums = sums;
  p.mean_rstd = mean_rstd;
  p.HW = HW; p.C = C; p.G = G; p.cpg = C / G; p.V = C / 8; p.eps = eps;
  p.R = p.V >= 256 ? 1 : 256 / p.V;
  if (p.R > HW) p.R = HW;
  threads = p.V * p.R;
  // one wave of the pair of kernels that will run (2 - 3 CTAs per SM by registers): every CTA starts at once and streams its
  // slab in several pipelined rounds; at least 2 rows per thread
  const int sms = sm_count() > 0 ? sm_count() : 148;
  int slabs = (sms * occupancy(threads)) / B;
  const int max_slabs = (HW + 2 * p.R - 1) / (2 * p.R);
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  p.rows_per_cta = (HW + slabs - 1) / slabs;
  p.rows_per_cta = (p.rows_per_cta + p.R - 1) / p.R * p.R;
  slabs = (HW + p.rows_per_cta - 1) / p.rows_per_cta;
  grid = dim3((unsigned)slabs, (unsigned)B, 1);
  return AQ_OK;
}

// ---------------------------------------------------------------- GEGLU (original_unet.py:708-729)
// Phi(g) = 0.5 (1 + erf(g / sqrt 2)) and exp(-g^2 / 2) from ONE exponential: erf by Abramowitz & Stegun 7.1.26,
//   erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2),  t = 1 / (1 + 0.3275911 z),  z >= 0,  |error| <= 1.5e-7
// -- far below the bf16 rounding of the outputs (2^-9) -- in ~14 instructions (2 SFU) against ~35 for erff() (+ a second
// exponential for the density in the backward).  The GEGLU kernels were bound by instruction issue, not by HBM (3.6 TB/s).
__device__ __forceinline__ void gauss_cdf_exp(float g, float& cdf, float& e) {
  const float z = fabsf(g) * 0.70710678118654752f;
  float t, ex;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(z * z * -1.4426950408889634f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float half_tail = 0.5f * poly * t * ex;      // 0.5 (1 - erf(z))
  cdf = g < 0.f ? half_tail : 1.f - half_tail;
  e = ex;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float cdf, e;
  gauss_cdf_exp(x, cdf, e);
  return x * cdf;
}

__global__ void geglu_fwd_kernel(const uint4* __restrict__ p, uint4* __restrict__ out, long long M, int FV, long long ldp_v) {
  const long long total = M * FV;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long m = idx / FV;
    const int fv = (int)(idx - m * FV);
    const uint4 hq = __ldg(p + m * ldp_v + fv), gq = __ldg(p + m * ldp_v + FV + fv);
    float h[8], g[8];
    unpack8(hq, h);
    unpack8(gq, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] *= gelu_erf(g[i]);
    out[idx] = pack8(h);
  }
}

__global__ void geglu_bwd_kernel(const uint4* __restrict__ p, const uint4* __restrict__ go, uint4* __restrict__ dp, long long M, int FV,
                                 long long ldp_v) {
  const long long total = M * FV;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long m = idx / FV;
    const int fv = (int)(idx - m * FV);
    const uint4 hq = __ldg(p + m * ldp_v + fv), gq = __ldg(p + m * ldp_v + FV + fv), oq = __ldg(go + idx);
    float h[8], g[8], o[8], dh[8], dg[8];
    unpack8(hq, h);
    unpack8(gq, g);
    unpack8(oq, o);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float cdf, e;
      gauss_cdf_exp(g[i], cdf, e);
      const float pdf = 0.3989422804014327f * e;
      dh[i] = o[i] * g[i] * cdf;
      dg[i] = o[i] * h[i] * fmaf(g[i], pdf, cdf);
    }
    dp[m * (2 * (long long)FV) + fv] = pack8(dh);
    dp[m * (2 * (long long)FV) + FV + fv] = pack8(dg);
  }
}

}  // namespace aq

extern "C" {

size_t aq_group_norm_workspace_bytes(int B, int G) { return (size_t)B * (size_t)G * 2 * sizeof(double); }

int aq_group_norm_nhwc_fwd(const void* x, const void* gamma, const void* beta, const void* add_bc, void* y, float* mean_rstd, int B,
                           int HW, int C, int G, float eps, int silu, void* ws, size_t ws_bytes, void* stream) {
  using namespace aq;
  GnParams p{};
  dim3 grid;
  int threads = 0;
  int (*occ)(int) = silu ? +[](int t) { return gn_ctas_per_sm(gn_fwd_stats_kernel, gn_fwd_apply_kernel<true>, t); }
                          : +[](int t) { return gn_ctas_per_sm(gn_fwd_stats_kernel, gn_fwd_apply_kernel<false>, t); };
  int rc = gn_setup(p, grid, threads, x, gamma, beta, add_bc, reinterpret_cast<double*>(ws), mean_rstd, B, HW, C, G, eps,
                    "aq_group_norm_nhwc_fwd", occ);
  if (rc) return rc;
  AQ_REQUIRE(ws_bytes >= aq_group_norm_workspace_bytes(B, G), AQ_ERR_WORKSPACE, "aq_group_norm_nhwc_fwd: workspace too small");
  AQ_REQUIRE(y != nullptr && (reinterpret_cast<uintptr_t>(y) & 15u) == 0, AQ_ERR_BAD_ALIGN, "aq_group_norm_nhwc_fwd: y must be 16-byte aligned");
  p.out = reinterpret_cast<uint4*>(y);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  AQ_CHECK_CUDA(cudaMemsetAsync(ws, 0, aq_group_norm_workspace_bytes(B, G), st));
  gn_fwd_stats_kernel<<<grid, threads, 0, st>>>(p);
  AQ_LAUNCHED();
  if (silu) gn_fwd_apply_kernel<true><<<grid, threads, 0, st>>>(p);
  else gn_fwd_apply_kernel<false><<<grid, threads, 0, st>>>(p);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_group_norm_nhwc_bwd(const void* dy, const void* x, const void* gamma, const void* beta, const void* add_bc,
                           const float* mean_rstd, void* dx, int B, int HW, int C, int G, float eps, int silu, void* ws, size_t ws_bytes,
                           void* stream) {
  using namespace aq;
  GnParams p{};
  dim3 grid;
  int threads = 0;
  int (*occ)(int) = silu ? +[](int t) { return gn_ctas_per_sm(gn_bwd_stats_kernel<true>, gn_bwd_apply_kernel<true>, t); }
                          : +[](int t) { return gn_ctas_per_sm(gn_bwd_stats_kernel<false>, gn_bwd_apply_kernel<false>, t); };
  int rc = gn_setup(p, grid, threads, x, gamma, beta, add_bc, reinterpret_cast<double*>(ws), const_cast<float*>(mean_rstd), B, HW, C, G,
                    eps, "aq_group_norm_nhwc_bwd", occ);
  if (rc) return rc;
  AQ_REQUIRE(ws_bytes >= aq_group_norm_workspace_bytes(B, G), AQ_ERR_WORKSPACE, "aq_group_norm_nhwc_bwd: workspace too small");
  AQ_REQUIRE(dy != nullptr && dx != nullptr && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15u) == 0,
             AQ_ERR_BAD_ALIGN, "aq_group_norm_nhwc_bwd: dy / dx must be 16-byte aligned");
  p.dy = reinterpret_cast<const uint4*>(dy);
  p.out = reinterpret_cast<uint4*>(dx);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  AQ_CHECK_CUDA(cudaMemsetAsync(ws, 0, aq_group_norm_workspace_bytes(B, G), st));
  if (silu) gn_bwd_stats_kernel<true><<<grid, threads, 0, st>>>(p);
  else gn_bwd_stats_kernel<false><<<grid, threads, 0, st>>>(p);
  AQ_LAUNCHED();
  if (silu) gn_bwd_apply_kernel<true><<<grid, threads, 0, st>>>(p);
  else gn_bwd_apply_kernel<false><<<grid, threads, 0, st>>>(p);
  AQ_LAUNCHED();
  return AQ_OK;
}

static int geglu_grid(long long total) {
  const int sms = aq::sm_count() > 0 ? aq::sm_count() : 148;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sms * 16;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

int aq_geglu_fwd(const void* proj, int64_t ldp, void* out, int64_t M, int F, void* stream) {
  using namespace aq;
  AQ_REQUIRE(M > 0 && F > 0 && F % 8 == 0 && ldp % 8 == 0 && ldp >= 2 * (int64_t)F, AQ_ERR_BAD_SHAPE,
             "aq_geglu_fwd: need M > 0, F %% 8 == 0, ldp %% 8 == 0, ldp >= 2F (M=%lld F=%d ldp=%lld)", (long long)M, F, (long long)ldp);
  AQ_REQUIRE(((reinterpret_cast<uintptr_t>(proj) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0 && proj != nullptr && out != nullptr,
             AQ_ERR_BAD_ALIGN, "aq_geglu_fwd: pointers must be non-NULL and 16-byte aligned");
  int rc = check_arch();
  if (rc) return rc;
  const int FV = F / 8;
  geglu_fwd_kernel<<<geglu_grid(M * FV), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(proj), reinterpret_cast<uint4*>(out), M, FV, ldp / 8);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_geglu_bwd(const void* proj, int64_t ldp, const void* g_out, void* g_proj, int64_t M, int F, void* stream) {
  using namespace aq;
  AQ_REQUIRE(M > 0 && F > 0 && F % 8 == 0 && ldp % 8 == 0 && ldp >= 2 * (int64_t)F, AQ_ERR_BAD_SHAPE,
             "aq_geglu_bwd: need M > 0, F %% 8 == 0, ldp %% 8 == 0, ldp >= 2F (M=%lld F=%d ldp=%lld)", (long long)M, F, (long long)ldp);
  AQ_REQUIRE(proj != nullptr && g_out != nullptr && g_proj != nullptr &&
                 ((reinterpret_cast<uintptr_t>(proj) | reinterpret_cast<uintptr_t>(g_out) | reinterpret_cast<uintptr_t>(g_proj)) & 15u) == 0,
             AQ_ERR_BAD_ALIGN, "aq_geglu_bwd: pointers must be non-NULL and 16-byte aligned");
  int rc = check_arch();
  if (rc) return rc;
  const int FV = F / 8;
  geglu_bwd_kernel<<<geglu_grid(M * FV), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(proj), reinterpret_cast<const uint4*>(g_out), reinterpret_cast<uint4*>(g_proj), M, FV, ldp / 8);
  AQ_LAUNCHED();
  return AQ_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- LayerNorm over bf16 token rows (BasicTransformerBlock.norm1/2/3,
// scripts/lib/original_unet.py:732-806).  One warp per row, the whole row in registers (C <= 2048): one read + one write, statistics
// by warp shuffles (two-pass variance).  The affine parameters are frozen: the backward returns dx only.
namespace aq {

template <int VPL>
__global__ void __launch_bounds__(256) layer_norm_fwd_kernel(const uint4* __restrict__ x, const __nv_bfloat16* __restrict__ gamma,
                                                            const __nv_bfloat16* __restrict__ beta, uint4* __restrict__ y,
                                                            float2* __restrict__ mean_rstd, long long M, int V, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  float gam[VPL][8], bet[VPL][8];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int v = lane + 32 * j;
    if (v < V) {
      load8_bf16(gamma + v * 8, gam[j]);
      load8_bf16(beta + v * 8, bet[j]);
    }
  }
  const float inv_c = 1.f / (float)(V * 8);
  auto load = [&](uint4 (&q)[VPL], long long row) {   // rows past M come back as zeros without touching memory
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int v = lane + 32 * j;
      q[j] = (v < V && row < M) ? __ldg(x + row * V + v) : make_uint4(0, 0, 0, 0);
    }
  };
  uint4 q[VPL], nq[VPL];
  load(q, warp0);
  for (long long row = warp0; row < M; row += nwarps) {
    load(nq, row + nwarps);   // the next row's loads are in flight while this one is reduced: one row per warp at a time left the
                              // kernel latency-bound at 1.5 TB/s (24 warps x 640 B in flight per SM)
    float f[VPL][8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      unpack8(q[j], f[j]);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += f[j][i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_c;
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      if (lane + 32 * j < V) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = f[j][i] - mean;
          ss = fmaf(d, d, ss);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss * inv_c + eps);
    if (lane == 0 && mean_rstd != nullptr) mean_rstd[row] = make_float2(mean, rstd);
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int v = lane + 32 * j;
      if (v < V) {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[j][i] = fmaf((f[j][i] - mean) * rstd, gam[j][i], bet[j][i]);
        y[row * V + v] = pack8(f[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < VPL; ++j) q[j] = nq[j];
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
template <int VPL>
__global__ void __launch_bounds__(256) layer_norm_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x,
                                                            const __nv_bfloat16* __restrict__ gamma, const float2* __restrict__ mean_rstd,
                                                            uint4* __restrict__ dx, long long M, int V) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  float gam[VPL][8];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int v = lane + 32 * j;
    if (v < V) load8_bf16(gamma + v * 8, gam[j]);
  }
  const float inv_c = 1.f / (float)(V * 8);
  auto load = [&](uint4 (&q)[VPL], uint4 (&d)[VPL], float2& mr, long long row) {
    const bool in = row < M;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int v = lane + 32 * j;
      q[j] = (v < V && in) ? __ldg(x + row * V + v) : make_uint4(0, 0, 0, 0);
      d[j] = (v < V && in) ? __ldg(dy + row * V + v) : make_uint4(0, 0, 0, 0);
    }
    mr = in ? __ldg(mean_rstd + row) : make_float2(0.f, 0.f);
  };
  uint4 q[VPL], d[VPL], nq[VPL], nd[VPL];
  float2 mr, nmr;
  load(q, d, mr, warp0);
  for (long long row = warp0; row < M; row += nwarps) {
    load(nq, nd, nmr, row + nwarps);
    float xh[VPL][8], g[VPL][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      unpack8(q[j], xh[j]);
      unpack8(d[j], g[j]);
      if (lane + 32 * j < V) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xh[j][i] = (xh[j][i] - mr.x) * mr.y;
          g[j][i] *= gam[j][i];
          s1 += g[j][i];
          s2 = fmaf(g[j][i], xh[j][i], s2);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 * inv_c, m2 = s2 * inv_c;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int v = lane + 32 * j;
      if (v < V) {
#pragma unroll
        for (int i = 0; i < 8; ++i) g[j][i] = mr.y * (g[j][i] - m1 - xh[j][i] * m2);
        dx[row * V + v] = pack8(g[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      q[j] = nq[j];
      d[j] = nd[j];
    }
    mr = nmr;
  }
}

static int ln_grid(long long M) {
  const int sms = sm_count() > 0 ? sm_count() : 148;
  const long long blocks = (M + 7) / 8;          // 8 warps per CTA, one row per warp at a time
  const long long cap = (long long)sms * 8;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

static int ln_check(const void* a, const void* b, const void* c, long long M, int C, const char* who) {
  AQ_REQUIRE(M > 0 && C > 0 && C % 8 == 0 && C <= 2048, AQ_ERR_BAD_SHAPE, "%s: need M > 0, C %% 8 == 0, C <= 2048 (M=%lld C=%d)", who, M, C);
  AQ_REQUIRE(a != nullptr && b != nullptr && c != nullptr &&
                 ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15u) == 0,
             AQ_ERR_BAD_ALIGN, "%s: pointers must be non-NULL and 16-byte aligned", who);
  return check_arch();
}

}  // namespace aq

extern "C" {

int aq_layer_norm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean_rstd, int64_t M, int C, float eps,
                      void* stream) {
  using namespace aq;
  int rc = ln_check(x, gamma, y, M, C, "aq_layer_norm_fwd");
  if (rc) return rc;
  AQ_REQUIRE(beta != nullptr && (reinterpret_cast<uintptr_t>(beta) & 15u) == 0, AQ_ERR_BAD_ALIGN, "aq_layer_norm_fwd: beta must be 16-byte aligned");
  const int V = C / 8, vpl = (V + 31) / 32;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = ln_grid(M);
#define AQ_LN_FWD(N)                                                                                                       \
  layer_norm_fwd_kernel<N><<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<const __nv_bfloat16*>(gamma), \
                                                 reinterpret_cast<const __nv_bfloat16*>(beta), reinterpret_cast<uint4*>(y),   \
                                                 reinterpret_cast<float2*>(mean_rstd), M, V, eps)
  if (vpl <= 1) AQ_LN_FWD(1);
  else if (vpl == 2) AQ_LN_FWD(2);
  else if (vpl == 3) AQ_LN_FWD(3);
  else if (vpl <= 5) AQ_LN_FWD(5);
  else AQ_LN_FWD(8);
#undef AQ_LN_FWD
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_layer_norm_bwd(const void* dy, const void* x, const void* gamma, const float* mean_rstd, void* dx, int64_t M, int C, void* stream) {
  using namespace aq;
  int rc = ln_check(dy, x, dx, M, C, "aq_layer_norm_bwd");
  if (rc) return rc;
  AQ_REQUIRE(gamma != nullptr && mean_rstd != nullptr && (reinterpret_cast<uintptr_t>(gamma) & 15u) == 0, AQ_ERR_BAD_ALIGN,
             "aq_layer_norm_bwd: gamma / mean_rstd must be non-NULL, gamma 16-byte aligned");
  const int V = C / 8, vpl = (V + 31) / 32;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = ln_grid(M);
#define AQ_LN_BWD(N)                                                                                                               \
  layer_norm_bwd_kernel<N><<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(x),             \
                                                 reinterpret_cast<const __nv_bfloat16*>(gamma), reinterpret_cast<const float2*>(mean_rstd), \
                                                 reinterpret_cast<uint4*>(dx), M, V)
  if (vpl <= 1) AQ_LN_BWD(1);
  else if (vpl == 2) AQ_LN_BWD(2);
  else if (vpl == 3) AQ_LN_BWD(3);
  else if (vpl <= 5) AQ_LN_BWD(5);
  else AQ_LN_BWD(8);
#undef AQ_LN_BWD
  AQ_LAUNCHED();
  return AQ_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- residual add with a per-channel bias
// out = a + b + bias[c] on channels-last bf16 rows: the closing `input + hidden` of ResnetBlock2D.forward
// (scripts/lib/original_unet.py:455-460) with the bias of conv2 (and of conv_shortcut) folded in, instead of cuDNN's separate
// broadcast bias pass after each convolution (2.5 ms per PPFT step at B = 16).
namespace aq {

__global__ void add_bias_rows_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, const __nv_bfloat16* __restrict__ bias,
                                     uint4* __restrict__ out, long long total_v, int V) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total_v; idx += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(idx % V);
    const uint4 qa = __ldg(a + idx), qb = __ldg(b + idx);
    float fa[8], fb[8], fc[8];
    unpack8(qa, fa);
    unpack8(qb, fb);
    load8_bf16(bias + v * 8, fc);
#pragma unroll
    for (int i = 0; i < 8; ++i) fa[i] = fa[i] + fb[i] + fc[i];
    out[idx] = pack8(fa);
  }
}

}  // namespace aq

extern "C" int aq_add_bias_rows(const void* a, const void* b, const void* bias, void* out, int64_t rows, int C, void* stream) {
  using namespace aq;
  AQ_REQUIRE(rows > 0 && C > 0 && C % 8 == 0, AQ_ERR_BAD_SHAPE, "aq_add_bias_rows: need rows > 0, C %% 8 == 0 (rows=%lld C=%d)", (long long)rows, C);
  AQ_REQUIRE(a && b && bias && out &&
                 ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(bias) |
                   reinterpret_cast<uintptr_t>(out)) & 15u) == 0,
             AQ_ERR_BAD_ALIGN, "aq_add_bias_rows: pointers must be non-NULL and 16-byte aligned");
  int rc = check_arch();
  if (rc) return rc;
  const long long total_v = rows * (C / 8);
  const int sms = sm_count() > 0 ? sm_count() : 148;
  long long blocks = (total_v + 255) / 256;
  if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
  add_bias_rows_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b), reinterpret_cast<const __nv_bfloat16*>(bias),
      reinterpret_cast<uint4*>(out), total_v, C / 8);
  AQ_LAUNCHED();
  return AQ_OK;
}
