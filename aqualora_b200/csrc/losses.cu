// Losses of the latent-watermark pretraining step (train/latent_wm_pretrain.py) as fused fp32 kernels:
//   PRVL_loss (:42-50)                      max over positions of the 32 x 32 box mean (zero padding 16) of mean_c |img1 - img2|
//   binary_cross_entropy_with_logits (:200) mean reduction, forward value + gradient in one pass
#include "aq_common.h"

namespace aq {

constexpr int kPrvlWin = 32;   // WINDOW_SIZE, train/latent_wm_pretrain.py:39

// d[b, y, x] = mean over the 3 channels of |a - b|
__global__ void prvl_diff_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d, int B, long long hw) {
  const long long n = (long long)B * hw;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const long long bi = idx / hw, px = idx % hw;
    const float* pa = a + bi * 3 * hw + px;
    const float* pb = b + bi * 3 * hw + px;
    d[idx] = (fabsf(pa[0] - pb[0]) + fabsf(pa[hw] - pb[hw]) + fabsf(pa[2 * hw] - pb[2 * hw])) / 3.f;
  }
}

// hs[b, y, ox] = sum_{x = ox - 16}^{ox + 15} d[b, y, x], ox in [0, W]   (conv2d padding = 16 -> W + 1 output columns)
__global__ void prvl_hsum_kernel(const float* __restrict__ d, float* __restrict__ hs, int B, int H, int W) {
  const int Wo = W + 1;
  const long long n = (long long)B * H * Wo;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % Wo);
    const long long row = idx / Wo;
    const float* src = d + row * W;
    float s = 0.f;
#pragma unroll 8
    for (int k = 0; k < kPrvlWin; ++k) {
      const int x = ox - kPrvlWin / 2 + k;
      if (x >= 0 && x < W) s += __ldg(src + x);
    }
    hs[idx] = s;
  }
}

// v[b, oy, ox] = (sum_{y = oy - 16}^{oy + 15} hs[b, y, ox]) / 1024 ; the global maximum and its position are kept as one packed
// 64-bit word (value bits << 32 | ~position): values are >= 0, so the unsigned order of the bit pattern is the float order; ties go
// to the smallest position.
__global__ void __launch_bounds__(256) prvl_vsum_max_kernel(const float* __restrict__ hs, unsigned long long* __restrict__ packed, int B,
                                                             int H, int W) {
  const int Wo = W + 1, Ho = H + 1;
  const long long n = (long long)B * Ho * Wo;
  unsigned long long best = 0ull;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % Wo);
    const int oy = (int)((idx / Wo) % Ho);
    const long long b = idx / ((long long)Wo * Ho);
    const float* src = hs + b * (long long)H * Wo + ox;
    float s = 0.f;
#pragma unroll 8
    for (int k = 0; k < kPrvlWin; ++k) {
      const int y = oy - kPrvlWin / 2 + k;
      if (y >= 0 && y < H) s += __ldg(src + (long long)y * Wo);
    }
    s *= 1.f / (kPrvlWin * kPrvlWin);
    const unsigned long long key = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)idx);
    best = key > best ? key : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  __shared__ unsigned long long wbest[8];
  if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) best = wbest[w] > best ? wbest[w] : best;
    atomicMax(packed, best);
  }
}

__global__ void prvl_finish_kernel(const unsigned long long* __restrict__ packed, float* __restrict__ loss) {
  loss[0] = __uint_as_float((unsigned)(packed[0] >> 32));
}

// gradient: only the 32 x 32 window at the arg-max position receives g / 1024, spread over the 3 channels (/ 3) with sign(a - b)
__global__ void prvl_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const unsigned long long* __restrict__ packed,
                                const float* __restrict__ g_loss, float* __restrict__ g_a, float* __restrict__ g_b, int B, int H, int W) {
  const int Wo = W + 1, Ho = H + 1;
  const unsigned pos = 0xFFFFFFFFu - (unsigned)(packed[0] & 0xFFFFFFFFull);
  const int ox = (int)(pos % Wo), oy = (int)((pos / Wo) % Ho), bb = (int)(pos / ((unsigned)Wo * Ho));
  const float g = g_loss[0] * (1.f / (kPrvlWin * kPrvlWin)) / 3.f;
  const long long hw = (long long)H * W;
  const long long n = (long long)B * 3 * hw;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    const int y = (int)((idx / W) % H);
    const int bi = (int)(idx / (3 * hw));
    float v = 0.f;
    if (bi == bb && y >= oy - kPrvlWin / 2 && y < oy + kPrvlWin / 2 && x >= ox - kPrvlWin / 2 && x < ox + kPrvlWin / 2) {
      const float df = a[idx] - b[idx];
      v = df > 0.f ? g : (df < 0.f ? -g : 0.f);
    }
    if (g_a != nullptr) g_a[idx] = v;
    if (g_b != nullptr) g_b[idx] = -v;
  }
}

// binary_cross_entropy_with_logits(x, y), mean reduction: loss += sum (max(x, 0) - x y + log1p(exp(-|x|))) / n ; g = (sigmoid(x) - y) / n
__global__ void bce_logits_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ loss, float* __restrict__ g,
                                  long long n) {
  float part = 0.f;
  const float inv = 1.f / (float)n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float xv = x[i], yv = y[i];
    part += fmaxf(xv, 0.f) - xv * yv + log1pf(expf(-fabsf(xv)));
    if (g != nullptr) g[i] = (1.f / (1.f + expf(-xv)) - yv) * inv;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss, part * inv);
}

static int grid_for(long long work, int block) {
  const int sms = sm_count() > 0 ? sm_count() : 148;
  long long blocks = (work + block - 1) / block;
  const long long cap = (long long)sms * 16;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace aq

using namespace aq;

extern "C" {

// workspace: d [B, H, W] + hs [B, H, W + 1] floats
size_t aq_prvl_workspace_bytes(int B, int H, int W) { return ((size_t)B * H * W + (size_t)B * H * (W + 1)) * sizeof(float); }

// loss[0] = PRVL(img1, img2); state[0] (8 bytes) keeps the packed (value, arg-max) word for the backward
int aq_prvl_loss_fwd(const float* img1, const float* img2, float* loss, void* state, int B, int H, int W, void* ws, size_t ws_bytes,
                     void* stream) {
  AQ_REQUIRE(img1 && img2 && loss && state && B > 0 && H > 0 && W > 0, AQ_ERR_BAD_SHAPE, "prvl_loss_fwd: bad arguments");
  AQ_REQUIRE((long long)B * (H + 1) * (W + 1) < 0xFFFFFFFFll, AQ_ERR_BAD_SHAPE, "prvl_loss_fwd: too many positions for the packed arg-max");
  AQ_REQUIRE(ws != nullptr && ws_bytes >= aq_prvl_workspace_bytes(B, H, W), AQ_ERR_WORKSPACE, "prvl_loss_fwd: workspace too small");
  AQ_REQUIRE((reinterpret_cast<uintptr_t>(state) & 7u) == 0, AQ_ERR_BAD_ALIGN, "prvl_loss_fwd: state must be 8-byte aligned");
  int rc = check_arch();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  float* d = reinterpret_cast<float*>(ws);
  float* hs = d + (size_t)B * H * W;
  const long long hw = (long long)H * W;
  prvl_diff_kernel<<<grid_for(B * hw, 256), 256, 0, st>>>(img1, img2, d, B, hw);
  AQ_LAUNCHED();
  prvl_hsum_kernel<<<grid_for((long long)B * H * (W + 1), 256), 256, 0, st>>>(d, hs, B, H, W);
  AQ_LAUNCHED();
  AQ_CHECK_CUDA(cudaMemsetAsync(state, 0, 8, st));
  prvl_vsum_max_kernel<<<grid_for((long long)B * (H + 1) * (W + 1), 256), 256, 0, st>>>(hs, reinterpret_cast<unsigned long long*>(state), B, H, W);
  AQ_LAUNCHED();
  prvl_finish_kernel<<<1, 1, 0, st>>>(reinterpret_cast<const unsigned long long*>(state), loss);
  AQ_LAUNCHED();
  return AQ_OK;
}

// g_img1 / g_img2 [B, 3, H, W] (either may be NULL) = g_loss[0] * dPRVL/dimg
int aq_prvl_loss_bwd(const float* img1, const float* img2, const void* state, const float* g_loss, float* g_img1, float* g_img2, int B,
                     int H, int W, void* stream) {
  AQ_REQUIRE(img1 && img2 && state && g_loss && (g_img1 || g_img2) && B > 0 && H > 0 && W > 0, AQ_ERR_BAD_SHAPE, "prvl_loss_bwd: bad arguments");
  int rc = check_arch();
  if (rc) return rc;
  prvl_bwd_kernel<<<grid_for((long long)B * 3 * H * W, 256), 256, 0, (cudaStream_t)stream>>>(
      img1, img2, reinterpret_cast<const unsigned long long*>(state), g_loss, g_img1, g_img2, B, H, W);
  AQ_LAUNCHED();
  return AQ_OK;
}

// loss[0] = mean BCE-with-logits; g (optional) = d loss / d logits
int aq_bce_logits(const float* logits, const float* targets, float* loss, float* g_logits, int64_t n, void* stream) {
  AQ_REQUIRE(logits && targets && loss && n > 0, AQ_ERR_BAD_SHAPE, "bce_logits: bad arguments");
  int rc = check_arch();
  if (rc) return rc;
  AQ_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream));
  bce_logits_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(logits, targets, loss, g_logits, n);
  AQ_LAUNCHED();
  return AQ_OK;
}

}  // extern "C"
