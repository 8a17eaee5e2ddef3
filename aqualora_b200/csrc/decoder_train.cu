// Convolutions of the message decoder's TRAIN path (train/latent_wm_pretrain.py:190,216; rob_enhance_finetune.py:1020-1034): what the
// eval chain of decoder.cu cannot serve because BatchNorm is not folded and every convolution needs its input / weight gradients.
// fp32, NHWC activations, weights in PyTorch's own layouts (no per-step repacking):
//
//   depthwise k x k, stride s      forward (raw: BN follows), input gradient (transposed stencil), weight gradient (per-channel
//                                  correlation over all pixels, fp32 atomics across pixel chunks)
//   pointwise 1 x 1                forward and input gradient run on the tensor cores through decoder_pw.cu (3 x TF32, fp32-faithful);
//                                  the weight gradient gW[n, k] = sum_m gz[m, n] x[m, k] (x optionally scaled by the SE gate of its image)
//                                  is the FFMA outer-product kernel below: tiny output, long reduction over the pixels
//   stem 3 x 3 stride 2 (3 -> 32)  forward from the NCHW image, input gradient back to the NCHW image, weight gradient
//
// All are HBM / L2 bound stencils except the pointwise weight gradient (2.6 GMAC per image on CUDA cores, ~0.3 ms per image).
#include "aq_common.h"

namespace aq {

static int tr_grid(long long work, int block) {
  const int sms = sm_count() > 0 ? sm_count() : 148;
  long long blocks = (work + block - 1) / block;
  const long long cap = (long long)sms * 32;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

// ---------------------------------------------------------------------------------------------------------------
// depthwise: x [B, H, W, C], w [C, k, k] (PyTorch [C, 1, k, k]), z [B, Ho, Wo, C]
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dw_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ z, int B,
                                                      int H, int W, int C, int Ho, int Wo, int k, int s) {
  const int C4 = C >> 2, p = (k - 1) / 2, kk = k * k;
  const long long n = (long long)B * Ho * Wo * C4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % C4);
    long long r = idx / C4;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const float* wc = w + (size_t)c4 * 4 * kk;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ky = 0; ky < k; ++ky) {
      const int iy = oy * s + ky - p;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ix = ox * s + kx - p;
        if (ix < 0 || ix >= W) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + (((size_t)b * H + iy) * W + ix) * C) + c4);
        const int t = ky * k + kx;
        acc[0] = fmaf(v.x, __ldg(wc + t), acc[0]);
        acc[1] = fmaf(v.y, __ldg(wc + kk + t), acc[1]);
        acc[2] = fmaf(v.z, __ldg(wc + 2 * kk + t), acc[2]);
        acc[3] = fmaf(v.w, __ldg(wc + 3 * kk + t), acc[3]);
      }
    }
    reinterpret_cast<float4*>(z)[idx] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// gx[b, iy, ix, c] = sum over taps with (iy + p - ky) = s * oy, (ix + p - kx) = s * ox of gz[b, oy, ox, c] w[c, ky, kx]
__global__ void __launch_bounds__(256) dw_dgrad_kernel(const float* __restrict__ gz, const float* __restrict__ w, float* __restrict__ gx, int B,
                                                        int H, int W, int C, int Ho, int Wo, int k, int s) {
  const int C4 = C >> 2, p = (k - 1) / 2, kk = k * k;
  const long long n = (long long)B * H * W * C4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % C4);
    long long r = idx / C4;
    const int ix = (int)(r % W); r /= W;
    const int iy = (int)(r % H);
    const int b = (int)(r / H);
    const float* wc = w + (size_t)c4 * 4 * kk;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ky = 0; ky < k; ++ky) {
      const int ty = iy + p - ky;
      if (ty < 0 || ty % s != 0) continue;
      const int oy = ty / s;
      if (oy >= Ho) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int tx = ix + p - kx;
        if (tx < 0 || tx % s != 0) continue;
        const int ox = tx / s;
        if (ox >= Wo) continue;
        const float4 g = __ldg(reinterpret_cast<const float4*>(gz + (((size_t)b * Ho + oy) * Wo + ox) * C) + c4);
        const int t = ky * k + kx;
        acc[0] = fmaf(g.x, __ldg(wc + t), acc[0]);
        acc[1] = fmaf(g.y, __ldg(wc + kk + t), acc[1]);
        acc[2] = fmaf(g.z, __ldg(wc + 2 * kk + t), acc[2]);
        acc[3] = fmaf(g.w, __ldg(wc + 3 * kk + t), acc[3]);
      }
    }
    reinterpret_cast<float4*>(gx)[idx] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// gw[c, ky, kx] += sum_{b, oy, ox} gz[b, oy, ox, c] x[b, oy s + ky - p, ox s + kx - p, c]
// block (32 channel vectors, k tap rows); grid (ceil(C4 / 32), pixel chunks); the k threads of a channel vector read the same gz
__global__ void dw_wgrad_kernel(const float* __restrict__ gz, const float* __restrict__ x, float* __restrict__ gw, int B, int H, int W, int C,
                                int Ho, int Wo, int k, int s, long long pix_per_block) {
  const int C4 = C >> 2, p = (k - 1) / 2, kk = k * k;
  const int c4 = blockIdx.x * 32 + threadIdx.x;
  const int ky = threadIdx.y;
  if (c4 >= C4) return;
  const long long npix = (long long)B * Ho * Wo;
  const long long pb = (long long)blockIdx.y * pix_per_block;
  const long long pe = pb + pix_per_block < npix ? pb + pix_per_block : npix;
  float acc[5][4];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[i][v] = 0.f;
  for (long long q = pb; q < pe; ++q) {
    const int ox = (int)(q % Wo);
    const int oy = (int)((q / Wo) % Ho);
    const int b = (int)(q / ((long long)Wo * Ho));
    const int iy = oy * s + ky - p;
    if (iy < 0 || iy >= H) continue;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gz + (size_t)q * C) + c4);
    const float* xrow = x + (((size_t)b * H + iy) * W) * C;
#pragma unroll
    for (int kx = 0; kx < 5; ++kx) {
      if (kx >= k) break;
      const int ix = ox * s + kx - p;
      if (ix < 0 || ix >= W) continue;
      const float4 v = __ldg(reinterpret_cast<const float4*>(xrow + (size_t)ix * C) + c4);
      acc[kx][0] = fmaf(g.x, v.x, acc[kx][0]);
      acc[kx][1] = fmaf(g.y, v.y, acc[kx][1]);
      acc[kx][2] = fmaf(g.z, v.z, acc[kx][2]);
      acc[kx][3] = fmaf(g.w, v.w, acc[kx][3]);
    }
  }
#pragma unroll
  for (int kx = 0; kx < 5; ++kx) {
    if (kx >= k) break;
#pragma unroll
    for (int v = 0; v < 4; ++v) atomicAdd(gw + (size_t)(c4 * 4 + v) * kk + ky * k + kx, acc[kx][v]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// pointwise weight gradient: gw[n, k] += sum_m gz[m, n] * x[m, k] * (se ? se[m / hw, k] : 1)
// 64 x 64 output tile per block, 256 threads x (4 x 4), rows streamed through shared memory 32 at a time; grid.z splits the rows
// ---------------------------------------------------------------------------------------------------------------
constexpr int kPgTile = 64, kPgRows = 32;

__global__ void __launch_bounds__(256) pw_wgrad_kernel(const float* __restrict__ gz, const float* __restrict__ x, const float* __restrict__ se,
                                                        float* __restrict__ gw, long long M, int N, int K, int hw, long long rows_per_block) {
  __shared__ float As[kPgRows][kPgTile + 4];
  __shared__ float Bs[kPgRows][kPgTile + 4];
  const int k0 = blockIdx.x * kPgTile, n0 = blockIdx.y * kPgTile;
  const long long mb = (long long)blockIdx.z * rows_per_block;
  const long long me = mb + rows_per_block < M ? mb + rows_per_block : M;
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (long long m0 = mb; m0 < me; m0 += kPgRows) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int e = threadIdx.x + i * 256;          // 512 float4 per operand tile
      const int r = e / 16, c = (e % 16) * 4;
      const long long m = m0 + r;
      float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < me) {
        if (n0 + c < N) a4 = __ldg(reinterpret_cast<const float4*>(gz + (size_t)m * N + n0 + c));
        if (k0 + c < K) {
          b4 = __ldg(reinterpret_cast<const float4*>(x + (size_t)m * K + k0 + c));
          if (se != nullptr) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(se + (size_t)(m / hw) * K + k0 + c));
            b4.x *= s4.x; b4.y *= s4.y; b4.z *= s4.z; b4.w *= s4.w;
          }
        }
      }
      *reinterpret_cast<float4*>(&As[r][c]) = a4;
      *reinterpret_cast<float4*>(&Bs[r][c]) = b4;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < kPgRows; ++r) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[r][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[r][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int n = n0 + ty * 4 + a;
    if (n >= N) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int k = k0 + tx * 4 + b;
      if (k < K) atomicAdd(gw + (size_t)n * K + k, acc[a][b]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// stem: x [B, 3, H, W] NCHW, w [32, 3, 3, 3] (PyTorch), z [B, Ho, Wo, 32] NHWC; 3 x 3, stride 2, padding 1
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) stem_fwd_raw_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ z,
                                                            int H, int W, int Ho, int Wo) {
  __shared__ float ws[27 * 32];     // [(ky, kx, ci)][co]
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) {
    const int co = i % 32, t = i / 32;
    const int ci = t % 3, kx = (t / 3) % 3, ky = t / 9;
    ws[i] = w[((co * 3 + ci) * 3 + ky) * 3 + kx];
  }
  __syncthreads();
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y, n = blockIdx.z;
  if (ox >= Wo) return;
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;
  const float* xn = x + (size_t)n * 3 * H * W;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 - 1 + ky;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 - 1 + kx;
      if (ix < 0 || ix >= W) continue;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float v = __ldg(xn + ((size_t)ci * H + iy) * W + ix);
        const float* wr = ws + ((ky * 3 + kx) * 3 + ci) * 32;
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = fmaf(v, wr[c], acc[c]);
      }
    }
  }
  float4* dst = reinterpret_cast<float4*>(z + (((size_t)n * Ho + oy) * Wo + ox) * 32);
#pragma unroll
  for (int c = 0; c < 8; ++c) dst[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
}

// gx[b, ci, iy, ix] = sum over (ky, kx) with iy + 1 - ky = 2 oy, ix + 1 - kx = 2 ox of sum_co gz[b, oy, ox, co] w[co, ci, ky, kx]
__global__ void __launch_bounds__(128) stem_dgrad_kernel(const float* __restrict__ gz, const float* __restrict__ w, float* __restrict__ gx,
                                                          int H, int W, int Ho, int Wo) {
  __shared__ float ws[27 * 32];     // [(ky, kx, ci)][co]
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) {
    const int co = i % 32, t = i / 32;
    const int ci = t % 3, kx = (t / 3) % 3, ky = t / 9;
    ws[i] = w[((co * 3 + ci) * 3 + ky) * 3 + kx];
  }
  __syncthreads();
  const int ix = blockIdx.x * blockDim.x + threadIdx.x, iy = blockIdx.y, n = blockIdx.z;
  if (ix >= W) return;
  float acc[3] = {0.f, 0.f, 0.f};
  for (int ky = 0; ky < 3; ++ky) {
    const int ty = iy + 1 - ky;
    if (ty < 0 || (ty & 1)) continue;
    const int oy = ty >> 1;
    if (oy >= Ho) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int tx = ix + 1 - kx;
      if (tx < 0 || (tx & 1)) continue;
      const int ox = tx >> 1;
      if (ox >= Wo) continue;
      const float4* g = reinterpret_cast<const float4*>(gz + (((size_t)n * Ho + oy) * Wo + ox) * 32);
      const float* w0 = ws + ((ky * 3 + kx) * 3) * 32;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 gv = __ldg(g + c);
        const float gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[0] = fmaf(gs[j], w0[c * 4 + j], acc[0]);
          acc[1] = fmaf(gs[j], w0[32 + c * 4 + j], acc[1]);
          acc[2] = fmaf(gs[j], w0[64 + c * 4 + j], acc[2]);
        }
      }
    }
  }
  float* dst = gx + (size_t)n * 3 * H * W + (size_t)iy * W + ix;
  dst[0] = acc[0]; dst[(size_t)H * W] = acc[1]; dst[2 * (size_t)H * W] = acc[2];
}

// gw[co, ci, ky, kx] += sum over pixels gz[b, oy, ox, co] x[b, ci, 2 oy + ky - 1, 2 ox + kx - 1]
// block 256 = (32 output channels, 8 pixel lanes); every thread keeps the 27 taps of its channel
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ gz, const float* __restrict__ x, float* __restrict__ gw, int B,
                                                          int H, int W, int Ho, int Wo, long long pix_per_block) {
  __shared__ float red[32 * 27];
  for (int i = threadIdx.x; i < 32 * 27; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int co = threadIdx.x % 32, lane8 = threadIdx.x / 32;
  const long long npix = (long long)B * Ho * Wo;
  const long long pb = (long long)blockIdx.x * pix_per_block;
  const long long pe = pb + pix_per_block < npix ? pb + pix_per_block : npix;
  float acc[27];
#pragma unroll
  for (int t = 0; t < 27; ++t) acc[t] = 0.f;
  for (long long q = pb + lane8; q < pe; q += 8) {
    const int ox = (int)(q % Wo);
    const int oy = (int)((q / Wo) % Ho);
    const int b = (int)(q / ((long long)Wo * Ho));
    const float g = __ldg(gz + (size_t)q * 32 + co);
    const float* xn = x + (size_t)b * 3 * H * W;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * 2 - 1 + ky;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = ox * 2 - 1 + kx;
          const float v = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(xn + ((size_t)ci * H + iy) * W + ix) : 0.f;
          acc[(ci * 3 + ky) * 3 + kx] = fmaf(g, v, acc[(ci * 3 + ky) * 3 + kx]);
        }
      }
  }
#pragma unroll
  for (int t = 0; t < 27; ++t) atomicAdd(&red[co * 27 + t], acc[t]);
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 27; i += blockDim.x) atomicAdd(gw + i, red[i]);     // gw [32][3][3][3] = [co][27]
}

}  // namespace aq

using namespace aq;

extern "C" {

int aq_dwconv_fwd(const float* x, const float* w, float* z, int B, int H, int W, int C, int k, int stride, void* stream) {
  AQ_REQUIRE(x && w && z && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, AQ_ERR_BAD_SHAPE, "dwconv_fwd: bad arguments (C %% 4 == 0 required)");
  AQ_REQUIRE((k == 3 || k == 5) && (stride == 1 || stride == 2), AQ_ERR_BAD_SHAPE, "dwconv_fwd: k=%d stride=%d unsupported", k, stride);
  int rc = check_arch();
  if (rc) return rc;
  const int p = (k - 1) / 2, Ho = (H + 2 * p - k) / stride + 1, Wo = (W + 2 * p - k) / stride + 1;
  const long long n = (long long)B * Ho * Wo * (C / 4);
  dw_fwd_kernel<<<tr_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(x, w, z, B, H, W, C, Ho, Wo, k, stride);
  AQ_LAUNCHED();
  return AQ_OK;
}

// gx [B, H, W, C] fully written; gw [C, k, k] ACCUMULATED into (either may be NULL)
int aq_dwconv_bwd(const float* gz, const float* x, const float* w, float* gx, float* gw, int B, int H, int W, int C, int k, int stride,
                  void* stream) {
  AQ_REQUIRE(gz && x && w && (gx || gw) && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, AQ_ERR_BAD_SHAPE, "dwconv_bwd: bad arguments");
  AQ_REQUIRE((k == 3 || k == 5) && (stride == 1 || stride == 2), AQ_ERR_BAD_SHAPE, "dwconv_bwd: k=%d stride=%d unsupported", k, stride);
  int rc = check_arch();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int p = (k - 1) / 2, Ho = (H + 2 * p - k) / stride + 1, Wo = (W + 2 * p - k) / stride + 1;
  if (gx != nullptr) {
    const long long n = (long long)B * H * W * (C / 4);
    dw_dgrad_kernel<<<tr_grid(n, 256), 256, 0, st>>>(gz, w, gx, B, H, W, C, Ho, Wo, k, stride);
    AQ_LAUNCHED();
  }
  if (gw != nullptr) {
    const int cgroups = (C / 4 + 31) / 32;
    const long long npix = (long long)B * Ho * Wo;
    const int sms = sm_count() > 0 ? sm_count() : 148;
    long long chunks = (8LL * sms + cgroups - 1) / cgroups;
    if (chunks > (npix + 63) / 64) chunks = (npix + 63) / 64;
    if (chunks < 1) chunks = 1;
    const long long per = (npix + chunks - 1) / chunks;
    dim3 grid(cgroups, (unsigned)((npix + per - 1) / per));
    dw_wgrad_kernel<<<grid, dim3(32, k), 0, st>>>(gz, x, gw, B, H, W, C, Ho, Wo, k, stride, per);
    AQ_LAUNCHED();
  }
  return AQ_OK;
}

// gw [N, K] += gz^T [N, M] (x (.) se) [M, K]; se [M / hw, K] or NULL
int aq_conv1x1_wgrad(const float* gz, const float* x, const float* se, float* gw, int64_t M, int K, int N, int hw, void* stream) {
  AQ_REQUIRE(gz && x && gw && M > 0 && K > 0 && N > 0, AQ_ERR_BAD_SHAPE, "conv1x1_wgrad: bad arguments");
  AQ_REQUIRE(K % 4 == 0 && N % 4 == 0 && (se == nullptr || hw > 0), AQ_ERR_BAD_SHAPE, "conv1x1_wgrad: K=%d and N=%d must be multiples of 4", K, N);
  int rc = check_arch();
  if (rc) return rc;
  const int tk = (K + kPgTile - 1) / kPgTile, tn = (N + kPgTile - 1) / kPgTile;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  long long splits = (4LL * sms + (long long)tk * tn - 1) / ((long long)tk * tn);
  const long long max_splits = (M + 4 * kPgRows - 1) / (4 * kPgRows);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  long long per = (M + splits - 1) / splits;
  per = (per + kPgRows - 1) / kPgRows * kPgRows;
  dim3 grid(tk, tn, (unsigned)((M + per - 1) / per));
  pw_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gz, x, se, gw, M, N, K, hw > 0 ? hw : 1, per);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_stem_conv_fwd(const float* x, const float* w, float* z, int B, int H, int W, void* stream) {
  AQ_REQUIRE(x && w && z && B > 0 && H > 0 && W > 0, AQ_ERR_BAD_SHAPE, "stem_conv_fwd: bad arguments");
  AQ_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15u) == 0, AQ_ERR_BAD_ALIGN, "stem_conv_fwd: z must be 16-byte aligned");
  int rc = check_arch();
  if (rc) return rc;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  dim3 grid((Wo + 127) / 128, Ho, B);
  stem_fwd_raw_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, w, z, H, W, Ho, Wo);
  AQ_LAUNCHED();
  return AQ_OK;
}

// gx [B, 3, H, W] fully written (may be NULL); gw [32, 3, 3, 3] ACCUMULATED into (may be NULL)
int aq_stem_conv_bwd(const float* gz, const float* x, const float* w, float* gx, float* gw, int B, int H, int W, void* stream) {
  AQ_REQUIRE(gz && x && w && (gx || gw) && B > 0 && H > 0 && W > 0, AQ_ERR_BAD_SHAPE, "stem_conv_bwd: bad arguments");
  int rc = check_arch();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  if (gx != nullptr) {
    dim3 grid((W + 127) / 128, H, B);
    stem_dgrad_kernel<<<grid, 128, 0, st>>>(gz, w, gx, H, W, Ho, Wo);
    AQ_LAUNCHED();
  }
  if (gw != nullptr) {
    const long long npix = (long long)B * Ho * Wo;
    const int sms = sm_count() > 0 ? sm_count() : 148;
    long long blocks = 4LL * sms;
    if (blocks > (npix + 255) / 256) blocks = (npix + 255) / 256;
    if (blocks < 1) blocks = 1;
    const long long per = (npix + blocks - 1) / blocks;
    stem_wgrad_kernel<<<(unsigned)((npix + per - 1) / per), 256, 0, st>>>(gz, x, gw, B, H, W, Ho, Wo, per);
    AQ_LAUNCHED();
  }
  return AQ_OK;
}

}  // extern "C"
