// SecretDecoder (utils/models.py:84-96 == evaluation/utils_eval.py:142-154): torchvision EfficientNet-B1 in eval mode with
// a Linear(1280, 2 * bits) head, as fp32 NHWC kernels with every BatchNorm folded into the preceding convolution:
//
//   stem      3x3 s2 conv (3 -> 32) + SiLU                                   NCHW image -> NHWC
//   MBConv    [1x1 expand + SiLU] -> depthwise kxk (+ SiLU, + squeeze sums) -> SE MLP -> 1x1 project (x SE scale, + residual)
//   head      1x1 (320 -> 1280) + SiLU + global average pool fused (the 1280-channel map is never written)
//   fc        Linear(1280 -> 2 * bits), bit = argmax over each logit pair
//
// Arithmetic is plain fp32 FFMA (no TF32): with random-initialised weights the logit margins are tiny, and
// the decoded bits must match the fp32 reference.  HBM-bound by design: activations make one round trip per layer, the
// elementwise work (bias, SiLU, SE scale, residual, pooling) rides in the producing / consuming kernel.
#include "aq_common.h"

namespace aq {

struct StageCfg { int expand, k, stride, cin, cout, layers; };
// torchvision efficientnet_b1: width 1.0, depth 1.1 (oracle/models_oracle.py:B1_STAGES)
static const StageCfg kStages[7] = {
    {1, 3, 1, 32, 16, 2}, {6, 3, 2, 16, 24, 3}, {6, 5, 2, 24, 40, 3}, {6, 3, 2, 40, 80, 4},
    {6, 5, 1, 80, 112, 4}, {6, 5, 2, 112, 192, 5}, {6, 3, 1, 192, 320, 2},
};
constexpr int kStemC = 32, kHeadC = 1280, kLastC = 320, kImg = 512;

static inline size_t pad4(size_t n) { return (n + 3) & ~(size_t)3; }

__device__ __forceinline__ float silu(float v) { return v / (1.f + expf(-v)); }

// ---------------------------------------------------------------------------------------------------------------
// stem: [B, 3, 512, 512] NCHW -> [B, 256, 256, 32] NHWC.  w [27][32] ((ky, kx, ci) major), b [32]
// ---------------------------------------------------------------------------------------------------------------
constexpr int kStemTile = 128;   // output pixels (one row segment) per block

__global__ void __launch_bounds__(kStemTile) stem_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ b, float* __restrict__ y, int H, int W,
                                                           int Ho, int Wo) {
  __shared__ float ws[27 * 32 + 32];
  __shared__ float outs[kStemTile][33];
  for (int i = threadIdx.x; i < 27 * 32 + 32; i += blockDim.x) ws[i] = i < 27 * 32 ? w[i] : b[i - 27 * 32];
  __syncthreads();
  const int ox = blockIdx.x * kStemTile + threadIdx.x, oy = blockIdx.y, n = blockIdx.z;
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = ws[27 * 32 + c];
  if (ox < Wo) {
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
  }
#pragma unroll
  for (int c = 0; c < 32; ++c) outs[threadIdx.x][c] = silu(acc[c]);
  __syncthreads();
  // coalesced NHWC store: the tile is 128 pixels x 32 channels = 4096 contiguous floats
  float* dst = y + (((size_t)n * Ho + oy) * Wo + (size_t)blockIdx.x * kStemTile) * 32;
  const int valid = min(kStemTile, Wo - blockIdx.x * kStemTile) * 32;
  for (int i = threadIdx.x; i < kStemTile * 32; i += blockDim.x)
    if (i < valid) dst[i] = outs[i >> 5][i & 31];
}

// ---------------------------------------------------------------------------------------------------------------
// pointwise (1x1) convolution = row GEMM over NHWC pixels:  Y[m, n] = epi( sum_k (X[m, k] * se[m / hw, k]) * Wt[k, n] + b[n] )
// X [M, K], Wt [K, N] (transposed at pack time), fp32 FFMA, 128 x BN x 16 tiles, 256 threads, 8 x TN register tiles
// ---------------------------------------------------------------------------------------------------------------
constexpr int kPwBM = 128, kPwBK = 16, kPwThreads = 256, kPwAStride = kPwBM + 4;
enum PwEpilogue { kEpiNone = 0, kEpiSilu = 1, kEpiResidual = 2, kEpiSiluPool = 3 };

struct PwArgs {
  const float* x; const float* wt; const float* bias; const float* se;   // se [B, K] or null
  const float* residual;                                                // [M, N] or null
  float* y;                                                             // [M, N]   (kEpiSiluPool: pooled sums [B, N])
  long long M; int K, N, hw, epi;
};

template <int BN>
__global__ void __launch_bounds__(kPwThreads) pointwise_kernel(const PwArgs a) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[kPwBK][kPwAStride];
  __shared__ __align__(16) float Bs[kPwBK][BN];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const long long m0 = (long long)blockIdx.x * kPwBM;
  const int n0 = blockIdx.y * BN;
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // A loader: 128 rows x 4 k-quads = 512 float4 -> 2 per thread;  B loader: 16 x BN/4 float4
  const int a_row[2] = {t >> 2, (t >> 2) + 64};
  const int a_kq = (t & 3) * 4;
  for (int k0 = 0; k0 < a.K; k0 += kPwBK) {
    float4 av[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long m = m0 + a_row[h];
      av[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < a.M && k0 + a_kq < a.K) {
        av[h] = __ldg(reinterpret_cast<const float4*>(a.x + m * a.K + k0 + a_kq));
        if (a.se != nullptr) {
          const float4 s = __ldg(reinterpret_cast<const float4*>(a.se + (m / a.hw) * a.K + k0 + a_kq));
          av[h].x *= s.x; av[h].y *= s.y; av[h].z *= s.z; av[h].w *= s.w;
        }
      }
    }
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    const int b_k = t / (BN / 4), b_n = (t % (BN / 4)) * 4;
    if (t < kPwBK * (BN / 4) && k0 + b_k < a.K && n0 + b_n < a.N)
      bv = __ldg(reinterpret_cast<const float4*>(a.wt + (size_t)(k0 + b_k) * a.N + n0 + b_n));
    __syncthreads();   // previous tile fully consumed
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      As[a_kq + 0][a_row[h]] = av[h].x; As[a_kq + 1][a_row[h]] = av[h].y;
      As[a_kq + 2][a_row[h]] = av[h].z; As[a_kq + 3][a_row[h]] = av[h].w;
    }
    if (t < kPwBK * (BN / 4)) *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = bv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPwBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float ar[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float br[TN];
      if constexpr (TN == 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        br[0] = b4.x; br[1] = b4.y; br[2] = b4.z; br[3] = b4.w;
      } else {
        const float2 b2 = *reinterpret_cast<const float2*>(&Bs[k][tx * 2]);
        br[0] = b2.x; br[1] = b2.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }

  const int nb = n0 + tx * TN;
  if (nb >= a.N) {
    if (a.epi != kEpiSiluPool) return;
  }
  float bias[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) bias[j] = nb + j < a.N ? __ldg(a.bias + nb + j) : 0.f;
  if (a.epi == kEpiSiluPool) {
    // column sums of SiLU(acc + b) over this tile's rows (all in one sample: hw % 128 == 0), then one atomic per column
    __shared__ float red[16][BN];
    float cs[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) cs[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long m = m0 + ty * 8 + i;
      if (m < a.M) {
#pragma unroll
        for (int j = 0; j < TN; ++j) cs[j] += silu(acc[i][j] + bias[j]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < TN; ++j) red[ty][tx * TN + j] = cs[j];
    __syncthreads();
    if (t < BN && n0 + t < a.N) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) s += red[r][t];
      atomicAdd(a.y + (m0 / a.hw) * a.N + n0 + t, s);
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + ty * 8 + i;
    if (m >= a.M) continue;
    float o[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      o[j] = acc[i][j] + bias[j];
      if (a.epi == kEpiSilu) o[j] = silu(o[j]);
    }
    float* dst = a.y + m * a.N + nb;
    if (a.epi == kEpiResidual) {
      const float* rs = a.residual + m * a.N + nb;
#pragma unroll
      for (int j = 0; j < TN; ++j) o[j] += __ldg(rs + j);
    }
    if constexpr (TN == 4) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    else *reinterpret_cast<float2*>(dst) = make_float2(o[0], o[1]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// depthwise k x k stride s (+ folded BN) + SiLU, NHWC, 4 channels per thread; squeeze sums for the SE block
// x [B, H, W, C], w [k*k][C], b [C], y [B, Ho, Wo, C], pooled [B, C] += sum over the tile's pixels
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDwTileW = 16;

template <int KS, int S>
__global__ void __launch_bounds__(256) depthwise_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ b, float* __restrict__ y,
                                                         float* __restrict__ pooled, int H, int W, int C, int Ho, int Wo) {
  constexpr int P = (KS - 1) / 2;
  const int n = blockIdx.z, oy = blockIdx.y, ox0 = blockIdx.x * kDwTileW;
  const int cq_count = C >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x) + (size_t)n * H * W * cq_count;
  const float4* w4 = reinterpret_cast<const float4*>(w);
  float4* y4 = reinterpret_cast<float4*>(y) + ((size_t)n * Ho + oy) * Wo * cq_count;
  for (int cq = threadIdx.x; cq < cq_count; cq += blockDim.x) {
    const float4 bias = __ldg(reinterpret_cast<const float4*>(b) + cq);
    float4 wreg[KS == 3 ? 9 : 1];
    if (KS == 3) {
#pragma unroll
      for (int i = 0; i < 9; ++i) wreg[i] = __ldg(w4 + i * cq_count + cq);
    }
    float4 pool = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int dx = 0; dx < kDwTileW; ++dx) {
      const int ox = ox0 + dx;
      if (ox >= Wo) break;
      float4 acc = bias;
#pragma unroll
      for (int ky = 0; ky < KS; ++ky) {
        const int iy = oy * S - P + ky;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          const int ix = ox * S - P + kx;
          if (ix < 0 || ix >= W) continue;
          const float4 v = __ldg(x4 + ((size_t)iy * W + ix) * cq_count + cq);
          const float4 ww = KS == 3 ? wreg[ky * 3 + kx] : __ldg(w4 + (ky * KS + kx) * cq_count + cq);
          acc.x = fmaf(v.x, ww.x, acc.x); acc.y = fmaf(v.y, ww.y, acc.y);
          acc.z = fmaf(v.z, ww.z, acc.z); acc.w = fmaf(v.w, ww.w, acc.w);
        }
      }
      acc.x = silu(acc.x); acc.y = silu(acc.y); acc.z = silu(acc.z); acc.w = silu(acc.w);
      y4[(size_t)ox * cq_count + cq] = acc;
      pool.x += acc.x; pool.y += acc.y; pool.z += acc.z; pool.w += acc.w;
    }
    float* pd = pooled + (size_t)n * C + cq * 4;
    atomicAdd(pd + 0, pool.x); atomicAdd(pd + 1, pool.y); atomicAdd(pd + 2, pool.z); atomicAdd(pd + 3, pool.w);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// SqueezeExcitation MLP: scale[n, c] = sigmoid(b2[c] + sum_j w2[c, j] * silu(b1[j] + sum_c' w1[j, c'] * mean[n, c']))
// one block per sample; pooled holds SUMS over hw pixels
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) se_kernel(const float* __restrict__ pooled, const float* __restrict__ w1,
                                                  const float* __restrict__ b1, const float* __restrict__ w2,
                                                  const float* __restrict__ b2, float* __restrict__ scale, int C, int SQ,
                                                  float inv_hw) {
  extern __shared__ float sm[];   // mean [C], s1 [SQ]
  float* mean = sm;
  float* s1 = sm + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mean[c] = pooled[(size_t)n * C + c] * inv_hw;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < SQ; j += 8) {
    const float* wr = w1 + (size_t)j * C;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(__ldg(wr + c), mean[c], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s1[j] = silu(acc + b1[j]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* wr = w2 + (size_t)c * SQ;
    float acc = b2[c];
    for (int j = 0; j < SQ; ++j) acc = fmaf(__ldg(wr + j), s1[j], acc);
    scale[(size_t)n * C + c] = 1.f / (1.f + expf(-acc));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// classifier: logits[n, o] = b[o] + sum_c w[o, c] * pooled_sum[n, c] / hw ;  bits[n, i] = argmax(logits[n, 2i], logits[n, 2i+1])
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fc_kernel(const float* __restrict__ pooled, const float* __restrict__ w,
                                                  const float* __restrict__ b, float* __restrict__ logits,
                                                  unsigned char* __restrict__ bits, int C, int O, float inv_hw) {
  extern __shared__ float sm[];   // mean [C], out [O]
  float* mean = sm;
  float* out = sm + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mean[c] = pooled[(size_t)n * C + c] * inv_hw;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < O; o += 8) {
    const float* wr = w + (size_t)o * C;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(__ldg(wr + c), mean[c], acc);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) {
      out[o] = acc + b[o];
      logits[(size_t)n * O + o] = out[o];
    }
  }
  __syncthreads();
  if (bits != nullptr)
    for (int i = threadIdx.x; i < O / 2; i += blockDim.x)
      bits[(size_t)n * (O / 2) + i] = out[2 * i + 1] > out[2 * i] ? 1 : 0;   // torch.argmax: first index wins a tie
}

// ---------------------------------------------------------------------------------------------------------------
// host: packed-parameter walk + launches
// ---------------------------------------------------------------------------------------------------------------
struct Walker {
  const float* base;
  size_t off = 0;
  const float* take(size_t n) {
    const float* p = base ? base + off : nullptr;
    off += pad4(n);
    return p;
  }
};

static size_t packed_floats(int out_features) {
  Walker wk{nullptr};
  wk.take(27 * 32); wk.take(32);
  for (int s = 0; s < 7; ++s) {
    const StageCfg& st = kStages[s];
    for (int l = 0; l < st.layers; ++l) {
      const int cin = l == 0 ? st.cin : st.cout, cexp = cin * st.expand, sq = cin / 4 > 1 ? cin / 4 : 1;
      if (st.expand != 1) { wk.take((size_t)cin * cexp); wk.take(cexp); }
      wk.take((size_t)st.k * st.k * cexp); wk.take(cexp);
      wk.take((size_t)sq * cexp); wk.take(sq); wk.take((size_t)cexp * sq); wk.take(cexp);
      wk.take((size_t)cexp * st.cout); wk.take(st.cout);
    }
  }
  wk.take((size_t)kLastC * kHeadC); wk.take(kHeadC);
  wk.take((size_t)out_features * kHeadC); wk.take(out_features);
  return wk.off;
}

struct Buffers { size_t act, exp, dwo, pooled_total, scale; };   // floats per image

static Buffers buffer_plan() {
  Buffers bf{};
  int hw = (kImg / 2) * (kImg / 2);
  bf.act = (size_t)hw * kStemC;
  int h = kImg / 2;
  size_t pooled = 0, scale = 0;
  for (int s = 0; s < 7; ++s) {
    const StageCfg& st = kStages[s];
    for (int l = 0; l < st.layers; ++l) {
      const int cin = l == 0 ? st.cin : st.cout, cexp = cin * st.expand, stride = l == 0 ? st.stride : 1;
      const int ho = (h + 2 * ((st.k - 1) / 2) - st.k) / stride + 1;
      if (st.expand != 1) bf.exp = bf.exp > (size_t)h * h * cexp ? bf.exp : (size_t)h * h * cexp;
      bf.dwo = bf.dwo > (size_t)ho * ho * cexp ? bf.dwo : (size_t)ho * ho * cexp;
      bf.act = bf.act > (size_t)ho * ho * st.cout ? bf.act : (size_t)ho * ho * st.cout;
      pooled += pad4(cexp);
      scale = scale > (size_t)cexp ? scale : (size_t)cexp;
      h = ho;
    }
  }
  bf.pooled_total = pooled + pad4(kHeadC);
  bf.scale = pad4(scale);
  return bf;
}

static int launch_pointwise(const PwArgs& a, cudaStream_t st) {
  const int bn = a.N <= 32 ? 32 : 64;
  dim3 grid((unsigned)((a.M + kPwBM - 1) / kPwBM), (a.N + bn - 1) / bn);
  if (bn == 32) pointwise_kernel<32><<<grid, kPwThreads, 0, st>>>(a);
  else pointwise_kernel<64><<<grid, kPwThreads, 0, st>>>(a);
  AQ_LAUNCHED();
  return AQ_OK;
}

static int launch_depthwise(const float* x, const float* w, const float* b, float* y, float* pooled, int B, int H, int C, int k,
                            int stride, int Ho, cudaStream_t st) {
  int threads = ((C / 4) + 31) / 32 * 32;
  if (threads > 256) threads = 256;
  dim3 grid((Ho + kDwTileW - 1) / kDwTileW, Ho, B);
  if (k == 3 && stride == 1) depthwise_kernel<3, 1><<<grid, threads, 0, st>>>(x, w, b, y, pooled, H, H, C, Ho, Ho);
  else if (k == 3 && stride == 2) depthwise_kernel<3, 2><<<grid, threads, 0, st>>>(x, w, b, y, pooled, H, H, C, Ho, Ho);
  else if (k == 5 && stride == 1) depthwise_kernel<5, 1><<<grid, threads, 0, st>>>(x, w, b, y, pooled, H, H, C, Ho, Ho);
  else if (k == 5 && stride == 2) depthwise_kernel<5, 2><<<grid, threads, 0, st>>>(x, w, b, y, pooled, H, H, C, Ho, Ho);
  else return fail(AQ_ERR_BAD_SHAPE, "depthwise: unsupported kernel %d stride %d", k, stride);
  AQ_LAUNCHED();
  return AQ_OK;
}

}  // namespace aq

using namespace aq;

extern "C" {

size_t aq_effnetb1_packed_floats(int out_features) { return packed_floats(out_features); }

size_t aq_effnetb1_workspace_bytes(int B) {
  const Buffers bf = buffer_plan();
  const size_t per_image = 2 * pad4(bf.act) + pad4(bf.exp) + pad4(bf.dwo) + bf.pooled_total + bf.scale;
  return (size_t)B * per_image * sizeof(float) + 256;
}

int aq_effnetb1_fwd(const float* x, const float* packed, float* logits, unsigned char* bits, int B, int out_features, void* ws,
                    size_t ws_bytes, void* stream) {
  AQ_REQUIRE(x && packed && logits && B > 0, AQ_ERR_BAD_SHAPE, "effnetb1_fwd: NULL operand or empty batch");
  AQ_REQUIRE(out_features > 0 && out_features % 2 == 0 && out_features <= 512, AQ_ERR_BAD_SHAPE,
             "effnetb1_fwd: out_features=%d must be an even number <= 512 (2 logits per bit)", out_features);
  AQ_REQUIRE(ws && ws_bytes >= aq_effnetb1_workspace_bytes(B), AQ_ERR_WORKSPACE, "effnetb1_fwd: workspace %zu bytes < required %zu",
             ws_bytes, aq_effnetb1_workspace_bytes(B));
  AQ_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(packed) | reinterpret_cast<uintptr_t>(ws)) & 15u) == 0,
             AQ_ERR_BAD_ALIGN, "effnetb1_fwd: x, packed and ws must be 16-byte aligned");
  int rc = check_arch();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const Buffers bf = buffer_plan();
  float* p = reinterpret_cast<float*>(ws);
  float* act[2] = {p, p + (size_t)B * pad4(bf.act)};
  p += 2 * (size_t)B * pad4(bf.act);
  float* expb = p; p += (size_t)B * pad4(bf.exp);
  float* dwo = p; p += (size_t)B * pad4(bf.dwo);
  float* pooled = p; p += (size_t)B * bf.pooled_total;
  float* scale = p;
  AQ_CHECK_CUDA(cudaMemsetAsync(pooled, 0, (size_t)B * bf.pooled_total * sizeof(float), st));

  Walker wk{packed};
  int h = kImg / 2;
  {
    const float* w = wk.take(27 * 32);
    const float* b = wk.take(32);
    dim3 grid((h + kStemTile - 1) / kStemTile, h, B);
    stem_kernel<<<grid, kStemTile, 0, st>>>(x, w, b, act[0], kImg, kImg, h, h);
    AQ_LAUNCHED();
  }
  int cur = 0;
  size_t pooled_off = 0;
  for (int s = 0; s < 7; ++s) {
    const StageCfg& sc = kStages[s];
    for (int l = 0; l < sc.layers; ++l) {
      const int cin = l == 0 ? sc.cin : sc.cout, cexp = cin * sc.expand, stride = l == 0 ? sc.stride : 1;
      const int sq = cin / 4 > 1 ? cin / 4 : 1;
      const int ho = (h + 2 * ((sc.k - 1) / 2) - sc.k) / stride + 1;
      const float* dw_in = act[cur];
      if (sc.expand != 1) {
        PwArgs a{};
        a.x = act[cur]; a.wt = wk.take((size_t)cin * cexp); a.bias = wk.take(cexp); a.se = nullptr; a.residual = nullptr; a.y = expb;
        a.M = (long long)B * h * h; a.K = cin; a.N = cexp; a.hw = h * h; a.epi = kEpiSilu;
        rc = launch_pointwise(a, st);
        if (rc) return rc;
        dw_in = expb;
      }
      float* pl = pooled + (size_t)B * pooled_off;
      pooled_off += pad4(cexp);
      {
        const float* w = wk.take((size_t)sc.k * sc.k * cexp);
        const float* b = wk.take(cexp);
        rc = launch_depthwise(dw_in, w, b, dwo, pl, B, h, cexp, sc.k, stride, ho, st);
        if (rc) return rc;
      }
      {
        const float* w1 = wk.take((size_t)sq * cexp);
        const float* b1 = wk.take(sq);
        const float* w2 = wk.take((size_t)cexp * sq);
        const float* b2 = wk.take(cexp);
        se_kernel<<<B, 256, (cexp + sq) * sizeof(float), st>>>(pl, w1, b1, w2, b2, scale, cexp, sq, 1.f / (float)(ho * ho));
        AQ_LAUNCHED();
      }
      {
        PwArgs a{};
        a.x = dwo; a.wt = wk.take((size_t)cexp * sc.cout); a.bias = wk.take(sc.cout); a.se = scale;
        const bool res = stride == 1 && cin == sc.cout;
        a.residual = res ? act[cur] : nullptr; a.y = act[cur ^ 1];
        a.M = (long long)B * ho * ho; a.K = cexp; a.N = sc.cout; a.hw = ho * ho; a.epi = res ? kEpiResidual : kEpiNone;
        rc = launch_pointwise(a, st);
        if (rc) return rc;
      }
      cur ^= 1;
      h = ho;
    }
  }
  float* head_pool = pooled + (size_t)B * pooled_off;
  {
    PwArgs a{};
    a.x = act[cur]; a.wt = wk.take((size_t)kLastC * kHeadC); a.bias = wk.take(kHeadC); a.se = nullptr; a.residual = nullptr; a.y = head_pool;
    a.M = (long long)B * h * h; a.K = kLastC; a.N = kHeadC; a.hw = h * h; a.epi = kEpiSiluPool;
    AQ_REQUIRE((h * h) % kPwBM == 0, AQ_ERR_BAD_SHAPE, "effnetb1_fwd: head map %d x %d is not a multiple of the row tile", h, h);
    rc = launch_pointwise(a, st);
    if (rc) return rc;
  }
  {
    const float* w = wk.take((size_t)out_features * kHeadC);
    const float* b = wk.take(out_features);
    fc_kernel<<<B, 256, (kHeadC + out_features) * sizeof(float), st>>>(head_pool, w, b, logits, bits, kHeadC, out_features, 1.f / (float)(h * h));
    AQ_LAUNCHED();
  }
  return AQ_OK;
}

}  // extern "C"
