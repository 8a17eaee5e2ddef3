// SecretDecoder (utils/models.py:84-96 == evaluation/utils_eval.py:142-154): torchvision EfficientNet-B1 in eval mode with
// a Linear(1280, 2 * bits) head, as fp32 NHWC kernels with every BatchNorm folded into the preceding convolution:
//
//   stem      3x3 s2 conv (3 -> 32) + SiLU                                   NCHW image -> NHWC
//   MBConv    [1x1 expand + SiLU] -> depthwise kxk (+ SiLU, + squeeze sums) -> SE MLP -> 1x1 project (x SE scale, + residual)
//   head      1x1 (320 -> 1280) + SiLU + global average pool fused (the 1280-channel map is never written)
//   fc        Linear(1280 -> 2 * bits), bit = argmax over each logit pair
//
// Arithmetic is fp32-faithful: with random-initialised weights the logit margins are tiny and the decoded bits
// must match the fp32 reference.  The pointwise convolutions (89 % of the FLOPs) run on the tensor cores as 3-term split-TF32
// products with fp32 accumulation (decoder_pw.cu); stem, depthwise, SE and classifier are fp32 FFMA.  HBM-bound by design:
// activations make one round trip per layer, the elementwise work (bias, SiLU, SE scale, residual, pooling) rides in the
// producing / consuming kernel.
#include <stdlib.h>
#include <string.h>

#include "aq_common.h"
#include "aq_ptx.cuh"
#include "decoder_fused.h"
#include "decoder_pw.h"

namespace aq {

struct StageCfg { int expand, k, stride, cin, cout, layers; };
// torchvision efficientnet_b1: width 1.0, depth 1.1 (oracle/models_oracle.py:B1_STAGES)
static const StageCfg kStages[7] = {
    {1, 3, 1, 32, 16, 2}, {6, 3, 2, 16, 24, 3}, {6, 5, 2, 24, 40, 3}, {6, 3, 2, 40, 80, 4},
    {6, 5, 1, 80, 112, 4}, {6, 5, 2, 112, 192, 5}, {6, 3, 1, 192, 320, 2},
};
constexpr int kStemC = 32, kHeadC = 1280, kLastC = 320, kImg = 512;

static inline size_t pad4(size_t n) { return (n + 3) & ~(size_t)3; }

__device__ __forceinline__ float silu(float v) { return v / (1.f + expf(-v)); }   // exact: SE MLP (tiny tensors)
// SFU exponential + reciprocal (relative error ~1e-6): the activation maps, where the exact form cost more than the convolution
__device__ __forceinline__ float silu_fast(float v) {
  float e, r;   // ex2.approx.ftz saturates cleanly at both ends; see decoder_pw.cu:pw_silu
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return v * r;
}

// ---------------------------------------------------------------------------------------------------------------
// stem: [B, 3, 512, 512] NCHW -> [B, 256, 256, 32] NHWC.  w [27][32] ((ky, kx, ci) major), b [32]
// ---------------------------------------------------------------------------------------------------------------
// One thread = FOUR adjacent output pixels x 16 channels (warps 0-1: channels 0-15, warps 2-3: channels 16-31).  The kernel is bound
// by the LSU pipe, not by FMA issue: every weight vector is a warp-wide LDS.128 broadcast that still occupies the pipe for 4 cycles
// (32 lanes x 16 bytes of register write-back).  One pixel x 32 channels per thread: 216 LDS per pixel (LSU and issue both at 75 %,
// ncu profiles/r02_ncu_decoder_kernels_v10.txt); two pixels: 108 per pixel (394 -> 289 us per 64 images); four pixels x 16
// channels: 27 LDS per pixel and channel half = 54 per pixel, the same 64 accumulators per thread.  The 3 x 9 input window per
// channel is two aligned float4 + one scalar per row (both channel halves load it: L1 hits).
constexpr int kStemThreads = 128;
constexpr int kStemPx = 4;                                  // output pixels per thread
constexpr int kStemTile = kStemPx * (kStemThreads / 2);     // 256 output pixels (one row segment) per block
constexpr int kStemInW = 2 * kStemTile;                     // 512 input columns per block (+ 1 halo column on the left)
constexpr int kStemRow = kStemInW + 8;                      // staged input row: [3] = halo column, [4 ... 4 + 512) = the segment

// The 9 input rows (3 ky x 3 channels) of the block are staged in shared memory first: each thread issues its 9 independent 16-byte
// loads back to back (one exposed DRAM latency per block).  Loading a row right before the FMA block that consumes it -- the form
// ptxas produces when the loads sit in the ky / ci loops behind the bounds branches -- exposed that latency nine times per thread
// (ncu: long-scoreboard 3.6 warps per issue, 41 % issue utilisation at 16 warps per SM).
__global__ void __launch_bounds__(kStemThreads) stem_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ b, float* __restrict__ y, int H, int W,
                                                              int Ho, int Wo) {
  __shared__ __align__(16) float ws[27 * 32 + 32];
  // output tile [256 pixels][8 chunks of 4 channels], chunk c of pixel px stored at slot c ^ ((px >> 2) & 7): the 8 lanes of a
  // quarter warp own pixels 4 lanes apart (128-byte rows, same banks) and land in 8 different slots.  The staged input rows
  // (9 x 520 floats) live in the same 32 KiB: they are dead once every thread has finished its FMAs (barrier below).
  __shared__ __align__(16) float outs[kStemTile * 32];
  static_assert(9 * kStemRow <= kStemTile * 32, "the staged input must fit in the output tile");
  float* xs = outs;
  const int oy = blockIdx.y, n = blockIdx.z;
  const int ixb = blockIdx.x * kStemInW;                    // first input column of the segment (a multiple of 512)
  {
    // W % 4 == 0 (host check): a float4 either lies inside the row or starts past it.  All nine loads are predicated, not branched,
    // and leave back to back; the halo column is ONE load of the threads 0 ... 8 (inside the row loop it shared a destination
    // register across the unrolled iterations and serialised the nine loads: 38 % of the kernel's stall samples)
    const float* xn = x + (size_t)n * 3 * H * W;
    const int ix = ixb + 4 * (int)threadIdx.x;
    float4 q[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {                           // r = ky * 3 + ci
      const int iy = oy * 2 - 1 + r / 3;
      const bool ok = iy >= 0 && iy < H && ix < W;
      const float* row = xn + ((size_t)(r % 3) * H + (ok ? iy : 0)) * W + (ok ? ix : 0);
      q[r] = __ldg(reinterpret_cast<const float4*>(row));
      if (!ok) q[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float hv = 0.f;
    if (threadIdx.x < 9) {
      const int r = threadIdx.x;
      const int iy = oy * 2 - 1 + r / 3;
      if (iy >= 0 && iy < H && ixb > 0) hv = __ldg(xn + ((size_t)(r % 3) * H + iy) * W + ixb - 1);
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) *reinterpret_cast<float4*>(&xs[r * kStemRow + 4 + 4 * threadIdx.x]) = q[r];
    if (threadIdx.x < 9) xs[threadIdx.x * kStemRow + 3] = hv;
  }
  for (int i = threadIdx.x; i < 27 * 32 + 32; i += blockDim.x) ws[i] = i < 27 * 32 ? w[i] : b[i - 27 * 32];
  __syncthreads();
  const int half = threadIdx.x / (kStemThreads / 2);        // warp-uniform
  const int pg = threadIdx.x % (kStemThreads / 2);
  float acc[kStemPx][16];
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const float bv = ws[27 * 32 + half * 16 + c];
#pragma unroll
    for (int px = 0; px < kStemPx; ++px) acc[px][c] = bv;
  }
  // rows above / below the image and columns past W are staged as zeros: their taps add 0 * w (exact), like the padding they are
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    const float* xr = xs + r * kStemRow + 4 + 8 * pg;       // input columns 8 pg - 1 ... 8 pg + 7 of the segment
    float v[2 * kStemPx + 1];
    v[0] = xr[-1];
    const float4 q0 = *reinterpret_cast<const float4*>(xr);
    const float4 q1 = *reinterpret_cast<const float4*>(xr + 4);
    v[1] = q0.x; v[2] = q0.y; v[3] = q0.z; v[4] = q0.w;
    v[5] = q1.x; v[6] = q1.y; v[7] = q1.z; v[8] = q1.w;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const float4* wr = reinterpret_cast<const float4*>(ws + (((r / 3) * 3 + kx) * 3 + (r % 3)) * 32 + half * 16);
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float4 wv = wr[c4];
        // packed FFMA2: 864 fused multiply-adds per pixel in 432 issue slots
#pragma unroll
        for (int px = 0; px < kStemPx; ++px) {
          ffma2(acc[px][4 * c4 + 0], acc[px][4 * c4 + 1], v[2 * px + kx], v[2 * px + kx], wv.x, wv.y);
          ffma2(acc[px][4 * c4 + 2], acc[px][4 * c4 + 3], v[2 * px + kx], v[2 * px + kx], wv.z, wv.w);
        }
      }
    }
  }
#pragma unroll
  for (int px = 0; px < kStemPx; ++px) {
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      silu2(acc[px][4 * c4], acc[px][4 * c4 + 1]);
      silu2(acc[px][4 * c4 + 2], acc[px][4 * c4 + 3]);
    }
  }
  __syncthreads();   // every thread is done reading the staged input: the buffer becomes the output tile
#pragma unroll
  for (int px = 0; px < kStemPx; ++px) {
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      const int slot = (half * 4 + c4) ^ (pg & 7);
      *reinterpret_cast<float4*>(&outs[(kStemPx * pg + px) * 32 + slot * 4]) =
          make_float4(acc[px][4 * c4], acc[px][4 * c4 + 1], acc[px][4 * c4 + 2], acc[px][4 * c4 + 3]);
    }
  }
  __syncthreads();
  // coalesced NHWC store: the tile is 256 pixels x 32 channels = 2048 contiguous float4 (slot i & 7 of pixel i >> 3 holds chunk
  // (i & 7) ^ ((i >> 5) & 7); the 8 lanes of a pixel still write its whole 128-byte line)
  float4* dst = reinterpret_cast<float4*>(y + (((size_t)n * Ho + oy) * Wo + (size_t)blockIdx.x * kStemTile) * 32);
  const int valid4 = min(kStemTile, Wo - blockIdx.x * kStemTile) * 8;
#pragma unroll 4
  for (int i = threadIdx.x; i < kStemTile * 8; i += kStemThreads)
    if (i < valid4) dst[(i & ~7) | ((i & 7) ^ ((i >> 5) & 7))] = *reinterpret_cast<const float4*>(&outs[i * 4]);
}

// ---------------------------------------------------------------------------------------------------------------
// depthwise k x k stride s (+ folded BN) + SiLU, NHWC; squeeze sums for the SE block
// x [B, H, W, C], w [k*k][C], b [C], y [B, Ho, Wo, C], pooled [B, C] += sum over pixels
// One thread = V channels x R output rows, sliding along x over a segment of TW outputs with the (R-1)*S+k by k input window
// and the k*k weights in registers: each step loads only the S new input columns (consecutive threads = consecutive channel
// vectors -> coalesced), so an input element is fetched (R-1+k)/(R*S) times (L1 / L2 hits) instead of k*k/S^2 times.
// ---------------------------------------------------------------------------------------------------------------
template <int V>
__device__ __forceinline__ void ldv(const float* p, float (&d)[V]) {
  if constexpr (V == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w;
  } else {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    d[0] = t.x; d[1] = t.y;
  }
}
template <int V>
__device__ __forceinline__ void ldsv(const float* p, float (&d)[V]) {   // plain (shared-memory) vector load
  if constexpr (V == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w;
  } else {
    const float2 t = *reinterpret_cast<const float2*>(p);
    d[0] = t.x; d[1] = t.y;
  }
}
template <int V>
__device__ __forceinline__ void stv(float* p, const float (&d)[V]) {
  if constexpr (V == 4) *reinterpret_cast<float4*>(p) = make_float4(d[0], d[1], d[2], d[3]);
  else *reinterpret_cast<float2*>(p) = make_float2(d[0], d[1]);
}

constexpr int kDwThreads = 128;

template <int KS, int S, int R, int V>
__global__ void __launch_bounds__(kDwThreads, 3) depthwise_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                   const float* __restrict__ b, float* __restrict__ y,
                                                                   float* __restrict__ pooled, int B, int H, int W, int C, int Ho, int Wo,
                                                                   int TW) {
  constexpr int P = (KS - 1) / 2, NR = (R - 1) * S + KS;
  constexpr int kBad = -(1 << 30);   // offset of an out-of-image row / column: the sum of two offsets stays negative
  const unsigned CV = (unsigned)(C / V);
  const unsigned nseg = (unsigned)((Wo + TW - 1) / TW), nrg = (unsigned)((Ho + R - 1) / R);
  unsigned idx = blockIdx.x * kDwThreads + threadIdx.x;   // < 2^32: B * Ho / R * Wo / TW * C / V threads
  const unsigned cv = idx % CV; idx /= CV;
  const unsigned seg = idx % nseg; idx /= nseg;
  const unsigned rg = idx % nrg;
  const unsigned n = idx / nrg;
  if (n >= (unsigned)B) return;
  const int c0 = (int)cv * V;
  float wr[KS * KS][V], bias[V], pool[V];
#pragma unroll
  for (int i = 0; i < KS * KS; ++i) ldv<V>(w + i * C + c0, wr[i]);
  ldv<V>(b + c0, bias);
#pragma unroll
  for (int v = 0; v < V; ++v) pool[v] = 0.f;
  const int oy0 = (int)rg * R, iy0 = oy0 * S - P;
  const int ox_begin = (int)seg * TW, ox_end = min(Wo, ox_begin + TW);
  const float* xn = x + (size_t)n * H * W * C + c0;                 // one 64-bit base per thread, 32-bit element offsets below
  float* yn = y + ((size_t)n * Ho + oy0) * Wo * C + c0;
  int roff[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int iy = iy0 + r;
    roff[r] = (iy >= 0 && iy < H) ? iy * W * C : kBad;
  }
  float win[NR][KS][V];
  auto load_col = [&](int ix, float (&dst)[NR][V]) {
    const int coff = (ix >= 0 && ix < W) ? ix * C : kBad;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int off = roff[r] + coff;
#pragma unroll
      for (int v = 0; v < V; ++v) dst[r][v] = 0.f;
      if (off >= 0) ldv<V>(xn + off, dst[r]);
    }
  };
  {
    float col[NR][V];
#pragma unroll
    for (int kx = 0; kx < KS; ++kx) {
      load_col(ox_begin * S - P + kx, col);
#pragma unroll
      for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int v = 0; v < V; ++v) win[r][kx][v] = col[r][v];
    }
  }
  const int WoC = Wo * C;
  for (int ox = ox_begin; ox < ox_end; ++ox) {
    // the S input columns the NEXT output needs are requested first: their latency overlaps this output's k*k*R FMAs
    // (an unrolled, shift-free variant with rotating window slots was slower: its body no longer fits the instruction cache)
    float nxt[S][NR][V];
    const bool more = ox + 1 < ox_end;
    if (more) {
#pragma unroll
      for (int s2 = 0; s2 < S; ++s2) load_col((ox + 1) * S - P + KS - S + s2, nxt[s2]);
    }
    const int yoff = ox * C;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float acc[V], acc2[V];   // even / odd taps: two independent FMA chains per channel (see depthwise_tma_kernel)
#pragma unroll
      for (int v = 0; v < V; ++v) {
        acc[v] = bias[v];
        acc2[v] = 0.f;
      }
#pragma unroll
      for (int ky = 0; ky < KS; ++ky)
#pragma unroll
        for (int kx = 0; kx < KS; ++kx)
#pragma unroll
          for (int v = 0; v < V; v += 2) {
            if (((ky * KS + kx) & 1) == 0)
              ffma2(acc[v], acc[v + 1], win[r * S + ky][kx][v], win[r * S + ky][kx][v + 1], wr[ky * KS + kx][v], wr[ky * KS + kx][v + 1]);
            else
              ffma2(acc2[v], acc2[v + 1], win[r * S + ky][kx][v], win[r * S + ky][kx][v + 1], wr[ky * KS + kx][v], wr[ky * KS + kx][v + 1]);
          }
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] += acc2[v];
      if (oy0 + r < Ho) {
#pragma unroll
        for (int v = 0; v < V; v += 2) {
          silu2(acc[v], acc[v + 1]);
          pool[v] += acc[v];
          pool[v + 1] += acc[v + 1];
        }
        stv<V>(yn + (yoff + r * WoC), acc);
      }
    }
    if (more) {
#pragma unroll
      for (int r = 0; r < NR; ++r) {
#pragma unroll
        for (int kx = 0; kx < KS - S; ++kx)
#pragma unroll
          for (int v = 0; v < V; ++v) win[r][kx][v] = win[r][kx + S][v];
#pragma unroll
        for (int s2 = 0; s2 < S; ++s2)
#pragma unroll
          for (int v = 0; v < V; ++v) win[r][KS - S + s2][v] = nxt[s2][r][v];
      }
    }
  }
  float* pd = pooled + (size_t)n * C + c0;
#pragma unroll
  for (int v = 0; v < V; ++v) atomicAdd(pd + v, pool[v]);
}

// ---------------------------------------------------------------------------------------------------------------
// depthwise k x k (stride 1 and 2) with TMA-staged input tiles.
// The register-window kernel above keeps only ~24 KiB of loads in flight per SM (12 warps x a handful of LDGs) and sat at
// 1 - 2.6 TB/s; here a producer warp streams ((8 - 1) S + k) x ((TW - 1) S + k) x 32-channel input boxes through a 3 - 6 stage
// mbarrier ring with cp.async.bulk.tensor.4d (zero-filled halo = the convolution's padding, no bounds checks in the math), so
// 100+ KiB are in flight per SM and the 8 consumer warps run the same sliding window out of shared memory.
//   x as a 4-D tensor {C, W, H, B}; box {32, TWin, THin, 1}
//   consumer thread = (V channels, worker); worker = 2 output rows x TW / WCOLS output columns
// Work distribution: CTA b owns ONE channel block (b % cblocks) and a CONTIGUOUS range of the spatial tiles (image, tile x,
// tile y; y fastest) -- group g = b / cblocks of G = SMs / cblocks takes tiles [g T / G, (g + 1) T / G).  The cblocks CTAs of a group walk the
// same tiles at the same time, so all channel blocks of a pixel (which share 128 B lines / 256 B L2 promotions when 4 C is not
// a multiple of 128) are still fetched together, and per CTA
//   * the k*k weights and the bias are loaded once (before: every item, the channel block was the fastest item index),
//   * the tile coordinates advance by increments (before: three integer divisions per item and thread),
//   * the squeeze sums stay in registers until the image changes (before: per item a shuffle reduction, shared-memory partials,
//     a 256-thread named barrier and an atomic per channel -- the barrier kept the 8 warps in lock step, i.e. no warp could hide
//     another's SFU / shared-memory latency across items; ncu: 37 % issue utilisation at 2 warps per scheduler).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDtTH = 8, kDtCB = 32;

struct DwTmaParams {
  CUtensorMap tmap_x;
  const float* w; const float* b; float* y; float* pooled;
  int B, H, W, C, Ho, Wo, tiles_x, tiles_y, cblocks, groups, ychunks, round_robin, stages;
  int l2_hint;   // L2 policy of the input loads: 0 default, 1 evict_last, 2 evict_first
};

// V = channels per thread, NCONS = consumer threads: 3 x 3 runs V = 2 with 16 consumer warps (<= 120 registers), 5 x 5 (50 weight
// + 60 window registers at V = 2) stays at 8 warps
template <int KS, int TW, int CB, int S, int V, int NCONS>
__global__ void __launch_bounds__(NCONS + 32, 1) depthwise_tma_kernel(const __grid_constant__ DwTmaParams p) {
  // S = stride (1 or 2): output tile kDtTH x TW, input box ((kDtTH - 1) S + k) x ((TW - 1) S + k); p.H / p.W = INPUT size, p.Ho / p.Wo = output size
  constexpr int P = (KS - 1) / 2, THin = (kDtTH - 1) * S + KS, TWin = (TW - 1) * S + KS;
  constexpr int CV = CB / V;                  // channel vectors per block
  constexpr int WORKERS = NCONS / CV;
  constexpr int WROWS = kDtTH / 2;               // 4 row pairs
  constexpr int WCOLS = WORKERS / WROWS;         // 8 or 4
  constexpr int CPW = TW / WCOLS;                // output columns per worker
  constexpr int R = 2, NR = (R - 1) * S + KS;
  constexpr uint32_t kTileBytes = THin * TWin * CB * 4;
  constexpr uint32_t kTileStride = (kTileBytes + 127u) & ~127u;
  static_assert(CPW >= 1 && TW % WCOLS == 0, "tile width must split evenly over the workers");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int stages = p.stages;
  const uint32_t bar_off = (uint32_t)stages * kTileStride;
  auto full_bar = [&](int s) { return smem_base + bar_off + 8u * s; };
  auto empty_bar = [&](int s) { return smem_base + bar_off + 8u * (8 + s); };
  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), NCONS / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();
  // this CTA's channel block and group; a group walks UNITS = (image, tile column, chunk of `ulen` consecutive tile rows), unit u of
  // group g: u = g, g + G, ... -- the groups work on neighbouring tile columns at the same time (the horizontal halo of a column is
  // fetched by the neighbour group within the same microseconds) and a CTA walks DOWN its column (the two halo rows it shares with
  // the next tile are in L2).  Contiguous per-group ranges (groups 2+ images apart) read 28 % more DRAM bytes than the tensor holds
  // on the 144-channel maps (profiles/r02_decoder_launches_v22_final.txt).
  // Small maps (tile columns of < 16 tiles) keep CONTIGUOUS unit ranges per group instead (p.round_robin = 0): there the interleaved
  // walk changes image every 1 - 2 tiles and measured 8 - 12 % slower, while their DRAM reads are at the algorithmic minimum anyway.
  const int cb = (int)blockIdx.x % p.cblocks, grp = (int)blockIdx.x / p.cblocks;
  const int units = p.B * p.tiles_x * p.ychunks, ulen = p.tiles_y / p.ychunks;
  const int u_begin = p.round_robin ? grp : (int)((long long)grp * units / p.groups);
  const int u_end = p.round_robin ? units : (int)((long long)(grp + 1) * units / p.groups);
  const int u_step = p.round_robin ? p.groups : 1;
  // (image, tile column, chunk) of the current unit: by division at the start and per round-robin step (units are long there), by
  // increments in the contiguous walk
  int chunk = u_begin % p.ychunks, n = (u_begin / p.ychunks) / p.tiles_x, tx = (u_begin / p.ychunks) % p.tiles_x;
  auto next_unit = [&](int u) {
    if (u_step == 1) {
      if (++chunk == p.ychunks) {
        chunk = 0;
        if (++tx == p.tiles_x) { tx = 0; ++n; }
      }
    } else {
      const int un = u + u_step, strip = un / p.ychunks;
      chunk = un % p.ychunks;
      n = strip / p.tiles_x;
      tx = strip % p.tiles_x;
    }
  };

  if (warp == NCONS / 32) {
    // =========================== TMA producer ===========================
    if (lane == 0) tma_prefetch_desc(&p.tmap_x);
    int stage = 0;
    uint32_t phase = 0;
    for (int u = u_begin; u < u_end; u += u_step) {
      const int ty0 = chunk * ulen;
      for (int ty = ty0; ty < ty0 + ulen; ++ty) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(full_bar(stage), kTileBytes);
          if (p.l2_hint != 0)
            tma_load_4d_hint(smem_base + (uint32_t)stage * kTileStride, &p.tmap_x, full_bar(stage), cb * CB, tx * TW * S - P, ty * kDtTH * S - P, n,
                             p.l2_hint == 1 ? kL2EvictLast : kL2EvictFirst);
          else
            tma_load_4d(smem_base + (uint32_t)stage * kTileStride, &p.tmap_x, full_bar(stage), cb * CB, tx * TW * S - P, ty * kDtTH * S - P, n);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
      next_unit(u);
    }
  } else {
    // =========================== consumers ===========================
    const int tid = threadIdx.x;
    const int cv = tid % CV, worker = tid / CV;
    const int wy = worker / WCOLS, wx = worker % WCOLS;
    const int r0 = 2 * wy, cbeg = wx * CPW;
    const int c0 = cb * CB + cv * V;
    const bool ch_ok = c0 < p.C;       // channels past C: zero-filled input, zero weights; nothing is stored or summed for them
    float wr[KS * KS][V], bias[V];
#pragma unroll
    for (int i = 0; i < KS * KS; ++i) {
      if (ch_ok) ldv<V>(p.w + i * p.C + c0, wr[i]);
      else {
#pragma unroll
        for (int v = 0; v < V; ++v) wr[i][v] = 0.f;
      }
    }
    if (ch_ok) ldv<V>(p.b + c0, bias);
    else {
#pragma unroll
      for (int v = 0; v < V; ++v) bias[v] = 0.f;
    }
    float pool[V];
#pragma unroll
    for (int v = 0; v < V; ++v) pool[v] = 0.f;
    // squeeze sums of image `img`: shuffle-reduce over the workers of this warp (lanes with equal cv), one fire-and-forget global
    // reduction per channel and warp -- once per image and CTA, not per item
    auto flush_pool = [&](int img) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
#pragma unroll
        for (int o = CV; o < 32; o <<= 1) pool[v] += __shfl_xor_sync(0xffffffffu, pool[v], o);
      }
      if (lane < CV && cb * CB + lane * V < p.C) {
        float* dst = p.pooled + (size_t)img * p.C + cb * CB + lane * V;
#pragma unroll
        for (int v = 0; v < V; ++v) atomicAdd(dst + v, pool[v]);
      }
#pragma unroll
      for (int v = 0; v < V; ++v) pool[v] = 0.f;
    };
    const int pix = p.C;
    int stage = 0, cur_n = -1;
    uint32_t phase = 0;
    for (int u = u_begin; u < u_end; u += u_step) {
      const int ty0 = chunk * ulen;
      if (n != cur_n) {
        if (cur_n >= 0) flush_pool(cur_n);     // the squeeze sums of the previous image leave when the image changes
        cur_n = n;
      }
      for (int ty = ty0; ty < ty0 + ulen; ++ty) {
      mbar_wait(full_bar(stage), phase);
      const float* tile = reinterpret_cast<const float*>(smem_gen + (uint32_t)stage * kTileStride) + cv * V;
      auto lds = [&](int row, int col, float (&d)[V]) { ldsv<V>(tile + (row * TWin + col) * CB, d); };
      // window slot of logical column kx at unrolled output column c: (kx + c) % KS -- nothing is ever shifted
      // input window of the worker: rows r0 S ..., columns (cbeg + c) S + kx; logical input column j lives in slot j % KS
      float win[NR][KS][V];
#pragma unroll
      for (int kx = 0; kx < KS; ++kx)
#pragma unroll
        for (int r = 0; r < NR; ++r) lds(r0 * S + r, cbeg * S + kx, win[r][kx]);
      const int oy0 = ty * kDtTH + r0, ox0 = tx * TW + cbeg;
      // one 64-bit base per output row, 32-bit element offsets per column: the full (r Wo + c) C product in 64 bits cost ~10 integer
      // instructions per store (16 stores per item in the 5 x 5 kernels)
      float* yrow[R];
      bool row_ok[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        yrow[r] = p.y + (((size_t)n * p.Ho + oy0 + r) * p.Wo + ox0) * p.C + c0;
        row_ok[r] = ch_ok && oy0 + r < p.Ho;
      }
#pragma unroll
      for (int c = 0; c < CPW; ++c) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          // two partial sums per channel (even / odd taps): a single accumulator is a chain of k*k dependent FMAs, and with 8 warps
          // per SM and 2 - 4 chains per thread the 5 x 5 kernels ran at a quarter of the FMA issue rate
          float acc[V], acc2[V];
#pragma unroll
          for (int v = 0; v < V; ++v) {
            acc[v] = bias[v];
            acc2[v] = 0.f;
          }
#pragma unroll
          for (int ky = 0; ky < KS; ++ky)
#pragma unroll
            for (int kx = 0; kx < KS; ++kx)
#pragma unroll
              for (int v = 0; v < V; v += 2) {   // packed FFMA2: the taps of a channel pair cost one issue slot each
                if (((ky * KS + kx) & 1) == 0)
                  ffma2(acc[v], acc[v + 1], win[r * S + ky][(kx + c * S) % KS][v], win[r * S + ky][(kx + c * S) % KS][v + 1], wr[ky * KS + kx][v], wr[ky * KS + kx][v + 1]);
                else
                  ffma2(acc2[v], acc2[v + 1], win[r * S + ky][(kx + c * S) % KS][v], win[r * S + ky][(kx + c * S) % KS][v + 1], wr[ky * KS + kx][v], wr[ky * KS + kx][v + 1]);
              }
#pragma unroll
          for (int v = 0; v < V; ++v) acc[v] += acc2[v];
          // SiLU outside the bounds test: inside it every output was its own branch region with the serial chain FMUL2 -> MUFU.EX2 ->
          // FADD2 -> MUFU.RCP -> FMUL2 -> STG exposed (16 regions per item in the 5 x 5 kernel, ~60 cycles each, two warps per
          // scheduler to hide them); unconditional, ptxas interleaves the SFU work with the FFMA2 stream of the next outputs.  The
          // squeeze sum takes the value times 1.0 / 0.0 (an exact add) and the store is a single predicated instruction.
#pragma unroll
          for (int v = 0; v < V; v += 2) silu2(acc[v], acc[v + 1]);
          const bool ok = row_ok[r] && ox0 + c < p.Wo;
          const float keep = ok ? 1.f : 0.f;
#pragma unroll
          for (int v = 0; v < V; ++v) pool[v] = fmaf(acc[v], keep, pool[v]);
          if (ok) stv<V>(yrow[r] + c * pix, acc);
        }
        if (c + 1 < CPW) {
          // the S input columns the next output column adds: logical columns c S + KS ... c S + KS + S - 1
#pragma unroll
          for (int s2 = 0; s2 < S; ++s2)
#pragma unroll
            for (int r = 0; r < NR; ++r) lds(r0 * S + r, cbeg * S + c * S + KS + s2, win[r][(c * S + KS + s2) % KS]);
        }
      }
      // this thread is done with the tile: hand the stage back to the producer
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(stage));
      if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
      next_unit(u);
    }
    if (cur_n >= 0) flush_pool(cur_n);
  }
}

// AQ_DW_PROMO = 0 / 128 / 256: L2 promotion of the depthwise input map (A/B measurements)
static int dw_l2_promotion() {
  static const int v = [] { const char* e = getenv("AQ_DW_PROMO"); return e == nullptr ? 256 : atoi(e); }();
  return v;
}

// AQ_DW_L2HINT = 0 / 1 / 2: L2 policy of the depthwise input loads (default, evict_last, evict_first; A/B measurements)
static int dw_l2_hint() {
  static const int v = [] { const char* e = getenv("AQ_DW_L2HINT"); return e == nullptr ? 0 : atoi(e); }();
  return v;
}

template <int KS, int TW, int CB, int S, int V, int NCONS>
static int launch_depthwise_tma_v(const float* x, const float* w, const float* b, float* y, float* pooled, int B, int H, int C,
                                  cudaStream_t st) {
  constexpr int THin = (kDtTH - 1) * S + KS, TWin = (TW - 1) * S + KS;
  const int Ho = (H + 2 * ((KS - 1) / 2) - KS) / S + 1;
  constexpr int kTileStride = (THin * TWin * CB * 4 + 127) & ~127;
  DwTmaParams p;
  memset(&p, 0, sizeof(p));
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)H, (uint64_t)H, (uint64_t)B};
  uint64_t str[3] = {(uint64_t)C * 4, (uint64_t)H * C * 4, (uint64_t)H * H * C * 4};
  uint32_t box[4] = {CB, TWin, THin, 1};
  int rc = make_tmap(&p.tmap_x, x, 4, 4, dims, str, box, kSwzNone, dw_l2_promotion());
  if (rc) return rc;
  p.w = w; p.b = b; p.y = y; p.pooled = pooled;
  p.B = B; p.H = H; p.W = H; p.C = C; p.Ho = Ho; p.Wo = Ho;
  p.tiles_x = (Ho + TW - 1) / TW;
  p.tiles_y = (Ho + kDtTH - 1) / kDtTH;
  p.cblocks = (C + CB - 1) / CB;
  const long long strips = (long long)B * p.tiles_x;
  AQ_REQUIRE(strips * p.tiles_y < (1ll << 31), AQ_ERR_BAD_SHAPE, "depthwise: too many tiles");
  int stages = (200 * 1024) / kTileStride;
  if (stages > 6) stages = 6;
  p.stages = stages;
  const int smem = stages * kTileStride + 16 * 8 + 128;
  const int sms = sm_count();
  if (sms <= 0) return fail(AQ_ERR_LAUNCH, "no CUDA device");
  // G groups of cblocks CTAs.  Large maps (tile columns of >= 16 tiles): units dealt round-robin, a unit = a tile column or 1 / 2,
  // 1 / 4, 1 / 8 of one -- the coarsest split within 2 % of the best last-round occupancy.  Small maps: contiguous ranges of units of
  // <= 2 tiles (balance within one unit).  AQ_DW_RR=0 / 1 forces one of the walks (A/B measurements).
  int groups = sms / p.cblocks;
  if (groups < 1) groups = 1;
  static const int rr_env = [] { const char* e = getenv("AQ_DW_RR"); return e == nullptr ? -1 : atoi(e); }();
  p.round_robin = rr_env >= 0 ? (rr_env != 0) : (p.tiles_y >= 16 ? 1 : 0);
  if (p.round_robin) {
    int best_chunks = 1;
    double best_eff = 0.0;
    for (int ch = 1; ch <= 8 && ch <= p.tiles_y; ch *= 2) {
      if (p.tiles_y % ch != 0) break;
      const long long units = strips * ch;
      const long long g = groups < units ? groups : units;
      const double eff = (double)units / (double)(((units + g - 1) / g) * g);
      if (eff > best_eff + 0.02) {
        best_eff = eff;
        best_chunks = ch;
      }
    }
    p.ychunks = best_chunks;
  } else {
    static const int ulen_env = [] { const char* e = getenv("AQ_DW_ULEN"); return e == nullptr ? 2 : atoi(e); }();
    const int ulen = (ulen_env >= 1 && p.tiles_y % ulen_env == 0) ? ulen_env : 1;     // tiles per unit (AQ_DW_ULEN, default 2)
    p.ychunks = p.tiles_y / ulen;
    if (p.ychunks < 1) p.ychunks = 1;
  }
  if ((long long)groups > strips * p.ychunks) groups = (int)(strips * p.ychunks);
  p.groups = groups;
  p.l2_hint = dw_l2_hint();
  const int grid = groups * p.cblocks;
  AQ_OPT_IN_SMEM((depthwise_tma_kernel<KS, TW, CB, S, V, NCONS>), 227 * 1024);
  depthwise_tma_kernel<KS, TW, CB, S, V, NCONS><<<grid, NCONS + 32, smem, st>>>(p);
  AQ_LAUNCHED();
  return AQ_OK;
}

// AQ_DW_V4=1: the 3 x 3 layers on 4 channels per thread and 8 consumer warps (the layout before the 16-warp version; A/B measurements)
static bool dw_v4() {
  static const bool on = [] { const char* e = getenv("AQ_DW_V4"); return e != nullptr && e[0] == '1'; }();
  return on;
}

template <int KS, int TW, int CB, int S = 1>
static int launch_depthwise_tma_t(const float* x, const float* w, const float* b, float* y, float* pooled, int B, int H, int C,
                                  cudaStream_t st) {
  if constexpr (KS == 3) {
    if (dw_v4()) return launch_depthwise_tma_v<KS, TW, CB, S, 4, 256>(x, w, b, y, pooled, B, H, C, st);
    return launch_depthwise_tma_v<KS, TW, CB, S, 2, 512>(x, w, b, y, pooled, B, H, C, st);
  } else {
    return launch_depthwise_tma_v<KS, TW, CB, S, 2, 256>(x, w, b, y, pooled, B, H, C, st);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// SqueezeExcitation MLP: scale[n, c] = sigmoid(b2[c] + sum_j w2t[j, c] * silu(b1[j] + sum_c' w1[j, c'] * mean[n, c']))
// One CLUSTER of 4 CTAs per image (grid (4, B), 512 threads): the kernel is a chain of two tiny matrix-vector products whose
// time is the latency of streaming w1 [SQ, C] and w2t [SQ, C] (up to 2 x 614 KB) through ONE SM, so the rows of w1 and the
// channels of w2t are split over the cluster: CTA r computes the squeeze entries j = r, r + 4, ... (a warp per row, float4 loads,
// 4 independent loads in flight per lane), stores them into the s1 array of every CTA of the cluster (st.shared::cluster), and
// after the cluster barrier produces the channels [r C/4, (r+1) C/4) (8 rows of w2t in flight per thread).  Earlier layouts:
// a block per 256 channels, each recomputing s1 (31 - 68 us per launch on the 1152 / 1920-channel layers); one block per 1024
// channels (38 - 63 us: the same latency chain on fewer SMs) -- profiles/r02_decoder_launches_v10/v11.txt.
// pooled holds SUMS over hw pixels.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSeThreads = 512, kSeCluster = 4;

__global__ void __cluster_dims__(kSeCluster, 1, 1) __launch_bounds__(kSeThreads)
se_kernel(const float* __restrict__ pooled, const float* __restrict__ w1, const float* __restrict__ b1,
          const float* __restrict__ w2t, const float* __restrict__ b2, float* __restrict__ scale, int C, int SQ, float inv_hw) {
  extern __shared__ __align__(16) float sm[];   // mean [C], s1 [SQ]
  float* mean = sm;
  float* s1 = sm + C;
  const int n = blockIdx.y;
  const uint32_t rank = cluster_ctarank();
  for (int c = threadIdx.x; c < C; c += kSeThreads) mean[c] = pooled[(size_t)n * C + c] * inv_hw;
  __syncthreads();
  cluster_sync_all();   // every CTA of the cluster is running (its shared memory exists) before remote stores target it
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarps = kSeThreads / 32;
  const int c4n = C >> 2;   // C % 4 == 0 (expanded widths of the network are multiples of 16)
  for (int j = (int)rank + kSeCluster * warp; j < SQ; j += kSeCluster * kWarps) {
    const float4* row = reinterpret_cast<const float4*>(w1 + (size_t)j * C);
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c4 = lane; c4 < c4n; c4 += 128) {
      float4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = c4 + 32 * k < c4n ? __ldg(row + c4 + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (c4 + 32 * k < c4n) {
          const float4 m = *reinterpret_cast<const float4*>(mean + 4 * (c4 + 32 * k));
          a[k] = fmaf(u[k].x, m.x, fmaf(u[k].y, m.y, fmaf(u[k].z, m.z, fmaf(u[k].w, m.w, a[k]))));
        }
      }
    }
    float acc = (a[0] + a[1]) + (a[2] + a[3]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane < kSeCluster) {
      const float v = silu(acc + b1[j]);
      const uint32_t remote = mapa_shared(smem_u32(s1 + j), (uint32_t)lane);
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
    }
  }
  cluster_sync_all();   // release / acquire at cluster scope: all of s1 is visible in every CTA
  const int cper = ((C + kSeCluster - 1) / kSeCluster + 3) & ~3;
  const int c = (int)rank * cper + threadIdx.x;
  if (threadIdx.x < cper && c < C) {
    float a0 = b2[c], a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int j = 0;
    for (; j + 7 < SQ; j += 8) {
      float w[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) w[k] = __ldg(w2t + (size_t)(j + k) * C + c);
      a0 = fmaf(w[0], s1[j], a0); a1 = fmaf(w[1], s1[j + 1], a1); a2 = fmaf(w[2], s1[j + 2], a2); a3 = fmaf(w[3], s1[j + 3], a3);
      a0 = fmaf(w[4], s1[j + 4], a0); a1 = fmaf(w[5], s1[j + 5], a1); a2 = fmaf(w[6], s1[j + 6], a2); a3 = fmaf(w[7], s1[j + 7], a3);
    }
    for (; j < SQ; ++j) a0 = fmaf(__ldg(w2t + (size_t)j * C + c), s1[j], a0);
    const float acc = (a0 + a1) + (a2 + a3);
    scale[(size_t)n * C + c] = 1.f / (1.f + expf(-acc));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// classifier: logits[n, o] = b[o] + sum_c w[o, c] * pooled_sum[n, c] / hw ;  bits[n, i] = argmax(logits[n, 2i], logits[n, 2i+1])
// ---------------------------------------------------------------------------------------------------------------
// 32 warps per image, 3 outputs each, 10 independent 16-byte weight loads per lane and output: with 8 warps x 12 outputs and a scalar
// load per step the kernel was one L2 latency chain (53 us for 7.9 MFLOP)
__global__ void __launch_bounds__(1024) fc_kernel(const float* __restrict__ pooled, const float* __restrict__ w,
                                                  const float* __restrict__ b, float* __restrict__ logits,
                                                  unsigned char* __restrict__ bits, int C, int O, float inv_hw) {
  extern __shared__ __align__(16) float sm[];   // mean [C], out [O]
  float* mean = sm;
  float* out = sm + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mean[c] = pooled[(size_t)n * C + c] * inv_hw;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  for (int o = warp; o < O; o += nwarps) {
    const float4* wr4 = reinterpret_cast<const float4*>(w + (size_t)o * C);       // C % 4 == 0, rows 16-byte aligned (host checks)
    const float4* mean4 = reinterpret_cast<const float4*>(mean);
    float acc = 0.f;
#pragma unroll 10
    for (int c4 = lane; c4 < (C >> 2); c4 += 32) {
      const float4 wv = __ldg(wr4 + c4);
      const float4 mv = mean4[c4];
      acc = fmaf(wv.x, mv.x, acc);
      acc = fmaf(wv.y, mv.y, acc);
      acc = fmaf(wv.z, mv.z, acc);
      acc = fmaf(wv.w, mv.w, acc);
    }
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
      if (st.expand != 1) { wk.take((size_t)cin * cexp); wk.take((size_t)cin * cexp); wk.take(cexp); }
      wk.take((size_t)st.k * st.k * cexp); wk.take(cexp);
      wk.take((size_t)sq * cexp); wk.take(sq); wk.take((size_t)cexp * sq); wk.take(cexp);
      wk.take((size_t)cexp * st.cout); wk.take((size_t)cexp * st.cout); wk.take(st.cout);
    }
  }
  wk.take((size_t)kLastC * kHeadC); wk.take((size_t)kLastC * kHeadC); wk.take(kHeadC);
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

// Which MBConv blocks run expand -> depthwise as one kernel (decoder_fused.cu).  Measured per 64 images (ncu launch lists,
// profiles/r02_decoder_launches_v12_fused_all.txt, ..._v14_tma_store.txt): 16 -> 96 k3 s2 at 256 x 256: 938 us fused vs 433 + 419 us
// unfused (597 + 420 before the pointwise kernel's TMA-store epilogue); 24 -> 144 k3 s1 at 128 x 128: 625 vs 152 + 283 us; 24 -> 144 k5
// s2: 846 vs 152 + 293 us.  The fused kernel removes the expanded map's HBM round trip (DRAM traffic of the first block 3.9 -> 0.65 GB)
// but pays the bias + SiLU epilogue on CUDA cores for the halo pixels too (1.2x ... 1.6x) at ~1.3 instructions per clock, and loses
// to the tensor-core pair at every shape today -> opt-in.  AQ_DEC_FUSED=1: the first stage-2 block, =2: every supported shape
// (A/B measurements; the kernel-level parity tests call aq_expand_dw_fused directly); default 0: never.
static int fused_mode() {
  static const int mode = [] { const char* e = getenv("AQ_DEC_FUSED"); return e == nullptr ? 0 : (e[0] == '1' ? 1 : (e[0] == '2' ? 2 : 0)); }();
  return mode;
}
static bool fused_enabled(int cin, int cexp, int k, int stride) {
  const int mode = fused_mode();
  if (mode == 0) return false;
  if (mode == 2) return true;
  return cin == 16 && cexp == 96 && k == 3 && stride == 2;
}

static bool dw_tw16() {
  static const bool on = [] { const char* e = getenv("AQ_DW_TW16"); return e != nullptr && e[0] == '1'; }();
  return on;
}

// AQ_DW_S2_TMA=0 keeps the stride-2 layers on the register-window kernel (A/B measurements)
static bool dw_s2_tma_enabled() {
  static const bool on = [] { const char* e = getenv("AQ_DW_S2_TMA"); return e == nullptr || e[0] != '0'; }();
  return on;
}

static int launch_depthwise(const float* x, const float* w, const float* b, float* y, float* pooled, int B, int H, int C, int k,
                            int stride, int Ho, cudaStream_t st) {
  // AQ_DW_CB16=1 (A/B measurement): channel counts that are an odd multiple of 16 (144 = 9 x 16) in 16-channel blocks -- no half-empty
  // last block, every box row is a whole 64-byte-aligned piece of the pixel -- instead of 32-channel blocks with a 64-byte-misaligned
  // 128-byte row on every other pixel
  static const bool cb16 = [] { const char* e = getenv("AQ_DW_CB16"); return e != nullptr && e[0] == '1'; }();
  if (cb16 && stride == 1 && k == 3 && C % 32 == 16 && C > 16 && Ho >= 64 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0)
    return launch_depthwise_tma_t<3, 64, 16>(x, w, b, y, pooled, B, H, C, st);
  if (stride == 1 && C >= kDtCB && C % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0) {
    const bool wide = Ho >= 32 && !dw_tw16();
    if (k == 3) return wide ? launch_depthwise_tma_t<3, 32, kDtCB>(x, w, b, y, pooled, B, H, C, st) : launch_depthwise_tma_t<3, 16, kDtCB>(x, w, b, y, pooled, B, H, C, st);
    if (k == 5) return wide ? launch_depthwise_tma_t<5, 32, kDtCB>(x, w, b, y, pooled, B, H, C, st) : launch_depthwise_tma_t<5, 16, kDtCB>(x, w, b, y, pooled, B, H, C, st);
  }
  // stride 2 through the same TMA-staged kernel (input boxes of 17 x 33 / 19 x 35 pixels for 8 x 16 outputs): the register-window
  // kernel below kept ~24 KiB of loads in flight per SM and ran the 5 x 5 stride-2 layers at 2.3 - 2.8 TB/s
  if (stride == 2 && C >= kDtCB && C % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0 && dw_s2_tma_enabled()) {
    if (k == 3) return launch_depthwise_tma_t<3, 16, kDtCB, 2>(x, w, b, y, pooled, B, H, C, st);
    if (k == 5) return launch_depthwise_tma_t<5, 16, kDtCB, 2>(x, w, b, y, pooled, B, H, C, st);
  }
  // the 16-channel depthwise of the second stage-1 block (256 x 256 maps): same kernel with a 16-channel block
  if (stride == 1 && C == 16 && k == 3 && Ho >= 32 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0)
    return Ho >= 64 ? launch_depthwise_tma_t<3, 64, 16>(x, w, b, y, pooled, B, H, C, st)    // 4 output columns per worker instead of 2
                    : launch_depthwise_tma_t<3, 32, 16>(x, w, b, y, pooled, B, H, C, st);
  const int R = 2;   // output rows per thread
  const int TW = Ho >= 64 ? 32 : 16;
  const int V = k == 3 ? 4 : 2;
  const long long threads = (long long)B * ((Ho + R - 1) / R) * ((Ho + TW - 1) / TW) * (C / V);
  AQ_REQUIRE(threads < (1ll << 32), AQ_ERR_BAD_SHAPE, "depthwise: %lld work items exceed 2^32", threads);
  const unsigned grid = (unsigned)((threads + kDwThreads - 1) / kDwThreads);
  if (k == 3 && stride == 1) depthwise_kernel<3, 1, 2, 4><<<grid, kDwThreads, 0, st>>>(x, w, b, y, pooled, B, H, H, C, Ho, Ho, TW);
  else if (k == 3 && stride == 2) depthwise_kernel<3, 2, 2, 4><<<grid, kDwThreads, 0, st>>>(x, w, b, y, pooled, B, H, H, C, Ho, Ho, TW);
  else if (k == 5 && stride == 1) depthwise_kernel<5, 1, 2, 2><<<grid, kDwThreads, 0, st>>>(x, w, b, y, pooled, B, H, H, C, Ho, Ho, TW);
  else if (k == 5 && stride == 2) depthwise_kernel<5, 2, 2, 2><<<grid, kDwThreads, 0, st>>>(x, w, b, y, pooled, B, H, H, C, Ho, Ho, TW);
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
    static_assert(kImg % 4 == 0, "stem_kernel loads the input rows as float4");
    dim3 grid((h + kStemTile - 1) / kStemTile, h, B);
    stem_kernel<<<grid, kStemThreads, 0, st>>>(x, w, b, act[0], kImg, kImg, h, h);
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
      float* pl = pooled + (size_t)B * pooled_off;
      pooled_off += pad4(cexp);
      // early blocks: expand 1x1 -> depthwise in ONE kernel, the 6x expanded map never reaches HBM (decoder_fused.cu)
      const bool fused = sc.expand != 1 && fused_enabled(cin, cexp, sc.k, stride) && fused_expand_dw_supported(cin, cexp, sc.k, stride, h);
      if (fused) {
        FusedArgs a{};
        a.x = act[cur]; a.w_hi = wk.take((size_t)cin * cexp); a.w_lo = wk.take((size_t)cin * cexp); a.b_e = wk.take(cexp);
        a.w_d = wk.take((size_t)sc.k * sc.k * cexp); a.b_d = wk.take(cexp);
        a.y = dwo; a.pooled = pl; a.B = B; a.H = h; a.cin = cin; a.cexp = cexp; a.k = sc.k; a.stride = stride;
        rc = launch_fused_expand_dw(a, st);
        if (rc) return rc;
      } else {
        if (sc.expand != 1) {
          PwTcArgs a{};
          a.x = act[cur]; a.w_hi = wk.take((size_t)cin * cexp); a.w_lo = wk.take((size_t)cin * cexp); a.bias = wk.take(cexp);
          a.se = nullptr; a.residual = nullptr; a.y = expb;
          a.M = (long long)B * h * h; a.K = cin; a.N = cexp; a.hw = h * h; a.epi = kPwSilu;
          rc = launch_pointwise_tc(a, st);
          if (rc) return rc;
          dw_in = expb;
        }
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
        AQ_REQUIRE(cexp <= kSeCluster * kSeThreads, AQ_ERR_BAD_SHAPE, "se: %d channels exceed %d", cexp, kSeCluster * kSeThreads);
        se_kernel<<<dim3(kSeCluster, B), kSeThreads, (cexp + sq) * sizeof(float), st>>>(pl, w1, b1, w2, b2, scale, cexp, sq, 1.f / (float)(ho * ho));
        AQ_LAUNCHED();
      }
      {
        PwTcArgs a{};
        a.x = dwo; a.w_hi = wk.take((size_t)cexp * sc.cout); a.w_lo = wk.take((size_t)cexp * sc.cout); a.bias = wk.take(sc.cout);
        a.se = scale;
        const bool res = stride == 1 && cin == sc.cout;
        a.residual = res ? act[cur] : nullptr; a.y = act[cur ^ 1];
        a.M = (long long)B * ho * ho; a.K = cexp; a.N = sc.cout; a.hw = ho * ho; a.epi = res ? kPwResidual : kPwNone;
        rc = launch_pointwise_tc(a, st);
        if (rc) return rc;
      }
      cur ^= 1;
      h = ho;
    }
  }
  float* head_pool = pooled + (size_t)B * pooled_off;
  {
    PwTcArgs a{};
    a.x = act[cur]; a.w_hi = wk.take((size_t)kLastC * kHeadC); a.w_lo = wk.take((size_t)kLastC * kHeadC); a.bias = wk.take(kHeadC);
    a.se = nullptr; a.residual = nullptr; a.y = head_pool;
    a.M = (long long)B * h * h; a.K = kLastC; a.N = kHeadC; a.hw = h * h; a.epi = kPwSiluPool;
    rc = launch_pointwise_tc(a, st);
    if (rc) return rc;
  }
  {
    const float* w = wk.take((size_t)out_features * kHeadC);
    const float* b = wk.take(out_features);
    fc_kernel<<<B, 1024, (kHeadC + out_features) * sizeof(float), st>>>(head_pool, w, b, logits, bits, kHeadC, out_features, 1.f / (float)(h * h));
    AQ_LAUNCHED();
  }
  return AQ_OK;
}

int aq_expand_dw_fused(const float* x, const float* w_hi, const float* w_lo, const float* b_e, const float* w_d, const float* b_d, float* y,
                       float* pooled, int B, int H, int cin, int cexp, int k, int stride, void* stream) {
  int rc = check_arch();
  if (rc) return rc;
  FusedArgs a{};
  a.x = x; a.w_hi = w_hi; a.w_lo = w_lo; a.b_e = b_e; a.w_d = w_d; a.b_d = b_d; a.y = y; a.pooled = pooled;
  a.B = B; a.H = H; a.cin = cin; a.cexp = cexp; a.k = k; a.stride = stride;
  return launch_fused_expand_dw(a, (cudaStream_t)stream);
}

int aq_conv1x1_tf32x3(const float* x, const float* w_hi, const float* w_lo, const float* bias, const float* se, const float* residual,
                      float* y, int64_t M, int K, int N, int hw, int epi, void* stream) {
  AQ_REQUIRE(x && w_hi && w_lo && bias && y, AQ_ERR_BAD_SHAPE, "conv1x1_tf32x3: NULL operand");
  AQ_REQUIRE(epi >= 0 && epi <= 3 && (epi != kPwResidual || residual != nullptr), AQ_ERR_BAD_SHAPE, "conv1x1_tf32x3: bad epilogue %d", epi);
  int rc = check_arch();
  if (rc) return rc;
  PwTcArgs a{};
  a.x = x; a.w_hi = w_hi; a.w_lo = w_lo; a.bias = bias; a.se = se; a.residual = residual; a.y = y;
  a.M = M; a.K = K; a.N = N; a.hw = hw; a.epi = epi;
  return launch_pointwise_tc(a, (cudaStream_t)stream);
}

int aq_depthwise_silu(const float* x, const float* w, const float* bias, float* y, float* pooled, int B, int H, int C, int k,
                      int stride, void* stream) {
  AQ_REQUIRE(x && w && bias && y && pooled && B > 0 && H > 0, AQ_ERR_BAD_SHAPE, "depthwise_silu: NULL operand or empty batch");
  AQ_REQUIRE(C % 4 == 0, AQ_ERR_BAD_SHAPE, "depthwise_silu: C=%d must be a multiple of 4", C);
  int rc = check_arch();
  if (rc) return rc;
  const int ho = (H + 2 * ((k - 1) / 2) - k) / stride + 1;
  return launch_depthwise(x, w, bias, y, pooled, B, H, C, k, stride, ho, (cudaStream_t)stream);
}

}  // extern "C"
