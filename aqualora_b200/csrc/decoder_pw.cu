// Pointwise (1x1) convolutions of the message decoder on the tcgen05 tensor cores with fp32-faithful arithmetic.
//
//   Y[m, n] = epi( sum_k (X[m, k] * se[m / hw, k]) * W[n, k] + b[n] )        X [M, K] NHWC pixels, W [N, K], fp32 in / fp32 out
//
// The reference runs these convolutions in fp32 (utils/models.py:84-96 never casts the decoder) and the decoded bits must agree
// with it, so a plain TF32 / bf16 product is not acceptable (logit margins of a random-init decoder go down to 5e-4).  Every
// product is therefore evaluated as the 3-term split
//   a * w  ~=  a_hi * w_hi + a_lo * w_hi + a_hi * w_lo ,   x_hi = x with the low 13 mantissa bits cleared (exactly a TF32 number),
//                                                            x_lo = x - x_hi (exact in fp32)
// on kind::tf32 MMAs with fp32 accumulation in TMEM; the dropped a_lo * w_lo term is <= 2^-22 |a w|.  Weights are split once at
// pack time; activations are split by 4 converter warps inside shared memory (in place for the hi part), which is also where the
// squeeze-excitation scale of the project convolutions is applied.
//
// One CTA per SM, persistent over (128-row tile, N tile <= 224) work items, 832 threads:
//   warps 0-15   epilogue   TMEM lane quadrant = warp % 4, 16-column pieces pc = warp / 4 (mod 4):
//                           TMEM -> +bias -> [SiLU] [+ residual] -> staging tile (box of Y's tensor map) -> one bulk-tensor store | SiLU + pooled sums
//   warps 16-23  converter  raw fp32 A tile (TMA, 128B swizzle) -> x * se -> hi (in place) / lo (second tile); 8 warps since round 2:
//                           with 4 the K-heavy, narrow project layers were bound by this stage
//   warp 24      TMA producer (A raw, W_hi, W_lo per 32-wide k chunk)
//   warp 25      TMEM allocator + MMA issuer (3 MMAs per 8-wide k step)
// Variations chosen per layer by launch_pointwise_tc (all measured, profiles/r02_decoder_launches_v19 ... v24):
//   tile_par   narrow outputs (N <= 48) and the 96-wide expand: a warp group takes WHOLE tiles in turn (4 / 8-slot accumulator ring)
//   rotation   otherwise the first 16-column piece of a warp group rotates with the tile (9 pieces = 3 + 2 + 2 + 2 per tile)
//   stack      N <= 48, K >= 32: [W_hi ; W_lo] as ONE B operand of 2 BN rows, 2 MMAs per k step, the epilogue adds the two halves
//   resident W single column tile with <= 96 KB of W hi / lo (<= 120 KB for the one 240-column tile): only A streams
//   CTA pairs  everything whose W is not resident (pointwise_tc_pair_kernel below)
// The expand convolutions (128 x 96 ... 240 outputs from a 16 ... 40-deep product) are bound by the epilogue's instruction issue
// (ncu: ~0.75 instructions per output element, the 8 epilogue warps of the previous version busy 90 % of the time, tensor pipe
// 4 % active), hence 16 epilogue warps -- 4 per scheduler -- an SFU SiLU and an accumulator ring of up to 8 tiles in TMEM.
#include <stdlib.h>
#include <string.h>

#include "aq_ptx.cuh"
#include "decoder_pw.h"

namespace aq {

constexpr int kPwThreads = 832;
constexpr int kPwEpiWarps = 16;
constexpr int kPwConvWarp0 = 16, kPwConvWarps = 8, kPwProducerWarp = 24, kPwMmaWarp = 25;
constexpr int kPwConvThreads = kPwConvWarps * 32, kPwConvIters = 1024 / kPwConvThreads;   // 1024 16-byte chunks per 128 x 32 fp32 tile
constexpr int kPwMaxBN = 224;                  // 2 pipeline stages + 16 epilogue staging buffers must fit in 227 KiB
constexpr int kPwBM = 128;
constexpr int kPwKC = 32;                      // fp32 elements per k chunk = one 128-byte swizzle row
constexpr int kPwATile = kPwBM * kPwKC * 4;    // 16 KiB
constexpr int kPwStgWarp = 32 * 64;             // per epilogue warp: one 32-row x 16-column fp32 piece = the box of Y's tensor map (64B swizzle)
constexpr int kPwSmemBudget = 232448 - 1024;
constexpr int kPwResidentMax = 120 * 1024;     // resident W of a single > 224-column tile: leaves two 32 KB stages beside the 32 KB of staging

struct PwTcParams {
  CUtensorMap tmap_x;     // X    [M, K] fp32   box {32, 128}  swizzle 128B
  CUtensorMap tmap_whi;   // W_hi [N, K] fp32   box {32, BN}   swizzle 128B
  CUtensorMap tmap_wlo;   // W_lo [N, K]
  CUtensorMap tmap_y;     // Y    [M, N] fp32   box {16, 32}   swizzle 64B (store; rows >= M / columns >= N are clipped)
  const float* bias;      // [N]
  const float* se;        // [M / hw, K] or null
  const float* residual;  // [M, N] or null
  float* y;               // [M, N]   (pool epilogue: [M / hw, N] sums, accumulated)
  long long M;
  int N, K, hw, epi;
  int BN, num_n_tiles, num_m_tiles, num_kc, stages;
  int w_resident;   // 1: all W chunks (hi + lo) stay in SMEM for the whole kernel (single N tile, small K): only A streams
  int tile_par;     // 1: narrow outputs (<= 3 pieces): the four epilogue warps of a TMEM quadrant take WHOLE tiles in turn (single-CTA kernel)
  int silu_nr;      // 1: SiLU with the reciprocal on the FMA pipe (silu2_nr: one SFU operation per element instead of two)
  int stack;        // 1: W_hi and W_lo form ONE B operand of 2 BN rows (single-CTA kernel): a_hi is read from shared memory once per k step
};

// SiLU with the SFU exponential and reciprocal (relative error ~1e-6, far inside the decoder's 1e-4 logit tolerance)
__device__ __forceinline__ float pw_silu(float v) {
  // 5 instructions: FMUL, MUFU.EX2, FADD, MUFU.RCP, FMUL.  ex2.approx.ftz saturates cleanly (v -> -inf: e = inf, 1 / inf = 0;
  // v -> +inf: e = 0), so __expf's denormal-range fix-up (3 more instructions per element) is not needed.
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return v * r;
}

__global__ void __launch_bounds__(kPwThreads, 1) pointwise_tc_kernel(const __grid_constant__ PwTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;

  const int BN = p.BN;
  const int stages = p.stages;
  const uint32_t w_bytes = (uint32_t)BN * 128u;                  // one W tile (hi or lo): BN rows x 128 B
  const bool wres = p.w_resident != 0;
  const uint32_t stage_bytes = 2u * kPwATile + (wres ? 0u : 2u * w_bytes);
  // layout: [resident W: num_kc x (hi | lo)] [stages x (A hi | A lo | W hi | W lo)] [epilogue staging, 16 warps] [barriers]
  const uint32_t wres_bytes = wres ? (uint32_t)p.num_kc * 2u * w_bytes : 0u;
  const uint32_t ring_off = wres_bytes;
  const uint32_t stg_off = ring_off + (uint32_t)stages * stage_bytes;
  const uint32_t bar_off = stg_off + (uint32_t)kPwEpiWarps * kPwStgWarp;
  const uint32_t bar_base = smem_base + bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                 // TMA bytes landed
  auto conv_bar = [&](int s) { return bar_base + 8u * (8 + s); };           // A split into hi / lo (4 converter warps)
  auto empty_bar = [&](int s) { return bar_base + 8u * (16 + s); };         // MMAs that read the stage completed
  auto acc_full_bar = [&](int b) { return bar_base + 8u * (24 + b); };
  auto acc_empty_bar = [&](int b) { return bar_base + 8u * (32 + b); };
  const uint32_t w_full_bar = bar_base + 8u * 40;                           // resident W landed (once)
  const uint32_t tmem_slot = bar_base + 8u * 41;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + bar_off + 8 * 41);
  // accumulator ring in TMEM: narrow layers (BN = 16 ... 64) keep up to 8 tiles in flight between the MMA thread and the epilogue
  // tile_par needs a ring size that is a multiple of 4 (a slot always belongs to the same epilogue warp group)
  // stack: the accumulator of a tile is [a_hi w_hi + a_lo w_hi | a_hi w_lo] (2 BN columns, summed by the epilogue).  The narrow project
  // layers are bound by shared-memory bandwidth (per 128 x 160 A tile: 80 KB TMA write, 80 KB converter read, 160 KB hi / lo write,
  // 3 x 80 KB operand reads by the MMAs -- the 16-cycle MMA of a 32-column tile waits for its 4 KB A read); one stacked MMA reads a_hi
  // once instead of twice
  const bool stack = p.stack != 0;
  const uint32_t acc_cols = stack ? 2u * (uint32_t)BN : (uint32_t)BN;
  const uint32_t nacc = p.tile_par ? (acc_cols <= 64 ? 8u : 4u) : min(8u, 512u / acc_cols);

  auto a_hi = [&](int s) { return smem_base + ring_off + (uint32_t)s * stage_bytes; };
  auto a_lo = [&](int s) { return smem_base + ring_off + (uint32_t)s * stage_bytes + kPwATile; };
  // W of k chunk kc in ring stage s, or in the resident region
  auto w_hi = [&](int s, int kc) { return wres ? smem_base + (uint32_t)kc * 2u * w_bytes : smem_base + ring_off + (uint32_t)s * stage_bytes + 2u * kPwATile; };
  auto w_lo = [&](int s, int kc) { return w_hi(s, kc) + w_bytes; };

  if (warp == kPwProducerWarp && lane == 0) {
    tma_prefetch_desc(&p.tmap_x);
    tma_prefetch_desc(&p.tmap_whi);
    tma_prefetch_desc(&p.tmap_wlo);
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(conv_bar(s), kPwConvWarps);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 8; ++b) {
      mbar_init(acc_full_bar(b), 1);
      mbar_init(acc_empty_bar(b), p.tile_par ? 4 : kPwEpiWarps);
    }
    mbar_init(w_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == kPwMmaWarp) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int total_items = p.num_m_tiles * p.num_n_tiles;
  const int num_kc = p.num_kc;

  if (warp == kPwProducerWarp) {
    // =========================== TMA producer ===========================
    int stage = 0;
    uint32_t phase = 0;
    if (wres) {
      // the whole (tiny) weight matrix once: re-fetching it per tile cost 60 % of the TMA traffic of the expand layers and
      // made every CTA hammer the same few L2 lines
      if (elect_one()) {
        mbar_arrive_expect_tx(w_full_bar, (uint32_t)num_kc * 2u * w_bytes);
        for (int kc = 0; kc < num_kc; ++kc) {
          tma_load_2d(w_hi(0, kc), &p.tmap_whi, w_full_bar, kc * kPwKC, 0);
          tma_load_2d(w_lo(0, kc), &p.tmap_wlo, w_full_bar, kc * kPwKC, 0);
        }
      }
      __syncwarp();
    }
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int m0 = (item / p.num_n_tiles) * kPwBM;
      const int n0 = (item % p.num_n_tiles) * BN;
      for (int kc = 0; kc < num_kc; ++kc) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(full_bar(stage), kPwATile + (wres ? 0u : 2u * w_bytes));
          tma_load_2d(a_hi(stage), &p.tmap_x, full_bar(stage), kc * kPwKC, m0);
          if (!wres) {
            tma_load_2d(w_hi(stage, kc), &p.tmap_whi, full_bar(stage), kc * kPwKC, n0);
            tma_load_2d(w_lo(stage, kc), &p.tmap_wlo, full_bar(stage), kc * kPwKC, n0);
          }
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == kPwMmaWarp) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc_tf32(kPwBM, (uint32_t)BN);
    const uint32_t idesc2 = make_idesc_tf32(kPwBM, 2u * (uint32_t)BN);   // stacked [W_hi ; W_lo] (contiguous in shared memory)
    constexpr uint64_t kDescHi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    auto desc = [&](uint32_t addr) { return kDescHi | (uint64_t)(((addr >> 4) & 0x3FFFu) | (1u << 16)); };
    int stage = 0;
    uint32_t phase = 0, acc_iter = 0;
    if (wres) {
      mbar_wait(w_full_bar, 0);
      tc_fence_after();
    }
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++acc_iter) {
      const uint32_t buf = acc_iter % nacc;
      const uint32_t tmem_acc = tmem_base + buf * acc_cols;
      mbar_wait(acc_empty_bar(buf), ((acc_iter / nacc) & 1u) ^ 1u);
      tc_fence_after();
      for (int kc = 0; kc < num_kc; ++kc) {
        mbar_wait(full_bar(stage), phase);   // W tiles (async proxy) ...
        mbar_wait(conv_bar(stage), phase);   // ... and the hi / lo split of A (generic proxy + fence.proxy.async)
        tc_fence_after();
        if (elect_one()) {
          const int ksteps = min(kPwKC, p.K - kc * kPwKC) >> 3;   // 8 tf32 per MMA; the zero-filled tail is skipped
          const uint64_t ah = desc(a_hi(stage)), al = desc(a_lo(stage)), wh = desc(w_hi(stage, kc)), wl = desc(w_lo(stage, kc));
          if (stack) {
            for (int k = 0; k < ksteps; ++k) {
              umma_tf32(tmem_acc, ah + 2 * k, wh + 2 * k, idesc2, (kc | k) != 0 ? 1u : 0u);   // [a_hi w_hi | a_hi w_lo]
              umma_tf32(tmem_acc, al + 2 * k, wh + 2 * k, idesc, 1u);                          // first half += a_lo w_hi
            }
          } else {
            for (int k = 0; k < ksteps; ++k) {
              // small terms first, the dominant hi * hi product last
              umma_tf32(tmem_acc, al + 2 * k, wh + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
              umma_tf32(tmem_acc, ah + 2 * k, wl + 2 * k, idesc, 1u);
              umma_tf32(tmem_acc, ah + 2 * k, wh + 2 * k, idesc, 1u);
            }
          }
          umma_commit(empty_bar(stage));
          if (kc == num_kc - 1) umma_commit(acc_full_bar(buf));
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= kPwConvWarp0) {
    // =========================== converter warps: A -> (A * se) hi / lo ===========================
    const int t = threadIdx.x - kPwConvWarp0 * 32;   // 0 .. kPwConvThreads - 1
    const int cphys = t & 7;                  // 16-byte chunk inside the 128-byte row (physical, swizzled)
    const int rlow = (t >> 3) & 7;            // row & 7 of every row this thread touches (rows t/8 + 16 j)
    const int clog = cphys ^ rlow;            // logical chunk = k offset / 4 inside the chunk
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const long long m0 = (long long)(item / p.num_n_tiles) * kPwBM;
      const float* se_row = p.se != nullptr ? p.se + (m0 / p.hw) * p.K : nullptr;   // a 128-row tile lies inside one image
      for (int kc = 0; kc < num_kc; ++kc) {
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
        const int k = kc * kPwKC + clog * 4;
        if (se_row != nullptr && k < p.K) sc = __ldg(reinterpret_cast<const float4*>(se_row + k));
        mbar_wait(full_bar(stage), phase);
        uint8_t* hi = smem_gen + (a_hi(stage) - smem_base);
        uint8_t* lo = smem_gen + (a_lo(stage) - smem_base);
        // columns at or past K are never read by the MMAs (the k-step count stops at K): for the 16 / 24 / 40-deep expand
        // layers that is up to half of the chunk's shared-memory traffic (the LSU data pipe was 65 % busy there, ncu)
        if (k < p.K) {
        float4 v[kPwConvIters];
#pragma unroll
        for (int j = 0; j < kPwConvIters; ++j) v[j] = *reinterpret_cast<const float4*>(hi + (t + kPwConvThreads * j) * 16);
#pragma unroll
        for (int j = 0; j < kPwConvIters; ++j) {
          float x[4] = {v[j].x * sc.x, v[j].y * sc.y, v[j].z * sc.z, v[j].w * sc.w};
          float h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            h[e] = __uint_as_float(__float_as_uint(x[e]) & 0xFFFFE000u);
            l[e] = x[e] - h[e];
          }
          *reinterpret_cast<float4*>(hi + (t + kPwConvThreads * j) * 16) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(lo + (t + kPwConvThreads * j) * 16) = make_float4(l[0], l[1], l[2], l[3]);
        }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(conv_bar(stage));
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // =========================== epilogue warps ===========================
    const int q = warp & 3;        // TMEM lane quadrant
    const int res = warp >> 2;     // this warp takes the 16-column pieces pc = res, res + 4, ... of every tile -- or, for narrow
                                   // outputs (tile_par: the 16 / 24 / 40-channel project layers, where 1 - 3 pieces would leave 12 - 4
                                   // of the 16 epilogue warps idle on the layers with the most rows), every 4th tile as a whole.
                                   // The accumulator ring then has 8 slots (a multiple of 4), so a slot always belongs to the
                                   // same warp group and its full / empty phases are observed in order.
    const bool tile_par = p.tile_par != 0;
    const int pc_step = tile_par ? 1 : 4;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint8_t* stg = smem_gen + stg_off + warp * kPwStgWarp;
    const uint32_t stg_u32 = smem_base + stg_off + warp * kPwStgWarp;
    uint32_t acc_iter = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++acc_iter) {
      if (tile_par && (int)(acc_iter & 3u) != res) continue;
      const long long m0 = (long long)(item / p.num_n_tiles) * kPwBM;
      const int n0 = (item % p.num_n_tiles) * BN;
      const uint32_t buf = acc_iter % nacc;
      const long long row_base = m0 + q * 32;
      const int pieces = min(BN, p.N - n0 + 15) >> 4;                  // 16-column pieces that hold at least one valid column
      // the first piece of a warp group rotates with the tile: with 9 pieces (the 144-wide expand layers) a fixed assignment gives
      // group 0 three pieces of EVERY tile and the others two (the tile is released when the slowest group is done); rotating, every
      // group carries the extra piece every fourth tile and the 3-slot accumulator ring absorbs the difference
      const int pc_first = tile_par ? 0 : (int)((res + acc_iter) & 3u);
      const int nmine = tile_par ? pieces : (pieces > pc_first ? (pieces - pc_first + 3) >> 2 : 0);    // this warp's pieces
      float4 rs[4];
      if (p.epi == kPwResidual && nmine > 0) {
        // residual of the first piece (this lane's row, 64 bytes): requested before the accumulator wait so its DRAM latency
        // hides behind the MMAs
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int c = n0 + pc_first * 16 + i4 * 4;
          rs[i4] = (row_base + lane < p.M && c < p.N) ? __ldg(reinterpret_cast<const float4*>(p.residual + (size_t)(row_base + lane) * p.N + c))
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      mbar_wait(acc_full_bar(buf), (acc_iter / nacc) & 1u);
      tc_fence_after();
      if (nmine == 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty_bar(buf));
        continue;
      }
      for (int g = 0; g < nmine; ++g) {
        const int pc = pc_first + pc_step * g;
        const int c0 = n0 + pc * 16;
        uint32_t tr[16];
        tmem_ld_32x16(tmem_base + lane_base + buf * acc_cols + pc * 16, tr);
        float4 b4[4];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4)
          b4[i4] = c0 + i4 * 4 < p.N ? __ldg(reinterpret_cast<const float4*>(p.bias + c0 + i4 * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (stack) {
          // the a_hi w_lo half of the stacked accumulator: bias joins the first half while the second is in flight, then the second
          // half takes the bias slot of the sum below
          tmem_wait_ld();
          uint32_t t2[16];
          tmem_ld_32x16(tmem_base + lane_base + buf * acc_cols + (uint32_t)BN + pc * 16, t2);
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            tr[i4 * 4 + 0] = __float_as_uint(__uint_as_float(tr[i4 * 4 + 0]) + b4[i4].x);
            tr[i4 * 4 + 1] = __float_as_uint(__uint_as_float(tr[i4 * 4 + 1]) + b4[i4].y);
            tr[i4 * 4 + 2] = __float_as_uint(__uint_as_float(tr[i4 * 4 + 2]) + b4[i4].z);
            tr[i4 * 4 + 3] = __float_as_uint(__uint_as_float(tr[i4 * 4 + 3]) + b4[i4].w);
          }
          tmem_wait_ld();
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4)
            b4[i4] = make_float4(__uint_as_float(t2[i4 * 4]), __uint_as_float(t2[i4 * 4 + 1]), __uint_as_float(t2[i4 * 4 + 2]), __uint_as_float(t2[i4 * 4 + 3]));
        }
        if (p.epi == kPwResidual && g > 0) {
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4)
            rs[i4] = (row_base + lane < p.M && c0 + i4 * 4 < p.N) ? __ldg(reinterpret_cast<const float4*>(p.residual + (size_t)(row_base + lane) * p.N + c0 + i4 * 4))
                                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tmem_wait_ld();
        if (g == nmine - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty_bar(buf));   // this warp has read everything it needs from the accumulator
        }
        float f[16];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          f[i4 * 4 + 0] = __uint_as_float(tr[i4 * 4 + 0]) + b4[i4].x;
          f[i4 * 4 + 1] = __uint_as_float(tr[i4 * 4 + 1]) + b4[i4].y;
          f[i4 * 4 + 2] = __uint_as_float(tr[i4 * 4 + 2]) + b4[i4].z;
          f[i4 * 4 + 3] = __uint_as_float(tr[i4 * 4 + 3]) + b4[i4].w;
        }
        if (p.epi == kPwSilu || p.epi == kPwSiluPool) {
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            if (p.silu_nr) silu2_nr(f[i], f[i + 1]);
            else silu2(f[i], f[i + 1]);
          }
        }
        if (p.epi == kPwSiluPool) {
          // column sums over this warp's 32 rows (all inside one image: hw % 128 == 0), one atomic per column
#pragma unroll
          for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
            const bool up = (lane & bit) != 0;
#pragma unroll
            for (int i = 0; i < w; ++i) {
              const float send = up ? f[i] : f[i + w];
              const float keep = up ? f[i + w] : f[i];
              f[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
            }
          }
          f[0] += __shfl_xor_sync(0xffffffffu, f[0], 1);   // lanes 2c and 2c+1 hold the two halves of column c
          const int col = c0 + (lane >> 1);
          if ((lane & 1) == 0 && col < p.N && row_base < p.M) atomicAdd(p.y + (row_base / p.hw) * p.N + col, f[0]);
          continue;
        }
        if (p.epi == kPwResidual) {
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            f[i4 * 4 + 0] += rs[i4].x; f[i4 * 4 + 1] += rs[i4].y; f[i4 * 4 + 2] += rs[i4].z; f[i4 * 4 + 3] += rs[i4].w;
          }
        }
        // the piece goes out as ONE bulk-tensor store from the warp's staging tile (the box of Y's tensor map, 64B swizzle: lane =
        // row, 16-byte chunk c at slot c ^ ((row >> 1) & 3) -> conflict-free); rows >= M and columns >= N are clipped by the map.
        // Before: a padded transpose tile, 4 LDS.128 + 4 predicated STG.128 per lane and two warp barriers per piece -- the LSU
        // data pipe was 65 % busy on the expand layers (ncu, profiles/r02_ncu_decoder_kernels_v10.txt).
        if (lane == 0) tma_store_wait_read<0>();   // the previous piece's store has read the tile
        __syncwarp();
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4)
          *reinterpret_cast<float4*>(stg + lane * 64 + ((i4 ^ ((lane >> 1) & 3)) << 4)) = make_float4(f[i4 * 4], f[i4 * 4 + 1], f[i4 * 4 + 2], f[i4 * 4 + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&p.tmap_y, stg_u32, c0, (int)row_base);
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_read<0>();   // every bulk store has read its staging tile before the shared memory goes away
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kPwMmaWarp) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------------------------
// CTA-pair edition (cluster of 2, tcgen05 cta_group::2) for the layers whose weights do NOT fit in shared memory -- the late,
// wide MBConv blocks (K = 480 ... 1920) and the head.  There the single-CTA kernel keeps ONE 32-deep k chunk in flight (A hi/lo
// 32 KB + W hi/lo up to 56 KB per stage = 2 stages) and every CTA streams the whole W (hi + lo) for its 128 rows: ncu
// (profiles/r02_ncu_decoder_kernels_v10.txt) shows the tensor pipe 39 - 47 % active, 42 warps per issue waiting, 5.6 TB/s of
// L2 -> SM traffic.  A pair shares W: each CTA loads half of the W tile (the MMA reads both halves), so a stage is 32 KB + <= 28 KB
// (3 stages), the W traffic per row halves, and M = 256 rows per MMA.
//   producer (both CTAs)   own 128 rows of A + own half of W_hi / W_lo, bytes on the CTA's OWN full barrier
//   converters (both CTAs) wait for the own full barrier (A and the W half have landed), split A, then arrive on the LEADER's
//                          conv barrier (8 arrivals = 4 warps x 2 CTAs): its completion covers both CTAs' operands
//   MMA (leader CTA)       3 x tcgen05.mma.cta_group::2.kind::tf32 per k-step, commits multicast to both CTAs' empty / acc_full barriers
//   epilogue (both CTAs)   own 128 TMEM lanes; "accumulator drained" arrives on the leader's barrier (32 arrivals)
// ------------------------------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPwThreads, 1) pointwise_tc_pair_kernel(const __grid_constant__ PwTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;

  const int BN = p.BN;
  const int stages = p.stages;
  const uint32_t w_bytes = (uint32_t)(BN / 2) * 128u;              // this CTA's half of one W tile (hi or lo)
  const uint32_t stage_bytes = 2u * kPwATile + 2u * w_bytes;
  const uint32_t stg_off = (uint32_t)stages * stage_bytes;
  const uint32_t bar_off = stg_off + (uint32_t)kPwEpiWarps * kPwStgWarp;
  const uint32_t bar_base = smem_base + bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                 // own TMA bytes landed (local)
  auto conv_bar = [&](int s) { return bar_base + 8u * (8 + s); };           // leader: both CTAs' operands ready (8 warp arrivals)
  auto empty_bar = [&](int s) { return bar_base + 8u * (16 + s); };         // both CTAs (multicast commit)
  auto acc_full_bar = [&](int b) { return bar_base + 8u * (24 + b); };      // both CTAs (multicast commit)
  auto acc_empty_bar = [&](int b) { return bar_base + 8u * (32 + b); };     // leader: 32 warp arrivals
  const uint32_t tmem_slot = bar_base + 8u * 41;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + bar_off + 8 * 41);
  const uint32_t nacc = min(8u, 512u / (uint32_t)BN);

  auto a_hi = [&](int s) { return smem_base + (uint32_t)s * stage_bytes; };
  auto a_lo = [&](int s) { return smem_base + (uint32_t)s * stage_bytes + kPwATile; };
  auto w_hi = [&](int s) { return smem_base + (uint32_t)s * stage_bytes + 2u * kPwATile; };
  auto w_lo = [&](int s) { return w_hi(s) + w_bytes; };

  if (warp == kPwProducerWarp && lane == 0) {
    tma_prefetch_desc(&p.tmap_x);
    tma_prefetch_desc(&p.tmap_whi);
    tma_prefetch_desc(&p.tmap_wlo);
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(conv_bar(s), 2 * kPwConvWarps);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 8; ++b) {
      mbar_init(acc_full_bar(b), 1);
      mbar_init(acc_empty_bar(b), 2 * kPwEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == kPwMmaWarp) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers exist before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int num_m_pairs = (p.num_m_tiles + 1) >> 1;
  const int total_items = num_m_pairs * p.num_n_tiles;
  const int num_kc = p.num_kc;
  const int item0 = blockIdx.x >> 1, item_step = gridDim.x >> 1;

  if (warp == kPwProducerWarp) {
    // =========================== TMA producer (both CTAs) ===========================
    int stage = 0;
    uint32_t phase = 0;
    for (int item = item0; item < total_items; item += item_step) {
      const int m0 = ((item / p.num_n_tiles) * 2 + (int)cta_rank) * kPwBM;
      const int n0 = (item % p.num_n_tiles) * BN + (int)cta_rank * (BN / 2);
      for (int kc = 0; kc < num_kc; ++kc) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(full_bar(stage), kPwATile + 2u * w_bytes);
          tma_load_2d(a_hi(stage), &p.tmap_x, full_bar(stage), kc * kPwKC, m0);
          tma_load_2d(w_hi(stage), &p.tmap_whi, full_bar(stage), kc * kPwKC, n0);
          tma_load_2d(w_lo(stage), &p.tmap_wlo, full_bar(stage), kc * kPwKC, n0);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == kPwMmaWarp) {
    // =========================== MMA issuer (leader CTA) ===========================
    if (leader) {
      const uint32_t idesc = make_idesc_tf32(2 * kPwBM, (uint32_t)BN);
      constexpr uint64_t kDescHi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      auto desc = [&](uint32_t addr) { return kDescHi | (uint64_t)(((addr >> 4) & 0x3FFFu) | (1u << 16)); };
      int stage = 0;
      uint32_t phase = 0, acc_iter = 0;
      for (int item = item0; item < total_items; item += item_step, ++acc_iter) {
        const uint32_t buf = acc_iter % nacc;
        const uint32_t tmem_acc = tmem_base + buf * (uint32_t)BN;
        mbar_wait(acc_empty_bar(buf), ((acc_iter / nacc) & 1u) ^ 1u);
        tc_fence_after();
        for (int kc = 0; kc < num_kc; ++kc) {
          mbar_wait(conv_bar(stage), phase);   // both CTAs: A split (generic proxy + fence.proxy.async) and W halves landed
          tc_fence_after();
          if (elect_one()) {
            const int ksteps = min(kPwKC, p.K - kc * kPwKC) >> 3;
            const uint64_t ah = desc(a_hi(stage)), al = desc(a_lo(stage)), wh = desc(w_hi(stage)), wl = desc(w_lo(stage));
            for (int k = 0; k < ksteps; ++k) {
              umma_tf32_pair(tmem_acc, al + 2 * k, wh + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
              umma_tf32_pair(tmem_acc, ah + 2 * k, wl + 2 * k, idesc, 1u);
              umma_tf32_pair(tmem_acc, ah + 2 * k, wh + 2 * k, idesc, 1u);
            }
            umma_commit_pair(empty_bar(stage), 0x3);
            if (kc == num_kc - 1) umma_commit_pair(acc_full_bar(buf), 0x3);
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= kPwConvWarp0) {
    // =========================== converter warps (both CTAs): own A rows -> (A * se) hi / lo ===========================
    const int t = threadIdx.x - kPwConvWarp0 * 32;
    const int cphys = t & 7;
    const int rlow = (t >> 3) & 7;
    const int clog = cphys ^ rlow;
    int stage = 0;
    uint32_t phase = 0;
    for (int item = item0; item < total_items; item += item_step) {
      const long long m0 = (long long)((item / p.num_n_tiles) * 2 + (int)cta_rank) * kPwBM;
      const float* se_row = (p.se != nullptr && m0 < p.M) ? p.se + (m0 / p.hw) * p.K : nullptr;
      for (int kc = 0; kc < num_kc; ++kc) {
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
        const int k = kc * kPwKC + clog * 4;
        if (se_row != nullptr && k < p.K) sc = __ldg(reinterpret_cast<const float4*>(se_row + k));
        mbar_wait(full_bar(stage), phase);
        uint8_t* hi = smem_gen + (a_hi(stage) - smem_base);
        uint8_t* lo = smem_gen + (a_lo(stage) - smem_base);
        if (k < p.K) {
          float4 v[kPwConvIters];
#pragma unroll
          for (int j = 0; j < kPwConvIters; ++j) v[j] = *reinterpret_cast<const float4*>(hi + (t + kPwConvThreads * j) * 16);
#pragma unroll
          for (int j = 0; j < kPwConvIters; ++j) {
            float x[4] = {v[j].x * sc.x, v[j].y * sc.y, v[j].z * sc.z, v[j].w * sc.w};
            float h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              h[e] = __uint_as_float(__float_as_uint(x[e]) & 0xFFFFE000u);
              l[e] = x[e] - h[e];
            }
            *reinterpret_cast<float4*>(hi + (t + kPwConvThreads * j) * 16) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(lo + (t + kPwConvThreads * j) * 16) = make_float4(l[0], l[1], l[2], l[3]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_shared(conv_bar(stage), 0));
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // =========================== epilogue warps (both CTAs; same as the single-CTA kernel on the CTA's own 128 rows) ===========
    const int q = warp & 3;
    const int res = warp >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint8_t* stg = smem_gen + stg_off + warp * kPwStgWarp;
    const uint32_t stg_u32 = smem_base + stg_off + warp * kPwStgWarp;
    uint32_t acc_iter = 0;
    for (int item = item0; item < total_items; item += item_step, ++acc_iter) {
      const long long m0 = (long long)((item / p.num_n_tiles) * 2 + (int)cta_rank) * kPwBM;
      const int n0 = (item % p.num_n_tiles) * BN;
      const uint32_t buf = acc_iter % nacc;
      const uint32_t acc_empty_remote = mapa_shared(acc_empty_bar(buf), 0);
      const long long row_base = m0 + q * 32;
      const int pieces = min(BN, p.N - n0 + 15) >> 4;
      const int nmine = pieces > res ? (pieces - res + 3) >> 2 : 0;
      float4 rs[4];
      if (p.epi == kPwResidual && nmine > 0) {
        // residual of the first piece (this lane's row, 64 bytes): requested before the accumulator wait so its DRAM latency
        // hides behind the MMAs
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int c = n0 + res * 16 + i4 * 4;
          rs[i4] = (row_base + lane < p.M && c < p.N) ? __ldg(reinterpret_cast<const float4*>(p.residual + (size_t)(row_base + lane) * p.N + c))
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      mbar_wait(acc_full_bar(buf), (acc_iter / nacc) & 1u);
      tc_fence_after();
      if (nmine == 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(acc_empty_remote);
        continue;
      }
      for (int g = 0; g < nmine; ++g) {
        const int pc = res + 4 * g;
        const int c0 = n0 + pc * 16;
        uint32_t tr[16];
        tmem_ld_32x16(tmem_base + lane_base + buf * (uint32_t)BN + pc * 16, tr);
        float4 b4[4];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4)
          b4[i4] = c0 + i4 * 4 < p.N ? __ldg(reinterpret_cast<const float4*>(p.bias + c0 + i4 * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.epi == kPwResidual && g > 0) {
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4)
            rs[i4] = (row_base + lane < p.M && c0 + i4 * 4 < p.N) ? __ldg(reinterpret_cast<const float4*>(p.residual + (size_t)(row_base + lane) * p.N + c0 + i4 * 4))
                                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tmem_wait_ld();
        if (g == nmine - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(acc_empty_remote);
        }
        float f[16];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          f[i4 * 4 + 0] = __uint_as_float(tr[i4 * 4 + 0]) + b4[i4].x;
          f[i4 * 4 + 1] = __uint_as_float(tr[i4 * 4 + 1]) + b4[i4].y;
          f[i4 * 4 + 2] = __uint_as_float(tr[i4 * 4 + 2]) + b4[i4].z;
          f[i4 * 4 + 3] = __uint_as_float(tr[i4 * 4 + 3]) + b4[i4].w;
        }
        if (p.epi == kPwSilu || p.epi == kPwSiluPool) {
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            if (p.silu_nr) silu2_nr(f[i], f[i + 1]);
            else silu2(f[i], f[i + 1]);
          }
        }
        if (p.epi == kPwSiluPool) {
#pragma unroll
          for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
            const bool up = (lane & bit) != 0;
#pragma unroll
            for (int i = 0; i < w; ++i) {
              const float send = up ? f[i] : f[i + w];
              const float keep = up ? f[i + w] : f[i];
              f[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
            }
          }
          f[0] += __shfl_xor_sync(0xffffffffu, f[0], 1);
          const int col = c0 + (lane >> 1);
          if ((lane & 1) == 0 && col < p.N && row_base < p.M) atomicAdd(p.y + (row_base / p.hw) * p.N + col, f[0]);
          continue;
        }
        if (p.epi == kPwResidual) {
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            f[i4 * 4 + 0] += rs[i4].x; f[i4 * 4 + 1] += rs[i4].y; f[i4 * 4 + 2] += rs[i4].z; f[i4 * 4 + 3] += rs[i4].w;
          }
        }
        // the piece goes out as ONE bulk-tensor store from the warp's staging tile (the box of Y's tensor map, 64B swizzle: lane =
        // row, 16-byte chunk c at slot c ^ ((row >> 1) & 3) -> conflict-free); rows >= M and columns >= N are clipped by the map.
        // Before: a padded transpose tile, 4 LDS.128 + 4 predicated STG.128 per lane and two warp barriers per piece -- the LSU
        // data pipe was 65 % busy on the expand layers (ncu, profiles/r02_ncu_decoder_kernels_v10.txt).
        if (lane == 0) tma_store_wait_read<0>();   // the previous piece's store has read the tile
        __syncwarp();
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4)
          *reinterpret_cast<float4*>(stg + lane * 64 + ((i4 ^ ((lane >> 1) & 3)) << 4)) = make_float4(f[i4 * 4], f[i4 * 4 + 1], f[i4 * 4 + 2], f[i4 * 4 + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&p.tmap_y, stg_u32, c0, (int)row_base);
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_read<0>();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA leaves while its pair may still read its shared memory or signal its barriers
  if (warp == kPwMmaWarp) tmem_dealloc_pair(tmem_base, 512);
}

// AQ_PW_PAIR=0 keeps every layer on the single-CTA kernel (A/B measurements)
static bool pw_pair_enabled() {
  static const bool on = [] { const char* e = getenv("AQ_PW_PAIR"); return e == nullptr || e[0] != '0'; }();
  return on;
}

// AQ_PW_STACK=0: three separate MMAs per k step everywhere (A/B measurements)
static bool pw_stack_enabled() {
  static const bool on = [] { const char* e = getenv("AQ_PW_STACK"); return e == nullptr || e[0] != '0'; }();
  return on;
}

// AQ_SILU_NR=0 / 1: SiLU epilogues with two SFU operations per element / with the Newton reciprocal (A/B measurements)
static bool pw_silu_nr() {
  static const bool on = [] { const char* e = getenv("AQ_SILU_NR"); return e != nullptr && e[0] == '1'; }();
  return on;
}

// AQ_PW_WIDE_RES=0: N = 240 layers as two tiles on the CTA-pair kernel (A/B measurements)
static bool pw_wide_resident() {
  static const bool on = [] { const char* e = getenv("AQ_PW_WIDE_RES"); return e == nullptr || e[0] != '0'; }();
  return on;
}

static int pick_pw_bn(int N) {
  const int n16 = (N + 15) / 16 * 16;
  if (n16 <= kPwMaxBN) return n16;
  // split into the fewest equal tiles of <= kPwMaxBN columns (multiples of 16)
  for (int parts = 2; parts <= 32; ++parts) {
    const int bn = ((N + parts - 1) / parts + 15) / 16 * 16;
    if (bn <= kPwMaxBN) return bn;
  }
  return kPwMaxBN;
}

int launch_pointwise_tc(const PwTcArgs& a, cudaStream_t st) {
  AQ_REQUIRE(a.M > 0 && a.K > 0 && a.N > 0, AQ_ERR_BAD_SHAPE, "pointwise: empty problem");
  AQ_REQUIRE(a.K % 8 == 0 && a.N % 4 == 0, AQ_ERR_BAD_SHAPE, "pointwise: K=%d must be a multiple of 8 and N=%d of 4", a.K, a.N);
  AQ_REQUIRE((a.se == nullptr && a.epi != kPwSiluPool) || a.hw % kPwBM == 0, AQ_ERR_BAD_SHAPE,
             "pointwise: %d pixels per image is not a multiple of the %d-row tile", a.hw, kPwBM);
  PwTcParams p;
  memset(&p, 0, sizeof(p));
  p.BN = pick_pw_bn(a.N);
  {
    // wide but shallow layers (40 -> 240): ONE column tile whose weights stay in shared memory (120 KB of W hi / lo + 2 stages of A)
    // instead of two 128-column tiles on the CTA-pair kernel, which re-streamed 122 KB of W per 256-row item against 64 KB of A
    const int n16 = (a.N + 15) / 16 * 16, kc = (a.K + kPwKC - 1) / kPwKC;
    if (n16 > kPwMaxBN && n16 <= 256 && kc * 2 * n16 * 128 <= kPwResidentMax && a.epi != kPwSiluPool && pw_wide_resident()) p.BN = n16;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.M};
    uint64_t str[1] = {(uint64_t)a.K * 4};
    uint32_t box[2] = {kPwKC, kPwBM};
    int rc = make_tmap(&p.tmap_x, a.x, 4, 2, dims, str, box, kSwz128);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.N};
    uint64_t str[1] = {(uint64_t)a.K * 4};
    uint32_t box[2] = {kPwKC, (uint32_t)p.BN};
    int rc = make_tmap(&p.tmap_whi, a.w_hi, 4, 2, dims, str, box, kSwz128);
    if (rc) return rc;
    rc = make_tmap(&p.tmap_wlo, a.w_lo, 4, 2, dims, str, box, kSwz128);
    if (rc) return rc;
  }
  p.bias = a.bias; p.se = a.se; p.residual = a.residual; p.y = a.y;
  p.M = a.M; p.N = a.N; p.K = a.K; p.hw = a.hw; p.epi = a.epi;
  if (a.epi != kPwSiluPool) {
    AQ_REQUIRE((reinterpret_cast<uintptr_t>(a.y) & 15u) == 0, AQ_ERR_BAD_ALIGN, "pointwise: y must be 16-byte aligned");
    uint64_t dims[2] = {(uint64_t)a.N, (uint64_t)a.M};
    uint64_t str[1] = {(uint64_t)a.N * 4};
    uint32_t box[2] = {16, 32};
    int rc = make_tmap(&p.tmap_y, a.y, 4, 2, dims, str, box, kSwz64);
    if (rc) return rc;
  }
  p.num_n_tiles = (a.N + p.BN - 1) / p.BN;
  p.num_m_tiles = (int)((a.M + kPwBM - 1) / kPwBM);
  p.num_kc = (a.K + kPwKC - 1) / kPwKC;
  const int w_tile_bytes = 2 * p.BN * 128;                                   // hi + lo of one k chunk
  const int fixed = kPwEpiWarps * kPwStgWarp + 512;
  p.w_resident = (p.num_n_tiles == 1 && p.num_kc * w_tile_bytes <= (p.BN > kPwMaxBN ? kPwResidentMax : 96 * 1024)) ? 1 : 0;
  // narrow outputs (<= 3 pieces: 8-slot ring) and the 96-wide expand (6 pieces over 4 warp groups = 2, 2, 1, 1 per tile; 4-slot ring)
  p.tile_par = ((p.BN <= 48 || p.BN == 96) && a.epi != kPwSiluPool) ? 1 : 0;
  p.silu_nr = pw_silu_nr() ? 1 : 0;
  // stacked B operand for the narrow layers with a deep product (the project convolutions): 2 BN <= 96 accumulator columns x 4 - 8 slots
  p.stack = (p.tile_par && p.BN <= 48 && a.K >= 32 && p.num_n_tiles == 1 && pw_stack_enabled()) ? 1 : 0;
  const int sms = sm_count();
  if (sms <= 0) return fail(AQ_ERR_LAUNCH, "no CUDA device");
  if (!p.w_resident && p.num_m_tiles >= 2 && pw_pair_enabled()) {
    // CTA pairs: each CTA holds half of the W tile (see pointwise_tc_pair_kernel)
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.N};
    uint64_t str[1] = {(uint64_t)a.K * 4};
    uint32_t box[2] = {kPwKC, (uint32_t)(p.BN / 2)};
    int rc = make_tmap(&p.tmap_whi, a.w_hi, 4, 2, dims, str, box, kSwz128);
    if (rc) return rc;
    rc = make_tmap(&p.tmap_wlo, a.w_lo, 4, 2, dims, str, box, kSwz128);
    if (rc) return rc;
    const int stage_bytes = 2 * kPwATile + w_tile_bytes / 2;
    int stages = (kPwSmemBudget - fixed) / stage_bytes;
    if (stages > 6) stages = 6;
    AQ_REQUIRE(stages >= 2, AQ_ERR_BAD_SHAPE, "pointwise: column tile %d leaves no room for a pipeline", p.BN);
    p.stages = stages;
    const int smem = stages * stage_bytes + fixed + 1024;
    const long long items = (long long)((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
    const int slots = sms / 2;
    const int grid = 2 * (int)(items < slots ? items : slots);
    AQ_OPT_IN_SMEM((pointwise_tc_pair_kernel), 232448);
    pointwise_tc_pair_kernel<<<grid, kPwThreads, smem, st>>>(p);
    AQ_LAUNCHED();
    return AQ_OK;
  }
  const int wres_bytes = p.w_resident ? p.num_kc * w_tile_bytes : 0;
  const int stage_bytes = 2 * kPwATile + (p.w_resident ? 0 : w_tile_bytes);
  int stages = (kPwSmemBudget - fixed - wres_bytes) / stage_bytes;
  if (stages > 6) stages = 6;
  AQ_REQUIRE(stages >= 2, AQ_ERR_BAD_SHAPE, "pointwise: column tile %d leaves no room for a pipeline", p.BN);
  p.stages = stages;
  const int smem = wres_bytes + stages * stage_bytes + fixed + 1024;
  const long long items = (long long)p.num_m_tiles * p.num_n_tiles;
  const int grid = (int)(items < sms ? items : sms);
  AQ_OPT_IN_SMEM((pointwise_tc_kernel), 232448);
  pointwise_tc_kernel<<<grid, kPwThreads, smem, st>>>(p);
  AQ_LAUNCHED();
  return AQ_OK;
}

}  // namespace aq
