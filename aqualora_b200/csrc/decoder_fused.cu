// Fused MBConv front half of the message decoder: expand 1x1 (+ folded BN + SiLU) -> depthwise k x k stride s (+ folded BN + SiLU)
// with the squeeze sums of the SE block, for the early blocks whose 6x expanded activation dominates the decoder's HBM traffic
// (torchvision MBConv.forward as used by utils/models.py:88-96).
//
//   E[p, c]  = SiLU(b_e[c] + sum_k x[p, k] * W_e[c, k])        for the pixels p of an output tile's input window (zero outside the image)
//   y[o, c]  = SiLU(b_d[c] + sum_{ky, kx} W_d[ky, kx, c] * E[o * s + (ky, kx) - pad, c]) ;   pooled[n, c] += sum_o y[o, c]
//
// The expanded tensor E never leaves the SM: per layer it was written once and read once through HBM (25 MB / image for the first
// stage-2 block at 256 x 256 x 96 fp32 -- 45 % of the decoder's algorithmic bytes over all blocks), and the unfused pair
// (pointwise_tc_kernel + depthwise kernel) took 1.02 ms per 64 images for that block alone (profiles/r02_decoder_launches_v10.txt).
//
// Arithmetic: the 1x1 product is evaluated exactly like the stand-alone pointwise kernel (decoder_pw.cu): 3-term TF32 split
// a_lo w_hi + a_hi w_lo + a_hi w_hi with fp32 accumulation -- here on warp-level mma.sync.m16n8k8 (the operands come from shared
// memory fragments and the result is consumed by the same warp's registers, which is what the depthwise stage needs; K is only
// 16 ... 24 deep, so the tensor work is ~2 % of the kernel and tcgen05 / TMEM staging would buy nothing).  The depthwise stage is
// fp32 FMA in the same (ky, kx) order as the stand-alone kernels.
//
// One persistent CTA walks a CONTIGUOUS range of output tiles (TH x TW pixels, all channels):
//   phase A  input window [(TH-1)s+k] x [(TW-1)s+k] pixels x CIN -> shared memory (zero outside the image)
//   phase B  a warp per 16-pixel row block: A fragments (hi / lo) in registers, loop over the 8-channel column blocks: 3 mma per
//            k-step, + bias, SiLU, zero for out-of-image pixels, float2 stores into the expanded window E [pixels][CEXP + 8]
//   phase C  a thread owns 4 channels (its depthwise weights live in registers for k = 3) and strides over the output pixels:
//            k*k float4 reads of E, SiLU, float4 store (consecutive threads = consecutive channel quads of a pixel: 384 / 576 B
//            runs), squeeze sums in registers, flushed with atomics only when the CTA's range crosses into the next image.
#include <string.h>

#include "aq_ptx.cuh"
#include "decoder_fused.h"

namespace aq {

namespace {

__device__ __forceinline__ float fz_silu(float v) {
  float e, r;   // same SFU form as decoder.cu / decoder_pw.cu
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return v * r;
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// exact TF32 split: hi = x with the low 13 mantissa bits cleared, lo = x - hi (exact in fp32)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

struct FusedParams {
  const float* x;      // [B, H, W, CIN]
  const float* w_hi;   // [CEXP, CIN]
  const float* w_lo;   // [CEXP, CIN]
  const float* b_e;    // [CEXP]
  const float* w_d;    // [k * k, CEXP]
  const float* b_d;    // [CEXP]
  float* y;            // [B, Ho, Wo, CEXP]
  float* pooled;       // [B, CEXP]
  int B, H, W, Ho, Wo;
  int tiles_x, tiles_per_img, total_tiles, tiles_per_cta;
};

template <int CIN, int CEXP, int KS, int S, int TH, int TW, int NT>
struct FusedCfg {
  static constexpr int PAD = (KS - 1) / 2;
  static constexpr int PH = (TH - 1) * S + KS, PW = (TW - 1) * S + KS, P = PH * PW;
  static constexpr int MT = (P + 15) / 16;      // 16-pixel row blocks of the expand product
  static constexpr int XS = CIN + 4;            // row stride (floats) of the input window and of W_e: conflict-free fragment loads
  static constexpr int ES = CEXP + 8;           // row stride of E: 8 lanes x 4 column pairs of a fragment store hit 32 distinct banks
  static constexpr int CQ = CEXP / 4, NW = NT / 32, KSTEPS = CIN / 8;
  static constexpr int kXsFloats = MT * 16 * XS, kWsFloats = 2 * CEXP * XS, kEsFloats = MT * 16 * ES, kWdFloats = KS * KS * CEXP;
  static constexpr int kSmemBytes = (kXsFloats + kWsFloats + kEsFloats + kWdFloats + 2 * CEXP) * 4;
  static_assert(CIN % 8 == 0 && CEXP % 8 == 0 && NT % 32 == 0 && NT % CQ == 0, "shape / thread-count constraints");
  static_assert((XS % 32 == 20 || XS % 32 == 28) && (ES % 32 == 8 || ES % 32 == 24), "bank-conflict-free strides");
};

template <int CIN, int CEXP, int KS, int S, int TH, int TW, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) expand_dw_kernel(const FusedParams p) {
  using C = FusedCfg<CIN, CEXP, KS, S, TH, TW, NT>;
  extern __shared__ __align__(16) float smem_f[];
  float* xs = smem_f;                         // [MT * 16][XS]
  float* ws = xs + C::kXsFloats;              // [2][CEXP][XS]   W_e as its TF32 hi / lo planes (the packed weights, row-padded)
  float* es = ws + C::kWsFloats;              // [P][ES]         expanded window
  float* wd = es + C::kEsFloats;              // [KS * KS][CEXP]
  float* be = wd + C::kWdFloats;              // [CEXP]
  float* bd = be + CEXP;                      // [CEXP]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;      // mma fragment coordinates: group id, thread in group

  for (int i = tid; i < CEXP * CIN; i += NT) {
    const int n = i / CIN, k = i - n * CIN;
    ws[n * C::XS + k] = __ldg(p.w_hi + i);
    ws[(CEXP + n) * C::XS + k] = __ldg(p.w_lo + i);
  }
  for (int i = tid; i < C::kWdFloats; i += NT) wd[i] = __ldg(p.w_d + i);
  for (int i = tid; i < CEXP; i += NT) {
    be[i] = __ldg(p.b_e + i);
    bd[i] = __ldg(p.b_d + i);
  }
  // rows of the input window past P (the last row block is padded to 16) stay zero for the whole kernel
  for (int i = tid; i < (C::MT * 16 - C::P) * C::XS; i += NT) xs[C::P * C::XS + i] = 0.f;
  __syncthreads();

  // phase C ownership: 4 channels per thread, fixed for the whole kernel
  const int cq = tid % C::CQ, op0 = tid / C::CQ;
  constexpr int kOpStep = NT / C::CQ;
  float4 pool = make_float4(0.f, 0.f, 0.f, 0.f);
  int pool_n = -1;
  auto flush_pool = [&]() {
    if (pool_n >= 0) {
      float* dst = p.pooled + (size_t)pool_n * CEXP + 4 * cq;
      atomicAdd(dst + 0, pool.x); atomicAdd(dst + 1, pool.y); atomicAdd(dst + 2, pool.z); atomicAdd(dst + 3, pool.w);
    }
    pool = make_float4(0.f, 0.f, 0.f, 0.f);
  };

  const int tile_begin = blockIdx.x * p.tiles_per_cta;
  const int tile_end = min(tile_begin + p.tiles_per_cta, p.total_tiles);
  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int n = tile / p.tiles_per_img, rem = tile - n * p.tiles_per_img;
    const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
    const int oy0 = ty * TH, ox0 = tx * TW;
    const int gy0 = oy0 * S - C::PAD, gx0 = ox0 * S - C::PAD;   // image coordinates of the window's first pixel
    if (n != pool_n) {
      flush_pool();
      pool_n = n;
    }

    // ---------------- phase A: input window -> shared memory ----------------
    {
      constexpr int C4 = CIN / 4;
      const float* xn = p.x + (size_t)n * p.H * p.W * CIN;
      for (int i = tid; i < C::P * C4; i += NT) {
        const int r = i / C4, c4 = i - r * C4;
        const int py = r / C::PW, px = r - py * C::PW;
        const int gy = gy0 + py, gx = gx0 + px;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) v = __ldg(reinterpret_cast<const float4*>(xn + ((size_t)gy * p.W + gx) * CIN + 4 * c4));
        *reinterpret_cast<float4*>(xs + r * C::XS + 4 * c4) = v;
      }
    }
    __syncthreads();

    // ---------------- phase B: E = SiLU(x W_e^T + b_e) on the window (zero outside the image) ----------------
    for (int mt = warp; mt < C::MT; mt += C::NW) {
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      // out-of-image pixels of the window are ZERO in E (the depthwise convolution pads the expanded map, not the input):
      // a 0 / 1 multiplier instead of a branch -- divergence bookkeeping was 13 % of this phase's instructions
      float m0, m1;
      {
        const int py0 = r0 / C::PW, px0 = r0 - py0 * C::PW, py1 = r1 / C::PW, px1 = r1 - py1 * C::PW;
        m0 = (r0 < C::P && gy0 + py0 >= 0 && gy0 + py0 < p.H && gx0 + px0 >= 0 && gx0 + px0 < p.W) ? 1.f : 0.f;
        m1 = (r1 < C::P && gy0 + py1 >= 0 && gy0 + py1 < p.H && gx0 + px1 >= 0 && gx0 + px1 < p.W) ? 1.f : 0.f;
      }
      uint32_t ahi[C::KSTEPS][4], alo[C::KSTEPS][4];
#pragma unroll
      for (int ks = 0; ks < C::KSTEPS; ++ks) {
        split_tf32(xs[r0 * C::XS + ks * 8 + t], ahi[ks][0], alo[ks][0]);
        split_tf32(xs[r1 * C::XS + ks * 8 + t], ahi[ks][1], alo[ks][1]);
        split_tf32(xs[r0 * C::XS + ks * 8 + t + 4], ahi[ks][2], alo[ks][2]);
        split_tf32(xs[r1 * C::XS + ks * 8 + t + 4], ahi[ks][3], alo[ks][3]);
      }
      float* e0 = es + r0 * C::ES + 2 * t;
      float* e1 = es + r1 * C::ES + 2 * t;
#pragma unroll 3
      for (int nt = 0; nt < CEXP / 8; ++nt) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        const uint32_t* whi = reinterpret_cast<const uint32_t*>(ws) + (nt * 8 + g) * C::XS + t;
        const uint32_t* wlo = whi + CEXP * C::XS;
#pragma unroll
        for (int ks = 0; ks < C::KSTEPS; ++ks) {
          const uint32_t bhi[2] = {whi[ks * 8], whi[ks * 8 + 4]};
          const uint32_t blo[2] = {wlo[ks * 8], wlo[ks * 8 + 4]};
          // small terms first, the dominant hi * hi product last (as in decoder_pw.cu)
          mma_tf32(c, alo[ks], bhi);
          mma_tf32(c, ahi[ks], blo);
          mma_tf32(c, ahi[ks], bhi);
        }
        const float2 b2 = *reinterpret_cast<const float2*>(be + nt * 8 + 2 * t);
        *reinterpret_cast<float2*>(e0 + nt * 8) = make_float2(fz_silu(c[0] + b2.x) * m0, fz_silu(c[1] + b2.y) * m0);
        *reinterpret_cast<float2*>(e1 + nt * 8) = make_float2(fz_silu(c[2] + b2.x) * m1, fz_silu(c[3] + b2.y) * m1);
      }
    }
    __syncthreads();

    // ---------------- phase C: depthwise k x k stride s + SiLU + squeeze sums ----------------
    {
      // the thread's depthwise weights (k = 3: registers for the whole phase; reloaded per tile so that they are not live across
      // the register-hungry phase B)
      float4 wreg[KS == 3 ? 9 : 1];
      if (KS == 3) {
#pragma unroll
        for (int i = 0; i < 9; ++i) wreg[i] = *reinterpret_cast<const float4*>(wd + i * CEXP + 4 * cq);
      }
      const float4 bd4 = *reinterpret_cast<const float4*>(bd + 4 * cq);
      float* yn = p.y + ((size_t)n * p.Ho * p.Wo) * CEXP + 4 * cq;
      for (int op = op0; op < TH * TW; op += kOpStep) {
        const int oy = op / TW, ox = op - oy * TW;
        const float* e0 = es + ((oy * S) * C::PW + ox * S) * C::ES + 4 * cq;
        float4 acc = bd4;
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
#pragma unroll
          for (int kx = 0; kx < KS; ++kx) {
            const float4 e = *reinterpret_cast<const float4*>(e0 + (ky * C::PW + kx) * C::ES);
            const float4 w = KS == 3 ? wreg[KS == 3 ? ky * 3 + kx : 0] : *reinterpret_cast<const float4*>(wd + (ky * KS + kx) * CEXP + 4 * cq);
            acc.x = fmaf(e.x, w.x, acc.x);
            acc.y = fmaf(e.y, w.y, acc.y);
            acc.z = fmaf(e.z, w.z, acc.z);
            acc.w = fmaf(e.w, w.w, acc.w);
          }
        }
        acc.x = fz_silu(acc.x); acc.y = fz_silu(acc.y); acc.z = fz_silu(acc.z); acc.w = fz_silu(acc.w);
        pool.x += acc.x; pool.y += acc.y; pool.z += acc.z; pool.w += acc.w;
        *reinterpret_cast<float4*>(yn + ((size_t)(oy0 + oy) * p.Wo + ox0 + ox) * CEXP) = acc;
      }
    }
    __syncthreads();   // E and the input window are free for the next tile
  }
  flush_pool();
}

template <int CIN, int CEXP, int KS, int S, int TH, int TW, int NT, int MINB>
int launch_cfg(const FusedArgs& a, cudaStream_t st) {
  using C = FusedCfg<CIN, CEXP, KS, S, TH, TW, NT>;
  const int Ho = (a.H + 2 * C::PAD - KS) / S + 1;
  AQ_REQUIRE(Ho % TH == 0 && Ho % TW == 0, AQ_ERR_BAD_SHAPE, "expand_dw: output %d x %d is not a multiple of the %d x %d tile", Ho, Ho, TH, TW);
  FusedParams p;
  memset(&p, 0, sizeof(p));
  p.x = a.x; p.w_hi = a.w_hi; p.w_lo = a.w_lo; p.b_e = a.b_e; p.w_d = a.w_d; p.b_d = a.b_d; p.y = a.y; p.pooled = a.pooled;
  p.B = a.B; p.H = a.H; p.W = a.H; p.Ho = Ho; p.Wo = Ho;
  p.tiles_x = Ho / TW;
  p.tiles_per_img = p.tiles_x * (Ho / TH);
  const long long total = (long long)a.B * p.tiles_per_img;
  AQ_REQUIRE(total < (1ll << 31), AQ_ERR_BAD_SHAPE, "expand_dw: too many tiles");
  p.total_tiles = (int)total;
  const int sms = sm_count();
  if (sms <= 0) return fail(AQ_ERR_LAUNCH, "no CUDA device");
  long long ctas = (long long)sms * MINB;
  if (ctas > total) ctas = total;
  p.tiles_per_cta = (int)((total + ctas - 1) / ctas);
  const int grid = (int)((total + p.tiles_per_cta - 1) / p.tiles_per_cta);
  AQ_OPT_IN_SMEM((expand_dw_kernel<CIN, CEXP, KS, S, TH, TW, NT, MINB>), C::kSmemBytes);
  expand_dw_kernel<CIN, CEXP, KS, S, TH, TW, NT, MINB><<<grid, NT, C::kSmemBytes, st>>>(p);
  AQ_LAUNCHED();
  return AQ_OK;
}

}  // namespace

bool fused_expand_dw_supported(int cin, int cexp, int k, int stride, int H) {
  if (cin == 16 && cexp == 96 && k == 3 && stride == 2) return (H / 2) % 8 == 0;
  if (cin == 24 && cexp == 144 && k == 3 && stride == 1) return H % 16 == 0;
  if (cin == 24 && cexp == 144 && k == 5 && stride == 2) return (H / 2) % 8 == 0;
  return false;
}

int launch_fused_expand_dw(const FusedArgs& a, cudaStream_t st) {
  AQ_REQUIRE(a.x && a.w_hi && a.w_lo && a.b_e && a.w_d && a.b_d && a.y && a.pooled && a.B > 0, AQ_ERR_BAD_SHAPE, "expand_dw: NULL operand or empty batch");
  AQ_REQUIRE(fused_expand_dw_supported(a.cin, a.cexp, a.k, a.stride, a.H), AQ_ERR_BAD_SHAPE,
             "expand_dw: no fused kernel for cin=%d cexp=%d k=%d stride=%d H=%d", a.cin, a.cexp, a.k, a.stride, a.H);
  AQ_REQUIRE(((reinterpret_cast<uintptr_t>(a.x) | reinterpret_cast<uintptr_t>(a.y) | reinterpret_cast<uintptr_t>(a.w_d) |
               reinterpret_cast<uintptr_t>(a.b_e) | reinterpret_cast<uintptr_t>(a.b_d)) & 15u) == 0,
             AQ_ERR_BAD_ALIGN, "expand_dw: operands must be 16-byte aligned");
  //                                                      CIN CEXP KS S TH TW  NT  CTAs/SM
  if (a.cin == 16 && a.cexp == 96) return launch_cfg<16, 96, 3, 2, 4, 8, 192, 2>(a, st);
  if (a.k == 3) return launch_cfg<24, 144, 3, 1, 8, 16, 576, 1>(a, st);
  return launch_cfg<24, 144, 5, 2, 4, 8, 576, 1>(a, st);
}

}  // namespace aq
