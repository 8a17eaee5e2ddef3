// Fused base projection + watermark-LoRA contraction for sm_100a.
//
//   H  = A Dn^T                                   (128 x 64 fp32 in TMEM, once per row block)
//   Hs = bf16(bf16(H) * scale[sample(row), :])    (epilogue warps: TMEM -> registers -> swizzled SMEM)
//   Y  = A W^T + bias + Hs Up^T                   (same TMEM accumulator; Hs is the A operand of a 65th k-block)
//
// Replaces the op sequence of utils/lora_modules.py:9-26 + :56-62 (Linear, down, diag_embed, bmm, up, add).
// The backward's dX uses the same kernel with (A, W, Dn, Up) = (G, W^T, Up^T, Dn^T) and a mid-epilogue that
// also emits dH, Hs and the per-sample dscale reduction (mode 1).
//
// Structure (one CTA per SM, persistent over work items = (row block, group of column tiles)):
//   warp 0    TMA producer      global -> SMEM ring (A, W, Dn tiles; 128B swizzle; mbarrier complete_tx)
//   warp 1    MMA issuer        one thread issues tcgen05.mma; tcgen05.commit releases ring slots
//   warp 2    TMEM allocator
//   warps 4-7 epilogue          tcgen05.ld -> bias/scale -> bf16 -> SMEM staging -> TMA store
#include <stdio.h>
#include <string.h>

#include "aq_ptx.cuh"
#include "lora_gemm.h"

namespace aq {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;          // 64 bf16 = 128 bytes = one 128B swizzle row
constexpr int kRankPad = 64;         // H tile width (r <= 64, zero padded by TMA)
constexpr int kThreads = 256;
constexpr int kEpiThreads = 128;
constexpr int kATileBytes = kBlockM * kBlockK * 2;   // 16 KiB
constexpr int kDnTileBytes = kRankPad * kBlockK * 2; // 8 KiB
constexpr int kHsBytes = kBlockM * kRankPad * 2;     // 16 KiB
constexpr int kStgBytesPerBuf = 32 * 64;             // one warp: 32 rows x 32 bf16
constexpr int kStgBytes = 4 * 2 * kStgBytesPerBuf;   // 4 warps x double buffer

struct LoraGemmParams {
  CUtensorMap tmap_a;    // A  [M, K]   box {64, 128} swizzle 128B
  CUtensorMap tmap_w;    // W  [N, K]   box {64, BN}  swizzle 128B
  CUtensorMap tmap_dn;   // Dn [r, K]   box {64, 64}  swizzle 128B
  CUtensorMap tmap_up;   // Up [N, r]   box {64, BN}  swizzle 128B
  CUtensorMap tmap_y;    // Y  [M, N]   box {32, 32}  swizzle 64B (store)
  const __nv_bfloat16* bias;   // [N] or null
  const float* scale;          // [num_samples, r]
  __nv_bfloat16* aux_out0;     // mode 0: H [M, r] (may be null); mode 1: dH [M, r]
  __nv_bfloat16* aux_out1;     // mode 1: Hs [M, r]
  const __nv_bfloat16* h_in;   // mode 1: H [M, r] saved by the forward
  float* g_scale;              // mode 1: [num_samples, r] accumulated (may be null)
  long long tokens;            // rows per sample
  int num_samples;
  int M, N, K, r;
  int num_m_tiles, num_n_tiles, group_size, num_groups;
  int mode;       // 0 forward, 1 backward (dX)
  int has_lora;   // 0: plain GEMM
  int has_main;   // 0: only the H phase + mid epilogue (backward of layers whose input needs no gradient)
};

template <int BN>
struct SmemLayout {
  static constexpr int kWTileBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kATileBytes + kWTileBytes + kDnTileBytes;
  static constexpr int kBudget = 232448 - 1024 /*alignment slack*/ - 512 /*barriers*/;
  static constexpr int kStagesRaw = (kBudget - kHsBytes - kStgBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
  static constexpr int kHsOff = kStages * kStageBytes;
  static constexpr int kStgOff = kHsOff + kHsBytes;
  static constexpr int kBarOff = kStgOff + kStgBytes;
  static constexpr int kTotal = kBarOff + 512 + 1024;
  static_assert(kStages >= 2, "not enough shared memory for a pipeline");
  static_assert(kWTileBytes % 1024 == 0, "W tile must keep 1024B alignment");
  static_assert(2 * BN + kRankPad <= 512, "TMEM budget");
};

// 64 values per lane, 32 lanes -> lane L ends with the column sums of columns 2L and 2L+1 in v[0], v[1].
__device__ __forceinline__ void warp_colsum64(float (&v)[64], int lane) {
#pragma unroll
  for (int w = 32, bit = 16; w >= 2; w >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = up ? v[i] : v[i + w];
      const float keep = up ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1) lora_gemm_kernel(const __grid_constant__ LoraGemmParams p) {
  using L = SmemLayout<BN>;
  constexpr int kStages = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 1024B alignment: required by the 128B swizzle atoms referenced through UMMA descriptors
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t bar_base = smem_base + L::kBarOff;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto acc_full_bar = [&](int b) { return bar_base + 8u * (2 * kStages + b); };
  auto acc_empty_bar = [&](int b) { return bar_base + 8u * (2 * kStages + 2 + b); };
  const uint32_t h_full_bar = bar_base + 8u * (2 * kStages + 4);
  const uint32_t hs_ready_bar = bar_base + 8u * (2 * kStages + 5);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 6);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + L::kBarOff + 8 * (2 * kStages + 6));

  auto a_tile = [&](int s) { return smem_base + s * L::kStageBytes; };
  auto w_tile = [&](int s) { return smem_base + s * L::kStageBytes + kATileBytes; };
  auto dn_tile = [&](int s) { return smem_base + s * L::kStageBytes + kATileBytes + L::kWTileBytes; };
  const uint32_t hs_tile = smem_base + L::kHsOff;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_a);
    if (p.has_main) {
      tma_prefetch_desc(&p.tmap_w);
      tma_prefetch_desc(&p.tmap_y);
    }
    if (p.has_lora) {
      tma_prefetch_desc(&p.tmap_dn);
      if (p.has_main) tma_prefetch_desc(&p.tmap_up);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full_bar(b), 1);
      mbar_init(acc_empty_bar(b), kEpiThreads);
    }
    mbar_init(h_full_bar, 1);
    mbar_init(hs_ready_bar, kEpiThreads);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int num_kb = (p.K + kBlockK - 1) / kBlockK;
  const int total_items = p.num_m_tiles * p.num_groups;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int m_tile = item % p.num_m_tiles;
        const int grp = item / p.num_m_tiles;
        const int m0 = m_tile * kBlockM;
        const int nt_begin = grp * p.group_size;
        const int nt_end = min(nt_begin + p.group_size, p.num_n_tiles);
        for (int nt = nt_begin; nt < nt_end; ++nt) {
          const int n0 = nt * BN;
          const bool first = (nt == nt_begin) && p.has_lora;
          if (p.has_main || first) {
            for (int kb = 0; kb < num_kb; ++kb) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              const uint32_t bytes = kATileBytes + (p.has_main ? L::kWTileBytes : 0) + (first ? kDnTileBytes : 0);
              mbar_arrive_expect_tx(full_bar(stage), bytes);
              tma_load_2d(a_tile(stage), &p.tmap_a, full_bar(stage), kb * kBlockK, m0);
              if (p.has_main) tma_load_2d(w_tile(stage), &p.tmap_w, full_bar(stage), kb * kBlockK, n0);
              if (first) tma_load_2d(dn_tile(stage), &p.tmap_dn, full_bar(stage), kb * kBlockK, 0);
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
          }
          if (p.has_lora && p.has_main) {
            // the rank-r "65th k-block": only the Up tile travels; its A operand (Hs) is produced on chip
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_arrive_expect_tx(full_bar(stage), L::kWTileBytes);
            tma_load_2d(w_tile(stage), &p.tmap_up, full_bar(stage), 0, n0);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc_main = make_idesc_bf16(kBlockM, BN, 0, 0);
      const uint32_t idesc_h = make_idesc_bf16(kBlockM, kRankPad, 0, 0);
      const uint32_t tmem_h = tmem_base + 2 * BN;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_iter = 0, item_iter = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++item_iter) {
        const int grp = item / p.num_m_tiles;
        const int nt_begin = grp * p.group_size;
        const int nt_end = min(nt_begin + p.group_size, p.num_n_tiles);
        for (int nt = nt_begin; nt < nt_end; ++nt) {
          const bool first = (nt == nt_begin) && p.has_lora;
          const uint32_t buf = acc_iter & 1u;
          const uint32_t tmem_acc = tmem_base + buf * BN;
          if (p.has_main) {
            mbar_wait(acc_empty_bar(buf), ((acc_iter >> 1) & 1u) ^ 1u);
            tc_fence_after();
          }
          if (p.has_main || first) {
            for (int kb = 0; kb < num_kb; ++kb) {
              mbar_wait(full_bar(stage), phase);
              tc_fence_after();
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                const uint64_t adesc = make_smem_desc(a_tile(stage) + k * 32, 16, 1024, 2);
                const uint32_t acc_flag = (kb | k) != 0 ? 1u : 0u;
                if (p.has_main) {
                  const uint64_t bdesc = make_smem_desc(w_tile(stage) + k * 32, 16, 1024, 2);
                  umma_f16(tmem_acc, adesc, bdesc, idesc_main, acc_flag);
                }
                if (first) {
                  const uint64_t ddesc = make_smem_desc(dn_tile(stage) + k * 32, 16, 1024, 2);
                  umma_f16(tmem_h, adesc, ddesc, idesc_h, acc_flag);
                }
              }
              umma_commit(empty_bar(stage));
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
          }
          if (first) {
            umma_commit(h_full_bar);                       // H accumulated -> wake the epilogue warps
            mbar_wait(hs_ready_bar, item_iter & 1u);       // Hs (bf16, swizzled) is in SMEM
            tc_fence_after();
          }
          if (p.has_lora && p.has_main) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < kRankPad / 16; ++k) {
              const uint64_t adesc = make_smem_desc(hs_tile + k * 32, 16, 1024, 2);
              const uint64_t bdesc = make_smem_desc(w_tile(stage) + k * 32, 16, 1024, 2);
              umma_f16(tmem_acc, adesc, bdesc, idesc_main, 1u);
            }
            umma_commit(empty_bar(stage));
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          if (p.has_main) {
            umma_commit(acc_full_bar(buf));
            ++acc_iter;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // =========================== epilogue warps ===========================
    const int q = warp - 4;  // TMEM lane quadrant == warp_id % 4
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t stg_base = smem_base + L::kStgOff + q * 2 * kStgBytesPerBuf;
    uint8_t* hs_gen = smem_gen + L::kHsOff;
    uint8_t* stg_gen = smem_gen + L::kStgOff + q * 2 * kStgBytesPerBuf;
    const int row_in_tile = q * 32 + lane;
    uint32_t acc_iter = 0, item_iter = 0, store_iter = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++item_iter) {
      const int m_tile = item % p.num_m_tiles;
      const int grp = item / p.num_m_tiles;
      const int m0 = m_tile * kBlockM;
      const int nt_begin = grp * p.group_size;
      const int nt_end = min(nt_begin + p.group_size, p.num_n_tiles);
      const long long grow = (long long)m0 + row_in_tile;
      const bool row_ok = grow < p.M;
      const bool aux_owner = row_ok && grp == 0;   // side outputs (H / dH / Hs / dscale) are emitted once per row block
      for (int nt = nt_begin; nt < nt_end; ++nt) {
        const int n0 = nt * BN;
        const bool first = (nt == nt_begin) && p.has_lora;
        if (first) {
          // ---------------- mid epilogue: H (TMEM, fp32) -> Hs (SMEM, bf16, 128B-swizzled K-major) ----------------
          long long sample = grow / p.tokens;
          if (sample > p.num_samples - 1) sample = p.num_samples - 1;
          const float* sp = p.scale + sample * p.r;
          const size_t aux_off = (size_t)grow * p.r;
          uint4 hin[8];
          if (p.mode == 1) {
#pragma unroll
            for (int j8 = 0; j8 < 8; ++j8) {
              hin[j8] = make_uint4(0, 0, 0, 0);
              if (row_ok && j8 * 8 < p.r) hin[j8] = __ldg(reinterpret_cast<const uint4*>(p.h_in + aux_off + j8 * 8));
            }
          }
          mbar_wait(h_full_bar, item_iter & 1u);
          tc_fence_after();
          float v[64];
          {
            uint32_t t0[32], t1[32];
            tmem_ld_32x32(tmem_base + lane_base + 2 * BN, t0);
            tmem_ld_32x32(tmem_base + lane_base + 2 * BN + 32, t1);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              v[i] = __uint_as_float(t0[i]);
              v[32 + i] = __uint_as_float(t1[i]);
            }
          }
#pragma unroll
          for (int j8 = 0; j8 < 8; ++j8) {
            uint4 to_smem = make_uint4(0, 0, 0, 0);
            if (j8 * 8 < p.r) {
              const float4 s0 = __ldg(reinterpret_cast<const float4*>(sp + j8 * 8));
              const float4 s1 = __ldg(reinterpret_cast<const float4*>(sp + j8 * 8 + 4));
              const float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
              float hb[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) hb[i] = bf16_round(v[j8 * 8 + i]);
              if (p.mode == 0) {
                to_smem.x = pack_bf16x2(hb[0] * s[0], hb[1] * s[1]);
                to_smem.y = pack_bf16x2(hb[2] * s[2], hb[3] * s[3]);
                to_smem.z = pack_bf16x2(hb[4] * s[4], hb[5] * s[5]);
                to_smem.w = pack_bf16x2(hb[6] * s[6], hb[7] * s[7]);
                if (aux_owner && p.aux_out0 != nullptr) {
                  uint4 hraw;
                  hraw.x = pack_bf16x2(hb[0], hb[1]);
                  hraw.y = pack_bf16x2(hb[2], hb[3]);
                  hraw.z = pack_bf16x2(hb[4], hb[5]);
                  hraw.w = pack_bf16x2(hb[6], hb[7]);
                  *reinterpret_cast<uint4*>(p.aux_out0 + aux_off + j8 * 8) = hraw;
                }
              } else {
                // hb = dHs (bf16-rounded like the reference's bf16 matmul output); hin = saved H
                const uint32_t hw[4] = {hin[j8].x, hin[j8].y, hin[j8].z, hin[j8].w};
                float hval[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  hval[2 * i] = bf16_lo(hw[i]);
                  hval[2 * i + 1] = bf16_hi(hw[i]);
                }
                to_smem.x = pack_bf16x2(hb[0] * s[0], hb[1] * s[1]);
                to_smem.y = pack_bf16x2(hb[2] * s[2], hb[3] * s[3]);
                to_smem.z = pack_bf16x2(hb[4] * s[4], hb[5] * s[5]);
                to_smem.w = pack_bf16x2(hb[6] * s[6], hb[7] * s[7]);
                if (aux_owner) {
                  *reinterpret_cast<uint4*>(p.aux_out0 + aux_off + j8 * 8) = to_smem;  // dH
                  uint4 hs;
                  hs.x = pack_bf16x2(hval[0] * s[0], hval[1] * s[1]);
                  hs.y = pack_bf16x2(hval[2] * s[2], hval[3] * s[3]);
                  hs.z = pack_bf16x2(hval[4] * s[4], hval[5] * s[5]);
                  hs.w = pack_bf16x2(hval[6] * s[6], hval[7] * s[7]);
                  *reinterpret_cast<uint4*>(p.aux_out1 + aux_off + j8 * 8) = hs;       // Hs
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) v[j8 * 8 + i] = hb[i] * hval[i];  // dscale integrand
              }
            } else if (p.mode == 1) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[j8 * 8 + i] = 0.f;
            }
            *reinterpret_cast<uint4*>(hs_gen + row_in_tile * 128 + ((j8 ^ (row_in_tile & 7)) << 4)) = to_smem;
          }
          tc_fence_before();
          fence_proxy_async_smem();   // generic-proxy SMEM writes -> visible to the tensor-core (async) proxy
          mbar_arrive(hs_ready_bar);
          if (p.mode == 1 && p.g_scale != nullptr && grp == 0) {
            // dscale[b, j] += sum over this tile's rows of dHs * H
            const long long first_row = (long long)m0 + q * 32;
            const bool uniform = (p.tokens % 32 == 0) && (first_row + 32 <= p.M);
            if (uniform) {
              warp_colsum64(v, lane);
              const int c = 2 * lane;
              if (c < p.r) {
                atomicAdd(p.g_scale + sample * p.r + c, v[0]);
                atomicAdd(p.g_scale + sample * p.r + c + 1, v[1]);
              }
            } else if (row_ok) {
#pragma unroll
              for (int j = 0; j < 64; ++j)
                if (j < p.r) atomicAdd(p.g_scale + sample * p.r + j, v[j]);
            }
          }
        }
        if (!p.has_main) continue;
        // ---------------- tile epilogue: accumulator -> (+bias) -> bf16 -> staging -> TMA store ----------------
        const uint32_t buf = acc_iter & 1u;
        mbar_wait(acc_full_bar(buf), (acc_iter >> 1) & 1u);
        tc_fence_after();
        const int cols_left = p.N - n0;
        const int nchunks = min(BN / 32, (cols_left + 31) / 32);
        for (int c = 0; c < nchunks; ++c) {
          uint32_t t[32];
          tmem_ld_32x32(tmem_base + lane_base + buf * BN + c * 32, t);
          tmem_wait_ld();
          if (c == nchunks - 1) {
            tc_fence_before();
            mbar_arrive(acc_empty_bar(buf));   // accumulator drained -> MMA may start the tile after next
          }
          const int col0 = n0 + c * 32;
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(t[i]);
          if (p.bias != nullptr) {
            if (col0 + 32 <= p.N) {
#pragma unroll
              for (int i4 = 0; i4 < 4; ++i4) {
                const uint4 bw = __ldg(reinterpret_cast<const uint4*>(p.bias + col0 + i4 * 8));
                const uint32_t bb[4] = {bw.x, bw.y, bw.z, bw.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  f[i4 * 8 + 2 * i] += bf16_lo(bb[i]);
                  f[i4 * 8 + 2 * i + 1] += bf16_hi(bb[i]);
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) f[i] += __bfloat162float(p.bias[col0 + i]);
            }
          }
          const uint32_t sb = store_iter & 1u;
          if (lane == 0) tma_store_wait_read<1>();   // the store that last used this buffer has read it
          __syncwarp();
          uint8_t* dst = stg_gen + sb * kStgBytesPerBuf + lane * 64;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint4 o;
            o.x = pack_bf16x2(f[ch * 8 + 0], f[ch * 8 + 1]);
            o.y = pack_bf16x2(f[ch * 8 + 2], f[ch * 8 + 3]);
            o.z = pack_bf16x2(f[ch * 8 + 4], f[ch * 8 + 5]);
            o.w = pack_bf16x2(f[ch * 8 + 6], f[ch * 8 + 7]);
            *reinterpret_cast<uint4*>(dst + ((ch ^ ((lane >> 1) & 3)) << 4)) = o;   // 64B swizzle
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&p.tmap_y, stg_base + sb * kStgBytesPerBuf, col0, m0 + q * 32);
            tma_store_commit();
          }
          ++store_iter;
        }
        ++acc_iter;
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int BN>
static int launch_bn(const LoraGemmArgs& a, cudaStream_t stream) {
  using L = SmemLayout<BN>;
  LoraGemmParams p;
  memset(&p, 0, sizeof(p));
  const int has_lora = a.dn != nullptr;
  {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.M};
    uint64_t str[1] = {(uint64_t)a.lda * 2};
    uint32_t box[2] = {kBlockK, kBlockM};
    int rc = make_tmap(&p.tmap_a, a.a, 2, 2, dims, str, box, kSwz128);
    if (rc) return rc;
  }
  if (a.has_main) {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.N};
    uint64_t str[1] = {(uint64_t)a.K * 2};
    uint32_t box[2] = {kBlockK, (uint32_t)BN};
    int rc = make_tmap(&p.tmap_w, a.w, 2, 2, dims, str, box, kSwz128);
    if (rc) return rc;
    uint64_t ydims[2] = {(uint64_t)a.N, (uint64_t)a.M};
    uint64_t ystr[1] = {(uint64_t)a.ldy * 2};
    uint32_t ybox[2] = {32, 32};
    rc = make_tmap(&p.tmap_y, a.y, 2, 2, ydims, ystr, ybox, kSwz64);
    if (rc) return rc;
  }
  if (has_lora) {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.r};
    uint64_t str[1] = {(uint64_t)a.K * 2};
    uint32_t box[2] = {kBlockK, kRankPad};
    int rc = make_tmap(&p.tmap_dn, a.dn, 2, 2, dims, str, box, kSwz128);
    if (rc) return rc;
    if (a.has_main) {
      uint64_t udims[2] = {(uint64_t)a.r, (uint64_t)a.N};
      uint64_t ustr[1] = {(uint64_t)a.r * 2};
      uint32_t ubox[2] = {kRankPad, (uint32_t)BN};
      rc = make_tmap(&p.tmap_up, a.up, 2, 2, udims, ustr, ubox, kSwz128);
      if (rc) return rc;
    }
  }
  p.bias = reinterpret_cast<const __nv_bfloat16*>(a.bias);
  p.scale = a.scale;
  p.aux_out0 = reinterpret_cast<__nv_bfloat16*>(a.aux_out0);
  p.aux_out1 = reinterpret_cast<__nv_bfloat16*>(a.aux_out1);
  p.h_in = reinterpret_cast<const __nv_bfloat16*>(a.h_in);
  p.g_scale = a.g_scale;
  p.tokens = a.tokens > 0 ? a.tokens : a.M;
  p.num_samples = (int)((a.M + p.tokens - 1) / p.tokens);
  p.M = (int)a.M; p.N = a.N; p.K = a.K; p.r = a.r;
  p.mode = a.mode; p.has_lora = has_lora; p.has_main = a.has_main;
  p.num_m_tiles = (int)((a.M + kBlockM - 1) / kBlockM);
  p.num_n_tiles = a.has_main ? (a.N + BN - 1) / BN : 1;

  const int sms = sm_count();
  if (sms <= 0) return fail(AQ_ERR_LAUNCH, "no CUDA device");
  // Column tiles per work item: the H phase is paid once per item, wave quantisation once per launch.
  int best_g = 1;
  if (a.force_group > 0) {
    best_g = a.force_group;
  } else {
    double best_cost = 1e30;
    const double kb = (double)((a.K + kBlockK - 1) / kBlockK);
    for (int g = 1; g <= p.num_n_tiles; ++g) {
      const int groups = (p.num_n_tiles + g - 1) / g;
      const long long items = (long long)groups * p.num_m_tiles;
      const long long waves = (items + sms - 1) / sms;
      // per item: g tiles of (kb + 1) k-blocks of width BN, plus the H phase (kb k-blocks of width 64) and its bubble
      const double item_cost = g * (kb + (has_lora ? 1.0 : 0.0)) * BN + (has_lora ? kb * kRankPad + 6.0 * BN : 0.0);
      const double cost = (double)waves * item_cost;
      if (cost < best_cost - 1e-9) { best_cost = cost; best_g = g; }
    }
  }
  p.group_size = best_g;
  p.num_groups = (p.num_n_tiles + best_g - 1) / best_g;
  const long long items = (long long)p.num_groups * p.num_m_tiles;
  const int grid = (int)(items < sms ? items : sms);

  static bool attr_set = false;   // benign race: idempotent
  if (!attr_set) {
    AQ_CHECK_CUDA(cudaFuncSetAttribute(lora_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  lora_gemm_kernel<BN><<<grid, kThreads, L::kTotal, stream>>>(p);
  AQ_LAUNCHED();
  return AQ_OK;
}

static int pick_bn(int N) {
  if (N % 160 == 0) return 160;
  if (N % 192 == 0) return 192;
  if (N % 128 == 0) return 128;
  if (N <= 64) return 64;
  // least padded columns, prefer the wider tile on ties
  int best = 192, best_pad = 1 << 30;
  const int cand[4] = {192, 160, 128, 64};
  for (int i = 0; i < 4; ++i) {
    const int pad = (N + cand[i] - 1) / cand[i] * cand[i] - N;
    if (pad < best_pad) { best_pad = pad; best = cand[i]; }
  }
  return best;
}

int launch_lora_gemm(const LoraGemmArgs& a, cudaStream_t stream) {
  AQ_REQUIRE(a.M > 0 && a.K > 0, AQ_ERR_BAD_SHAPE, "lora_gemm: empty problem M=%lld K=%d", (long long)a.M, a.K);
  AQ_REQUIRE(a.M < (1ll << 31), AQ_ERR_BAD_SHAPE, "lora_gemm: M=%lld exceeds 2^31-1 rows", (long long)a.M);
  AQ_REQUIRE(a.K % 8 == 0, AQ_ERR_BAD_SHAPE, "lora_gemm: K=%d must be a multiple of 8", a.K);
  AQ_REQUIRE(!a.has_main || (a.N > 0 && a.N % 8 == 0), AQ_ERR_BAD_SHAPE, "lora_gemm: N=%d must be a positive multiple of 8", a.N);
  if (a.dn != nullptr) {
    AQ_REQUIRE(a.r >= 8 && a.r <= kRankPad && a.r % 8 == 0, AQ_ERR_BAD_SHAPE,
               "lora_gemm: rank r=%d unsupported (need 8 <= r <= 64, r %% 8 == 0)", a.r);
    AQ_REQUIRE(a.scale != nullptr, AQ_ERR_BAD_SHAPE, "lora_gemm: scale is NULL");
    AQ_REQUIRE(!a.has_main || a.up != nullptr, AQ_ERR_BAD_SHAPE, "lora_gemm: up is NULL");
  }
  AQ_REQUIRE(a.has_main || a.dn != nullptr, AQ_ERR_BAD_SHAPE, "lora_gemm: nothing to compute");
  AQ_REQUIRE(a.lda % 8 == 0 && (!a.has_main || a.ldy % 8 == 0), AQ_ERR_BAD_ALIGN, "lora_gemm: leading dimensions must be multiples of 8 elements");
  int rc = check_arch();
  if (rc) return rc;
  const int bn = a.force_bn > 0 ? a.force_bn : pick_bn(a.has_main ? a.N : 64);
  switch (bn) {
    case 64: return launch_bn<64>(a, stream);
    case 128: return launch_bn<128>(a, stream);
    case 160: return launch_bn<160>(a, stream);
    case 192: return launch_bn<192>(a, stream);
    default: return fail(AQ_ERR_BAD_SHAPE, "lora_gemm: unsupported column tile %d", bn);
  }
}

}  // namespace aq
