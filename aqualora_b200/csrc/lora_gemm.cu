// Fused base projection + watermark-LoRA contraction for sm_100a, CTA-pair (tcgen05 cta_group::2) edition.
//
//   H  = A Dn^T                                   (256 x 64 fp32 in the pair's TMEM, once per row block)
//   Hs = bf16(bf16(H) * scale[sample(row), :])    (epilogue warps: TMEM -> registers -> swizzled SMEM)
//   Y  = A W^T + bias + Hs Up^T                   (same TMEM accumulator; Hs is the A operand of one extra k-block)
//
// Replaces the op sequence of utils/lora_modules.py:9-26 + :56-62 (Linear, down, diag_embed, bmm, up, add).
// The backward's dX uses the same kernel with (A, W, Dn, Up) = (G, W^T, Up^T, Dn^T) and a mid-epilogue that
// also emits dH, Hs and the per-sample dscale reduction (mode 1).
//
// Why pairs: measured on B200 the 1-CTA 128 x 160 version moved 36 KiB L2->SMEM per 2.6 MFLOP k-block and sat at the
// ~11 TB/s L2->SMEM ceiling (profiles/r01_*).  A pair shares the W tile (each CTA loads half of it), so the same k-block
// costs 26 KiB (BN=160) / 28 KiB (BN=192, 3.1 MFLOP) per CTA.
//
// Structure: cluster of 2 CTAs = one 256-row block, one CTA pair per SM pair, persistent over work items
// (row block, group of column tiles); 320 threads per CTA:
//   warps 0-7  epilogue          TMEM lane quadrant = warp % 4, column half = warp / 4:
//                                tcgen05.ld -> bias -> bf16 -> staging tile (box of Y's tensor map) -> cp.async.bulk.tensor store
//   warp 8     TMA producer      both CTAs: own A rows, own half of W / Dn / Up; bytes signalled on the LEADER's barrier
//   warp 9     TMEM allocator; in the leader CTA one thread issues every tcgen05.mma of the pair
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "aq_ptx.cuh"
#include "lora_gemm.h"

namespace aq {

#ifdef AQ_GEMM_TRACE
// developer build only (tools/gemm_trace.py): clock64() stamps of CTA 0's roles (32 slots per work item) and per-CTA
// globaltimer stamps (entry, after the prologue, exit, SM id)
__device__ unsigned long long g_trace[64 * 32];
__device__ unsigned long long g_cta_times[256 * 4];
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned smid() {
  unsigned r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}
#define AQ_TRACE(iter, slot)                                                                        \
  do {                                                                                              \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (iter) < 64u) g_trace[(iter) * 32 + (slot)] = clock64(); \
  } while (0)
#define AQ_CTA_STAMP(k)                                                                                \
  do {                                                                                                 \
    if (threadIdx.x == 0 && blockIdx.x < 256) g_cta_times[blockIdx.x * 4 + (k)] = globaltimer_ns();    \
  } while (0)
#else
#define AQ_TRACE(iter, slot) do {} while (0)
#define AQ_CTA_STAMP(k) do {} while (0)
#endif

constexpr int kBlockM = 128;         // rows per CTA (the pair covers 256)
constexpr int kPairM = 2 * kBlockM;
constexpr int kBlockK = 64;          // 64 bf16 = 128 bytes = one 128B swizzle row
constexpr int kRankPad = 64;         // H tile width (r <= 64, zero padded by TMA)
constexpr int kThreads = 320;
constexpr int kEpiWarps = 8;
constexpr int kProducerWarp = 8;
constexpr int kMmaWarp = 9;
constexpr int kATileBytes = kBlockM * kBlockK * 2;          // 16 KiB
constexpr int kDnHalfBytes = (kRankPad / 2) * kBlockK * 2;  // 4 KiB: this CTA's 32 rows of Dn
constexpr int kHsBytes = kBlockM * kRankPad * 2;            // 16 KiB
constexpr uint16_t kBothCtas = 0x3;
constexpr int kMaxGroup = 32;       // projections per grouped launch (32 cross-attention K / V projections in an SD U-Net)

// One projection sharing the launch's A operand.  A grouped launch (NP > 1) walks the work items of several projections of the
// SAME input -- the 32 cross-attention K / V projections of a U-Net forward all read the text context, q / k / v of a
// self-attention read the same normed tokens -- so their launch + prologue floor (14 us against < 3 us of work at M ~ 1 k rows)
// is paid once.
struct LoraProblem {
  CUtensorMap tmap_w;    // W  [N, K]   box {64, BN/2}   swizzle 128B
  CUtensorMap tmap_dn;   // Dn [r, K]   box {64, 32}     swizzle 128B
  CUtensorMap tmap_up;   // Up [N, r]   box {64, BN/2}   swizzle 128B
  CUtensorMap tmap_y;    // Y  [M, N]   box {kPassCols, 32}  (store; swizzle 64B when kPassCols == 32)
  const __nv_bfloat16* bias;   // [N] or null
  const __nv_bfloat16* res;    // [M, ldres] or null: residual added in the tile epilogue (Y = ... + res)
  long long ldres;
  __nv_bfloat16* y;            // [M, ldy]
  __nv_bfloat16* aux_out0;     // mode 0: H [M, r] (may be null); mode 1: dH [M, r]
  long long ldy;
  int N, num_n_tiles, group_size, num_groups;
  int item_begin;              // first work item of this projection
  int pad_;
};

template <int NP>
struct LoraGemmParams {
  CUtensorMap tmap_a;    // A  [M, K]   box {64, 128}    swizzle 128B
  const float* scale;          // [num_samples, r]
  __nv_bfloat16* aux_out1;     // mode 1: Hs [M, r]
  const __nv_bfloat16* h_in;   // mode 1: H [M, r] saved by the forward
  float* g_scale;              // mode 1: [num_samples, r] accumulated (may be null)
  long long tokens;            // rows per sample
  int num_samples;
  int M, K, r;
  int num_m_pairs;
  int num_problems, total_items;
  int mode;       // 0 forward, 1 backward (dX)
  int has_lora;   // 0: plain GEMM
  int has_main;   // 0: only the H phase + mid epilogue (backward of layers whose input needs no gradient)
  int skip_base;  // 1: a later rank chunk -- no A W^T term, tiles start from Hs Up^T alone
  int accum_y;    // 1: the tile epilogue adds the Y already in memory
  long long ld_r; // row stride of scale / H / dH / Hs (>= r: the rank chunk is a column slice of [*, r_total] tensors)
  LoraProblem prob[NP];
};

// DUAL: a pipeline stage carries the W tiles of TWO column tiles, so that one pass over A feeds both accumulators (K-heavy shapes:
// the kernel is bound by operand traffic L2 -> SM there, and A is the larger part of it; see launch_bn)
template <int BN, bool DUAL = false>
struct SmemLayout {
  static constexpr int kNC = BN / 2;                         // accumulator columns per epilogue warp
  static constexpr int kWHalfBytes = (BN / 2) * kBlockK * 2;
  static constexpr int kStageBytes = kATileBytes + (DUAL ? 2 : 1) * kWHalfBytes + kDnHalfBytes;
  // Output staging for the TMA-store epilogue: each epilogue warp owns two buffers of 32 rows x kPassCols bf16 in the layout of
  // the Y tensor map's box; a pass = convert kPassCols accumulator columns, st.shared, fence, one cp.async.bulk.tensor store.
  // 80-byte rows (BN = 160) are conflict-free as they are (8 consecutive lanes -> 8 distinct 16-byte bank groups); 64-byte rows
  // use the 64B TMA swizzle.
  static constexpr int kPassCols = (BN == 160) ? 40 : 32;
  static constexpr int kPasses = kNC / kPassCols;
  static constexpr bool kStgSwizzle64 = (kPassCols == 32);
  static constexpr int kStgRowBytes = kPassCols * 2;
  static constexpr int kStgBufBytes = 32 * kStgRowBytes;
  static constexpr int kStgWarpBytes = 2 * kStgBufBytes;
  static constexpr int kStgBytes = kEpiWarps * kStgWarpBytes;
  static constexpr int kBiasWarpBytes = 256;                 // this warp's slice of the bias row (kNC bf16 <= 192 B)
  static constexpr int kBiasBytes = kEpiWarps * kBiasWarpBytes;
  static constexpr int kBudget = 232448 - 1024 /*alignment slack*/ - 512 /*barriers*/;
  static constexpr int kStagesRaw = (kBudget - kHsBytes - kStgBytes - kBiasBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
  static constexpr int kHsOff = kStages * kStageBytes;
  static constexpr int kStgOff = kHsOff + kHsBytes;
  static constexpr int kBiasOff = kStgOff + kStgBytes;
  static constexpr int kBarOff = kBiasOff + kBiasBytes;
  static constexpr int kTotal = kBarOff + 512 + 1024;
  static_assert(kStages >= 3, "not enough shared memory for a pipeline");
  static_assert(kWHalfBytes % 1024 == 0, "W tile must keep 1024B alignment");
  static_assert(BN % 32 == 0, "two epilogue column halves of whole 16-column TMEM loads");
  static_assert(kNC % kPassCols == 0 && kStgBufBytes % 512 == 0, "store passes tile the warp's slice; buffers keep the swizzle phase");
  static_assert(2 * BN + kRankPad <= 512, "TMEM budget");
  static_assert((2 * kStages + 7) * 8 <= 512, "barrier block");
};

template <bool B>
struct BoolC { static constexpr bool value = B; };

// N values per lane, 32 lanes -> lane L ends with the column sums of columns L*N/32 ... in v[0 .. N/32).
template <int N>
__device__ __forceinline__ void warp_colsum(float (&v)[N], int lane) {
#pragma unroll
  for (int w = N / 2, bit = 16; bit >= 1; w >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = up ? v[i] : v[i + w];
      const float keep = up ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
}

// MODE (0 forward, 1 backward dX) is a template parameter: each kernel carries only its own mid-epilogue (smaller instruction
// footprint: ncu attributes a visible share of the epilogue warps' compute stalls to instruction fetch), and the projection /
// launch parameters the epilogue loops need are read into registers once per item (a constant-bank read behind a uniform branch
// inside an unrolled loop is a dependent LDCU -> compare -> branch chain per 8 columns).
template <int BN, int NP, int MODE, bool DUAL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) lora_gemm_kernel(const __grid_constant__ LoraGemmParams<NP> p) {
  using L = SmemLayout<BN, DUAL>;
  constexpr int kStages = L::kStages;
  constexpr int NC = L::kNC;
  extern __shared__ uint8_t smem_raw[];
  // 1024B alignment: required by the 128B swizzle atoms referenced through UMMA descriptors
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  AQ_CTA_STAMP(0);
  griddep_launch_dependents();   // PDL: the next kernel's CTAs may queue behind ours now (they wait for this grid before any global access)
  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;

  const uint32_t bar_base = smem_base + L::kBarOff;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                       // waited on in the leader only
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };          // both CTAs (multicast commit)
  auto acc_full_bar = [&](int b) { return bar_base + 8u * (2 * kStages + b); };   // both CTAs (multicast commit)
  auto acc_empty_bar = [&](int b) { return bar_base + 8u * (2 * kStages + 2 + b); };  // leader: 16 warp arrivals
  const uint32_t h_full_bar = bar_base + 8u * (2 * kStages + 4);                  // both CTAs (multicast commit)
  const uint32_t hs_ready_bar = bar_base + 8u * (2 * kStages + 5);                // leader: 16 warp arrivals
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 6);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + L::kBarOff + 8 * (2 * kStages + 6));

  auto a_tile = [&](int s) { return smem_base + s * L::kStageBytes; };
  auto w_tile = [&](int s) { return smem_base + s * L::kStageBytes + kATileBytes; };
  auto w2_tile = [&](int s) { return smem_base + s * L::kStageBytes + kATileBytes + L::kWHalfBytes; };   // DUAL only
  auto dn_tile = [&](int s) { return smem_base + s * L::kStageBytes + kATileBytes + (DUAL ? 2 : 1) * L::kWHalfBytes; };
  const uint32_t hs_tile = smem_base + L::kHsOff;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&p.tmap_a);
    if (p.has_main) {
      if (!p.skip_base) tma_prefetch_desc(&p.prob[0].tmap_w);
      tma_prefetch_desc(&p.prob[0].tmap_y);
    }
    if (p.has_lora) {
      tma_prefetch_desc(&p.prob[0].tmap_dn);
      if (p.has_main) tma_prefetch_desc(&p.prob[0].tmap_up);
    }
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full_bar(b), 1);
      mbar_init(acc_empty_bar(b), 2 * kEpiWarps);
    }
    mbar_init(h_full_bar, 1);
    mbar_init(hs_ready_bar, 2 * kEpiWarps);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / complete_tx
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  griddep_wait();                // PDL: everything above overlapped the previous kernel's tail; from here on we read / write global memory
  AQ_CTA_STAMP(1);

  const int num_kb = (p.K + kBlockK - 1) / kBlockK;
  const int total_items = p.total_items;
  // work item -> (projection, row pair, column-tile group); items of one CTA pair ascend, so the projection index only moves forward
  auto locate = [&](int item, int& pi) -> const LoraProblem& {
    while (pi + 1 < p.num_problems && item >= p.prob[pi + 1].item_begin) ++pi;
    return p.prob[pi];
  };
  const int item0 = blockIdx.x >> 1, item_step = gridDim.x >> 1;

  if (warp == kProducerWarp) {
    // =========================== TMA producer (both CTAs; converged warp, one elected lane issues) ===========================
    int stage = 0;
    uint32_t phase = 0;
    const int w_row_off = (int)cta_rank * (BN / 2);
    const int dn_row_off = (int)cta_rank * (kRankPad / 2);
    const bool load_w = p.has_main && !p.skip_base;
    auto k_loads = [&](const LoraProblem& q, int m0, int n0, bool first) {
      const uint32_t bytes = 2u * (kATileBytes + (load_w ? L::kWHalfBytes : 0) + (first ? kDnHalfBytes : 0));
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          const uint32_t fb = mapa_shared(full_bar(stage), 0);
          if (leader) mbar_arrive_expect_tx(full_bar(stage), bytes);
          tma_load_2d_pair(a_tile(stage), &p.tmap_a, fb, kb * kBlockK, m0);
          if (load_w) tma_load_2d_pair(w_tile(stage), &q.tmap_w, fb, kb * kBlockK, n0 + w_row_off);
          if (first) tma_load_2d_pair(dn_tile(stage), &q.tmap_dn, fb, kb * kBlockK, dn_row_off);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    };
    // DUAL: one pass over A for the column tiles at n0 and n1
    auto k_loads_dual = [&](const LoraProblem& q, int m0, int n0, int n1, bool first) {
      const uint32_t bytes = 2u * (kATileBytes + 2u * L::kWHalfBytes + (first ? kDnHalfBytes : 0));
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          const uint32_t fb = mapa_shared(full_bar(stage), 0);
          if (leader) mbar_arrive_expect_tx(full_bar(stage), bytes);
          tma_load_2d_pair(a_tile(stage), &p.tmap_a, fb, kb * kBlockK, m0);
          tma_load_2d_pair(w_tile(stage), &q.tmap_w, fb, kb * kBlockK, n0 + w_row_off);
          tma_load_2d_pair(w2_tile(stage), &q.tmap_w, fb, kb * kBlockK, n1 + w_row_off);
          if (first) tma_load_2d_pair(dn_tile(stage), &q.tmap_dn, fb, kb * kBlockK, dn_row_off);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    };
    auto up_load = [&](const LoraProblem& q, int n0) {
      // the rank-r extra k-block: only the Up tile travels; its A operand (Hs) is produced on chip
      mbar_wait(empty_bar(stage), phase ^ 1u);
      if (elect_one()) {
        const uint32_t fb = mapa_shared(full_bar(stage), 0);
        if (leader) mbar_arrive_expect_tx(full_bar(stage), 2u * L::kWHalfBytes);
        tma_load_2d_pair(w_tile(stage), &q.tmap_up, fb, 0, n0 + w_row_off);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    };
    int pi = 0;
    uint32_t trace_iter = 0;
    for (int item = item0; item < total_items; item += item_step, ++trace_iter) {
      const LoraProblem& q = locate(item, pi);
      const int local = item - q.item_begin;
      const int m_pair = local % p.num_m_pairs;
      const int grp = local / p.num_m_pairs;
      const int m0 = m_pair * kPairM + (int)cta_rank * kBlockM;
      const int nt_begin = grp * q.group_size;
      const int nt_end = min(nt_begin + q.group_size, q.num_n_tiles);
      const bool fused = p.has_lora && p.has_main;
      AQ_TRACE(trace_iter, 0);
      if constexpr (DUAL) {
        // tiles in pairs (the host guarantees an even tile count per group, a main product and no rank-chunk continuation)
        for (int nt = nt_begin; nt < nt_end; nt += 2) {
          k_loads_dual(q, m0, nt * BN, (nt + 1) * BN, nt == nt_begin && p.has_lora);
          if (fused) {
            up_load(q, nt * BN);
            up_load(q, (nt + 1) * BN);
          }
        }
      } else
      if (fused && nt_end - nt_begin >= 2 && !p.skip_base) {
        // deferred order: the Hs.Up k-blocks of the first two tiles follow the second tile's main loop
        k_loads(q, m0, nt_begin * BN, true);
        k_loads(q, m0, (nt_begin + 1) * BN, false);
        up_load(q, nt_begin * BN);
        up_load(q, (nt_begin + 1) * BN);
        for (int nt = nt_begin + 2; nt < nt_end; ++nt) {
          k_loads(q, m0, nt * BN, false);
          up_load(q, nt * BN);
        }
      } else {
        for (int nt = nt_begin; nt < nt_end; ++nt) {
          const bool first = (nt == nt_begin) && p.has_lora;
          if (load_w || first) k_loads(q, m0, nt * BN, first);
          if (fused) up_load(q, nt * BN);
        }
      }
      AQ_TRACE(trace_iter, 1);
    }
  } else if (warp == kMmaWarp) {
    // =========================== MMA issuer (leader CTA; converged warp, one elected lane issues) ===========================
    if (leader) {
      const uint32_t idesc_main = make_idesc_bf16(kPairM, BN, 0, 0);
      const uint32_t idesc_h = make_idesc_bf16(kPairM, kRankPad, 0, 0);
      const uint32_t tmem_h = tmem_base + 2 * BN;
      // K-major 128B-swizzled operand tile: LBO 16 B, SBO 1024 B (8 rows); a 16-element k-step advances the start by 32 B
      constexpr uint64_t kDescHi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      auto desc = [&](uint32_t addr) { return kDescHi | (uint64_t)(((addr >> 4) & 0x3FFFu) | (1u << 16)); };
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_iter = 0, item_iter = 0;
      auto acquire_acc = [&](uint32_t it) {
        mbar_wait(acc_empty_bar(it & 1u), ((it >> 1) & 1u) ^ 1u);
        tc_fence_after();
      };
      auto commit = [&](uint32_t bar) {
        if (elect_one()) umma_commit_pair(bar, kBothCtas);
        __syncwarp();
      };
      auto k_mmas_t = [&](uint32_t tmem_acc, auto main_c, auto first_c) {
        constexpr bool kMain = decltype(main_c)::value;
        constexpr bool kFirst = decltype(first_c)::value;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (kFirst && kb == 0) AQ_TRACE(item_iter, 3);
          if (elect_one()) {
            const uint64_t ad = desc(a_tile(stage));
            const uint64_t wd = desc(w_tile(stage));
            const uint64_t dd = desc(dn_tile(stage));
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              const uint32_t acc_flag = (kb | k) != 0 ? 1u : 0u;
              if constexpr (kMain) umma_f16_pair(tmem_acc, ad + 2 * k, wd + 2 * k, idesc_main, acc_flag);
              if constexpr (kFirst) umma_f16_pair(tmem_h, ad + 2 * k, dd + 2 * k, idesc_h, acc_flag);
            }
            umma_commit_pair(empty_bar(stage), kBothCtas);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      };
      // DUAL: both column tiles (and H on the first pair) from the same A stage
      auto k_mmas_dual = [&](uint32_t acc_a, uint32_t acc_b, bool first) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = desc(a_tile(stage));
            const uint64_t wd = desc(w_tile(stage));
            const uint64_t w2d = desc(w2_tile(stage));
            const uint64_t dd = desc(dn_tile(stage));
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              const uint32_t acc_flag = (kb | k) != 0 ? 1u : 0u;
              umma_f16_pair(acc_a, ad + 2 * k, wd + 2 * k, idesc_main, acc_flag);
              umma_f16_pair(acc_b, ad + 2 * k, w2d + 2 * k, idesc_main, acc_flag);
              if (first) umma_f16_pair(tmem_h, ad + 2 * k, dd + 2 * k, idesc_h, acc_flag);
            }
            umma_commit_pair(empty_bar(stage), kBothCtas);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      };
      using TrueC = BoolC<true>;
      using FalseC = BoolC<false>;
      const bool load_w = p.has_main && !p.skip_base;
      auto k_mmas = [&](uint32_t tmem_acc, bool first) {
        if (!load_w) k_mmas_t(tmem_acc, FalseC{}, TrueC{});
        else if (first) k_mmas_t(tmem_acc, TrueC{}, TrueC{});
        else k_mmas_t(tmem_acc, TrueC{}, FalseC{});
      };
      auto up_mmas = [&](uint32_t tmem_acc) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = desc(hs_tile);
          const uint64_t wd = desc(w_tile(stage));
          const uint32_t keep = p.skip_base ? 0u : 1u;   // a later rank chunk has no base product underneath: overwrite
#pragma unroll
          for (int k = 0; k < kRankPad / 16; ++k) umma_f16_pair(tmem_acc, ad + 2 * k, wd + 2 * k, idesc_main, k == 0 ? keep : 1u);
          umma_commit_pair(empty_bar(stage), kBothCtas);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      };
      int pi = 0;
      for (int item = item0; item < total_items; item += item_step, ++item_iter) {
        const LoraProblem& q = locate(item, pi);
        const int grp = (item - q.item_begin) / p.num_m_pairs;
        const int nt_begin = grp * q.group_size;
        const int nt_end = min(nt_begin + q.group_size, q.num_n_tiles);
        const bool fused = p.has_lora && p.has_main;
        if constexpr (DUAL) {
          for (int nt = nt_begin; nt < nt_end; nt += 2) {
            const bool first = nt == nt_begin && p.has_lora;
            const uint32_t it0 = acc_iter, it1 = acc_iter + 1;
            const uint32_t acc0 = tmem_base + (it0 & 1u) * BN, acc1 = tmem_base + (it1 & 1u) * BN;
            acquire_acc(it0);
            acquire_acc(it1);
            k_mmas_dual(acc0, acc1, first);
            if (first) {
              commit(h_full_bar);
              mbar_wait(hs_ready_bar, item_iter & 1u);
              tc_fence_after();
            }
            if (fused) up_mmas(acc0);
            commit(acc_full_bar(it0 & 1u));
            if (fused) up_mmas(acc1);
            commit(acc_full_bar(it1 & 1u));
            acc_iter += 2;
          }
        } else
        if (fused && nt_end - nt_begin >= 2 && !p.skip_base) {
          const uint32_t it0 = acc_iter, it1 = acc_iter + 1;
          const uint32_t acc0 = tmem_base + (it0 & 1u) * BN, acc1 = tmem_base + (it1 & 1u) * BN;
          acquire_acc(it0);
          AQ_TRACE(item_iter, 2);
          k_mmas(acc0, true);
          commit(h_full_bar);                                // H accumulated -> wake the epilogue warps of both CTAs
          AQ_TRACE(item_iter, 4);
          acquire_acc(it1);
          AQ_TRACE(item_iter, 5);
          k_mmas(acc1, false);                               // the tensor core stays busy while Hs is being produced
          AQ_TRACE(item_iter, 6);
          mbar_wait(hs_ready_bar, item_iter & 1u);           // Hs (bf16, swizzled) is in both CTAs' SMEM
          tc_fence_after();
          AQ_TRACE(item_iter, 7);
          up_mmas(acc0);
          commit(acc_full_bar(it0 & 1u));
          AQ_TRACE(item_iter, 8);
          up_mmas(acc1);
          commit(acc_full_bar(it1 & 1u));
          AQ_TRACE(item_iter, 9);
          acc_iter += 2;
          for (int nt = nt_begin + 2; nt < nt_end; ++nt) {
            const uint32_t acc = tmem_base + (acc_iter & 1u) * BN;
            acquire_acc(acc_iter);
            k_mmas(acc, false);
            up_mmas(acc);
            commit(acc_full_bar(acc_iter & 1u));
            ++acc_iter;
          }
        } else {
          for (int nt = nt_begin; nt < nt_end; ++nt) {
            const bool first = (nt == nt_begin) && p.has_lora;
            const uint32_t acc = tmem_base + (acc_iter & 1u) * BN;
            if (p.has_main) acquire_acc(acc_iter);
            if (load_w || first) k_mmas(acc, first);
            if (first) {
              commit(h_full_bar);
              mbar_wait(hs_ready_bar, item_iter & 1u);
              tc_fence_after();
            }
            if (fused) up_mmas(acc);
            if (p.has_main) {
              commit(acc_full_bar(acc_iter & 1u));
              ++acc_iter;
            }
          }
        }
      }
    }
  } else {
    // =========================== epilogue warps (both CTAs) ===========================
    const int q = warp & 3;       // TMEM lane quadrant == warp_id % 4
    const int half = warp >> 2;   // which half of the tile's columns (and of H's 64 columns)
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint8_t* hs_gen = smem_gen + L::kHsOff;
    uint8_t* stg_gen = smem_gen + L::kStgOff + warp * L::kStgWarpBytes;
    const uint32_t stg_u32 = smem_base + L::kStgOff + warp * L::kStgWarpBytes;
    uint32_t stg_iter = 0;   // store passes issued by this warp (buffer = parity)
    uint32_t* bias_sm = reinterpret_cast<uint32_t*>(smem_gen + L::kBiasOff + warp * L::kBiasWarpBytes);
    const int row_in_tile = q * 32 + lane;
    const uint32_t acc_empty_remote[2] = {mapa_shared(acc_empty_bar(0), 0), mapa_shared(acc_empty_bar(1), 0)};
    const uint32_t hs_ready_remote = mapa_shared(hs_ready_bar, 0);
    uint32_t acc_iter = 0, item_iter = 0;
    int pi = 0;
    for (int item = item0; item < total_items; item += item_step, ++item_iter) {
      const LoraProblem& q_ = locate(item, pi);
      const int local = item - q_.item_begin;
      const int m_pair = local % p.num_m_pairs;
      const int grp = local / p.num_m_pairs;
      const int m0 = m_pair * kPairM + (int)cta_rank * kBlockM;
      const int nt_begin = grp * q_.group_size;
      const int nt_end = min(nt_begin + q_.group_size, q_.num_n_tiles);
      const long long grow = (long long)m0 + row_in_tile;
      const bool row_ok = grow < p.M;
      const bool aux_owner = row_ok && grp == 0;   // side outputs (H / dH / Hs / dscale) are emitted once per row block
      if (warp == 0) AQ_TRACE(item_iter, 10);
      if (p.has_lora) {
        // ---------------- mid epilogue: H (TMEM, fp32) -> Hs (SMEM, bf16, 128B-swizzled K-major) ----------------
        // this warp: 32 rows x H columns [32 * half, 32 * half + 32)
        const int hc0 = 32 * half;
        long long sample = grow / p.tokens;
        if (sample > p.num_samples - 1) sample = p.num_samples - 1;
        const float* sp = p.scale + sample * p.ld_r + hc0;
        const size_t aux_off = (size_t)grow * p.ld_r + hc0;
        const int rr = p.r;
        __nv_bfloat16* const aux0 = q_.aux_out0;
        __nv_bfloat16* const aux1 = p.aux_out1;
        const bool store_h = aux_owner && aux0 != nullptr;
        uint4 hin[4];
        if (MODE == 1) {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            hin[j8] = make_uint4(0, 0, 0, 0);
            if (row_ok && hc0 + j8 * 8 < rr) hin[j8] = __ldg(reinterpret_cast<const uint4*>(p.h_in + aux_off + j8 * 8));
          }
        }
        float4 sc4[8];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          sc4[j4] = (hc0 + j4 * 4 < rr) ? __ldg(reinterpret_cast<const float4*>(sp + j4 * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        mbar_wait(h_full_bar, item_iter & 1u);
        tc_fence_after();
        if (warp == 0) AQ_TRACE(item_iter, 11);
        float v[32];
        {
          uint32_t t0[32];
          tmem_ld_32x16(tmem_base + lane_base + 2 * BN + hc0, t0);
          tmem_ld_32x16(tmem_base + lane_base + 2 * BN + hc0 + 16, t0 + 16);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(t0[i]);
        }
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          uint4 to_smem = make_uint4(0, 0, 0, 0);
          if (hc0 + j8 * 8 < rr) {
            const float4 s0 = sc4[2 * j8], s1 = sc4[2 * j8 + 1];
            const float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            // H rounded to bf16 as packed pairs (one F2FP per two values), unpacked by shift / mask
            uint32_t hp[4];
            float hb[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              hp[i] = pack_bf16x2(v[j8 * 8 + 2 * i], v[j8 * 8 + 2 * i + 1]);
              hb[2 * i] = bf16_lo(hp[i]);
              hb[2 * i + 1] = bf16_hi(hp[i]);
            }
            to_smem.x = pack_bf16x2(hb[0] * s[0], hb[1] * s[1]);
            to_smem.y = pack_bf16x2(hb[2] * s[2], hb[3] * s[3]);
            to_smem.z = pack_bf16x2(hb[4] * s[4], hb[5] * s[5]);
            to_smem.w = pack_bf16x2(hb[6] * s[6], hb[7] * s[7]);
            if (MODE == 0) {
              if (store_h) {
                *reinterpret_cast<uint4*>(aux0 + aux_off + j8 * 8) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
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
              if (aux_owner) {
                *reinterpret_cast<uint4*>(aux0 + aux_off + j8 * 8) = to_smem;  // dH
                uint4 hs;
                hs.x = pack_bf16x2(hval[0] * s[0], hval[1] * s[1]);
                hs.y = pack_bf16x2(hval[2] * s[2], hval[3] * s[3]);
                hs.z = pack_bf16x2(hval[4] * s[4], hval[5] * s[5]);
                hs.w = pack_bf16x2(hval[6] * s[6], hval[7] * s[7]);
                *reinterpret_cast<uint4*>(aux1 + aux_off + j8 * 8) = hs;       // Hs
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) v[j8 * 8 + i] = row_ok ? hb[i] * hval[i] : 0.f;  // dscale integrand
            }
          } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[j8 * 8 + i] = 0.f;
          }
          const int chunk = 4 * half + j8;   // 16-byte chunk inside the 128-byte swizzled row
          *reinterpret_cast<uint4*>(hs_gen + row_in_tile * 128 + ((chunk ^ (row_in_tile & 7)) << 4)) = to_smem;
        }
        tc_fence_before();
        fence_proxy_async_smem();   // generic-proxy SMEM writes -> visible to the tensor-core (async) proxy
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(hs_ready_remote);
        if (warp == 0) AQ_TRACE(item_iter, 12);
        if (MODE == 1 && p.g_scale != nullptr && grp == 0) {
          // dscale[b, j] += sum over this tile's rows of dHs * H
          const long long first_row = (long long)m0 + q * 32;
          const bool uniform = (p.tokens % 32 == 0) && (first_row + 32 <= p.M);
          if (uniform) {
            warp_colsum<32>(v, lane);
            if (hc0 + lane < p.r) atomicAdd(p.g_scale + sample * p.ld_r + hc0 + lane, v[0]);
          } else if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (hc0 + j < p.r) atomicAdd(p.g_scale + sample * p.ld_r + hc0 + j, v[j]);
          }
        }
      }
      if (!p.has_main) continue;
      const __nv_bfloat16* const bias_p = q_.bias;
      const int n_cols = q_.N;
      const __nv_bfloat16* const addend = p.accum_y ? q_.y : q_.res;
      const long long ld_add = p.accum_y ? q_.ldy : q_.ldres;
      const CUtensorMap* const tmap_y = &q_.tmap_y;
      for (int nt = nt_begin; nt < nt_end; ++nt) {
        // ---------------- tile epilogue: accumulator -> (+bias) -> bf16 -> staging tile -> bulk-tensor (TMA) stores ----------------
        const int col0 = nt * BN + half * NC;   // first global column of this warp's slice
        const uint32_t buf = acc_iter & 1u;
        // this warp's NC bias values: global loads issued before the accumulator wait, parked in SMEM after the TMEM loads
        // are in flight (a dependent load between tcgen05.ld and the stores was 11-16 % of the epilogue's stall samples)
        uint32_t bias_r[(NC / 2 + 31) / 32] = {};
        if (bias_p != nullptr) {
          const uint32_t* bw = reinterpret_cast<const uint32_t*>(bias_p + col0);
#pragma unroll
          for (int j = 0; j < (NC / 2 + 31) / 32; ++j) {
            const int w = j * 32 + lane;
            bias_r[j] = (w < NC / 2 && col0 + 2 * w < n_cols) ? __ldg(bw + w) : 0u;
          }
        }
        mbar_wait(acc_full_bar(buf), (acc_iter >> 1) & 1u);
        tc_fence_after();
        if (warp == 0 && nt - nt_begin < 4) AQ_TRACE(item_iter, 13 + 3 * (nt - nt_begin));
        uint32_t t[NC];
#pragma unroll
        for (int c = 0; c < NC / 16; ++c) tmem_ld_32x16(tmem_base + lane_base + buf * BN + half * NC + c * 16, t + c * 16);
        if (bias_p != nullptr) {
#pragma unroll
          for (int j = 0; j < (NC / 2 + 31) / 32; ++j) {
            const int w = j * 32 + lane;
            if (w < NC / 2) bias_sm[w] = bias_r[j];
          }
        }
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(acc_empty_remote[buf]);   // drained -> the MMA thread may reuse this accumulator
        if (warp == 0 && nt - nt_begin < 4) AQ_TRACE(item_iter, 14 + 3 * (nt - nt_begin));
        ++acc_iter;
        if (col0 >= n_cols) continue;   // whole slice past the last column (partial last tile)
        constexpr int PC = L::kPassCols;   // columns per store pass
        constexpr int kCpr = PC / 8;       // 16-byte chunks per row and pass
        const int row_base = m0 + q * 32;
#pragma unroll
        for (int pass = 0; pass < L::kPasses; ++pass) {
          const int pcol0 = col0 + pass * PC;
          if (pcol0 >= n_cols) break;        // warp-uniform: the rest of the slice lies past the last column
          const uint32_t buf = stg_iter & 1u;
          // the bulk store that read this buffer two passes ago has finished reading it (at most one store stays in flight)
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
#pragma unroll
          for (int c8 = 0; c8 < kCpr; ++c8) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(t[pass * PC + c8 * 8 + i]);
            if (bias_p != nullptr) {
              const uint4 bw = *reinterpret_cast<const uint4*>(bias_sm + (pass * PC + c8 * 8) / 2);
              const uint32_t bb[4] = {bw.x, bw.y, bw.z, bw.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                f[2 * i] += bf16_lo(bb[i]);
                f[2 * i + 1] += bf16_hi(bb[i]);
              }
            }
            // addend read from memory: the Y of the previous rank chunk (Y += this chunk's Hs Up^T), or the residual stream the
            // caller adds to this projection (x + to_out(...), x + ff(...), x + proj_out(...) of a transformer block) -- the lane
            // owns row `grow`; 16 bytes = 8 columns of it
            if (addend != nullptr) {
              const int c = pcol0 + c8 * 8;
              if (row_ok && c < n_cols) {
                const uint4 yw = __ldg(reinterpret_cast<const uint4*>(addend + (size_t)grow * ld_add + c));
                const uint32_t yy[4] = {yw.x, yw.y, yw.z, yw.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  f[2 * i] += bf16_lo(yy[i]);
                  f[2 * i + 1] += bf16_hi(yy[i]);
                }
              }
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]);
            o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]);
            o.w = pack_bf16x2(f[6], f[7]);
            const int chunk = L::kStgSwizzle64 ? (c8 ^ ((lane >> 1) & 3)) : c8;
            *reinterpret_cast<uint4*>(stg_gen + buf * L::kStgBufBytes + lane * L::kStgRowBytes + chunk * 16) = o;
          }
          fence_proxy_async_smem();        // generic-proxy writes -> visible to the bulk-copy (async) proxy
          __syncwarp();
          if (warp == 0 && nt == nt_begin) AQ_TRACE(item_iter, 25 + 3 * pass);
          if (lane == 0) {
            // rows >= M and columns >= N are clipped by the tensor map; the warp moves on while the copy engine drains the
            // buffer (measured before: 5 st.global.v4 per pass stalled ~0.36 us on the L2 write burst of all CTAs)
            tma_store_2d(tmap_y, stg_u32 + buf * L::kStgBufBytes, pcol0, row_base);
            tma_store_commit();
          }
          ++stg_iter;
          if (warp == 0 && nt == nt_begin) AQ_TRACE(item_iter, 27 + 3 * pass);
        }
        if (warp == 0 && nt - nt_begin < 4) AQ_TRACE(item_iter, 15 + 3 * (nt - nt_begin));
      }
    }
    // every bulk store of this warp has READ its staging buffer before the shared memory goes away; the global writes themselves
    // complete asynchronously and are flushed by the end of the grid (waiting for full completion here cost ~0.5 us per CTA tail)
    if (lane == 0) tma_store_wait_read<0>();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA leaves while its pair may still read its SMEM / signal its barriers
  if (warp == kMmaWarp) tmem_dealloc_pair(tmem_base, 512);
  AQ_CTA_STAMP(2);
#ifdef AQ_GEMM_TRACE
  if (threadIdx.x == 0 && blockIdx.x < 256) g_cta_times[blockIdx.x * 4 + 3] = smid();
#endif
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Column tiles per work item: the H phase is paid once per item, wave quantisation once per launch.
static int pick_group(int num_n_tiles, int num_m_pairs, int K, int BN, bool has_lora, int slots) {
  int best_g = 1;
  double best_cost = 1e30;
  const double kb = (double)((K + kBlockK - 1) / kBlockK);
  // A [M, K] is streamed once per column-tile GROUP (the groups of a row block are num_m_pairs items apart, far beyond what the
  // 126 MB L2 keeps when A is large): every extra group re-reads A from HBM.  In the cost unit below (one k-block of one column =
  // 2 clocks of the pair's MMA rate) a byte costs 1.9e9 / 2 / 6.5e12 units.  Measured: (65 536, 1280 -> 320) without LoRA 69.8 us with
  // one tile per item, 59.2 us with both tiles in one item (profiles/r02_gemm_sweep_all_plain.log).
  const double a_bytes = (double)num_m_pairs * kPairM * (double)K * 2.0;
  const double a_reread = a_bytes > 48e6 ? a_bytes * (1.9e9 / 2.0 / 6.5e12) : 0.0;
  for (int g = 1; g <= num_n_tiles; ++g) {
    const int groups = (num_n_tiles + g - 1) / g;
    const long long items = (long long)groups * num_m_pairs;
    const long long waves = (items + slots - 1) / slots;
    // per item: g tiles of (kb + 1) k-blocks of width BN, plus the H phase (kb k-blocks of width 64); a single-tile
    // item cannot hide the H -> Hs round trip behind the next tile's main loop
    const double bubble = has_lora ? (g == 1 ? 6.0 * BN : 1.0 * BN) : 0.0;
    const double item_cost = g * (kb + (has_lora ? 1.0 : 0.0)) * BN + (has_lora ? kb * kRankPad : 0.0) + bubble;
    const double cost = (double)waves * item_cost + (groups - 1) * a_reread;
    if (cost < best_cost - 1e-9) { best_cost = cost; best_g = g; }
  }
  return best_g;
}

// `probs` are projections of the same A (same M, K, r, scale, tokens, mode); args[0] carries the shared operands.
template <int BN, int NP, int MODE, bool DUAL = false>
static int launch_bn(const LoraGemmArgs* probs, int nprob, cudaStream_t stream) {
  using L = SmemLayout<BN, DUAL>;
  const LoraGemmArgs& a = probs[0];
  static thread_local LoraGemmParams<NP> p;   // 14 KiB for NP = 32: kept off the stack
  memset(&p, 0, sizeof(p));
  const int has_lora = a.dn != nullptr;
  {
    uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.M};
    uint64_t str[1] = {(uint64_t)a.lda * 2};
    uint32_t box[2] = {kBlockK, kBlockM};
    int rc = make_tmap(&p.tmap_a, a.a, 2, 2, dims, str, box, kSwz128);
    if (rc) return rc;
  }
  p.scale = a.scale;
  p.aux_out1 = reinterpret_cast<__nv_bfloat16*>(a.aux_out1);
  p.h_in = reinterpret_cast<const __nv_bfloat16*>(a.h_in);
  p.g_scale = a.g_scale;
  p.tokens = a.tokens > 0 ? a.tokens : a.M;
  p.num_samples = (int)((a.M + p.tokens - 1) / p.tokens);
  p.M = (int)a.M; p.K = a.K; p.r = a.r;
  p.mode = a.mode; p.has_lora = has_lora; p.has_main = a.has_main;
  p.skip_base = a.skip_base; p.accum_y = a.accum_y;
  p.ld_r = a.ld_r > 0 ? a.ld_r : a.r;
  p.num_m_pairs = (int)((a.M + kPairM - 1) / kPairM);
  p.num_problems = nprob;

  const int sms = sm_count();
  if (sms < 2) return fail(AQ_ERR_LAUNCH, "no CUDA device");
  const int slots = sms / 2;   // CTA pairs resident at once
  long long items = 0;
  for (int i = 0; i < nprob; ++i) {
    const LoraGemmArgs& b = probs[i];
    LoraProblem& q = p.prob[i];
    if (a.has_main && !a.skip_base) {
      uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)b.N};
      uint64_t str[1] = {(uint64_t)a.K * 2};
      uint32_t box[2] = {kBlockK, (uint32_t)(BN / 2)};
      int rc = make_tmap(&q.tmap_w, b.w, 2, 2, dims, str, box, kSwz128);
      if (rc) return rc;
    }
    if (has_lora) {
      uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.r};
      uint64_t str[1] = {(uint64_t)a.K * 2};
      uint32_t box[2] = {kBlockK, kRankPad / 2};
      int rc = make_tmap(&q.tmap_dn, b.dn, 2, 2, dims, str, box, kSwz128);
      if (rc) return rc;
      if (a.has_main) {
        uint64_t udims[2] = {(uint64_t)a.r, (uint64_t)b.N};
        uint64_t ustr[1] = {(uint64_t)(a.ld_r > 0 ? a.ld_r : a.r) * 2};
        uint32_t ubox[2] = {kRankPad, (uint32_t)(BN / 2)};
        rc = make_tmap(&q.tmap_up, b.up, 2, 2, udims, ustr, ubox, kSwz128);
        if (rc) return rc;
      }
    }
    if (a.has_main) {
      uint64_t dims[2] = {(uint64_t)b.N, (uint64_t)a.M};
      uint64_t str[1] = {(uint64_t)b.ldy * 2};
      uint32_t box[2] = {(uint32_t)L::kPassCols, 32};
      int rc = make_tmap(&q.tmap_y, b.y, 2, 2, dims, str, box, L::kStgSwizzle64 ? kSwz64 : kSwzNone);
      if (rc) return rc;
    }
    q.bias = reinterpret_cast<const __nv_bfloat16*>(b.bias);
    q.res = reinterpret_cast<const __nv_bfloat16*>(b.res);
    q.ldres = b.ldres;
    q.y = reinterpret_cast<__nv_bfloat16*>(b.y);
    q.ldy = b.ldy;
    q.aux_out0 = reinterpret_cast<__nv_bfloat16*>(b.aux_out0);
    q.N = b.N;
    q.num_n_tiles = a.has_main ? (b.N + BN - 1) / BN : 1;
    if (a.force_group > 0) q.group_size = a.force_group;
    else if (nprob > 1) q.group_size = q.num_n_tiles < 2 ? q.num_n_tiles : 2;   // many projections fill the waves; 2 tiles hide the Hs bubble
    else q.group_size = pick_group(q.num_n_tiles, p.num_m_pairs, a.K, BN, has_lora != 0, slots);
    if (DUAL) {
      q.group_size = (q.group_size + 1) & ~1;   // tiles are processed in pairs
      if (q.group_size > q.num_n_tiles) q.group_size = q.num_n_tiles;
    }
    q.num_groups = (q.num_n_tiles + q.group_size - 1) / q.group_size;
    q.item_begin = (int)items;
    items += (long long)q.num_groups * p.num_m_pairs;
  }
  AQ_REQUIRE(items < (1ll << 31), AQ_ERR_BAD_SHAPE, "lora_gemm: too many work items");
  p.total_items = (int)items;
  const int grid = 2 * (int)(items < slots ? items : slots);

  AQ_OPT_IN_SMEM((lora_gemm_kernel<BN, NP, MODE, DUAL>), L::kTotal);
  PdlLaunch launch(dim3(grid), dim3(kThreads), L::kTotal, stream);
  AQ_CHECK_CUDA(cudaLaunchKernelEx(&launch.cfg, lora_gemm_kernel<BN, NP, MODE, DUAL>, p));
  AQ_LAUNCHED();
  return AQ_OK;
}

// AQ_GEMM_DUAL=0 disables the two-tiles-per-A-pass mode (A/B measurements)
static bool dual_enabled() {
  static const bool on = [] { const char* e = getenv("AQ_GEMM_DUAL"); return e == nullptr || e[0] != '0'; }();
  return on;
}

static int pick_bn(int N) {
  if (N <= 64) return 64;
  // Wide projections (the feed-forward layers, N >= 2560): the 192-column tile moves 28 KiB per 3.1 MFLOP k-block instead of
  // 26 KiB per 2.6 MFLOP and measured 4 - 12 % faster than 160 despite <= 5 % padded columns (profiles/r01_lora_kernel_check_v2.log).
  if (N >= 2560) return 192;
  // otherwise: fewest computed columns; ties go to the wider tile
  int best = 192, best_cols = 1 << 30;
  const int cand[3] = {192, 160, 128};
  for (int i = 0; i < 3; ++i) {
    const int cols = (N + cand[i] - 1) / cand[i] * cand[i];
    if (cols < best_cols) { best_cols = cols; best = cand[i]; }
  }
  return best;
}

static int validate(const LoraGemmArgs& a) {
  AQ_REQUIRE(a.M > 0 && a.K > 0, AQ_ERR_BAD_SHAPE, "lora_gemm: empty problem M=%lld K=%d", (long long)a.M, a.K);
  AQ_REQUIRE(a.M < (1ll << 31) - kPairM, AQ_ERR_BAD_SHAPE, "lora_gemm: M=%lld exceeds 2^31-1 rows", (long long)a.M);
  AQ_REQUIRE(a.K % 8 == 0, AQ_ERR_BAD_SHAPE, "lora_gemm: K=%d must be a multiple of 8", a.K);
  AQ_REQUIRE(!a.has_main || (a.N > 0 && a.N % 8 == 0), AQ_ERR_BAD_SHAPE, "lora_gemm: N=%d must be a positive multiple of 8", a.N);
  if (a.dn != nullptr) {
    AQ_REQUIRE(a.r >= 8 && a.r <= kRankPad && a.r % 8 == 0, AQ_ERR_BAD_SHAPE,
               "lora_gemm: rank chunk r=%d unsupported (one launch covers 8 <= r <= 64, r %% 8 == 0; larger ranks are chunked by the ABI layer)", a.r);
    AQ_REQUIRE(a.scale != nullptr, AQ_ERR_BAD_SHAPE, "lora_gemm: scale is NULL");
    AQ_REQUIRE(!a.has_main || a.up != nullptr, AQ_ERR_BAD_SHAPE, "lora_gemm: up is NULL");
  }
  AQ_REQUIRE(a.has_main || a.dn != nullptr, AQ_ERR_BAD_SHAPE, "lora_gemm: nothing to compute");
  AQ_REQUIRE(a.ld_r == 0 || (a.ld_r >= a.r && a.ld_r % 8 == 0), AQ_ERR_BAD_SHAPE, "lora_gemm: ld_r=%lld must be a multiple of 8 >= r", (long long)a.ld_r);
  AQ_REQUIRE(!a.skip_base || (a.dn != nullptr && a.has_main), AQ_ERR_BAD_SHAPE, "lora_gemm: skip_base needs LoRA operands and an output");
  AQ_REQUIRE(a.res == nullptr || (a.has_main && a.ldres % 8 == 0 && (reinterpret_cast<uintptr_t>(a.res) & 15u) == 0 && !a.accum_y), AQ_ERR_BAD_ALIGN,
             "lora_gemm: residual needs an output, a 16-byte aligned base and a leading dimension that is a multiple of 8");
  AQ_REQUIRE(a.lda % 8 == 0 && (!a.has_main || a.ldy % 8 == 0), AQ_ERR_BAD_ALIGN, "lora_gemm: leading dimensions must be multiples of 8 elements");
  AQ_REQUIRE(!a.has_main || (reinterpret_cast<uintptr_t>(a.y) & 15u) == 0, AQ_ERR_BAD_ALIGN, "lora_gemm: y must be 16-byte aligned");
  return AQ_OK;
}

int launch_lora_gemm(const LoraGemmArgs& a, cudaStream_t stream) {
  int rc = validate(a);
  if (rc) return rc;
  rc = check_arch();
  if (rc) return rc;
  const int bn = a.force_bn > 0 ? a.force_bn : pick_bn(a.has_main ? a.N : 64);
  switch (bn) {
    case 64: return a.mode ? launch_bn<64, 1, 1>(&a, 1, stream) : launch_bn<64, 1, 0>(&a, 1, stream);
    case 128: return a.mode ? launch_bn<128, 1, 1>(&a, 1, stream) : launch_bn<128, 1, 0>(&a, 1, stream);
    case 160:
      // K-heavy projections with an even number of 160-column tiles (N = 320 / 640 / 1280 against K >= 1024: the feed-forward output
      // projections, the dX of the feed-forward input projections): one pass over A feeds two column tiles -- 40 KB instead of 56 KB of
      // operands per pair of k-blocks and CTA.  The price: both accumulators belong to one pass, so the tile epilogues no longer
      // overlap the next pass; with >= 16 k-blocks per pass that is a few per cent against -29 % of L2 -> SM traffic.
      if (dual_enabled() && a.has_main && !a.skip_base && a.K >= 1024 && ((a.N + 159) / 160) % 2 == 0 && (a.force_group == 0 || a.force_group % 2 == 0))
        return a.mode ? launch_bn<160, 1, 1, true>(&a, 1, stream) : launch_bn<160, 1, 0, true>(&a, 1, stream);
      return a.mode ? launch_bn<160, 1, 1>(&a, 1, stream) : launch_bn<160, 1, 0>(&a, 1, stream);
    case 192: return a.mode ? launch_bn<192, 1, 1>(&a, 1, stream) : launch_bn<192, 1, 0>(&a, 1, stream);
    default: return fail(AQ_ERR_BAD_SHAPE, "lora_gemm: unsupported column tile %d", bn);
  }
}

// Several forward projections of ONE input in one launch (mode 0).  Every entry repeats the shared operands (a, lda, M, K, r,
// tokens, scale); w / bias / dn / up / y / ldy / aux_out0 / N differ.  All projections use the column tile of the first.
int launch_lora_gemm_grouped(const LoraGemmArgs* probs, int nprob, cudaStream_t stream) {
  AQ_REQUIRE(probs != nullptr && nprob >= 1 && nprob <= kMaxGroup, AQ_ERR_BAD_SHAPE, "lora_gemm_grouped: 1 ... %d projections, got %d", kMaxGroup, nprob);
  for (int i = 0; i < nprob; ++i) {
    int rc = validate(probs[i]);
    if (rc) return rc;
    AQ_REQUIRE(probs[i].mode == 0 && probs[i].has_main == 1, AQ_ERR_BAD_SHAPE, "lora_gemm_grouped: forward projections only");
    AQ_REQUIRE(probs[i].a == probs[0].a && probs[i].M == probs[0].M && probs[i].K == probs[0].K && probs[i].r == probs[0].r &&
                   probs[i].lda == probs[0].lda && probs[i].tokens == probs[0].tokens && probs[i].scale == probs[0].scale &&
                   (probs[i].dn != nullptr) == (probs[0].dn != nullptr),
               AQ_ERR_BAD_SHAPE, "lora_gemm_grouped: projection %d does not share the input / rank / scale of projection 0", i);
  }
  int rc = check_arch();
  if (rc) return rc;
  if (nprob == 1) return launch_lora_gemm(probs[0], stream);
  // one column tile for the whole group: 160 divides every SD width (320 / 640 / 1280 and their multiples)
  int bn = probs[0].force_bn > 0 ? probs[0].force_bn : 160;
  for (int i = 0; i < nprob && probs[0].force_bn <= 0; ++i)
    if (probs[i].N % 160 != 0) bn = 128;
  if (nprob <= 4) {
    if (bn == 160) return launch_bn<160, 4, 0>(probs, nprob, stream);
    return launch_bn<128, 4, 0>(probs, nprob, stream);
  }
  if (bn == 160) return launch_bn<160, kMaxGroup, 0>(probs, nprob, stream);
  return launch_bn<128, kMaxGroup, 0>(probs, nprob, stream);
}

}  // namespace aq

#ifdef AQ_GEMM_TRACE
extern "C" int aq_debug_gemm_cta_times(unsigned long long* host_out, int clear) {
  cudaDeviceSynchronize();
  int rc = (int)cudaMemcpyFromSymbol(host_out, aq::g_cta_times, sizeof(unsigned long long) * 256 * 4);
  if (clear) {
    static unsigned long long zeros[64 * 32];
    cudaMemcpyToSymbol(aq::g_cta_times, zeros, sizeof(unsigned long long) * 256 * 4);
    cudaMemcpyToSymbol(aq::g_trace, zeros, sizeof(unsigned long long) * 64 * 32);
  }
  return rc;
}
extern "C" int aq_debug_gemm_trace(unsigned long long* host_out, int n) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host_out, aq::g_trace, sizeof(unsigned long long) * (size_t)(n < 64 * 32 ? n : 64 * 32));
}
#endif
