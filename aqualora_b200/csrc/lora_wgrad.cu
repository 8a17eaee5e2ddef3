// Skinny weight-gradient contraction for the LoRA backward on sm_100a:
//
//     C[i, j] += sum_m P[m, i] * Q[m, j]        P [M, I] bf16,  Q [M, J<=64] bf16,  C fp32
//
// dUp  = G^T  (H (.) s)   -> P = G  [M, dout],  Q = Hs [M, r],  C = g_up   [dout, r]
// dDn  = dH^T X           -> P = X  [M, din],   Q = dH [M, r],  C = g_down [r, din]  (transpose_out)
// (autograd of utils/lora_modules.py:12-19.)  The contraction runs over the token axis, so both operands are
// "MN-major" for tcgen05: the TMA box is [64 rows of m] x [64 contiguous i] and the UMMA descriptors walk it
// with LBO = one 64-wide column atom, SBO = 8 rows.  HBM-bound: P is streamed exactly once; split over row
// chunks (grid.y) with fp32 vector reductions into C.
#include <stdio.h>
#include <string.h>

#include "aq_ptx.cuh"
#include "lora_gemm.h"

namespace aq {

constexpr int kWgThreads = 256;
constexpr int kWgBlockI = 128;
constexpr int kWgBlockK = 64;   // rows of m per pipeline stage
constexpr int kWgJ = 64;
constexpr int kWgPBytes = kWgBlockK * kWgBlockI * 2;  // 16 KiB (two 64-column atoms)
constexpr int kWgQBytes = kWgBlockK * kWgJ * 2;       // 8 KiB
constexpr int kWgStageBytes = kWgPBytes + kWgQBytes;
constexpr int kWgStages = 4;
constexpr int kWgSmemBytes = kWgStages * kWgStageBytes + 256 + 1024;

struct WgradParams {
  CUtensorMap tmap_p;   // dims {I, M}, box {64, 64}, swizzle 128B
  CUtensorMap tmap_q;   // dims {J, M}, box {64, 64}, swizzle 128B
  float* c;
  long long ldc;
  int M, I, J;
  int kb_per_split;     // k-blocks (of 64 rows) per grid.y slice
  int transpose_out;
  int i_tiles, splits;  // extent of this problem inside the launch grid
};

// Many independent contractions per launch: the backward of one LoRA layer needs dUp = G^T Hs and dDn = dH^T X, each a few
// microseconds of work against a launch floor of 5 - 8 us, and nothing downstream depends on them (they only feed the flat
// gradient buffer) -- so the weight gradients of several layers are queued and leave in ONE launch.  The grid is one-dimensional:
// CTA -> (problem, row-tile of the output, slice of the reduction) through the prefix table cta_begin.
constexpr int kWgMaxProblems = 32;
struct WgradLaunch {
  WgradParams prob[kWgMaxProblems];
  int cta_begin[kWgMaxProblems + 1];
  int num_problems;
};

__global__ void __launch_bounds__(kWgThreads) lora_wgrad_kernel(const __grid_constant__ WgradLaunch launch) {
  griddep_launch_dependents();   // PDL (see aq_ptx.cuh)
  int pi = 0;
  while (pi + 1 < launch.num_problems && (int)blockIdx.x >= launch.cta_begin[pi + 1]) ++pi;
  const WgradParams& p = launch.prob[pi];
  const int local_cta = (int)blockIdx.x - launch.cta_begin[pi];
  const int cta_i = local_cta % p.i_tiles, cta_split = local_cta / p.i_tiles;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t bar_base = smem_base + kWgStages * kWgStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kWgStages + s); };
  const uint32_t acc_full_bar = bar_base + 8u * (2 * kWgStages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kWgStages + 1);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kWgStages * kWgStageBytes + 8 * (2 * kWgStages + 1));
  auto p_tile = [&](int s) { return smem_base + s * kWgStageBytes; };
  auto q_tile = [&](int s) { return smem_base + s * kWgStageBytes + kWgPBytes; };

  const int num_kb_total = (p.M + kWgBlockK - 1) / kWgBlockK;
  const int kb_begin = cta_split * p.kb_per_split;
  const int kb_end = min(kb_begin + p.kb_per_split, num_kb_total);
  const int i0 = cta_i * kWgBlockI;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_p);
    tma_prefetch_desc(&p.tmap_q);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  griddep_wait();                // PDL: the prologue above overlapped the previous kernel (the dX kernel that writes dH / Hs)

  if (kb_begin < kb_end) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), kWgStageBytes);
          const int m0 = kb * kWgBlockK;
          tma_load_2d(p_tile(stage), &p.tmap_p, full_bar(stage), i0, m0);
          tma_load_2d(p_tile(stage) + kWgPBytes / 2, &p.tmap_p, full_bar(stage), i0 + 64, m0);
          tma_load_2d(q_tile(stage), &p.tmap_q, full_bar(stage), 0, m0);
          if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = make_idesc_bf16(kWgBlockI, kWgJ, 1, 1);  // both operands MN-major
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kWgBlockK / 16; ++k) {
            // 16 rows of m per MMA = two 8-row swizzle groups = 2048 bytes
            const uint64_t adesc = make_smem_desc(p_tile(stage) + k * 2048, kWgPBytes / 2, 1024, 2);
            const uint64_t bdesc = make_smem_desc(q_tile(stage) + k * 2048, kWgQBytes, 1024, 2);
            umma_f16(tmem_base, adesc, bdesc, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(acc_full_bar);
      }
    } else if (warp >= 4) {
      const int q = warp - 4;
      mbar_wait(acc_full_bar, 0);
      tc_fence_after();
      const int i = i0 + q * 32 + lane;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t t[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + half * 32, t);
        tmem_wait_ld();
        if (i < p.I) {
          if (!p.transpose_out) {
            float* dst = p.c + (long long)i * p.ldc + half * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (half * 32 + j < p.J)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(t[j])),
                             "f"(__uint_as_float(t[j + 1])), "f"(__uint_as_float(t[j + 2])), "f"(__uint_as_float(t[j + 3]))
                             : "memory");
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (half * 32 + j < p.J) atomicAdd(p.c + (long long)(half * 32 + j) * p.ldc + i, __uint_as_float(t[j]));
          }
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 64);
}

static int fill_wgrad(WgradParams& p, const void* pm, int64_t ldp, const void* qm, int64_t ldq, float* c, int64_t ldc, int64_t M, int I,
                      int J, int transpose_out) {
  AQ_REQUIRE(M > 0 && I > 0 && J > 0, AQ_ERR_BAD_SHAPE, "wgrad: empty problem");
  AQ_REQUIRE(M < (1ll << 31), AQ_ERR_BAD_SHAPE, "wgrad: M too large");
  AQ_REQUIRE(J <= kWgJ && J % 8 == 0, AQ_ERR_BAD_SHAPE, "wgrad: J=%d must be a multiple of 8 and <= 64", J);
  AQ_REQUIRE(I % 8 == 0 && ldp % 8 == 0 && ldq % 8 == 0, AQ_ERR_BAD_ALIGN, "wgrad: I, ldp, ldq must be multiples of 8");
  AQ_REQUIRE(transpose_out || ldc % 4 == 0, AQ_ERR_BAD_ALIGN, "wgrad: ldc must be a multiple of 4");
  AQ_REQUIRE((reinterpret_cast<uintptr_t>(c) & 15u) == 0, AQ_ERR_BAD_ALIGN, "wgrad: C must be 16-byte aligned");
  int rc;
  {
    uint64_t dims[2] = {(uint64_t)I, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)ldp * 2};
    uint32_t box[2] = {64, kWgBlockK};
    rc = make_tmap(&p.tmap_p, pm, 2, 2, dims, str, box, kSwz128);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)J, (uint64_t)M};
    uint64_t str[1] = {(uint64_t)ldq * 2};
    uint32_t box[2] = {64, kWgBlockK};
    rc = make_tmap(&p.tmap_q, qm, 2, 2, dims, str, box, kSwz128);
    if (rc) return rc;
  }
  p.c = c; p.ldc = ldc; p.M = (int)M; p.I = I; p.J = J; p.transpose_out = transpose_out;
  p.i_tiles = (I + kWgBlockI - 1) / kWgBlockI;
  const int num_kb = (int)((M + kWgBlockK - 1) / kWgBlockK);
  const int sms = sm_count();
  if (sms <= 0) return fail(AQ_ERR_LAUNCH, "no CUDA device");
  // enough CTAs to keep HBM busy (2 per SM), but at least 8 k-blocks each so the reduction traffic stays small
  int splits = (2 * sms + p.i_tiles - 1) / p.i_tiles;
  const int max_splits = (num_kb + 7) / 8;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.kb_per_split = (num_kb + splits - 1) / splits;
  p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
  return AQ_OK;
}

static int launch_wgrad_n(WgradLaunch& l, int n, cudaStream_t stream) {
  int rc = check_arch();
  if (rc) return rc;
  AQ_OPT_IN_SMEM((lora_wgrad_kernel), kWgSmemBytes);
  int total = 0;
  for (int i = 0; i < n; ++i) {
    l.cta_begin[i] = total;
    total += l.prob[i].i_tiles * l.prob[i].splits;
  }
  l.cta_begin[n] = total;
  l.num_problems = n;
  PdlLaunch launch(dim3(total), dim3(kWgThreads), kWgSmemBytes, stream);
  AQ_CHECK_CUDA(cudaLaunchKernelEx(&launch.cfg, lora_wgrad_kernel, l));
  AQ_LAUNCHED();
  return AQ_OK;
}

int launch_wgrad(const void* pm, int64_t ldp, const void* qm, int64_t ldq, float* c, int64_t ldc, int64_t M, int I, int J,
                 int transpose_out, cudaStream_t stream) {
  static thread_local WgradLaunch l;
  memset(&l, 0, sizeof(l));
  int rc = fill_wgrad(l.prob[0], pm, ldp, qm, ldq, c, ldc, M, I, J, transpose_out);
  if (rc) return rc;
  return launch_wgrad_n(l, 1, stream);
}

int launch_wgrad_pair(const void* p0, int64_t ldp0, const void* q0, int64_t ldq0, float* c0, int64_t ldc0, int I0, int J0, int t0,
                      const void* p1, int64_t ldp1, const void* q1, int64_t ldq1, float* c1, int64_t ldc1, int I1, int J1, int t1,
                      int64_t M, cudaStream_t stream) {
  static thread_local WgradLaunch l;
  memset(&l, 0, sizeof(l));
  int rc = fill_wgrad(l.prob[0], p0, ldp0, q0, ldq0, c0, ldc0, M, I0, J0, t0);
  if (rc) return rc;
  rc = fill_wgrad(l.prob[1], p1, ldp1, q1, ldq1, c1, ldc1, M, I1, J1, t1);
  if (rc) return rc;
  return launch_wgrad_n(l, 2, stream);
}

// the weight gradients of several layers in as few launches as kWgMaxProblems allows
int launch_wgrad_jobs(const WgradJob* jobs, int njobs, cudaStream_t stream) {
  static thread_local WgradLaunch l;
  memset(&l, 0, sizeof(l));
  int n = 0;
  for (int j = 0; j < njobs; ++j) {
    const WgradJob& w = jobs[j];
    if (n + 2 > kWgMaxProblems) {
      int rc = launch_wgrad_n(l, n, stream);
      if (rc) return rc;
      memset(&l, 0, sizeof(l));
      n = 0;
    }
    int rc = fill_wgrad(l.prob[n], w.p0, w.ldp0, w.q0, w.ldq0, w.c0, w.ldc0, w.M, w.I0, w.J, 0);
    if (rc) return rc;
    rc = fill_wgrad(l.prob[n + 1], w.p1, w.ldp1, w.q1, w.ldq1, w.c1, w.ldc1, w.M, w.I1, w.J, 1);
    if (rc) return rc;
    n += 2;
  }
  return n > 0 ? launch_wgrad_n(l, n, stream) : AQ_OK;
}

}  // namespace aq
