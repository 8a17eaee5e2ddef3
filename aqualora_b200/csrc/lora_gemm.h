// Internal launch interface of the fused projection + LoRA kernel (lora_gemm.cu) and the weight-gradient
// contraction (lora_wgrad.cu).
#pragma once
#include "aq_common.h"

namespace aq {

struct LoraGemmArgs {
  const void* a; int64_t lda;       // [M, K] bf16
  const void* w;                    // [N, K] bf16
  const void* bias;                 // [N] bf16 or null
  const void* res = nullptr;        // [M, ldres] bf16 or null: residual added in the epilogue
  int64_t ldres = 0;
  const void* dn;                   // [r, K] bf16 or null (null: plain projection)
  const void* up;                   // [N, r] bf16
  const float* scale;               // [num_samples, r] fp32
  void* y; int64_t ldy;             // [M, N] bf16 (unused when has_main == 0)
  void* aux_out0;                   // mode 0: H [M, r] (optional);  mode 1: dH [M, r]
  void* aux_out1;                   // mode 1: Hs [M, r]
  const void* h_in;                 // mode 1: H [M, r] saved by the forward
  float* g_scale;                   // mode 1: [num_samples, r] accumulated (optional)
  int64_t M, tokens; int K, N, r;
  int mode;                         // 0 forward, 1 backward
  int has_main;                     // 0: H phase + mid epilogue only
  int force_bn, force_group;        // 0 = heuristic
  // rank chunking (r > 64 runs as ceil(r / 64) launches over 64-wide slices of the LoRA operands, see lora_abi.cu):
  int64_t ld_r = 0;                 // row stride (elements) of scale / H / dH / Hs and of `up` [N, r]; 0 = r
  int skip_base = 0;                // 1: no A W^T term (a later rank chunk): the tile starts from the LoRA product alone ...
  int accum_y = 0;                  // 1: ... and the epilogue adds the Y already in memory (Y += chunk)
};

int launch_lora_gemm(const LoraGemmArgs& a, cudaStream_t stream);
int launch_lora_gemm_grouped(const LoraGemmArgs* probs, int nprob, cudaStream_t stream);
int launch_wgrad(const void* pm, int64_t ldp, const void* qm, int64_t ldq, float* c, int64_t ldc, int64_t M, int I, int J,
                 int transpose_out, cudaStream_t stream);
// two contractions over the same M rows in one launch (problem 0 / 1: operands p, q, output c, widths I, J, transpose_out)
int launch_wgrad_pair(const void* p0, int64_t ldp0, const void* q0, int64_t ldq0, float* c0, int64_t ldc0, int I0, int J0, int t0,
                      const void* p1, int64_t ldp1, const void* q1, int64_t ldq1, float* c1, int64_t ldc1, int I1, int J1, int t1,
                      int64_t M, cudaStream_t stream);

// one layer's (one rank chunk's) pair of weight-gradient contractions over the same M rows:
//   c0 [I0, J] += p0^T [I0, M] q0 [M, J]            (dUp = G^T Hs)          c1 [J, I1] += (p1^T [I1, M] q1 [M, J])^T   (dDn = dH^T X)
struct WgradJob {
  const void* p0; int64_t ldp0; const void* q0; int64_t ldq0; float* c0; int64_t ldc0; int I0;
  const void* p1; int64_t ldp1; const void* q1; int64_t ldq1; float* c1; int64_t ldc1; int I1;
  int J; int64_t M;
};
int launch_wgrad_jobs(const WgradJob* jobs, int njobs, cudaStream_t stream);

}  // namespace aq
