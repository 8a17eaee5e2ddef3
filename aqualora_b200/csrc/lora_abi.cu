// extern "C" entry points of the watermark-LoRA path (include/aqualora_b200.h).
#include "lora_gemm.h"

namespace aq {

static thread_local int g_force_bn = 0;
static thread_local int g_force_group = 0;
constexpr int kRankChunk = 64;   // rank slice one launch of the fused kernel covers (TMEM: 2 accumulators + one 64-column H tile)

static inline const void* bf16_at(const void* base, int64_t elems) { return reinterpret_cast<const uint16_t*>(base) + elems; }
static inline void* bf16_at(void* base, int64_t elems) { return reinterpret_cast<uint16_t*>(base) + elems; }

}  // namespace aq

using namespace aq;

extern "C" {

// Test / tuning hook: pin the column-tile width (64/128/160/192) and the column tiles per work item of the
// calling thread's next launches; 0 restores the heuristics.
int aq_lora_set_tuning(int block_n, int group_size) {
  g_force_bn = block_n;
  g_force_group = group_size;
  return AQ_OK;
}

int aq_lora_linear_fwd(const void* x, int64_t ldx, const void* w, const void* bias, const void* down, const void* up,
                       const float* scale, void* y, int64_t ldy, void* h_save, int64_t M, int64_t tokens_per_sample, int din,
                       int dout, int r, void* stream) {
  return aq_lora_linear_fwd_residual(x, ldx, w, bias, down, up, scale, nullptr, 0, y, ldy, h_save, M, tokens_per_sample, din, dout, r, stream);
}

int aq_lora_linear_fwd_residual(const void* x, int64_t ldx, const void* w, const void* bias, const void* down, const void* up,
                                const float* scale, const void* residual, int64_t ldres, void* y, int64_t ldy, void* h_save, int64_t M,
                                int64_t tokens_per_sample, int din, int dout, int r, void* stream) {
  AQ_REQUIRE(x && w && y, AQ_ERR_BAD_SHAPE, "lora_linear_fwd: x, w and y must be non-NULL");
  AQ_REQUIRE(tokens_per_sample > 0 || down == nullptr, AQ_ERR_BAD_SHAPE, "lora_linear_fwd: tokens_per_sample must be > 0");
  AQ_REQUIRE(down == nullptr || (r >= 8 && r % 8 == 0), AQ_ERR_BAD_SHAPE,
             "lora_linear_fwd: rank r=%d must be a multiple of 8, >= 8 (pad smaller / odd ranks with zero rows)", r);
  // Ranks above 64 (the reference's released recipe is rank 320: train/README.md:34-48) run as ceil(r / 64) launches over 64-wide
  // slices of the LoRA operands: the first is the fused kernel as usual, each further one adds its slice's Hs Up^T to Y in place.
  const int chunks = down == nullptr ? 1 : (r + kRankChunk - 1) / kRankChunk;
  for (int c = 0; c < chunks; ++c) {
    const int r0 = c * kRankChunk, rc_ = down == nullptr ? 0 : (r - r0 < kRankChunk ? r - r0 : kRankChunk);
    LoraGemmArgs a;
    a.a = x; a.lda = ldx; a.w = w; a.bias = c == 0 ? bias : nullptr; a.scale = scale ? scale + r0 : nullptr; a.y = y; a.ldy = ldy;
    a.res = c == 0 ? residual : nullptr; a.ldres = ldres;
    a.dn = down ? bf16_at(down, (int64_t)r0 * din) : nullptr;        // rows r0 ... of down [r, din]
    a.up = up ? bf16_at(up, r0) : nullptr;                            // columns r0 ... of up [dout, r]
    a.aux_out0 = h_save ? bf16_at(h_save, r0) : nullptr; a.aux_out1 = nullptr; a.h_in = nullptr; a.g_scale = nullptr;
    a.M = M; a.tokens = tokens_per_sample; a.K = din; a.N = dout; a.r = rc_; a.mode = 0; a.has_main = 1;
    a.ld_r = r; a.skip_base = c > 0; a.accum_y = c > 0;
    a.force_bn = g_force_bn; a.force_group = g_force_group;
    int rc = launch_lora_gemm(a, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return AQ_OK;
}

int aq_lora_linear_fwd_grouped(const void* x, int64_t ldx, const aq_lora_projection* proj, int nproj, const float* scale, int64_t M,
                               int64_t tokens_per_sample, int din, int r, void* stream) {
  AQ_REQUIRE(x && proj && nproj >= 1 && nproj <= 32, AQ_ERR_BAD_SHAPE, "lora_linear_fwd_grouped: x / proj NULL or nproj=%d outside 1 ... 32", nproj);
  AQ_REQUIRE(r <= kRankChunk, AQ_ERR_BAD_SHAPE, "lora_linear_fwd_grouped: rank %d > %d runs through aq_lora_linear_fwd per projection", r, kRankChunk);
  LoraGemmArgs args[32];
  for (int i = 0; i < nproj; ++i) {
    AQ_REQUIRE(proj[i].w && proj[i].y, AQ_ERR_BAD_SHAPE, "lora_linear_fwd_grouped: projection %d has a NULL w / y", i);
    AQ_REQUIRE(tokens_per_sample > 0 || proj[i].down == nullptr, AQ_ERR_BAD_SHAPE, "lora_linear_fwd_grouped: tokens_per_sample must be > 0");
    LoraGemmArgs& a = args[i];
    a.a = x; a.lda = ldx; a.w = proj[i].w; a.bias = proj[i].bias; a.dn = proj[i].down; a.up = proj[i].up; a.scale = scale;
    a.y = proj[i].y; a.ldy = proj[i].ldy; a.aux_out0 = proj[i].h_save; a.aux_out1 = nullptr; a.h_in = nullptr; a.g_scale = nullptr;
    a.M = M; a.tokens = tokens_per_sample; a.K = din; a.N = proj[i].dout; a.r = r; a.mode = 0; a.has_main = 1;
    a.force_bn = g_force_bn; a.force_group = g_force_group;
  }
  return launch_lora_gemm_grouped(args, nproj, (cudaStream_t)stream);
}

size_t aq_lora_linear_bwd_workspace_bytes(int64_t M, int r) {
  // dH [M, r] bf16 + Hs [M, r] bf16, each padded to 256 bytes
  const size_t one = ((size_t)M * (size_t)r * 2 + 255) & ~(size_t)255;
  return 2 * one;
}

// dX (+ dH / Hs / dscale side outputs into the workspace) of one layer: part 1 of the backward.  The workspace must stay alive
// until the matching aq_lora_wgrad_batch job has been launched.
int aq_lora_linear_bwd_dx(const void* gy, int64_t ldgy, const void* w_t, const void* down_t, const void* up_t, const float* scale,
                          const void* h_save, void* gx, int64_t ldgx, float* g_scale, int64_t M, int64_t tokens_per_sample, int din,
                          int dout, int r, void* ws, size_t ws_bytes, void* stream) {
  AQ_REQUIRE(gy && down_t && up_t && scale && h_save, AQ_ERR_BAD_SHAPE, "lora_linear_bwd_dx: gy, down_t, up_t, scale, h_save must be non-NULL");
  AQ_REQUIRE((w_t == nullptr) == (gx == nullptr), AQ_ERR_BAD_SHAPE, "lora_linear_bwd_dx: pass both w_t and gx, or neither");
  AQ_REQUIRE(tokens_per_sample > 0, AQ_ERR_BAD_SHAPE, "lora_linear_bwd_dx: tokens_per_sample must be > 0");
  const size_t need = aq_lora_linear_bwd_workspace_bytes(M, r);
  AQ_REQUIRE(ws != nullptr && ws_bytes >= need, AQ_ERR_WORKSPACE, "lora_linear_bwd_dx: workspace %zu bytes < required %zu", ws_bytes, need);
  AQ_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255u) == 0, AQ_ERR_BAD_ALIGN, "lora_linear_bwd_dx: workspace must be 256-byte aligned");
  AQ_REQUIRE(r >= 8 && r % 8 == 0, AQ_ERR_BAD_SHAPE, "lora_linear_bwd_dx: rank r=%d must be a multiple of 8, >= 8", r);
  uint8_t* dh = reinterpret_cast<uint8_t*>(ws);
  uint8_t* hs = dh + need / 2;
  const int chunks = (r + kRankChunk - 1) / kRankChunk;
  for (int c = 0; c < chunks; ++c) {
    const int r0 = c * kRankChunk, rc_ = r - r0 < kRankChunk ? r - r0 : kRankChunk;
    // dX (+)= G W + ((G Up_c) (.) s_c) Dn_c, with dH_c / Hs_c / dscale_c produced by the mid-epilogue of the same kernel
    LoraGemmArgs a;
    a.a = gy; a.lda = ldgy; a.w = w_t; a.bias = nullptr; a.scale = scale + r0; a.y = gx; a.ldy = ldgx;
    a.dn = bf16_at(up_t, (int64_t)r0 * dout);          // rows r0 ... of Up^T [r, dout]
    a.up = bf16_at(down_t, r0);                        // columns r0 ... of Dn^T [din, r]
    a.aux_out0 = bf16_at(dh, r0); a.aux_out1 = bf16_at(hs, r0); a.h_in = bf16_at(h_save, r0); a.g_scale = g_scale ? g_scale + r0 : nullptr;
    a.M = M; a.tokens = tokens_per_sample; a.K = dout; a.N = din; a.r = rc_; a.mode = 1; a.has_main = (gx != nullptr);
    a.ld_r = r; a.skip_base = (gx != nullptr && c > 0); a.accum_y = a.skip_base;
    a.force_bn = g_force_bn; a.force_group = g_force_group;
    int rc = launch_lora_gemm(a, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return AQ_OK;
}

// dUp += G^T Hs and dDn += dH^T X for a batch of layers (part 2 of their backward) in as few launches as possible: nothing
// downstream of a layer's backward depends on these, so callers queue them and flush once per group of layers.
int aq_lora_wgrad_batch(const aq_wgrad_job* jobs, int njobs, void* stream) {
  AQ_REQUIRE(jobs != nullptr && njobs >= 1 && njobs <= 64, AQ_ERR_BAD_SHAPE, "lora_wgrad_batch: 1 ... 64 jobs, got %d", njobs);
  static thread_local WgradJob flat[64 * 8];
  int n = 0;
  for (int j = 0; j < njobs; ++j) {
    const aq_wgrad_job& w = jobs[j];
    AQ_REQUIRE(w.gy && w.x && w.ws && w.g_down && w.g_up && w.M > 0 && w.r >= 8 && w.r % 8 == 0 && w.r <= 8 * kRankChunk, AQ_ERR_BAD_SHAPE,
               "lora_wgrad_batch: job %d has a NULL operand or an unsupported rank %d", j, w.r);
    const size_t half = aq_lora_linear_bwd_workspace_bytes(w.M, w.r) / 2;
    const uint8_t* dh = reinterpret_cast<const uint8_t*>(w.ws);
    const uint8_t* hs = dh + half;
    const int chunks = (w.r + kRankChunk - 1) / kRankChunk;
    for (int c = 0; c < chunks; ++c) {
      const int r0 = c * kRankChunk, rc_ = w.r - r0 < kRankChunk ? w.r - r0 : kRankChunk;
      WgradJob& f = flat[n++];
      f.p0 = w.gy; f.ldp0 = w.ldgy; f.q0 = bf16_at(hs, r0); f.ldq0 = w.r; f.c0 = w.g_up + r0; f.ldc0 = w.r; f.I0 = w.dout;
      f.p1 = w.x; f.ldp1 = w.ldx; f.q1 = bf16_at(dh, r0); f.ldq1 = w.r; f.c1 = w.g_down + (int64_t)r0 * w.din; f.ldc1 = w.din; f.I1 = w.din;
      f.J = rc_; f.M = w.M;
    }
  }
  return launch_wgrad_jobs(flat, n, (cudaStream_t)stream);
}

int aq_lora_linear_bwd(const void* gy, int64_t ldgy, const void* x, int64_t ldx, const void* w_t, const void* down_t,
                       const void* up_t, const float* scale, const void* h_save, void* gx, int64_t ldgx, float* g_down,
                       float* g_up, float* g_scale, int64_t M, int64_t tokens_per_sample, int din, int dout, int r, void* ws,
                       size_t ws_bytes, void* stream) {
  AQ_REQUIRE(x && g_down && g_up, AQ_ERR_BAD_SHAPE, "lora_linear_bwd: x, g_down, g_up must be non-NULL");
  int rc = aq_lora_linear_bwd_dx(gy, ldgy, w_t, down_t, up_t, scale, h_save, gx, ldgx, g_scale, M, tokens_per_sample, din, dout, r, ws,
                                 ws_bytes, stream);
  if (rc) return rc;
  aq_wgrad_job job;
  job.gy = gy; job.ldgy = ldgy; job.x = x; job.ldx = ldx; job.ws = ws; job.g_down = g_down; job.g_up = g_up; job.M = M;
  job.din = din; job.dout = dout; job.r = r;
  return aq_lora_wgrad_batch(&job, 1, stream);
}

int aq_wgrad_tn(const void* p, int64_t ldp, const void* q, int64_t ldq, float* c, int64_t ldc, int64_t M, int I, int J,
                int transpose_out, void* stream) {
  AQ_REQUIRE(p && q && c, AQ_ERR_BAD_SHAPE, "wgrad_tn: NULL operand");
  return launch_wgrad(p, ldp, q, ldq, c, ldc, M, I, J, transpose_out, (cudaStream_t)stream);
}

}  // extern "C"
