// extern "C" entry points of the watermark-LoRA path (include/aqualora_b200.h).
#include "lora_gemm.h"

namespace aq {

static thread_local int g_force_bn = 0;
static thread_local int g_force_group = 0;

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
  AQ_REQUIRE(x && w && y, AQ_ERR_BAD_SHAPE, "lora_linear_fwd: x, w and y must be non-NULL");
  AQ_REQUIRE(tokens_per_sample > 0 || down == nullptr, AQ_ERR_BAD_SHAPE, "lora_linear_fwd: tokens_per_sample must be > 0");
  LoraGemmArgs a;
  a.a = x; a.lda = ldx; a.w = w; a.bias = bias; a.dn = down; a.up = up; a.scale = scale; a.y = y; a.ldy = ldy;
  a.aux_out0 = h_save; a.aux_out1 = nullptr; a.h_in = nullptr; a.g_scale = nullptr;
  a.M = M; a.tokens = tokens_per_sample; a.K = din; a.N = dout; a.r = r; a.mode = 0; a.has_main = 1;
  a.force_bn = g_force_bn; a.force_group = g_force_group;
  return launch_lora_gemm(a, (cudaStream_t)stream);
}

int aq_lora_linear_fwd_grouped(const void* x, int64_t ldx, const aq_lora_projection* proj, int nproj, const float* scale, int64_t M,
                               int64_t tokens_per_sample, int din, int r, void* stream) {
  AQ_REQUIRE(x && proj && nproj >= 1 && nproj <= 32, AQ_ERR_BAD_SHAPE, "lora_linear_fwd_grouped: x / proj NULL or nproj=%d outside 1 ... 32", nproj);
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

int aq_lora_linear_bwd(const void* gy, int64_t ldgy, const void* x, int64_t ldx, const void* w_t, const void* down_t,
                       const void* up_t, const float* scale, const void* h_save, void* gx, int64_t ldgx, float* g_down,
                       float* g_up, float* g_scale, int64_t M, int64_t tokens_per_sample, int din, int dout, int r, void* ws,
                       size_t ws_bytes, void* stream) {
  AQ_REQUIRE(gy && x && down_t && up_t && scale && h_save && g_down && g_up, AQ_ERR_BAD_SHAPE,
             "lora_linear_bwd: gy, x, down_t, up_t, scale, h_save, g_down, g_up must be non-NULL");
  AQ_REQUIRE((w_t == nullptr) == (gx == nullptr), AQ_ERR_BAD_SHAPE, "lora_linear_bwd: pass both w_t and gx, or neither");
  AQ_REQUIRE(tokens_per_sample > 0, AQ_ERR_BAD_SHAPE, "lora_linear_bwd: tokens_per_sample must be > 0");
  const size_t need = aq_lora_linear_bwd_workspace_bytes(M, r);
  AQ_REQUIRE(ws != nullptr && ws_bytes >= need, AQ_ERR_WORKSPACE, "lora_linear_bwd: workspace %zu bytes < required %zu", ws_bytes, need);
  AQ_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255u) == 0, AQ_ERR_BAD_ALIGN, "lora_linear_bwd: workspace must be 256-byte aligned");
  uint8_t* dh = reinterpret_cast<uint8_t*>(ws);
  uint8_t* hs = dh + need / 2;
  cudaStream_t st = (cudaStream_t)stream;
  // 1) dX = G W + ((G Up) (.) s) Dn, with dH / Hs / dscale produced by the mid-epilogue of the same kernel
  LoraGemmArgs a;
  a.a = gy; a.lda = ldgy; a.w = w_t; a.bias = nullptr; a.dn = up_t; a.up = down_t; a.scale = scale; a.y = gx; a.ldy = ldgx;
  a.aux_out0 = dh; a.aux_out1 = hs; a.h_in = h_save; a.g_scale = g_scale;
  a.M = M; a.tokens = tokens_per_sample; a.K = dout; a.N = din; a.r = r; a.mode = 1; a.has_main = (gx != nullptr);
  a.force_bn = g_force_bn; a.force_group = g_force_group;
  int rc = launch_lora_gemm(a, st);
  if (rc) return rc;
  // 2) dUp[dout, r] += G^T Hs  and  dDn[r, din] += dH^T X, one launch (grid.z = 2)
  return launch_wgrad_pair(gy, ldgy, hs, r, g_up, r, dout, r, 0, x, ldx, dh, r, g_down, din, din, r, 1, M, st);
}

int aq_wgrad_tn(const void* p, int64_t ldp, const void* q, int64_t ldq, float* c, int64_t ldc, int64_t M, int I, int J,
                int transpose_out, void* stream) {
  AQ_REQUIRE(p && q && c, AQ_ERR_BAD_SHAPE, "wgrad_tn: NULL operand");
  return launch_wgrad(p, ldp, q, ldq, c, ldc, M, I, J, transpose_out, (cudaStream_t)stream);
}

}  // extern "C"
