// Message -> deployable weights: the two row-scale / rank-r update steps that follow PPFT training.
//   aq_lora_fold_down : down'[i, :] = (down[i, :] * m[i]) * scale         scripts/create_wm_lora.py:23-41 (diag(mapper(msg)) @ down * scale)
//   aq_lora_merge     : W[o, i]   += coef * sum_j up[o, j] * down[j, i]   scripts/merge_lora.py:98-120 (W + ratio * (U @ D) * alpha / dim)
// Both are HBM-bound fp32 kernels (one read + one write of the result matrix); the rank-r product runs on FFMA from shared memory.
#include "aq_common.h"

namespace aq {

__global__ void __launch_bounds__(256) fold_down_kernel(const float* __restrict__ down, const float* __restrict__ m,
                                                         float* __restrict__ out, int r, long long cols, float scale) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)r * cols) return;
  const int row = (int)(idx / cols);
  // two roundings, in the reference's order: (diag(m) @ down) is exactly m[i] * down[i, j], then "* scale"
  out[idx] = (down[idx] * m[row]) * scale;
}

constexpr int kMergeTile = 64, kMergeRank = 64;

// 64 x 64 tile of W per block, 256 threads x (4 x 4) outputs; up tile [64 x r] and down tile [r x 64] staged in SMEM
__global__ void __launch_bounds__(256) merge_kernel(float* __restrict__ w, const float* __restrict__ up, const float* __restrict__ down,
                                                     int dout, int din, int r, float coef) {
  __shared__ float us[kMergeTile][kMergeRank + 1];
  __shared__ float ds[kMergeRank][kMergeTile + 4];
  const int o0 = blockIdx.y * kMergeTile, i0 = blockIdx.x * kMergeTile;
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
  float acc[4][4] = {};
  // the rank is walked in tiles of 64 (the reference's released LoRAs are rank 320: scripts/create_wm_lora.py:19); the fp32 sum runs
  // over j in ascending order whatever the tiling
  for (int j0 = 0; j0 < r; j0 += kMergeRank) {
    if (j0 > 0) __syncthreads();
    for (int t = threadIdx.x; t < kMergeTile * kMergeRank; t += 256) {
      const int a = t / kMergeRank, j = t % kMergeRank;
      us[a][j] = (o0 + a < dout && j0 + j < r) ? up[(size_t)(o0 + a) * r + j0 + j] : 0.f;
      const int jj = t / kMergeTile, b = t % kMergeTile;
      ds[jj][b] = (j0 + jj < r && i0 + b < din) ? down[(size_t)(j0 + jj) * din + i0 + b] : 0.f;
    }
    __syncthreads();
    const int jn = min(kMergeRank, r - j0);
    for (int j = 0; j < jn; ++j) {
      const float4 d4 = *reinterpret_cast<const float4*>(&ds[j][tx * 4]);
      const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float u = us[ty * 4 + a][j];
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(u, dv[b], acc[a][b]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int o = o0 + ty * 4 + a;
    if (o >= dout) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = i0 + tx * 4 + b;
      if (i < din) w[(size_t)o * din + i] += coef * acc[a][b];
    }
  }
}

}  // namespace aq

using namespace aq;

extern "C" {

int aq_lora_fold_down(const float* down, const float* m, float* out, int r, int64_t cols, float scale, void* stream) {
  AQ_REQUIRE(down && m && out && r > 0 && cols > 0, AQ_ERR_BAD_SHAPE, "lora_fold_down: NULL operand or empty matrix");
  int rc = check_arch();
  if (rc) return rc;
  const long long n = (long long)r * cols;
  fold_down_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(down, m, out, r, cols, scale);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_lora_merge(float* w, const float* up, const float* down, int dout, int din, int r, float coef, void* stream) {
  AQ_REQUIRE(w && up && down && dout > 0 && din > 0, AQ_ERR_BAD_SHAPE, "lora_merge: NULL operand or empty matrix");
  AQ_REQUIRE(r > 0, AQ_ERR_BAD_SHAPE, "lora_merge: rank %d must be positive", r);
  int rc = check_arch();
  if (rc) return rc;
  dim3 grid((din + kMergeTile - 1) / kMergeTile, (dout + kMergeTile - 1) / kMergeTile);
  merge_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, up, down, dout, din, r, coef);
  AQ_LAUNCHED();
  return AQ_OK;
}

}  // extern "C"
