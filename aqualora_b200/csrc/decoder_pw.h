// Launch interface of the tensor-core pointwise convolution (decoder_pw.cu) used by the decoder chain (decoder.cu).
#pragma once
#include "aq_common.h"

namespace aq {

enum PwEpilogue { kPwNone = 0, kPwSilu = 1, kPwResidual = 2, kPwSiluPool = 3 };

struct PwTcArgs {
  const float* x;         // [M, K] NHWC pixels
  const float* w_hi;      // [N, K] weights with the low 13 mantissa bits cleared (BatchNorm folded)
  const float* w_lo;      // [N, K] w - w_hi
  const float* bias;      // [N]
  const float* se;        // [M / hw, K] squeeze-excitation scale applied to x, or null
  const float* residual;  // [M, N] or null
  float* y;               // [M, N]; kPwSiluPool: [M / hw, N] sums (accumulated with atomics)
  long long M;
  int K, N, hw, epi;
};

int launch_pointwise_tc(const PwTcArgs& a, cudaStream_t st);

}  // namespace aq
