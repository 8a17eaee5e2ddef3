// Launch interface of the fused expand 1x1 -> depthwise front half of an MBConv block (decoder_fused.cu).
#pragma once
#include "aq_common.h"

namespace aq {

struct FusedArgs {
  const float* x;      // [B, H, H, cin] NHWC
  const float* w_hi;   // [cexp, cin]   TF32 split of the BatchNorm-folded expand weight
  const float* w_lo;   // [cexp, cin]
  const float* b_e;    // [cexp]
  const float* w_d;    // [k * k, cexp] BatchNorm-folded depthwise weight
  const float* b_d;    // [cexp]
  float* y;            // [B, Ho, Ho, cexp]
  float* pooled;       // [B, cexp]  += sum over output pixels (squeeze of the SE block)
  int B, H, cin, cexp, k, stride;
};

// true for the MBConv shapes that have a fused kernel (the early EfficientNet-B1 blocks: 16 -> 96 k3 s2, 24 -> 144 k3 s1, 24 -> 144 k5 s2)
bool fused_expand_dw_supported(int cin, int cexp, int k, int stride, int H);
int launch_fused_expand_dw(const FusedArgs& a, cudaStream_t st);

}  // namespace aq
