// SecretEncoder (utils/models.py:51-81) as two small kernels:
//   msg [B, bits] -> Linear(bits, base^2) -> SiLU -> [B, 1, base, base] -> repeat to 4 channels -> nearest x(res/base)
//       -> Conv3x3(4 -> 4, pad 1)                                   => c_res [B, 4, res, res]        (kernel 1)
//   bilinear resize (align_corners=False) to the latent size and x_out = x + c                          (kernel 2)
// Work is microseconds; the point is one launch per stage instead of six library calls.
#include "aq_common.h"

namespace aq {

// one block per (sample, 8-row band of the res x res map)
__global__ void secret_encoder_map_kernel(const float* __restrict__ msg, const float* __restrict__ w1,
                                          const float* __restrict__ b1, const float* __restrict__ wc,
                                          const float* __restrict__ bc, float* __restrict__ c_res, int bits, int base,
                                          int res) {
  extern __shared__ float sm[];
  const int b = blockIdx.y;
  const int up = res / base;
  const int band = 8;
  const int y0 = blockIdx.x * band;
  // hidden rows needed: res rows y0-1 .. y0+band  -> base rows
  const int by0 = max(y0 - 1, 0) / up;
  const int by1 = min(y0 + band, res - 1) / up;
  const int nrows = by1 - by0 + 1;
  float* hid = sm;  // [nrows][base]
  const float* m = msg + (size_t)b * bits;
  for (int idx = threadIdx.x; idx < nrows * base; idx += blockDim.x) {
    const int o = (by0 + idx / base) * base + idx % base;
    float acc = b1[o];
    const float* wr = w1 + (size_t)o * bits;
    for (int i = 0; i < bits; ++i) acc = fmaf(wr[i], m[i], acc);
    hid[idx] = acc / (1.f + __expf(-acc));   // SiLU
  }
  __syncthreads();
  // the 4 input channels are identical copies, so the 4x4x3x3 conv collapses to 4 output filters of 3x3
  __shared__ float wsum[4][9];
  if (threadIdx.x < 36) {
    const int co = threadIdx.x / 9, t = threadIdx.x % 9;
    float s = 0.f;
    for (int ci = 0; ci < 4; ++ci) s += wc[(co * 4 + ci) * 9 + t];
    wsum[co][t] = s;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < band * res; idx += blockDim.x) {
    const int y = y0 + idx / res, x = idx % res;
    if (y >= res) continue;
    float v[9];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int yy = y + dy, xx = x + dx;
        float t = 0.f;
        if (yy >= 0 && yy < res && xx >= 0 && xx < res) t = hid[(yy / up - by0) * base + xx / up];
        v[(dy + 1) * 3 + (dx + 1)] = t;
      }
#pragma unroll
    for (int co = 0; co < 4; ++co) {
      float acc = bc[co];
#pragma unroll
      for (int t = 0; t < 9; ++t) acc = fmaf(wsum[co][t], v[t], acc);
      c_res[(((size_t)b * 4 + co) * res + y) * res + x] = acc;
    }
  }
}

__global__ void secret_encoder_resize_add_kernel(const float* __restrict__ c_res, const float* __restrict__ x,
                                                 float* __restrict__ c_out, float* __restrict__ x_out, int planes, int res,
                                                 int H, int W) {
  const long long n = (long long)planes * H * W;
  const float sy = (float)res / (float)H, sx = (float)res / (float)W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int xo = (int)(idx % W);
    const int yo = (int)((idx / W) % H);
    const long long pl = idx / ((long long)W * H);
    float fy = sy * (yo + 0.5f) - 0.5f, fx = sx * (xo + 0.5f) - 0.5f;   // F.interpolate(bilinear, align_corners=False)
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < res - 1 ? 1 : 0), x1 = x0 + (x0 < res - 1 ? 1 : 0);
    const float ly = fy - y0, lx = fx - x0;
    const float* src = c_res + pl * res * res;
    const float v = (1.f - ly) * ((1.f - lx) * src[y0 * res + x0] + lx * src[y0 * res + x1]) +
                    ly * ((1.f - lx) * src[y1 * res + x0] + lx * src[y1 * res + x1]);
    c_out[idx] = v;
    if (x_out != nullptr) x_out[idx] = x[idx] + v;
  }
}


// ------------------------------------------------------------------------------------------------------------------
// backward (train/latent_wm_pretrain.py:174,216: the encoder is trained through the VAE decoder, the noise layer and the
// message decoder).  g_c = dL/dc at the latent size; outputs accumulate (+=) into the parameter gradients.
// ------------------------------------------------------------------------------------------------------------------
// adjoint of the bilinear resize: scatter into the zeroed res x res map
__global__ void secret_encoder_resize_bwd_kernel(const float* __restrict__ g_c, float* __restrict__ g_cres, int planes, int res, int H,
                                                 int W) {
  const long long n = (long long)planes * H * W;
  const float sy = (float)res / (float)H, sx = (float)res / (float)W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int xo = (int)(idx % W);
    const int yo = (int)((idx / W) % H);
    const long long pl = idx / ((long long)W * H);
    float fy = sy * (yo + 0.5f) - 0.5f, fx = sx * (xo + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < res - 1 ? 1 : 0), x1 = x0 + (x0 < res - 1 ? 1 : 0);
    const float ly = fy - y0, lx = fx - x0;
    float* dst = g_cres + pl * res * res;
    const float g = g_c[idx];
    atomicAdd(dst + y0 * res + x0, g * (1.f - ly) * (1.f - lx));
    atomicAdd(dst + y0 * res + x1, g * (1.f - ly) * lx);
    atomicAdd(dst + y1 * res + x0, g * ly * (1.f - lx));
    atomicAdd(dst + y1 * res + x1, g * ly * lx);
  }
}

// one block per sample: conv weight / bias gradients and the gradient of the pre-activation a = W1 m + b1
__global__ void __launch_bounds__(256) secret_encoder_bwd_kernel(const float* __restrict__ g_cres, const float* __restrict__ msg,
                                                                 const float* __restrict__ w1, const float* __restrict__ b1,
                                                                 const float* __restrict__ wc, float* __restrict__ g_a,
                                                                 float* __restrict__ g_wc, float* __restrict__ g_bc, int bits, int base,
                                                                 int res) {
  extern __shared__ float sm[];
  const int b = blockIdx.x;
  const int up = res / base;
  const int nb = base * base;
  float* a_s = sm;         // [base^2] pre-activation
  float* u_s = sm + nb;    // [base^2] SiLU(a)
  __shared__ float wsum[4][9];
  __shared__ float red[40];
  const float* m = msg + (size_t)b * bits;
  for (int j = threadIdx.x; j < nb; j += blockDim.x) {
    float acc = b1[j];
    const float* wr = w1 + (size_t)j * bits;
    for (int i = 0; i < bits; ++i) acc = fmaf(wr[i], m[i], acc);
    a_s[j] = acc;
    u_s[j] = acc / (1.f + __expf(-acc));
  }
  if (threadIdx.x < 36) {
    const int co = threadIdx.x / 9, t = threadIdx.x % 9;
    float s = 0.f;
    for (int ci = 0; ci < 4; ++ci) s += wc[(co * 4 + ci) * 9 + t];
    wsum[co][t] = s;
  }
  if (threadIdx.x < 40) red[threadIdx.x] = 0.f;
  __syncthreads();
  const float* gb = g_cres + (size_t)b * 4 * res * res;
  // (1) dL/dwc[co, ci, t] = sum_{y, x} g[co, y, x] * in[y + dy, x + dx] (identical for the 4 input channels: they are copies)
  float pw[4][9], pb[4];
#pragma unroll
  for (int co = 0; co < 4; ++co) {
    pb[co] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) pw[co][t] = 0.f;
  }
  for (int idx = threadIdx.x; idx < res * res; idx += blockDim.x) {
    const int y = idx / res, x = idx % res;
    float v[9];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int yy = y + dy, xx = x + dx;
        v[(dy + 1) * 3 + (dx + 1)] = (yy >= 0 && yy < res && xx >= 0 && xx < res) ? u_s[(yy / up) * base + xx / up] : 0.f;
      }
#pragma unroll
    for (int co = 0; co < 4; ++co) {
      const float g = gb[(size_t)co * res * res + idx];
      pb[co] += g;
#pragma unroll
      for (int t = 0; t < 9; ++t) pw[co][t] = fmaf(g, v[t], pw[co][t]);
    }
  }
#pragma unroll
  for (int co = 0; co < 4; ++co) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float v = pw[co][t];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(&red[co * 9 + t], v);
    }
    float v = pb[co];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[36 + co], v);
  }
  __syncthreads();
  if (threadIdx.x < 36) {
    const int co = threadIdx.x / 9, t = threadIdx.x % 9;
    for (int ci = 0; ci < 4; ++ci) atomicAdd(g_wc + (co * 4 + ci) * 9 + t, red[threadIdx.x]);
  } else if (threadIdx.x < 40) {
    atomicAdd(g_bc + (threadIdx.x - 36), red[threadIdx.x]);
  }
  // (2) gradient of the hidden map: in[yy, xx] feeds out[yy - dy, xx - dx] through tap (dy, dx) of every output channel; the
  // nearest upsampling and the channel repeat sum their copies
  for (int j = threadIdx.x; j < nb; j += blockDim.x) {
    const int by = j / base, bx = j % base;
    float s = 0.f;
    for (int fy = 0; fy < up; ++fy)
      for (int fx = 0; fx < up; ++fx) {
        const int yy = by * up + fy, xx = bx * up + fx;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const int oy = yy - dy, ox = xx - dx;
            if (oy < 0 || oy >= res || ox < 0 || ox >= res) continue;
            const int t = (dy + 1) * 3 + (dx + 1);
#pragma unroll
            for (int co = 0; co < 4; ++co) s = fmaf(wsum[co][t], gb[((size_t)co * res + oy) * res + ox], s);
          }
      }
    const float a = a_s[j];
    const float sg = 1.f / (1.f + __expf(-a));
    g_a[(size_t)b * nb + j] = s * sg * (1.f + a * (1.f - sg));     // d SiLU(a) / da
  }
}

// dL/dW1[j, i] += sum_b g_a[b, j] msg[b, i];  dL/db1[j] += sum_b g_a[b, j]   (column i == bits is the bias)
__global__ void secret_encoder_wgrad_kernel(const float* __restrict__ g_a, const float* __restrict__ msg, float* __restrict__ g_w1,
                                            float* __restrict__ g_b1, int B, int bits, int nb) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nb * (bits + 1)) return;
  const int j = idx / (bits + 1), i = idx % (bits + 1);
  float s = 0.f;
  for (int b = 0; b < B; ++b) s = fmaf(g_a[(size_t)b * nb + j], i < bits ? msg[(size_t)b * bits + i] : 1.f, s);
  if (i < bits) g_w1[(size_t)j * bits + i] += s;
  else g_b1[j] += s;
}

}  // namespace aq

using namespace aq;

extern "C" {

size_t aq_secret_encoder_workspace_bytes(int B, int res) { return (size_t)B * 4 * res * res * sizeof(float); }

int aq_secret_encoder_fwd(const float* msg, const float* w1, const float* b1, const float* wc, const float* bc, const float* x,
                          float* c_out, float* x_out, int B, int bits, int base, int res, int H, int W, void* ws, void* stream) {
  AQ_REQUIRE(B > 0 && bits > 0 && base > 0 && res >= base && res % base == 0 && H > 0 && W > 0, AQ_ERR_BAD_SHAPE,
             "secret_encoder_fwd: bad shape B=%d bits=%d base=%d res=%d H=%d W=%d", B, bits, base, res, H, W);
  AQ_REQUIRE(msg && w1 && b1 && wc && bc && c_out, AQ_ERR_BAD_SHAPE, "secret_encoder_fwd: NULL operand");
  AQ_REQUIRE((x == nullptr) == (x_out == nullptr), AQ_ERR_BAD_SHAPE, "secret_encoder_fwd: pass both x and x_out, or neither");
  int rc = check_arch();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool same = (H == res && W == res);
  float* c_res = same && x == nullptr ? c_out : reinterpret_cast<float*>(ws);
  AQ_REQUIRE(c_res != nullptr, AQ_ERR_WORKSPACE, "secret_encoder_fwd: workspace required");
  const int up = res / base;
  dim3 grid((res + 7) / 8, B);
  const size_t smem = (size_t)(8 / up + 3) * base * sizeof(float);
  secret_encoder_map_kernel<<<grid, 256, smem, st>>>(msg, w1, b1, wc, bc, c_res, bits, base, res);
  AQ_LAUNCHED();
  if (c_res != c_out) {
    const long long n = (long long)B * 4 * H * W;
    int blocks = (int)((n + 255) / 256);
    const int cap = (sm_count() > 0 ? sm_count() : 148) * 8;
    if (blocks > cap) blocks = cap;
    secret_encoder_resize_add_kernel<<<blocks, 256, 0, st>>>(c_res, x, c_out, x_out, B * 4, res, H, W);
    AQ_LAUNCHED();
  }
  return AQ_OK;
}

size_t aq_secret_encoder_bwd_workspace_bytes(int B, int base, int res) {
  return ((size_t)B * 4 * res * res + (size_t)B * base * base) * sizeof(float);
}

// g_c [B, 4, H, W] -> g_w1 [base^2, bits] +=, g_b1 [base^2] +=, g_wc [4, 4, 3, 3] +=, g_bc [4] +=
int aq_secret_encoder_bwd(const float* g_c, const float* msg, const float* w1, const float* b1, const float* wc, float* g_w1, float* g_b1,
                          float* g_wc, float* g_bc, int B, int bits, int base, int res, int H, int W, void* ws, size_t ws_bytes,
                          void* stream) {
  AQ_REQUIRE(B > 0 && bits > 0 && base > 0 && res >= base && res % base == 0 && H > 0 && W > 0, AQ_ERR_BAD_SHAPE,
             "secret_encoder_bwd: bad shape B=%d bits=%d base=%d res=%d H=%d W=%d", B, bits, base, res, H, W);
  AQ_REQUIRE(g_c && msg && w1 && b1 && wc && g_w1 && g_b1 && g_wc && g_bc, AQ_ERR_BAD_SHAPE, "secret_encoder_bwd: NULL operand");
  AQ_REQUIRE(ws != nullptr && ws_bytes >= aq_secret_encoder_bwd_workspace_bytes(B, base, res), AQ_ERR_WORKSPACE,
             "secret_encoder_bwd: workspace too small");
  int rc = check_arch();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  float* g_cres_ws = reinterpret_cast<float*>(ws);
  float* g_a = g_cres_ws + (size_t)B * 4 * res * res;
  const float* g_cres = g_c;
  if (!(H == res && W == res)) {
    AQ_CHECK_CUDA(cudaMemsetAsync(g_cres_ws, 0, (size_t)B * 4 * res * res * sizeof(float), st));
    const long long n = (long long)B * 4 * H * W;
    int blocks = (int)((n + 255) / 256);
    const int cap = (sm_count() > 0 ? sm_count() : 148) * 8;
    if (blocks > cap) blocks = cap;
    secret_encoder_resize_bwd_kernel<<<blocks, 256, 0, st>>>(g_c, g_cres_ws, B * 4, res, H, W);
    AQ_LAUNCHED();
    g_cres = g_cres_ws;
  }
  const size_t smem = (size_t)2 * base * base * sizeof(float);
  AQ_REQUIRE(smem <= 48 * 1024, AQ_ERR_BAD_SHAPE, "secret_encoder_bwd: base=%d too large", base);
  secret_encoder_bwd_kernel<<<B, 256, smem, st>>>(g_cres, msg, w1, b1, wc, g_a, g_wc, g_bc, bits, base, res);
  AQ_LAUNCHED();
  const int n = base * base * (bits + 1);
  secret_encoder_wgrad_kernel<<<(n + 255) / 256, 256, 0, st>>>(g_a, msg, g_w1, g_b1, B, bits, base * base);
  AQ_LAUNCHED();
  return AQ_OK;
}

}  // extern "C"
