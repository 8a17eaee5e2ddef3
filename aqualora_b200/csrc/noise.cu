// noise_layers distortion stack (utils/noise_layers/*) as HBM-bound fp32 kernels over [B, 3, H, W] NCHW images in [-1, 1].
// Every random quantity a reference layer draws internally is an explicit argument (no hidden RNG), so the CPU oracle and
// these kernels consume identical parameters.  One read + one write of the image per layer:
//   jpeg         jpeg_compression.py:130-162   pad -> YUV -> 8x8 DCT -> keep (25, 9, 9) zig-zag coefficients -> IDCT -> RGB
//   crop_resize  noises.py:46-57               crop -> bilinear -> bilinear, fused (the intermediate image is never stored)
//   gauss_blur   noises.py:67-70               separable (3 x 9) Gaussian, per-sample sigma, reflect border
//   gauss_noise  noises.py:80-85               x + std * N(0, 1), Philox4x32-10 + Box-Muller in the kernel
//   color_jiggle noises.py:95-104              brightness / contrast / saturation / hue in a sampled order
#include "aq_common.h"

namespace aq {

// ================================================================================================================
// JPEG mask
// ================================================================================================================
// cos(pi * m / 16), m = 0..31
__host__ __device__ constexpr float cos16(int m) {
  switch (m & 31) {
    case 0: return 1.0f;
    case 1: return 0.98078528040323043f;
    case 2: return 0.92387953251128674f;
    case 3: return 0.83146961230254524f;
    case 4: return 0.70710678118654757f;
    case 5: return 0.55557023301960229f;
    case 6: return 0.38268343236508984f;
    case 7: return 0.19509032201612833f;
    case 8: return 0.0f;
    case 9: return -0.19509032201612819f;
    case 10: return -0.38268343236508973f;
    case 11: return -0.55557023301960196f;
    case 12: return -0.70710678118654746f;
    case 13: return -0.83146961230254535f;
    case 14: return -0.92387953251128674f;
    case 15: return -0.98078528040323043f;
    case 16: return -1.0f;
    case 17: return -0.98078528040323043f;
    case 18: return -0.92387953251128685f;
    case 19: return -0.83146961230254546f;
    case 20: return -0.70710678118654768f;
    case 21: return -0.55557023301960218f;
    case 22: return -0.38268343236509034f;
    case 23: return -0.19509032201612866f;
    case 24: return 0.0f;
    case 25: return 0.19509032201612830f;
    case 26: return 0.38268343236509000f;
    case 27: return 0.55557023301960184f;
    case 28: return 0.70710678118654735f;
    case 29: return 0.83146961230254524f;
    case 30: return 0.92387953251128652f;
    default: return 0.98078528040323032f;
  }
}
// forward 1-D factor D[k][n] = cos(pi/8 (n + 1/2) k)                     (jpeg_compression.py:8-18)
__host__ __device__ constexpr float dct_f(int k, int n) { return cos16(k * (2 * n + 1)); }
// inverse 1-D factor I[n][k] = ((k == 0 ? -1/2 : 0) + cos(..)) * sqrt(1/16)   (jpeg_compression.py:44-50)
__host__ __device__ constexpr float dct_i(int n, int k) { return k == 0 ? 0.125f : 0.25f * cos16(k * (2 * n + 1)); }
// rank of coefficient (a, b) in the reference's zig-zag order (jpeg_compression.py:31-41): diagonals a + b ascending;
// inside an odd diagonal b descending, inside an even one b ascending.  Only diagonals <= 7 matter for keep <= 36.
__host__ __device__ constexpr int zz_rank(int a, int b) {
  const int s = a + b;
  if (s > 7) return 64;
  return s * (s + 1) / 2 + ((s & 1) ? (s - b) : b);
}
__host__ __device__ constexpr int keep_count(int c) { return c == 0 ? 25 : 9; }   // yuv_keep_weights = (25, 9, 9)
__host__ __device__ constexpr bool kept(int c, int a, int b) { return zz_rank(a, b) < keep_count(c); }
// number of leading second-index values that hold any kept coefficient
__host__ __device__ constexpr int kb_max(int c) { return c == 0 ? 6 : 4; }
static_assert(kept(0, 6, 0) && kept(0, 3, 3) && kept(0, 0, 5) && !kept(0, 0, 6) && !kept(0, 2, 4) && kept(1, 0, 3) && kept(1, 2, 1) &&
                  !kept(1, 3, 0) && !kept(1, 0, 4),
              "zig-zag table (jpeg_compression.py:31-41)");

constexpr int kJpegBlocks = 64;                 // 8x8 blocks per CTA strip (512 columns)
constexpr int kJpegCols = kJpegBlocks * 8;
constexpr int kJpegThreads = kJpegBlocks * 3;   // one thread per (block, channel)
constexpr int kJpegRowStride = kJpegCols + 4;   // +4 floats: de-phases the 8 rows of a block over the banks

template <int C>
__device__ __forceinline__ void jpeg_block(float* __restrict__ plane, int bx) {
  // plane: [8][kJpegRowStride] for this channel; the thread owns columns bx*8 .. bx*8+7
  constexpr int KB = kb_max(C);
  float t[8][KB];   // t[y][b] = sum_x D[b][x] v[y][x]      (second index of the reference's coefficient = x frequency)
#pragma unroll
  for (int y = 0; y < 8; ++y) {
    const float4 lo = *reinterpret_cast<const float4*>(plane + y * kJpegRowStride + bx * 8);
    const float4 hi = *reinterpret_cast<const float4*>(plane + y * kJpegRowStride + bx * 8 + 4);
    const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int b = 0; b < KB; ++b) {
      float acc = 0.f;
#pragma unroll
      for (int x = 0; x < 8; ++x) acc = fmaf(dct_f(b, x), v[x], acc);
      t[y][b] = acc;
    }
  }
  // kept coefficients c[a][b] = sum_y D[a][y] t[y][b]; then the column IDCT u[y][b] = sum_a I[y][a] c[a][b]
  float u[8][KB];
#pragma unroll
  for (int b = 0; b < KB; ++b) {
    float c[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      c[a] = 0.f;
      if (kept(C, a, b)) {
#pragma unroll
        for (int y = 0; y < 8; ++y) c[a] = fmaf(dct_f(a, y), t[y][b], c[a]);
      }
    }
#pragma unroll
    for (int y = 0; y < 8; ++y) {
      float acc = 0.f;
#pragma unroll
      for (int a = 0; a < 8; ++a)
        if (kept(C, a, b)) acc = fmaf(dct_i(y, a), c[a], acc);
      u[y][b] = acc;
    }
  }
#pragma unroll
  for (int y = 0; y < 8; ++y) {
    float o[8];
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      float acc = 0.f;
#pragma unroll
      for (int b = 0; b < KB; ++b) acc = fmaf(dct_i(x, b), u[y][b], acc);
      o[x] = acc;
    }
    *reinterpret_cast<float4*>(plane + y * kJpegRowStride + bx * 8) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(plane + y * kJpegRowStride + bx * 8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
}

// grid: (ceil(Wp / 512), Hp / 8, B); block: 192 threads.
// BWD = false: y = T x with T = unpad . C2 . P . C1 . pad (C1 = RGB->YUV, P = per-channel masked block DCT round trip, C2 = YUV->RGB).
// BWD = true : gx = T^T gy.  P = sum over kept (a, b) of w_a w_b (d_a (x) d_b)(d_a (x) d_b)^T is symmetric (I = D^T diag(1/8, 1/4, ...)),
// pad^T = unpad and unpad^T = zero-pad, so T^T = unpad . C1^T . P . C2^T . pad: the SAME kernel with the two colour matrices
// transposed and swapped.
template <bool BWD>
__global__ void __launch_bounds__(kJpegThreads) jpeg_mask_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W) {
  extern __shared__ float sm[];   // [3][8][kJpegRowStride]
  const int col0 = blockIdx.x * kJpegCols;
  const int row0 = blockIdx.y * 8;
  const size_t plane_sz = (size_t)H * W;
  const float* xb = x + (size_t)blockIdx.z * 3 * plane_sz;
  float* yb = y + (size_t)blockIdx.z * 3 * plane_sz;
  const bool vec_ok = (W % 4 == 0);
  // ---- load RGB (zero beyond the image = the reference's ZeroPad2d), convert to YUV (jpeg_compression.py:53-57) ----
  for (int i = threadIdx.x; i < 8 * (kJpegCols / 4); i += kJpegThreads) {
    const int r = i / (kJpegCols / 4), c4 = (i % (kJpegCols / 4)) * 4;
    const int gr = row0 + r, gc = col0 + c4;
    float R[4] = {0, 0, 0, 0}, G[4] = {0, 0, 0, 0}, Bv[4] = {0, 0, 0, 0};
    if (gr < H && gc < W) {
      const size_t off = (size_t)gr * W + gc;
      if (vec_ok && gc + 3 < W) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(xb + off));
        const float4 b = __ldg(reinterpret_cast<const float4*>(xb + plane_sz + off));
        const float4 c = __ldg(reinterpret_cast<const float4*>(xb + 2 * plane_sz + off));
        R[0] = a.x; R[1] = a.y; R[2] = a.z; R[3] = a.w;
        G[0] = b.x; G[1] = b.y; G[2] = b.z; G[3] = b.w;
        Bv[0] = c.x; Bv[1] = c.y; Bv[2] = c.z; Bv[3] = c.w;
      } else {
        for (int k = 0; k < 4; ++k)
          if (gc + k < W) {
            R[k] = xb[off + k]; G[k] = xb[plane_sz + off + k]; Bv[k] = xb[2 * plane_sz + off + k];
          }
      }
    }
    float Y[4], U[4], V[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if constexpr (!BWD) {
        Y[k] = 0.299f * R[k] + 0.587f * G[k] + 0.114f * Bv[k];
        U[k] = -0.14713f * R[k] + -0.28886f * G[k] + 0.436f * Bv[k];
        V[k] = 0.615f * R[k] + -0.51499f * G[k] + -0.10001f * Bv[k];
      } else {   // C2^T (gR, gG, gB)
        Y[k] = R[k] + G[k] + Bv[k];
        U[k] = -0.39465f * G[k] + 2.03211f * Bv[k];
        V[k] = 1.13983f * R[k] + -0.58060f * G[k];
      }
    }
    float* d = sm + r * kJpegRowStride + c4;
    *reinterpret_cast<float4*>(d) = make_float4(Y[0], Y[1], Y[2], Y[3]);
    *reinterpret_cast<float4*>(d + 8 * kJpegRowStride) = make_float4(U[0], U[1], U[2], U[3]);
    *reinterpret_cast<float4*>(d + 16 * kJpegRowStride) = make_float4(V[0], V[1], V[2], V[3]);
  }
  __syncthreads();
  // ---- one thread per (8x8 block, channel): DCT -> mask -> IDCT entirely in registers ----
  {
    const int ch = threadIdx.x / kJpegBlocks, bx = threadIdx.x % kJpegBlocks;   // warps are channel-uniform
    if (col0 + bx * 8 < W) {
      if (ch == 0) jpeg_block<0>(sm, bx);
      else if (ch == 1) jpeg_block<1>(sm + 8 * kJpegRowStride, bx);
      else jpeg_block<2>(sm + 16 * kJpegRowStride, bx);
    }
  }
  __syncthreads();
  // ---- YUV -> RGB (jpeg_compression.py:60-64), un-pad, store ----
  for (int i = threadIdx.x; i < 8 * (kJpegCols / 4); i += kJpegThreads) {
    const int r = i / (kJpegCols / 4), c4 = (i % (kJpegCols / 4)) * 4;
    const int gr = row0 + r, gc = col0 + c4;
    if (gr >= H || gc >= W) continue;
    const float* s = sm + r * kJpegRowStride + c4;
    const float4 Y = *reinterpret_cast<const float4*>(s);
    const float4 U = *reinterpret_cast<const float4*>(s + 8 * kJpegRowStride);
    const float4 V = *reinterpret_cast<const float4*>(s + 16 * kJpegRowStride);
    const float yv[4] = {Y.x, Y.y, Y.z, Y.w}, uv[4] = {U.x, U.y, U.z, U.w}, vv[4] = {V.x, V.y, V.z, V.w};
    float R[4], G[4], Bv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if constexpr (!BWD) {
        R[k] = yv[k] + 1.13983f * vv[k];
        G[k] = yv[k] + -0.39465f * uv[k] + -0.58060f * vv[k];
        Bv[k] = yv[k] + 2.03211f * uv[k];
      } else {   // C1^T (gY, gU, gV)
        R[k] = 0.299f * yv[k] + -0.14713f * uv[k] + 0.615f * vv[k];
        G[k] = 0.587f * yv[k] + -0.28886f * uv[k] + -0.51499f * vv[k];
        Bv[k] = 0.114f * yv[k] + 0.436f * uv[k] + -0.10001f * vv[k];
      }
    }
    const size_t off = (size_t)gr * W + gc;
    if (vec_ok && gc + 3 < W) {
      *reinterpret_cast<float4*>(yb + off) = make_float4(R[0], R[1], R[2], R[3]);
      *reinterpret_cast<float4*>(yb + plane_sz + off) = make_float4(G[0], G[1], G[2], G[3]);
      *reinterpret_cast<float4*>(yb + 2 * plane_sz + off) = make_float4(Bv[0], Bv[1], Bv[2], Bv[3]);
    } else {
      for (int k = 0; k < 4; ++k)
        if (gc + k < W) {
          yb[off + k] = R[k]; yb[plane_sz + off + k] = G[k]; yb[2 * plane_sz + off + k] = Bv[k];
        }
    }
  }
}

// ================================================================================================================
// crop + bilinear + bilinear
// ================================================================================================================
struct CropResizeArgs {
  int H, W, top, left, ch, cw, rh, rw, oh, ow;
  float sy1, sx1, sy2, sx2;   // input/output size ratios of the two resizes (ATen area_pixel_compute_scale, align_corners=False)
};

__device__ __forceinline__ void bilinear_src(float scale, int dst, int in_size, int& i0, int& i1, float& l1) {
  float s = scale * (dst + 0.5f) - 0.5f;     // ATen area_pixel_compute_source_index
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - i0;
}

__device__ __forceinline__ float stage1_pixel(const float* __restrict__ src, const CropResizeArgs& a, int iy, int ix) {
  // pixel (iy, ix) of the intermediate (rh x rw) image = bilinear sample of the crop
  int y0, y1, x0, x1;
  float ly, lx;
  bilinear_src(a.sy1, iy, a.ch, y0, y1, ly);
  bilinear_src(a.sx1, ix, a.cw, x0, x1, lx);
  const float* r0 = src + (size_t)(a.top + y0) * a.W + a.left;
  const float* r1 = src + (size_t)(a.top + y1) * a.W + a.left;
  const float hy = 1.f - ly, hx = 1.f - lx;
  return hy * (hx * __ldg(r0 + x0) + lx * __ldg(r0 + x1)) + ly * (hx * __ldg(r1 + x0) + lx * __ldg(r1 + x1));
}

__global__ void crop_resize_kernel(const float* __restrict__ x, float* __restrict__ y, CropResizeArgs a, int planes) {
  const long long n = (long long)planes * a.oh * a.ow;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % a.ow);
    const int oy = (int)((idx / a.ow) % a.oh);
    const long long pl = idx / ((long long)a.ow * a.oh);
    const float* src = x + pl * (long long)a.H * a.W;
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_src(a.sy2, oy, a.rh, y0, y1, ly);
    bilinear_src(a.sx2, ox, a.rw, x0, x1, lx);
    const float p00 = stage1_pixel(src, a, y0, x0), p01 = stage1_pixel(src, a, y0, x1);
    const float p10 = stage1_pixel(src, a, y1, x0), p11 = stage1_pixel(src, a, y1, x1);
    const float hy = 1.f - ly, hx = 1.f - lx;
    y[idx] = hy * (hx * p00 + lx * p01) + ly * (hx * p10 + lx * p11);
  }
}

// adjoint of crop_resize_kernel: every output pixel scatters its gradient over the (up to) 16 crop pixels it read.  fp32 atomics:
// the summation order (not the set of terms) varies between runs.
__device__ __forceinline__ void stage1_scatter(float* __restrict__ dst, const CropResizeArgs& a, int iy, int ix, float g) {
  int y0, y1, x0, x1;
  float ly, lx;
  bilinear_src(a.sy1, iy, a.ch, y0, y1, ly);
  bilinear_src(a.sx1, ix, a.cw, x0, x1, lx);
  float* r0 = dst + (size_t)(a.top + y0) * a.W + a.left;
  float* r1 = dst + (size_t)(a.top + y1) * a.W + a.left;
  const float hy = 1.f - ly, hx = 1.f - lx;
  atomicAdd(r0 + x0, g * hy * hx);
  atomicAdd(r0 + x1, g * hy * lx);
  atomicAdd(r1 + x0, g * ly * hx);
  atomicAdd(r1 + x1, g * ly * lx);
}

__global__ void crop_resize_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, CropResizeArgs a, int planes) {
  const long long n = (long long)planes * a.oh * a.ow;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % a.ow);
    const int oy = (int)((idx / a.ow) % a.oh);
    const long long pl = idx / ((long long)a.ow * a.oh);
    float* dst = gx + pl * (long long)a.H * a.W;
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_src(a.sy2, oy, a.rh, y0, y1, ly);
    bilinear_src(a.sx2, ox, a.rw, x0, x1, lx);
    const float g = __ldg(gy + idx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    stage1_scatter(dst, a, y0, x0, g * hy * hx);
    stage1_scatter(dst, a, y0, x1, g * hy * lx);
    stage1_scatter(dst, a, y1, x0, g * ly * hx);
    stage1_scatter(dst, a, y1, x1, g * ly * lx);
  }
}

// ================================================================================================================
// Gaussian blur (ky x kx) with per-sample sigma, reflect border
// ================================================================================================================
constexpr int kBlurTileH = 16, kBlurTileW = 128, kBlurMaxK = 15;

__device__ __forceinline__ int reflect(int i, int n) {   // torch 'reflect' padding (no edge repeat)
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

// grid: (ceil(W / 128), ceil(H / 16), B * 3); block 256
__global__ void __launch_bounds__(256) gauss_blur_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          const float* __restrict__ sigmas, int H, int W, int ky, int kx) {
  extern __shared__ float sm[];
  const int ry = ky / 2, rx = kx / 2;
  const int in_h = kBlurTileH + 2 * ry, in_w = kBlurTileW + 2 * rx;
  float* tin = sm;                       // [in_h][in_w]
  float* tmid = sm + in_h * in_w;        // [in_h][kBlurTileW]   after the horizontal pass
  __shared__ float taps_x[kBlurMaxK], taps_y[kBlurMaxK];
  const int plane = blockIdx.z;
  const float sigma = sigmas[plane / 3];
  if (threadIdx.x < 32) {
    // normalised Gaussian taps exp(-d^2 / (2 sigma^2)) / sum  (kornia get_gaussian_kernel1d)
    const int t = threadIdx.x;
    float gx = 0.f, gy = 0.f;
    if (t < kx) { const float d = (float)(t - rx); gx = expf(-(d * d) / (2.f * sigma * sigma)); }
    if (t < ky) { const float d = (float)(t - ry); gy = expf(-(d * d) / (2.f * sigma * sigma)); }
    float sx = gx, sy = gy;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
    }
    if (t < kx) taps_x[t] = gx / sx;
    if (t < ky) taps_y[t] = gy / sy;
  }
  const float* src = x + (size_t)plane * H * W;
  float* dst = y + (size_t)plane * H * W;
  const int r0 = blockIdx.y * kBlurTileH - ry, c0 = blockIdx.x * kBlurTileW - rx;
  for (int i = threadIdx.x; i < in_h * in_w; i += blockDim.x) {
    const int r = i / in_w, c = i % in_w;
    const int gr = reflect(r0 + r, H), gc = reflect(c0 + c, W);
    tin[i] = (gr >= 0 && gr < H && gc >= 0 && gc < W) ? __ldg(src + (size_t)gr * W + gc) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < in_h * kBlurTileW; i += blockDim.x) {
    const int r = i / kBlurTileW, c = i % kBlurTileW;
    float acc = 0.f;
    for (int k = 0; k < kx; ++k) acc = fmaf(taps_x[k], tin[r * in_w + c + k], acc);
    tmid[i] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kBlurTileH * kBlurTileW; i += blockDim.x) {
    const int r = i / kBlurTileW, c = i % kBlurTileW;
    const int gr = blockIdx.y * kBlurTileH + r, gc = blockIdx.x * kBlurTileW + c;
    if (gr >= H || gc >= W) continue;
    float acc = 0.f;
    for (int k = 0; k < ky; ++k) acc = fmaf(taps_y[k], tmid[(r + k) * kBlurTileW + c], acc);
    dst[(size_t)gr * W + gc] = acc;
  }
}

// adjoint of gauss_blur_kernel as a gather.  Forward: y[r, c] = sum_ij ty[i] tx[j] x[refl(r + i - ry), refl(c + j - rx)].  Input pixel
// (p, q) is read through the direct index and, within a radius of the border, through the reflections -p and 2 (n - 1) - p:
// gx[p, q] = sum_ij ty[i] tx[j] sum over the (<= 3 x 3) index variants (pv, qv) of gy[pv - i + ry, qv - j + rx] inside the image.
// grid: (ceil(W / 128), ceil(H / 4), B * 3); block 128 x 4
__global__ void __launch_bounds__(512) gauss_blur_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx,
                                                              const float* __restrict__ sigmas, int H, int W, int ky, int kx) {
  __shared__ float taps_x[kBlurMaxK], taps_y[kBlurMaxK];
  const int ry = ky / 2, rx = kx / 2;
  const int plane = blockIdx.z;
  const float sigma = sigmas[plane / 3];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  if (tid < 32) {
    float g1 = 0.f, g2 = 0.f;
    if (tid < kx) { const float d = (float)(tid - rx); g1 = expf(-(d * d) / (2.f * sigma * sigma)); }
    if (tid < ky) { const float d = (float)(tid - ry); g2 = expf(-(d * d) / (2.f * sigma * sigma)); }
    float sx = g1, sy = g2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
    }
    if (tid < kx) taps_x[tid] = g1 / sx;
    if (tid < ky) taps_y[tid] = g2 / sy;
  }
  __syncthreads();
  const int q = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y * blockDim.y + threadIdx.y;
  if (p >= H || q >= W) return;
  const float* src = gy + (size_t)plane * H * W;
  // index variants through which (p, q) is read: direct, left/top reflection, right/bottom reflection
  int pv[3], qv[3], np = 1, nq = 1;
  pv[0] = p; qv[0] = q;
  if (p >= 1 && p <= ry) pv[np++] = -p;
  if (p <= H - 2 && p >= H - 1 - ry) pv[np++] = 2 * (H - 1) - p;
  if (q >= 1 && q <= rx) qv[nq++] = -q;
  if (q <= W - 2 && q >= W - 1 - rx) qv[nq++] = 2 * (W - 1) - q;
  float acc = 0.f;
  for (int a = 0; a < np; ++a)
    for (int i = 0; i < ky; ++i) {
      const int r = pv[a] - i + ry;
      if (r < 0 || r >= H) continue;
      float row = 0.f;
      for (int b = 0; b < nq; ++b)
        for (int j = 0; j < kx; ++j) {
          const int c = qv[b] - j + rx;
          if (c >= 0 && c < W) row = fmaf(taps_x[j], __ldg(src + (size_t)r * W + c), row);
        }
      acc = fmaf(taps_y[i], row, acc);
    }
  gx[(size_t)plane * H * W + (size_t)p * W + q] = acc;
}

// ================================================================================================================
// Gaussian noise: Philox4x32-10 counter RNG + Box-Muller
// ================================================================================================================
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// element e of the stream (seed, offset): counter = offset + e / 4, 4 normals per counter
__device__ __forceinline__ void philox_normal4(unsigned long long seed, unsigned long long ctr, float (&z)[4]) {
  uint32_t r[4];
  philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  // uniforms in (0, 1]: (u + 1) * 2^-32 ; Box-Muller pairs
  const float u0 = ((float)r[0] + 1.0f) * 2.3283064365386963e-10f, u1 = ((float)r[1] + 1.0f) * 2.3283064365386963e-10f;
  const float u2 = ((float)r[2] + 1.0f) * 2.3283064365386963e-10f, u3 = ((float)r[3] + 1.0f) * 2.3283064365386963e-10f;
  const float ra = sqrtf(-2.f * logf(u0)), rb = sqrtf(-2.f * logf(u2));
  float s, c;
  sincospif(2.f * u1, &s, &c);
  z[0] = ra * c; z[1] = ra * s;
  sincospif(2.f * u3, &s, &c);
  z[2] = rb * c; z[3] = rb * s;
}

// y = (x ? x : 0) + std * N(0,1);  x == nullptr writes the unit noise itself (std = 1): the tensor the oracle consumes
__global__ void gauss_noise_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float stdv,
                                   unsigned long long seed, unsigned long long offset) {
  const long long n4 = (n + 3) >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float z[4];
    philox_normal4(seed, offset + (unsigned long long)i, z);
    const long long e = i << 2;
    if (e + 3 < n && (n & 3) == 0) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (x != nullptr) v = __ldg(reinterpret_cast<const float4*>(x) + i);
      v.x += stdv * z[0]; v.y += stdv * z[1]; v.z += stdv * z[2]; v.w += stdv * z[3];
      reinterpret_cast<float4*>(y)[i] = v;
    } else {
      for (int k = 0; k < 4 && e + k < n; ++k) y[e + k] = (x != nullptr ? x[e + k] : 0.f) + stdv * z[k];
    }
  }
}

// ================================================================================================================
// Color jiggle (kornia ColorJiggle semantics restated in oracle/noise_oracle.py)
// ================================================================================================================
struct JiggleOrder { int op[4]; };
constexpr float kTwoPi = 6.283185307179586f;

// Forward-mode dual number carrying the three partial derivatives w.r.t. the input (r, g, b): running the SAME per-pixel code on
// duals yields the exact 3 x 3 Jacobian of the arithmetic the forward executes (clamps, hue wrap, the piecewise HSV maps included),
// and the backward is gx = J^T gy.  floor / integer selections are piecewise constant: zero derivative.
struct Dual3 {
  float v, d0, d1, d2;
};
__device__ __forceinline__ Dual3 mk(float v) { return Dual3{v, 0.f, 0.f, 0.f}; }
__device__ __forceinline__ Dual3 operator+(Dual3 a, Dual3 b) { return Dual3{a.v + b.v, a.d0 + b.d0, a.d1 + b.d1, a.d2 + b.d2}; }
__device__ __forceinline__ Dual3 operator-(Dual3 a, Dual3 b) { return Dual3{a.v - b.v, a.d0 - b.d0, a.d1 - b.d1, a.d2 - b.d2}; }
__device__ __forceinline__ Dual3 operator+(Dual3 a, float b) { return Dual3{a.v + b, a.d0, a.d1, a.d2}; }
__device__ __forceinline__ Dual3 operator-(Dual3 a, float b) { return Dual3{a.v - b, a.d0, a.d1, a.d2}; }
__device__ __forceinline__ Dual3 operator-(float a, Dual3 b) { return Dual3{a - b.v, -b.d0, -b.d1, -b.d2}; }
__device__ __forceinline__ Dual3 operator*(Dual3 a, float b) { return Dual3{a.v * b, a.d0 * b, a.d1 * b, a.d2 * b}; }
__device__ __forceinline__ Dual3 operator*(float a, Dual3 b) { return b * a; }
__device__ __forceinline__ Dual3 operator/(Dual3 a, float b) { return Dual3{a.v / b, a.d0 / b, a.d1 / b, a.d2 / b}; }
__device__ __forceinline__ Dual3 operator*(Dual3 a, Dual3 b) {
  return Dual3{a.v * b.v, a.d0 * b.v + a.v * b.d0, a.d1 * b.v + a.v * b.d1, a.d2 * b.v + a.v * b.d2};
}
__device__ __forceinline__ Dual3 operator/(Dual3 a, Dual3 b) {
  const float q = a.v / b.v, ib = 1.f / b.v;
  return Dual3{q, (a.d0 - q * b.d0) * ib, (a.d1 - q * b.d1) * ib, (a.d2 - q * b.d2) * ib};
}
__device__ __forceinline__ float val(float a) { return a; }
__device__ __forceinline__ float val(Dual3 a) { return a.v; }
__device__ __forceinline__ float t_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ float t_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ Dual3 t_max(Dual3 a, Dual3 b) { return a.v >= b.v ? a : b; }
__device__ __forceinline__ Dual3 t_min(Dual3 a, Dual3 b) { return a.v <= b.v ? a : b; }
__device__ __forceinline__ float t_const(float, float c) { return c; }
__device__ __forceinline__ Dual3 t_const(Dual3, float c) { return mk(c); }
// x shifted by a piecewise-constant amount (floor / fmod wrap): same derivative as x
__device__ __forceinline__ float t_shift(float x, float newv) { return newv; }
__device__ __forceinline__ Dual3 t_shift(Dual3 x, float newv) { return Dual3{newv, x.d0, x.d1, x.d2}; }
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }
__device__ __forceinline__ Dual3 clamp01(Dual3 a) { return (a.v < 0.f) ? mk(0.f) : ((a.v > 1.f) ? mk(1.f) : a); }

template <typename T>
__device__ __forceinline__ void rgb2hsv(T r, T g, T b, T& h, T& s, T& v) {
  const T maxc = t_max(r, t_max(g, b)), minc = t_min(r, t_min(g, b));
  v = maxc;
  const T delta = maxc - minc;
  s = delta / (maxc + 1e-8f);
  const T dz = val(delta) == 0.f ? t_const(delta, 1.f) : delta;
  const T rc = maxc - r, gc = maxc - g, bc = maxc - b;
  T hh = (val(maxc) == val(r)) ? (bc - gc) : ((val(maxc) == val(g)) ? (2.f * dz + rc - bc) : (4.f * dz + gc - rc));
  hh = hh / dz / 6.f;
  hh = t_shift(hh, val(hh) - floorf(val(hh)));            // python % 1.0
  h = hh * kTwoPi;
}

template <typename T>
__device__ __forceinline__ void hsv2rgb(T h, T s, T v, T& r, T& g, T& b) {
  const T h1 = h / kTwoPi;
  const T h6 = h1 * 6.f;
  float hi = floorf(val(h6));
  hi = hi - 6.f * floorf(hi / 6.f);          // floor(h*6) % 6
  const T m6 = t_shift(h6, val(h6) - 6.f * floorf(val(h6) / 6.f));   // (h*6) % 6
  const T f = m6 - hi;
  const T p = v * (1.f - s), q = v * (1.f - f * s), t = v * (1.f - (1.f - f) * s);
  const int i = (int)hi;
  switch (i) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

// the four ColorJiggle ops in the sampled order on one pixel in [0, 1]
template <typename T>
__device__ __forceinline__ void jiggle_pixel(T& r, T& g, T& bl, float pb, float pc, float ps, float ph, const JiggleOrder& order) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int op = order.op[k];
    if (op == 0) {
      const float d = pb - 1.f;
      r = clamp01(r + d); g = clamp01(g + d); bl = clamp01(bl + d);
    } else if (op == 1) {
      r = clamp01(r * pc); g = clamp01(g * pc); bl = clamp01(bl * pc);
    } else {
      T h, s, v;
      rgb2hsv(r, g, bl, h, s, v);
      if (op == 2) {
        s = clamp01(s * ps);
      } else {
        const float hv = val(h) + ph * kTwoPi;
        h = t_shift(h, fmodf(hv, kTwoPi));
      }
      hsv2rgb(h, s, v, r, g, bl);
    }
  }
}

// params [B, 4] = (brightness, contrast, saturation, hue) per sample
__global__ void color_jiggle_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ params,
                                    JiggleOrder order, int B, long long hw) {
  const long long n = (long long)B * hw;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const long long b = idx / hw, px = idx % hw;
    const float* src = x + b * 3 * hw + px;
    float r = src[0] / 2.f + 0.5f, g = src[hw] / 2.f + 0.5f, bl = src[2 * hw] / 2.f + 0.5f;   // [-1, 1] -> [0, 1]
    jiggle_pixel<float>(r, g, bl, params[b * 4 + 0], params[b * 4 + 1], params[b * 4 + 2], params[b * 4 + 3], order);
    float* dst = y + b * 3 * hw + px;
    dst[0] = r * 2.f - 1.f; dst[hw] = g * 2.f - 1.f; dst[2 * hw] = bl * 2.f - 1.f;
  }
}

// gx = J^T gy per pixel (the [-1, 1] <-> [0, 1] maps contribute 1/2 * 2 = 1)
__global__ void color_jiggle_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                                        const float* __restrict__ params, JiggleOrder order, int B, long long hw) {
  const long long n = (long long)B * hw;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const long long b = idx / hw, px = idx % hw;
    const float* src = x + b * 3 * hw + px;
    Dual3 r{src[0] / 2.f + 0.5f, 1.f, 0.f, 0.f}, g{src[hw] / 2.f + 0.5f, 0.f, 1.f, 0.f}, bl{src[2 * hw] / 2.f + 0.5f, 0.f, 0.f, 1.f};
    jiggle_pixel<Dual3>(r, g, bl, params[b * 4 + 0], params[b * 4 + 1], params[b * 4 + 2], params[b * 4 + 3], order);
    const float* gs = gy + b * 3 * hw + px;
    const float g0 = gs[0], g1 = gs[hw], g2 = gs[2 * hw];
    float* dst = gx + b * 3 * hw + px;
    dst[0] = g0 * r.d0 + g1 * g.d0 + g2 * bl.d0;
    dst[hw] = g0 * r.d1 + g1 * g.d1 + g2 * bl.d1;
    dst[2 * hw] = g0 * r.d2 + g1 * g.d2 + g2 * bl.d2;
  }
}

static int ew_grid(long long work, int block) {
  const int sms = sm_count() > 0 ? sm_count() : 148;
  long long blocks = (work + block - 1) / block;
  const long long cap = (long long)sms * 16;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace aq

using namespace aq;

extern "C" {

int aq_noise_jpeg(const float* x, float* y, int B, int H, int W, void* stream) {
  AQ_REQUIRE(x && y && B > 0 && H > 0 && W > 0, AQ_ERR_BAD_SHAPE, "noise_jpeg: bad arguments B=%d H=%d W=%d", B, H, W);
  AQ_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0, AQ_ERR_BAD_ALIGN,
             "noise_jpeg: pointers must be 16-byte aligned");
  int rc = check_arch();
  if (rc) return rc;
  const int smem = 3 * 8 * kJpegRowStride * (int)sizeof(float);
  AQ_OPT_IN_SMEM((jpeg_mask_kernel<false>), smem);
  dim3 grid((W + kJpegCols - 1) / kJpegCols, (H + 7) / 8, B);
  jpeg_mask_kernel<false><<<grid, kJpegThreads, smem, (cudaStream_t)stream>>>(x, y, H, W);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_noise_jpeg_bwd(const float* gy, float* gx, int B, int H, int W, void* stream) {
  AQ_REQUIRE(gy && gx && B > 0 && H > 0 && W > 0, AQ_ERR_BAD_SHAPE, "noise_jpeg_bwd: bad arguments B=%d H=%d W=%d", B, H, W);
  AQ_REQUIRE(((reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(gx)) & 15u) == 0, AQ_ERR_BAD_ALIGN,
             "noise_jpeg_bwd: pointers must be 16-byte aligned");
  int rc = check_arch();
  if (rc) return rc;
  const int smem = 3 * 8 * kJpegRowStride * (int)sizeof(float);
  AQ_OPT_IN_SMEM((jpeg_mask_kernel<true>), smem);
  dim3 grid((W + kJpegCols - 1) / kJpegCols, (H + 7) / 8, B);
  jpeg_mask_kernel<true><<<grid, kJpegThreads, smem, (cudaStream_t)stream>>>(gy, gx, H, W);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_noise_crop_resize(const float* x, float* y, int B, int H, int W, int top, int left, int crop_h, int crop_w, int resize_h,
                         int resize_w, int out_h, int out_w, void* stream) {
  AQ_REQUIRE(x && y && B > 0 && H > 0 && W > 0, AQ_ERR_BAD_SHAPE, "noise_crop_resize: bad arguments");
  AQ_REQUIRE(top >= 0 && left >= 0 && crop_h > 0 && crop_w > 0 && top + crop_h <= H && left + crop_w <= W, AQ_ERR_BAD_SHAPE,
             "noise_crop_resize: crop box (%d, %d, %d, %d) outside %d x %d", top, left, crop_h, crop_w, H, W);
  AQ_REQUIRE(resize_h > 0 && resize_w > 0 && out_h > 0 && out_w > 0, AQ_ERR_BAD_SHAPE, "noise_crop_resize: bad sizes");
  int rc = check_arch();
  if (rc) return rc;
  CropResizeArgs a;
  a.H = H; a.W = W; a.top = top; a.left = left; a.ch = crop_h; a.cw = crop_w; a.rh = resize_h; a.rw = resize_w; a.oh = out_h; a.ow = out_w;
  a.sy1 = (float)crop_h / (float)resize_h; a.sx1 = (float)crop_w / (float)resize_w;
  a.sy2 = (float)resize_h / (float)out_h; a.sx2 = (float)resize_w / (float)out_w;
  const long long n = (long long)B * 3 * out_h * out_w;
  crop_resize_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, a, B * 3);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_noise_crop_resize_bwd(const float* gy, float* gx, int B, int H, int W, int top, int left, int crop_h, int crop_w, int resize_h,
                             int resize_w, int out_h, int out_w, void* stream) {
  AQ_REQUIRE(gy && gx && B > 0 && H > 0 && W > 0, AQ_ERR_BAD_SHAPE, "noise_crop_resize_bwd: bad arguments");
  AQ_REQUIRE(top >= 0 && left >= 0 && crop_h > 0 && crop_w > 0 && top + crop_h <= H && left + crop_w <= W, AQ_ERR_BAD_SHAPE,
             "noise_crop_resize_bwd: crop box (%d, %d, %d, %d) outside %d x %d", top, left, crop_h, crop_w, H, W);
  AQ_REQUIRE(resize_h > 0 && resize_w > 0 && out_h > 0 && out_w > 0, AQ_ERR_BAD_SHAPE, "noise_crop_resize_bwd: bad sizes");
  int rc = check_arch();
  if (rc) return rc;
  CropResizeArgs a;
  a.H = H; a.W = W; a.top = top; a.left = left; a.ch = crop_h; a.cw = crop_w; a.rh = resize_h; a.rw = resize_w; a.oh = out_h; a.ow = out_w;
  a.sy1 = (float)crop_h / (float)resize_h; a.sx1 = (float)crop_w / (float)resize_w;
  a.sy2 = (float)resize_h / (float)out_h; a.sx2 = (float)resize_w / (float)out_w;
  AQ_CHECK_CUDA(cudaMemsetAsync(gx, 0, (size_t)B * 3 * H * W * sizeof(float), (cudaStream_t)stream));   // pixels outside the crop: zero
  const long long n = (long long)B * 3 * out_h * out_w;
  crop_resize_bwd_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(gy, gx, a, B * 3);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_noise_gauss_blur_bwd(const float* gy, float* gx, const float* sigmas, int B, int H, int W, int ky, int kx, void* stream) {
  AQ_REQUIRE(gy && gx && sigmas && B > 0 && H > 0 && W > 0, AQ_ERR_BAD_SHAPE, "noise_gauss_blur_bwd: bad arguments");
  AQ_REQUIRE(ky % 2 == 1 && kx % 2 == 1 && ky <= kBlurMaxK && kx <= kBlurMaxK && ky / 2 < H && kx / 2 < W, AQ_ERR_BAD_SHAPE,
             "noise_gauss_blur_bwd: kernel (%d, %d) must be odd, <= %d and smaller than the image", ky, kx, kBlurMaxK);
  int rc = check_arch();
  if (rc) return rc;
  dim3 block(128, 4);
  dim3 grid((W + 127) / 128, (H + 3) / 4, B * 3);
  gauss_blur_bwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(gy, gx, sigmas, H, W, ky, kx);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_noise_gauss_blur(const float* x, float* y, const float* sigmas, int B, int H, int W, int ky, int kx, void* stream) {
  AQ_REQUIRE(x && y && sigmas && B > 0 && H > 0 && W > 0, AQ_ERR_BAD_SHAPE, "noise_gauss_blur: bad arguments");
  AQ_REQUIRE(ky % 2 == 1 && kx % 2 == 1 && ky <= kBlurMaxK && kx <= kBlurMaxK && ky / 2 < H && kx / 2 < W, AQ_ERR_BAD_SHAPE,
             "noise_gauss_blur: kernel (%d, %d) must be odd, <= %d and smaller than the image", ky, kx, kBlurMaxK);
  int rc = check_arch();
  if (rc) return rc;
  const int in_h = kBlurTileH + 2 * (ky / 2), in_w = kBlurTileW + 2 * (kx / 2);
  const size_t smem = (size_t)(in_h * in_w + in_h * kBlurTileW) * sizeof(float);
  dim3 grid((W + kBlurTileW - 1) / kBlurTileW, (H + kBlurTileH - 1) / kBlurTileH, B * 3);
  gauss_blur_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, y, sigmas, H, W, ky, kx);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_noise_gauss_noise(const float* x, float* y, int64_t n, float std, uint64_t seed, uint64_t offset, void* stream) {
  AQ_REQUIRE(y && n > 0, AQ_ERR_BAD_SHAPE, "noise_gauss_noise: bad arguments");
  AQ_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0, AQ_ERR_BAD_ALIGN,
             "noise_gauss_noise: pointers must be 16-byte aligned");
  int rc = check_arch();
  if (rc) return rc;
  gauss_noise_kernel<<<ew_grid((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n, std, seed, offset);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_noise_color_jiggle(const float* x, float* y, const float* params, const int* order_host, int B, int H, int W, void* stream) {
  AQ_REQUIRE(x && y && params && order_host && B > 0 && H > 0 && W > 0, AQ_ERR_BAD_SHAPE, "noise_color_jiggle: bad arguments");
  JiggleOrder o;
  int seen = 0;
  for (int k = 0; k < 4; ++k) {
    AQ_REQUIRE(order_host[k] >= 0 && order_host[k] < 4, AQ_ERR_BAD_SHAPE, "noise_color_jiggle: order must be a permutation of 0..3");
    o.op[k] = order_host[k];
    seen |= 1 << order_host[k];
  }
  AQ_REQUIRE(seen == 15, AQ_ERR_BAD_SHAPE, "noise_color_jiggle: order must be a permutation of 0..3");
  int rc = check_arch();
  if (rc) return rc;
  const long long hw = (long long)H * W;
  color_jiggle_kernel<<<ew_grid((long long)B * hw, 256), 256, 0, (cudaStream_t)stream>>>(x, y, params, o, B, hw);
  AQ_LAUNCHED();
  return AQ_OK;
}

int aq_noise_color_jiggle_bwd(const float* x, const float* gy, float* gx, const float* params, const int* order_host, int B, int H,
                              int W, void* stream) {
  AQ_REQUIRE(x && gy && gx && params && order_host && B > 0 && H > 0 && W > 0, AQ_ERR_BAD_SHAPE, "noise_color_jiggle_bwd: bad arguments");
  JiggleOrder o;
  int seen = 0;
  for (int k = 0; k < 4; ++k) {
    AQ_REQUIRE(order_host[k] >= 0 && order_host[k] < 4, AQ_ERR_BAD_SHAPE, "noise_color_jiggle_bwd: order must be a permutation of 0..3");
    o.op[k] = order_host[k];
    seen |= 1 << order_host[k];
  }
  AQ_REQUIRE(seen == 15, AQ_ERR_BAD_SHAPE, "noise_color_jiggle_bwd: order must be a permutation of 0..3");
  int rc = check_arch();
  if (rc) return rc;
  const long long hw = (long long)H * W;
  color_jiggle_bwd_kernel<<<ew_grid((long long)B * hw, 256), 256, 0, (cudaStream_t)stream>>>(x, gy, gx, params, o, B, hw);
  AQ_LAUNCHED();
  return AQ_OK;
}

}  // extern "C"
