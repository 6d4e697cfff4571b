// Training kernels of the tensor-core path (config 5, bf16 storage): everything BETWEEN the tcgen05 convolutions.
//
// Reference: MilAttentionFineTuneModel._forward_bags / .train (models/mil_attention_finetune.py:135-162, 164-253) -- backbone in
// TRAIN mode, BatchNorm batch statistics per 16-slice chunk of one bag, loss.backward() through torchvision's ResNet.  On this
// path activations, pre-BatchNorm convolution outputs and activation gradients are STORED in bf16 (what the tensor-core kernels
// read and write); every statistic, every normalisation and every parameter gradient is computed in f32 / f64 -- the numerics of
// torch's bf16 autocast, which the tests use as the calibration run (tests/test_gpu_training.py).  The FP32 twins of these
// kernels (train.cu) stay the 1e-5-class parity path.
//
// The step is bound by HBM between the convolutions (ResNet50 at 224^2 stores ~11 M activations per slice): bf16 storage halves
// every one of these passes against the f32 tape, and the kernels move 16 bytes per thread and access.
//
//   bn16_stats / bn16_apply          per-(group, channel) float64 sums over row slabs -> mean / invstd; y = bn(x) (+res) (ReLU)
//   bn16_bwd_reduce / bn16_bwd_apply sums of g and g*xhat; dx, and the residual branch's gradient (accumulated in place)
//   maxpool (train) / avgpool bwd    the stem's 3x3/2 max pool with the winner's position recorded (first maximum in window order)
//                                    and its gather backward; global average pool backward
//   stem_im2col3                     f32 NHWC3 network input -> bf16 patch matrix [M, 192] (7x7x3 = 147 columns, zero padded):
//                                    the stem's forward and weight gradient become 1x1 tensor-core GEMMs
//   dilate2 / scatter_add2           stride-2 data gradients: zero-dilated dY for the 3x3 convolutions (the transposed convolution
//                                    becomes a stride-1 one), scatter of the 1x1 downsample's gradient onto the even pixels
//   pack_conv_weights                f32 master weights [K,C,R,S] -> bf16 [K][R][S][C] (forward / wgrad layout) and the rotated,
//                                    transposed [C][R][S][K] (data-gradient operand), one launch per convolution
#include <algorithm>

#include "common.cuh"
#include "ops.cuh"

namespace pdf {

constexpr int kBn16Rows = 512;    // rows per reduction slab (small: the 64-channel layers have one channel block per group)

__device__ __forceinline__ float2 bf2_to_f2(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__device__ __forceinline__ uint32_t f2_to_bf2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  float2 t;
  t = bf2_to_f2(v.x); f[0] = t.x; f[1] = t.y;
  t = bf2_to_f2(v.y); f[2] = t.x; f[3] = t.y;
  t = bf2_to_f2(v.z); f[4] = t.x; f[5] = t.y;
  t = bf2_to_f2(v.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  v.x = f2_to_bf2(f[0], f[1]); v.y = f2_to_bf2(f[2], f[3]); v.z = f2_to_bf2(f[4], f[5]); v.w = f2_to_bf2(f[6], f[7]);
  return v;
}

// grid (C/64, groups, slabs); a lane owns two adjacent channels (one 4-byte load per row), 8 row lanes per block
__global__ void __launch_bounds__(256)
bn16_stats_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ goff, int C, double* __restrict__ acc) {
  __shared__ double red[4][8][33];
  const int g = blockIdx.y, lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + lane * 2;
  const int r0 = goff[g] + blockIdx.z * kBn16Rows;
  const int r1 = min(goff[g + 1], r0 + kBn16Rows);
  if (r0 >= r1) return;
  double s0 = 0.0, s1 = 0.0, q0 = 0.0, q1 = 0.0;
  if (c < C) {
    // strength-reduced addressing (one pointer, one constant stride) and float32 partial sums over the 8 rows of an iteration,
    // promoted to float64 once per iteration: the first version spent its time on 64-bit index arithmetic and conversions, not on
    // memory (42 us per layer against a 17 us HBM floor).  Eight bf16-derived values sum exactly enough in float32.
    const size_t step = (size_t)8 * C / 2;                              // 32-bit words per 8 rows
    const uint32_t* xp = reinterpret_cast<const uint32_t*>(x) + (((size_t)(r0 + ry) * C + c) >> 1);
    int left = (r1 - r0 - ry + 7) >> 3;                                   // rows this thread owns (r0+ry, +8, ...)
    for (; left >= 8; left -= 8, xp += 8 * step) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = xp[j * step];
      float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 f = bf2_to_f2(v[j]);
        a0 += f.x; a1 += f.y; b0 = fmaf(f.x, f.x, b0); b1 = fmaf(f.y, f.y, b1);
      }
      s0 += (double)a0; s1 += (double)a1; q0 += (double)b0; q1 += (double)b1;
    }
    for (; left > 0; --left, xp += step) {
      const float2 f = bf2_to_f2(*xp);
      s0 += (double)f.x; s1 += (double)f.y; q0 += (double)f.x * (double)f.x; q1 += (double)f.y * (double)f.y;
    }
  }
  red[0][ry][lane] = s0; red[1][ry][lane] = s1; red[2][ry][lane] = q0; red[3][ry][lane] = q1;
  __syncthreads();
  if (ry < 4 && c < C) {                       // warp ry finishes quantity ry: (sum c, sum c+1, sumsq c, sumsq c+1)
    double a = 0.0;
    for (int i = 0; i < 8; ++i) a += red[ry][i][lane];
    atomicAdd(acc + ((size_t)g * C + c + (ry & 1)) * 2 + (ry >> 1), a);
  }
}

__global__ void bn16_stats_finalize_kernel(const double* __restrict__ acc, const int* __restrict__ goff, int G, int C, float eps,
                                           float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ var_unbiased) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * C) return;
  const int g = i / C;
  const int rows = goff[g + 1] - goff[g];
  const double m = (double)max(rows, 1);
  const double mu = acc[(size_t)i * 2] / m;
  const double v = fmax(acc[(size_t)i * 2 + 1] - m * mu * mu, 0.0);
  const double var = v / m;
  mean[i] = (float)mu;
  invstd[i] = (float)(1.0 / sqrt(var + (double)eps));
  var_unbiased[i] = (float)(rows > 1 ? v / (double)(rows - 1) : var);
}

// y = (x - mean) * invstd * gamma + beta (+ residual) (ReLU), 8 channels (16 bytes) per thread and access.  C/8 is a power of two
// <= 256 and the grid stride a multiple of 256, so a thread keeps the SAME 8 channels over its whole loop: scale / shift live in
// registers and the loop body is load - 8 FMAs - store.
__global__ void __launch_bounds__(256)
bn16_apply_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ goff, int C, const float* __restrict__ mean,
                  const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                  const __nv_bfloat16* __restrict__ residual, int relu, __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ mask8) {
  const int g = blockIdx.y;
  const size_t lo = (size_t)goff[g] * C / 8, hi = (size_t)goff[g + 1] * C / 8;
  const int c = (int)(threadIdx.x % (C / 8)) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = invstd[(size_t)g * C + c + j] * __ldg(gamma + c + j);
    sh[j] = __ldg(beta + c + j) - mean[(size_t)g * C + c + j] * sc[j];
  }
  for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (size_t)gridDim.x * blockDim.x) {
    float f[8], r[8];
    unpack8(reinterpret_cast<const uint4*>(x)[i], f);
    if (residual) unpack8(reinterpret_cast<const uint4*>(residual)[i], r);
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = fmaf(f[j], sc[j], sh[j]);
      if (residual) v += r[j];
      if (relu) { bits |= (v > 0.f ? 1u : 0u) << j; v = fmaxf(v, 0.f); }
      f[j] = v;
    }
    reinterpret_cast<uint4*>(y)[i] = pack8(f);
    if (relu && mask8) mask8[i] = (uint8_t)bits;     // one bit per element: the backward passes read 1/16 of y's bytes for the ReLU mask
  }
}

__global__ void __launch_bounds__(256)
bn16_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ mask8, const __nv_bfloat16* __restrict__ x,
                       const int* __restrict__ goff, int C, const float* __restrict__ mean, const float* __restrict__ invstd, int relu,
                       double* __restrict__ acc) {
  __shared__ double red[4][8][33];
  const int g = blockIdx.y, lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + lane * 2;
  const int r0 = goff[g] + blockIdx.z * kBn16Rows;
  const int r1 = min(goff[g + 1], r0 + kBn16Rows);
  if (r0 >= r1) return;
  double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
  if (c < C) {
    const float mu0 = mean[(size_t)g * C + c], mu1 = mean[(size_t)g * C + c + 1];
    const float is0 = invstd[(size_t)g * C + c], is1 = invstd[(size_t)g * C + c + 1];
    const int sh = c & 7;                                   // this lane's two channels sit in bits sh, sh+1 of their mask byte
    const size_t step = (size_t)8 * C / 2;                  // 32-bit words per 8 rows (same addressing as bn16_stats_kernel)
    const size_t first = ((size_t)(r0 + ry) * C + c) >> 1;
    const uint32_t* dp = reinterpret_cast<const uint32_t*>(dy) + first;
    const uint32_t* xp = reinterpret_cast<const uint32_t*>(x) + first;
    const uint8_t* mp = mask8 + (first >> 2);
    const size_t mstep = step >> 2;
    int left = (r1 - r0 - ry + 7) >> 3;
    for (; left >= 4; left -= 4, dp += 4 * step, xp += 4 * step, mp += 4 * mstep) {
      uint32_t d[4], xv[4], m[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { d[j] = dp[j * step]; xv[j] = xp[j * step]; m[j] = relu ? ((uint32_t)mp[j * mstep] >> sh) : 3u; }
      float ga = 0.f, gb = 0.f, ha = 0.f, hb = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 gr = bf2_to_f2(d[j]);
        const float2 xf = bf2_to_f2(xv[j]);
        if (!(m[j] & 1u)) gr.x = 0.f;
        if (!(m[j] & 2u)) gr.y = 0.f;
        ga += gr.x; gb += gr.y;
        ha = fmaf(gr.x, (xf.x - mu0) * is0, ha); hb = fmaf(gr.y, (xf.y - mu1) * is1, hb);
      }
      a0 += (double)ga; a1 += (double)gb; b0 += (double)ha; b1 += (double)hb;
    }
    for (; left > 0; --left, dp += step, xp += step, mp += mstep) {
      float2 gr = bf2_to_f2(*dp);
      const float2 xf = bf2_to_f2(*xp);
      if (relu) { const uint32_t m = (uint32_t)*mp >> sh; if (!(m & 1u)) gr.x = 0.f; if (!(m & 2u)) gr.y = 0.f; }
      a0 += (double)gr.x; a1 += (double)gr.y;
      b0 += (double)gr.x * (double)((xf.x - mu0) * is0); b1 += (double)gr.y * (double)((xf.y - mu1) * is1);
    }
  }
  red[0][ry][lane] = a0; red[1][ry][lane] = a1; red[2][ry][lane] = b0; red[3][ry][lane] = b1;
  __syncthreads();
  if (ry < 4 && c < C) {
    double a = 0.0;
    for (int i = 0; i < 8; ++i) a += red[ry][i][lane];
    atomicAdd(acc + ((size_t)g * C + c + (ry & 1)) * 2 + (ry >> 1), a);
  }
}

__global__ void bn16_bwd_finalize_kernel(const double* __restrict__ acc, int G, int C, float* __restrict__ sums, float* __restrict__ dgamma,
                                         float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double a = 0.0, b = 0.0;
  for (int g = 0; g < G; ++g) {
    const double s0 = acc[((size_t)g * C + c) * 2], s1 = acc[((size_t)g * C + c) * 2 + 1];
    sums[(size_t)g * C + c] = (float)s0;
    sums[(size_t)(G + g) * C + c] = (float)s1;
    a += s0; b += s1;
  }
  dbeta[c] += (float)a;
  dgamma[c] += (float)b;
}

// dx = gamma * invstd * (g - sum_g/m - xhat * sum_gx/m);  dres (+)= g.  Per-thread channel constants as in bn16_apply_kernel:
// dx = k0 * g - k1 - k2 * x  with k0 = gamma*invstd, k2 = k0 * invstd * sum_gx/m, k1 = k0 * sum_g/m - k2 * mean
__global__ void __launch_bounds__(256)
bn16_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ mask8, const __nv_bfloat16* __restrict__ x,
                      const int* __restrict__ goff, int C, const float* __restrict__ mean, const float* __restrict__ invstd,
                      const float* __restrict__ gamma, int relu, const float* __restrict__ sum_g, const float* __restrict__ sum_gx,
                      __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dres, int dres_accumulate) {
  const int g = blockIdx.y;
  const float inv_m = 1.0f / (float)max(goff[g + 1] - goff[g], 1);
  const size_t lo = (size_t)goff[g] * C / 8, hi = (size_t)goff[g + 1] * C / 8;
  const int c = (int)(threadIdx.x % (C / 8)) * 8;
  float k0[8], k1[8], k2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const size_t gc = (size_t)g * C + c + j;
    const float is = invstd[gc];
    k0[j] = __ldg(gamma + c + j) * is;
    k2[j] = k0[j] * is * sum_gx[gc] * inv_m;
    k1[j] = k0[j] * sum_g[gc] * inv_m - k2[j] * mean[gc];
  }
  for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (size_t)gridDim.x * blockDim.x) {
    float gr[8], xv[8], o[8];
    unpack8(reinterpret_cast<const uint4*>(dy)[i], gr);
    unpack8(reinterpret_cast<const uint4*>(x)[i], xv);
    if (relu) {
      const uint32_t m = mask8[i];
#pragma unroll
      for (int j = 0; j < 8; ++j) if (!((m >> j) & 1u)) gr[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(k0[j], gr[j], -fmaf(k2[j], xv[j], k1[j]));
    reinterpret_cast<uint4*>(dx)[i] = pack8(o);
    if (dres) {
      if (dres_accumulate) {
        float d[8];
        unpack8(reinterpret_cast<const uint4*>(dres)[i], d);
#pragma unroll
        for (int j = 0; j < 8; ++j) gr[j] += d[j];
      }
      reinterpret_cast<uint4*>(dres)[i] = pack8(gr);
    }
  }
}

// 3x3 stride-2 pad-1 max pool, training form: the forward also records WHICH window element won (r*3 + s of the first maximum in
// row-major order, one byte per output element), so the backward is a gather: one thread per input pixel and 8 channels looks at
// the <= 4 windows that contain it and takes their gradient where the recorded winner is its own position -- no atomics, no
// memset, no re-reading of the 9-element windows.
__global__ void __launch_bounds__(256)
maxpool16_fwd_idx_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ idx, int N, int H, int W,
                         int C, int Ho, int Wo) {
  const int C8 = C / 8;
  const size_t total = (size_t)N * Ho * Wo * C8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    size_t t = i / C8;
    const int q = (int)(t % Wo); t /= Wo;
    const int p = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float best[8];
    uint32_t win[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; win[j] = 0; }
    for (int r = 0; r < 3; ++r) {
      const int hh = 2 * p - 1 + r;
      if (hh < 0 || hh >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int ww = 2 * q - 1 + s;
        if (ww < 0 || ww >= W) continue;
        float v[8];
        unpack8(reinterpret_cast<const uint4*>(x)[(((size_t)n * H + hh) * W + ww) * C8 + c8], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) if (v[j] > best[j]) { best[j] = v[j]; win[j] = (uint32_t)(r * 3 + s); }
      }
    }
    reinterpret_cast<uint4*>(y)[i] = pack8(best);
    uint2 o;
    o.x = win[0] | (win[1] << 8) | (win[2] << 16) | (win[3] << 24);
    o.y = win[4] | (win[5] << 8) | (win[6] << 16) | (win[7] << 24);
    reinterpret_cast<uint2*>(idx)[i] = o;
  }
}

__global__ void __launch_bounds__(256)
maxpool16_bwd_kernel(const uint8_t* __restrict__ idx, const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N, int H,
                     int W, int C, int Ho, int Wo) {
  const int C8 = C / 8;
  const size_t total = (size_t)N * H * W * C8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    size_t t = i / C8;
    const int iw = (int)(t % W); t /= W;
    const int ih = (int)(t % H);
    const int n = (int)(t / H);
    float out[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int p = ih / 2; p <= (ih + 1) / 2; ++p) {               // pooled rows whose window [2p-1, 2p+1] contains ih
      if (p >= Ho) continue;
      for (int q = iw / 2; q <= (iw + 1) / 2; ++q) {
        if (q >= Wo) continue;
        const uint32_t me = (uint32_t)((ih - (2 * p - 1)) * 3 + (iw - (2 * q - 1)));
        const size_t o = (((size_t)n * Ho + p) * Wo + q) * C8 + c8;
        const uint2 w = reinterpret_cast<const uint2*>(idx)[o];
        float gr[8];
        unpack8(reinterpret_cast<const uint4*>(dy)[o], gr);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (((w.x >> (8 * j)) & 0xffu) == me) out[j] += gr[j];
          if (((w.y >> (8 * j)) & 0xffu) == me) out[4 + j] += gr[4 + j];
        }
      }
    }
    reinterpret_cast<uint4*>(dx)[i] = pack8(out);
  }
}

__global__ void avgpool16_bwd_kernel(const float* __restrict__ demb, __nv_bfloat16* __restrict__ dx, int HW, int C, size_t total8) {
  const int C8 = C / 8;
  const float inv = 1.f / (float)HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    const size_t n = i / ((size_t)HW * C8);
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = demb[n * C + c + j] * inv;
    reinterpret_cast<uint4*>(dx)[i] = pack8(f);
  }
}

// f32 NHWC3 image -> bf16 patch matrix [N*Ho*Wo, 192] of the 7x7 stride-2 pad-3 stem: column (r*7 + s)*3 + ch, columns 147..191 zero.
// One thread writes 8 columns (16 bytes).
// Block = 8 output pixels of one row x 24 column groups (192 threads... rounded to 256: x = column group, y = pixel); a thread's
// 8 columns map to fixed (r, s, ch) triples, decoded ONCE per thread (the first version divided by 3 and 7 per element).
__global__ void __launch_bounds__(256)
stem_im2col3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int H, int W, int Ho, int Wo) {
  const int c8 = threadIdx.x;                        // 0..23 (threads 24..31 of each row idle)
  if (c8 >= 24) return;
  int dr[8], ds[8];                                  // tap row, (tap column * 3 + channel); -1 = zero padding column
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = c8 * 8 + j;
    if (col < 147) { const int rs = col / 3; dr[j] = rs / 7; ds[j] = (rs - dr[j] * 7) * 3 + (col - rs * 3); }
    else { dr[j] = -1; ds[j] = 0; }
  }
  const size_t pixels = (size_t)N * Ho * Wo;
  for (size_t pix = (size_t)blockIdx.x * blockDim.y + threadIdx.y; pix < pixels; pix += (size_t)gridDim.x * blockDim.y) {
    const int q = (int)(pix % Wo);
    const size_t t = pix / Wo;
    const int p = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const float* img = x + (size_t)n * H * W * 3;
    const int ih0 = 2 * p - 3, iw0 = (2 * q - 3) * 3;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ih = ih0 + dr[j], iw3 = iw0 + ds[j];
      f[j] = (dr[j] >= 0 && ih >= 0 && ih < H && iw3 >= 0 && iw3 < W * 3) ? __ldg(img + (size_t)ih * W * 3 + iw3) : 0.f;
    }
    reinterpret_cast<uint4*>(out)[pix * 24 + c8] = pack8(f);
  }
}

// zero-dilated copy for stride 2: out[n, y, x, :] = (y, x both even and y/2 < Ho, x/2 < Wo) ? src[n, y/2, x/2, :] : 0
__global__ void __launch_bounds__(256)
dilate2_16_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ out, int N, int Ho, int Wo, int K, int Hd, int Wd) {
  const int K8 = K / 8;
  const size_t total = (size_t)N * Hd * Wd * K8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k8 = (int)(i % K8);
    size_t t = i / K8;
    const int xx = (int)(t % Wd); t /= Wd;
    const int yy = (int)(t % Hd);
    const int n = (int)(t / Hd);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (!((yy | xx) & 1) && (yy >> 1) < Ho && (xx >> 1) < Wo)
      v = reinterpret_cast<const uint4*>(src)[(((size_t)n * Ho + (yy >> 1)) * Wo + (xx >> 1)) * K8 + k8];
    reinterpret_cast<uint4*>(out)[i] = v;
  }
}

// dx[n, 2p, 2q, :] += t[n, p, q, :]  (data gradient of a stride-2 1x1 convolution: only the even pixels were read)
__global__ void __launch_bounds__(256)
scatter_add2_16_kernel(const __nv_bfloat16* __restrict__ t, __nv_bfloat16* __restrict__ dx, int N, int Ho, int Wo, int C, int H, int W) {
  const int C8 = C / 8;
  const size_t total = (size_t)N * Ho * Wo * C8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    size_t r = i / C8;
    const int q = (int)(r % Wo); r /= Wo;
    const int p = (int)(r % Ho);
    const int n = (int)(r / Ho);
    uint4* dst = reinterpret_cast<uint4*>(dx) + (((size_t)n * H + 2 * p) * W + 2 * q) * C8 + c8;
    float a[8], b[8];
    unpack8(reinterpret_cast<const uint4*>(t)[i], a);
    unpack8(*dst, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    *dst = pack8(a);
  }
}

// w f32 [K][C][R][S] (torchvision) -> wk [K][R][S][C] bf16 and wrot [C][R][S][K] bf16 with wrot[c][r][s][k] = w[k][c][R-1-r][S-1-s]
__global__ void pack_conv_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wk, __nv_bfloat16* __restrict__ wrot, int K,
                                         int C, int R, int S) {
  const size_t total = (size_t)K * C * R * S;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    // i indexes the [K][R][S][C] output (coalesced writes of wk)
    const int c = (int)(i % C);
    size_t t = i / C;
    const int s = (int)(t % S); t /= S;
    const int r = (int)(t % R);
    const int k = (int)(t / R);
    const float v = w[(((size_t)k * C + c) * R + r) * S + s];
    const __nv_bfloat16 b = __float2bfloat16(v);
    wk[i] = b;
    if (wrot) wrot[(((size_t)c * R + (R - 1 - r)) * S + (S - 1 - s)) * K + k] = b;
  }
}

static inline int ew_blocks16(size_t n) { return (int)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, (size_t)num_sms() * 16)); }

}  // namespace pdf

using namespace pdf;
typedef __nv_bfloat16 bf16_t;

extern "C" int pdf_bn_train_forward_bf16(int n_groups, const int32_t* d_goff, int max_group_rows, int C, const void* d_x, const float* d_gamma,
                                         const float* d_beta, float eps, const void* d_residual, int relu, void* d_y, uint8_t* d_relu_mask,
                                         float* d_mean, float* d_invstd, float* d_var_unbiased, double* d_scratch, pdf_stream_t stream) {
  PDF_REQUIRE(n_groups > 0 && d_goff && max_group_rows > 0 && C > 0 && C % 8 == 0 && d_x && d_gamma && d_beta && d_y && d_mean && d_invstd &&
              d_var_unbiased && d_scratch && (!relu || d_relu_mask),
              "pdf_bn_train_forward_bf16: bad arguments (C %% 8 == 0, scratch of 2*groups*C doubles, a mask buffer when relu)");
  cudaStream_t s = as_stream(stream);
  PDF_CHECK_CUDA(cudaMemsetAsync(d_scratch, 0, (size_t)2 * n_groups * C * sizeof(double), s));
  bn16_stats_kernel<<<dim3(ceil_div(C, 64), n_groups, ceil_div(max_group_rows, kBn16Rows)), 256, 0, s>>>((const bf16_t*)d_x, d_goff, C, d_scratch);
  PDF_CHECK_LAUNCH();
  bn16_stats_finalize_kernel<<<ceil_div(n_groups * C, 256), 256, 0, s>>>(d_scratch, d_goff, n_groups, C, eps, d_mean, d_invstd, d_var_unbiased);
  PDF_CHECK_LAUNCH();
  bn16_apply_kernel<<<dim3(std::max(1, 8 * num_sms() / n_groups), n_groups), 256, 0, s>>>((const bf16_t*)d_x, d_goff, C, d_mean, d_invstd, d_gamma,
                                                                                         d_beta, (const bf16_t*)d_residual, relu, (bf16_t*)d_y, d_relu_mask);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

extern "C" int pdf_bn_train_backward_bf16(int n_groups, const int32_t* d_goff, int max_group_rows, int C, const void* d_dy,
                                          const uint8_t* d_relu_mask, const void* d_x, const float* d_gamma, const float* d_mean, const float* d_invstd, int relu,
                                          double* d_scratch, void* d_dx, void* d_dres, int dres_accumulate, float* d_dgamma, float* d_dbeta,
                                          pdf_stream_t stream) {
  PDF_REQUIRE(n_groups > 0 && d_goff && max_group_rows > 0 && C > 0 && C % 8 == 0 && d_dy && (!relu || d_relu_mask) && d_x && d_gamma && d_mean &&
              d_invstd && d_scratch && d_dx && d_dgamma && d_dbeta,
              "pdf_bn_train_backward_bf16: bad arguments (C %% 8 == 0, scratch of 3*groups*C doubles, the forward's mask when relu)");
  cudaStream_t s = as_stream(stream);
  double* acc = d_scratch;
  float* sums = reinterpret_cast<float*>(d_scratch + (size_t)2 * n_groups * C);
  PDF_CHECK_CUDA(cudaMemsetAsync(acc, 0, (size_t)2 * n_groups * C * sizeof(double), s));
  bn16_bwd_reduce_kernel<<<dim3(ceil_div(C, 64), n_groups, ceil_div(max_group_rows, kBn16Rows)), 256, 0, s>>>(
      (const bf16_t*)d_dy, d_relu_mask, (const bf16_t*)d_x, d_goff, C, d_mean, d_invstd, relu, acc);
  PDF_CHECK_LAUNCH();
  bn16_bwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, s>>>(acc, n_groups, C, sums, d_dgamma, d_dbeta);
  PDF_CHECK_LAUNCH();
  bn16_bwd_apply_kernel<<<dim3(std::max(1, 8 * num_sms() / n_groups), n_groups), 256, 0, s>>>(
      (const bf16_t*)d_dy, d_relu_mask, (const bf16_t*)d_x, d_goff, C, d_mean, d_invstd, d_gamma, relu, sums,
      sums + (size_t)n_groups * C, (bf16_t*)d_dx, (bf16_t*)d_dres, dres_accumulate);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

extern "C" int pdf_maxpool_train_forward_bf16(int n, int h, int w, int c, const void* d_x, void* d_y, uint8_t* d_idx, pdf_stream_t stream) {
  PDF_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 && d_x && d_y && d_idx, "pdf_maxpool_train_forward_bf16: bad arguments");
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  maxpool16_fwd_idx_kernel<<<ew_blocks16((size_t)n * ho * wo * c / 8), 256, 0, as_stream(stream)>>>((const bf16_t*)d_x, (bf16_t*)d_y, d_idx, n, h, w, c,
                                                                                                  ho, wo);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

extern "C" int pdf_maxpool_backward_bf16(int n, int h, int w, int c, const uint8_t* d_idx, const void* d_dy, void* d_dx, pdf_stream_t stream) {
  PDF_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 && d_idx && d_dy && d_dx, "pdf_maxpool_backward_bf16: bad arguments");
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  maxpool16_bwd_kernel<<<ew_blocks16((size_t)n * h * w * c / 8), 256, 0, as_stream(stream)>>>(d_idx, (const bf16_t*)d_dy, (bf16_t*)d_dx, n, h, w, c, ho, wo);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

extern "C" int pdf_avgpool_backward_bf16(int n, int hw, int c, const float* d_demb, void* d_dx, pdf_stream_t stream) {
  PDF_REQUIRE(n > 0 && hw > 0 && c > 0 && c % 8 == 0 && d_demb && d_dx, "pdf_avgpool_backward_bf16: bad arguments");
  const size_t total8 = (size_t)n * hw * c / 8;
  avgpool16_bwd_kernel<<<ew_blocks16(total8), 256, 0, as_stream(stream)>>>(d_demb, (bf16_t*)d_dx, hw, c, total8);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

extern "C" int pdf_stem_im2col3_bf16(int n, int h, int w, const float* d_x, void* d_out, pdf_stream_t stream) {
  PDF_REQUIRE(n > 0 && h > 0 && w > 0 && d_x && d_out, "pdf_stem_im2col3_bf16: bad arguments");
  const int ho = (h + 6 - 7) / 2 + 1, wo = (w + 6 - 7) / 2 + 1;
  stem_im2col3_kernel<<<ew_blocks16((size_t)n * ho * wo * 32), dim3(32, 8), 0, as_stream(stream)>>>(d_x, (bf16_t*)d_out, n, h, w, ho, wo);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

extern "C" int pdf_dilate2_bf16(int n, int ho, int wo, int k, int hd, int wd, const void* d_src, void* d_out, pdf_stream_t stream) {
  PDF_REQUIRE(n > 0 && ho > 0 && wo > 0 && k > 0 && k % 8 == 0 && hd >= 2 * (ho - 1) + 1 && wd >= 2 * (wo - 1) + 1 && d_src && d_out,
              "pdf_dilate2_bf16: bad arguments");
  dilate2_16_kernel<<<ew_blocks16((size_t)n * hd * wd * k / 8), 256, 0, as_stream(stream)>>>((const bf16_t*)d_src, (bf16_t*)d_out, n, ho, wo, k, hd, wd);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

extern "C" int pdf_scatter_add2_bf16(int n, int ho, int wo, int c, int h, int w, const void* d_t, void* d_dx, pdf_stream_t stream) {
  PDF_REQUIRE(n > 0 && ho > 0 && wo > 0 && c > 0 && c % 8 == 0 && h >= 2 * (ho - 1) + 1 && w >= 2 * (wo - 1) + 1 && d_t && d_dx,
              "pdf_scatter_add2_bf16: bad arguments");
  scatter_add2_16_kernel<<<ew_blocks16((size_t)n * ho * wo * c / 8), 256, 0, as_stream(stream)>>>((const bf16_t*)d_t, (bf16_t*)d_dx, n, ho, wo, c, h, w);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

extern "C" int pdf_pack_conv_weights(int k, int c, int r, int s, const float* d_w, void* d_wk, void* d_wrot, pdf_stream_t stream) {
  PDF_REQUIRE(k > 0 && c > 0 && r > 0 && s > 0 && d_w && d_wk, "pdf_pack_conv_weights: bad arguments");
  pack_conv_weights_kernel<<<ew_blocks16((size_t)k * c * r * s), 256, 0, as_stream(stream)>>>(d_w, (bf16_t*)d_wk, (bf16_t*)d_wrot, k, c, r, s);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}
