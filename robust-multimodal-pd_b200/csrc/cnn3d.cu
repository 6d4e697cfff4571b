// f4 ("cnn3d" feature mode): the pieces of the 3-D convolutional auto-encoder that are not already a 2-D kernel.
//
// Reference: scripts/build_cnn3d_embeddings.py:28-88 (load_volume's per-volume standardisation, Simple3DAE: three Conv3d 3x3x3 +
// ReLU + MaxPool3d(2), two Linear layers, three ConvTranspose3d 2x2x2 stride 2) and its training loop :132-142 (MSE, Adam).
// Layout here is channels-last [N, D, H, W, C] f32.  The 3x3x3 convolutions -- forward, data gradient and weight gradient -- are
// DEPTH-DECOMPOSED on the host (pd_fusion_b200/cnn3d.py): out[:, d] = sum_kd conv2d(x[:, d + kd - 1], W[kd]), i.e. three passes of
// the FP32 implicit-GEMM 2-D kernels (conv_fp32.cu, train.cu) over depth-shifted slice ranges, accumulating in place.  A stride-2
// 2x2x2 transposed convolution is a GEMM [voxels, Cin] x [Cin, 8*Cout] (pdf_gemm_f32) followed by a pixel shuffle.  What is left
// for this file: max pool 2x2x2 with recorded winner + its backward, the pixel shuffle and its inverse (bias, ReLU and the ReLU
// mask fused), the MSE loss with its gradient, an in-place ReLU, and the standardisation of a volume by the mean / std of its
// positive voxels (statistics from pdf_simple_stats).
#include <algorithm>

#include "common.cuh"

namespace pdf {

static inline int ew3(size_t n) { return (int)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, (size_t)num_sms() * 16)); }

// MaxPool3d(2): the first maximum in (d, h, w) scan order wins (torch's `>` update rule); idx = i*4 + j*2 + k of the winner
__global__ void maxpool3d_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ idx, int N, int D, int H, int W,
                                     int C) {
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)N * Do * Ho * Wo * C;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    size_t r = t / C;
    const int w = (int)(r % Wo); r /= Wo;
    const int h = (int)(r % Ho); r /= Ho;
    const int d = (int)(r % Do);
    const int n = (int)(r / Do);
    float best = -INFINITY;
    int win = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float v = x[((((size_t)n * D + 2 * d + i) * H + 2 * h + j) * W + 2 * w + k) * C + c];
          if (v > best || (i | j | k) == 0) { best = v; win = i * 4 + j * 2 + k; }
        }
    y[t] = best;
    idx[t] = (uint8_t)win;
  }
}

__global__ void maxpool3d_bwd_kernel(const uint8_t* __restrict__ idx, const float* __restrict__ dy, float* __restrict__ dx, int N, int D, int H,
                                     int W, int C) {
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)N * Do * Ho * Wo * C;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    size_t r = t / C;
    const int w = (int)(r % Wo); r /= Wo;
    const int h = (int)(r % Ho); r /= Ho;
    const int d = (int)(r % Do);
    const int n = (int)(r / Do);
    const int win = idx[t];
    const float g = dy[t];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 2; ++k)
          dx[((((size_t)n * D + 2 * d + i) * H + 2 * h + j) * W + 2 * w + k) * C + c] = (win == i * 4 + j * 2 + k) ? g : 0.f;
  }
}

// ConvTranspose3d(k=2, s=2) epilogue: y[n, 2d+i, 2h+j, 2w+k, co] = act(t[v, (i*4+j*2+k)*Cout + co] + bias[co]), v = (n, d, h, w)
__global__ void shuffle2_3d_kernel(const float* __restrict__ t, const float* __restrict__ bias, float* __restrict__ y, int N, int D, int H, int W,
                                   int Cout, int relu) {
  const size_t total = (size_t)N * D * H * W * 8 * Cout;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(e % Cout);
    size_t r = e / Cout;
    const int ijk = (int)(r % 8); r /= 8;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H); r /= H;
    const int d = (int)(r % D);
    const int n = (int)(r / D);
    float v = t[e] + (bias ? __ldg(bias + co) : 0.f);
    if (relu) v = fmaxf(v, 0.f);
    const int i = ijk >> 2, j = (ijk >> 1) & 1, k = ijk & 1;
    y[((((size_t)n * 2 * D + 2 * d + i) * 2 * H + 2 * h + j) * 2 * W + 2 * w + k) * Cout + co] = v;
  }
}

// its backward: dt[v, ijk*Cout + co] = dy[...] * (y[...] > 0 if relu)
__global__ void unshuffle2_3d_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dt, int N, int D, int H, int W,
                                     int Cout, int relu) {
  const size_t total = (size_t)N * D * H * W * 8 * Cout;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(e % Cout);
    size_t r = e / Cout;
    const int ijk = (int)(r % 8); r /= 8;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H); r /= H;
    const int d = (int)(r % D);
    const int n = (int)(r / D);
    const int i = ijk >> 2, j = (ijk >> 1) & 1, k = ijk & 1;
    const size_t o = ((((size_t)n * 2 * D + 2 * d + i) * 2 * H + 2 * h + j) * 2 * W + 2 * w + k) * Cout + co;
    float g = dy[o];
    if (relu && !(y[o] > 0.f)) g = 0.f;
    dt[e] = g;
  }
}

__global__ void relu_kernel(float* __restrict__ x, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = fmaxf(x[i], 0.f);
}

// nn.MSELoss (mean): loss += sum (p - t)^2 / n ; dp = 2 (p - t) / n
__global__ void __launch_bounds__(256) mse_train_kernel(const float* __restrict__ p, const float* __restrict__ t, size_t n, float* __restrict__ loss,
                                                        float* __restrict__ dp) {
  __shared__ double red[8];
  const double inv = 1.0 / (double)n;
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float d = p[i] - t[i];
    s += (double)d * (double)d;
    if (dp) dp[i] = (float)(2.0 * (double)d * inv);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += red[i];
    atomicAdd(loss, (float)(tot * inv));
  }
}

// load_volume's standardisation: (v - mean) / (std + 1e-6) over the positive voxels' statistics, skipped when there is none.
// stats: pdf_simple_stats rows (0 n | 1 sum | 4 sum (v-mean)^2 | 12 all-voxel fallback flag)
__global__ void standardize_kernel(const float* __restrict__ x, const double* __restrict__ stats, int stride, size_t voxels, float* __restrict__ y) {
  const double* st = stats + (size_t)blockIdx.y * stride;
  const bool skip = st[12] != 0.0;
  const float mean = (float)(st[1] / st[0]);
  const float sd = __fadd_rn((float)sqrt(st[4] / st[0]), 1e-6f);
  const float* xi = x + (size_t)blockIdx.y * voxels;
  float* yi = y + (size_t)blockIdx.y * voxels;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < voxels; i += (size_t)gridDim.x * blockDim.x)
    yi[i] = skip ? xi[i] : __fdiv_rn(__fsub_rn(xi[i], mean), sd);
}

}  // namespace pdf

using namespace pdf;

extern "C" int pdf_maxpool3d_forward(int n, int d, int h, int w, int c, const float* d_x, float* d_y, uint8_t* d_idx, pdf_stream_t stream) {
  PDF_REQUIRE(n > 0 && d > 1 && h > 1 && w > 1 && c > 0 && d_x && d_y && d_idx, "pdf_maxpool3d_forward: bad arguments");
  maxpool3d_fwd_kernel<<<ew3((size_t)n * (d / 2) * (h / 2) * (w / 2) * c), 256, 0, as_stream(stream)>>>(d_x, d_y, d_idx, n, d, h, w, c);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

extern "C" int pdf_maxpool3d_backward(int n, int d, int h, int w, int c, const uint8_t* d_idx, const float* d_dy, float* d_dx, pdf_stream_t stream) {
  PDF_REQUIRE(n > 0 && d > 1 && h > 1 && w > 1 && d % 2 == 0 && h % 2 == 0 && w % 2 == 0 && c > 0 && d_idx && d_dy && d_dx,
              "pdf_maxpool3d_backward: bad arguments (even extents)");
  maxpool3d_bwd_kernel<<<ew3((size_t)n * (d / 2) * (h / 2) * (w / 2) * c), 256, 0, as_stream(stream)>>>(d_idx, d_dy, d_dx, n, d, h, w, c);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

extern "C" int pdf_shuffle2_3d(int n, int d, int h, int w, int cout, const float* d_t, const float* d_bias, int relu, float* d_y, pdf_stream_t stream) {
  PDF_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0 && cout > 0 && d_t && d_y, "pdf_shuffle2_3d: bad arguments");
  shuffle2_3d_kernel<<<ew3((size_t)n * d * h * w * 8 * cout), 256, 0, as_stream(stream)>>>(d_t, d_bias, d_y, n, d, h, w, cout, relu);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

extern "C" int pdf_unshuffle2_3d(int n, int d, int h, int w, int cout, const float* d_dy, const float* d_y, int relu, float* d_dt, pdf_stream_t stream) {
  PDF_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0 && cout > 0 && d_dy && d_dt && (!relu || d_y), "pdf_unshuffle2_3d: bad arguments");
  unshuffle2_3d_kernel<<<ew3((size_t)n * d * h * w * 8 * cout), 256, 0, as_stream(stream)>>>(d_dy, d_y, d_dt, n, d, h, w, cout, relu);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

extern "C" int pdf_relu_f32(float* d_x, size_t n, pdf_stream_t stream) {
  PDF_REQUIRE(d_x && n > 0, "pdf_relu_f32: bad arguments");
  relu_kernel<<<ew3(n), 256, 0, as_stream(stream)>>>(d_x, n);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

extern "C" int pdf_mse_train(size_t n, const float* d_pred, const float* d_target, float* d_loss, float* d_dpred, pdf_stream_t stream) {
  PDF_REQUIRE(n > 0 && d_pred && d_target && d_loss, "pdf_mse_train: bad arguments");
  mse_train_kernel<<<ew3(n), 256, 0, as_stream(stream)>>>(d_pred, d_target, n, d_loss, d_dpred);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

extern "C" int pdf_standardize_volume(int batch, size_t voxels, const float* d_x, const double* d_stats, int stats_stride, float* d_y,
                                      pdf_stream_t stream) {
  PDF_REQUIRE(batch > 0 && voxels > 0 && d_x && d_stats && stats_stride > 12 && d_y, "pdf_standardize_volume: bad arguments");
  standardize_kernel<<<dim3(std::max(1, ew3(voxels) / 4), batch), 256, 0, as_stream(stream)>>>(d_x, d_stats, stats_stride, voxels, d_y);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}
