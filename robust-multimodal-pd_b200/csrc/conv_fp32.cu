// FP32 CUDA-core implicit-GEMM convolution (the 1e-5 parity path of K2) and the pooling / im2col helpers
// shared by both precisions.  Reference: torchvision ResNet forward called at
// data/openneuro_features.py:257-262 (F.conv2d + eval BatchNorm + ReLU + residual, max_pool2d, adaptive_avg_pool2d).
#include "common.cuh"
#include "ops.cuh"
#include <algorithm>

namespace pdf {

constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256)
conv_f32_kernel(const float* __restrict__ in, const float* __restrict__ wgt, const float* __restrict__ scale,
                const float* __restrict__ bias, const float* __restrict__ residual, float* __restrict__ out,
                int N, int H, int W, int C, int K, int R, int S, int stride, int pad, int Ho, int Wo, int relu) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int M = N * Ho * Wo;
  const int Kg = R * S * C;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  // A-load assignment: thread -> pixel (tid/4), 4 consecutive kk
  const int am = tid >> 2, ak = (tid & 3) * 4;
  const int m = m0 + am;
  int pn = 0, ih0 = 0, iw0 = 0;
  const bool mvalid = m < M;
  if (mvalid) {
    pn = m / (Ho * Wo);
    const int rem = m - pn * Ho * Wo;
    const int p = rem / Wo, q = rem - p * Wo;
    ih0 = p * stride - pad;
    iw0 = q * stride - pad;
  }
  // B-load assignment: thread -> row kk (tid/16), 4 consecutive columns
  const int bk = tid >> 4, bn = (tid & 15) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int kg0 = 0; kg0 < Kg; kg0 += BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int kg = kg0 + ak + e;
      float v = 0.f;
      if (mvalid && kg < Kg) {
        const int c = kg % C;
        const int rs = kg / C;
        const int s = rs % S, r = rs / S;
        const int ih = ih0 + r, iw = iw0 + s;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = __ldg(in + (((size_t)pn * H + ih) * W + iw) * C + c);
      }
      As[ak + e][am] = v;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int kg = kg0 + bk, col = n0 + bn + e;
      Bs[bk][bn + e] = (kg < Kg && col < K) ? __ldg(wgt + (size_t)kg * K + col) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int mm = m0 + ty * 4 + i;
    if (mm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= K) continue;
      float v = acc[i][j];
      if (scale) v = v * __ldg(scale + col);
      if (bias) v = v + __ldg(bias + col);
      if (residual) v = v + __ldg(residual + (size_t)mm * K + col);
      if (relu) v = fmaxf(v, 0.f);
      out[(size_t)mm * K + col] = v;
    }
  }
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// 3x3 stride-2 pad-1 max pooling, NHWC. Padding never wins (max over valid taps only), as F.max_pool2d.
template <typename T>
__global__ void maxpool_kernel(const T* __restrict__ in, T* __restrict__ out, int N, int H, int W, int C, int Ho, int Wo) {
  const size_t total = (size_t)N * Ho * Wo * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    size_t t = i / C;
    const int q = (int)(t % Wo); t /= Wo;
    const int p = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float m = -INFINITY;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ih = p * 2 - 1 + r;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int iw = q * 2 - 1 + s;
        if (iw < 0 || iw >= W) continue;
        m = fmaxf(m, to_f<T>(in[(((size_t)n * H + ih) * W + iw) * C + c]));
      }
    }
    out[i] = from_f<T>(m);
  }
}

// global average pool: [N, HW, C] -> [N, C] f32
template <typename T>
__global__ void avgpool_kernel(const T* __restrict__ in, float* __restrict__ out, int HW, int C) {
  const int n = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const T* p = in + (size_t)n * HW * C + c;
  float s = 0.f;
  for (int i = 0; i < HW; ++i) s += to_f<T>(p[(size_t)i * C]);
  out[(size_t)n * C + c] = s / (float)HW;
}

// 7x7 stride-2 pad-3 patch matrix of a one-channel bf16 image: row = output pixel, col = r*8+s (s=7 and r>=7 columns are zero)
__global__ void stem_im2col_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int N, int H, int W,
                                   int Ho, int Wo, int kpad) {
  const int chunks = kpad / 8;
  const size_t total = (size_t)N * Ho * Wo * chunks;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % chunks);
    size_t t = i / chunks;
    const int q = (int)(t % Wo); t /= Wo;
    const int p = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const __nv_bfloat16* img = in + (size_t)n * H * W;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = ch * 8 + e;
      __nv_bfloat16 x = __float2bfloat16(0.f);
      const int r = k >> 3, s = k & 7;          // K index = r*8 + s (s = 7 / r = 7 are zero columns), as stem_tc.cu
      if (r < 7 && s < 7) {
        const int ih = p * 2 - 3 + r, iw = q * 2 - 3 + s;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) x = img[(size_t)ih * W + iw];
      }
      v[e] = x;
    }
    *reinterpret_cast<uint4*>(out + i * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

__global__ void slice_mean_kernel(const float* __restrict__ emb, const int32_t* __restrict__ nvalid, float* __restrict__ out,
                                  int L, int D) {
  const int b = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const int n = nvalid ? min(max(nvalid[b], 1), L) : L;
  float s = 0.f;
  for (int l = 0; l < n; ++l) s += emb[((size_t)b * L + l) * D + d];
  out[(size_t)b * D + d] = s / (float)n;
}

int launch_conv_f32(const pdf_op& op, cudaStream_t s) {
  const int M = op.n * op.ho * op.wo;
  dim3 grid(ceil_div(M, BM), ceil_div(op.k, BN));
  conv_f32_kernel<<<grid, 256, 0, s>>>((const float*)op.d_in, (const float*)op.d_weight, op.d_scale, op.d_bias,
                                       (const float*)op.d_residual, (float*)op.d_out, op.n, op.h, op.w, op.c, op.k, op.r, op.s,
                                       op.stride, op.pad, op.ho, op.wo, op.relu);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

int launch_maxpool(const pdf_op& op, cudaStream_t s) {
  const size_t total = (size_t)op.n * op.ho * op.wo * op.c;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)num_sms() * 16);
  if (op.precision == PDF_PREC_F32)
    maxpool_kernel<float><<<blocks, 256, 0, s>>>((const float*)op.d_in, (float*)op.d_out, op.n, op.h, op.w, op.c, op.ho, op.wo);
  else
    maxpool_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>((const __nv_bfloat16*)op.d_in, (__nv_bfloat16*)op.d_out, op.n, op.h,
                                                         op.w, op.c, op.ho, op.wo);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

int launch_avgpool(const pdf_op& op, cudaStream_t s) {
  dim3 grid(ceil_div(op.c, 128), op.n);
  // input dtype: f32 path, or bf16 path whose producer stored f32 (op.out_f32 set on THIS op to say "input is f32")
  if (op.precision == PDF_PREC_F32 || op.out_f32)
    avgpool_kernel<float><<<grid, 128, 0, s>>>((const float*)op.d_in, (float*)op.d_out, op.h * op.w, op.c);
  else
    avgpool_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>((const __nv_bfloat16*)op.d_in, (float*)op.d_out, op.h * op.w, op.c);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

int launch_stem_im2col(const pdf_op& op, cudaStream_t s) {
  const size_t total = (size_t)op.n * op.ho * op.wo * (op.k / 8);
  const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)num_sms() * 16);
  stem_im2col_kernel<<<blocks, 256, 0, s>>>((const __nv_bfloat16*)op.d_in, (__nv_bfloat16*)op.d_out, op.n, op.h, op.w, op.ho,
                                            op.wo, op.k);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

}  // namespace pdf

extern "C" int pdf_slice_mean(int batch, int L, int D, const float* d_emb, const int32_t* d_nvalid, float* d_out,
                              pdf_stream_t stream) {
  PDF_REQUIRE(batch > 0 && L > 0 && D > 0 && d_emb && d_out, "pdf_slice_mean: bad arguments");
  dim3 grid(pdf::ceil_div(D, 128), batch);
  pdf::slice_mean_kernel<<<grid, 128, 0, pdf::as_stream(stream)>>>(d_emb, d_nvalid, d_out, L, D);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}
