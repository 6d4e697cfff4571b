// K2 stem, fused: 7x7 stride-2 pad-3 convolution (one folded input channel) + bias + ReLU + 3x3 stride-2 pad-1
// max-pool in ONE kernel.  Per CTA: a 7x7 tile of pooled pixels <- 15x15 conv pixels <- 35x35 input patch.
//
//   1. input patch (bf16, zero padded) -> shared memory
//   2. the 225 x 64 im2col matrix (K = r*8+s: 49 taps + 15 zero columns) is BUILT in shared memory in the K-major
//      SWIZZLE_128B layout the tensor core reads (never touches HBM); weights [64 x 64] likewise
//   3. 2 x 4 tcgen05.mma (M=128, N=64, K=16) -> two f32 accumulators in TMEM
//   4. epilogue: tcgen05.ld -> +bias, ReLU -> bf16 conv tile in shared memory (overlays the A matrix)
//   5. 3x3/2 max over the conv tile -> [n, P, P, 64] bf16 NHWC
//
// HBM traffic per image: S*S*2 bytes in (+halo re-reads from L2), P*P*64*2 bytes out -- the unfused sequence
// (stem_im2col_kernel -> conv_tc_kernel -> maxpool_kernel) moves ~6.8 MB per 224x224 image instead of 0.5 MB.
// Replaces conv1/bn1/relu/maxpool of torchvision's ResNet (`model(batch)`, data/openneuro_features.py:260).
#include "tc_common.cuh"
#include "ops.cuh"

namespace pdf {

constexpr int kTP = 7;             // pooled tile side
constexpr int kCT = 2 * kTP + 1;   // conv tile side (15)
constexpr int kIT = 2 * kCT + 5;   // input patch side (35)
constexpr int kITP = kIT + 1;      // padded row pitch
constexpr int kConvPix = kCT * kCT;  // 225 valid rows of the 256-row A matrix

struct StemParams {
  const __nv_bfloat16* in;    // [n, S, S]
  const __nv_bfloat16* w;     // [64, 64]  (cout, r*8+s; s = 7 and r = 7 columns are zero)
  const float* bias;          // [64]
  __nv_bfloat16* out;         // [n, P, P, 64]
  int S, H1, P, tiles;
};

__global__ void __launch_bounds__(256)
stem_fused_kernel(const __grid_constant__ CUtensorMap tmap_w, const StemParams p) {
  // [ A: 256 rows x 128 B | W: 64 rows x 128 B | input patch | barrier, tmem slot ]
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  uint8_t* sA = smem;
  __nv_bfloat16* sIn = reinterpret_cast<__nv_bfloat16*>(smem + 256 * 128 + 64 * 128);
  const uint32_t bar = base + 256 * 128 + 64 * 128 + ((kIT * kITP * 2 + 15) & ~15);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (bar - base) + 32);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.y;
  const int ty = blockIdx.x / p.tiles, tx = blockIdx.x - ty * p.tiles;
  const int tp0 = ty * kTP, tq0 = tx * kTP;          // pooled-tile origin
  const int cy0 = 2 * tp0 - 1, cx0 = 2 * tq0 - 1;    // conv-tile origin (may be -1)
  const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;    // input-patch origin

  const uint32_t bar_w = bar + 16;
  if (warp == 0 && elect_one()) {
    mbar_init(bar, 1);
    mbar_init(bar_w, 1);
    fence_barrier_init();
    // weights [64 x 64] bf16 -> 128B-swizzled K-major rows, written by the TMA unit (one instruction per CTA)
    mbar_expect_tx(bar_w, 64 * 128);
    tma_load_2d(base + 256 * 128, &tmap_w, bar_w, 0, 0);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 128);

  // 1. input patch.  Patch column 0 is global x = ix0 (odd), so columns 1.. pair up into 4-byte aligned global words:
  //    one 2-byte load for column 0 and 17 word loads per row instead of 35 scalar loads.
  const __nv_bfloat16* img = p.in + (size_t)n * p.S * p.S;
  const unsigned short* img16 = reinterpret_cast<const unsigned short*>(img);
  unsigned short* sIn16 = reinterpret_cast<unsigned short*>(sIn);
  const bool even = (p.S & 1) == 0;
  for (int i = tid; i < kIT * 18; i += 256) {
    const int y = i / 18, k = i - y * 18;
    const int gy = iy0 + y;
    const bool rowin = gy >= 0 && gy < p.S;
    if (k == 0) {
      sIn16[y * kITP] = (rowin && ix0 >= 0 && ix0 < p.S) ? img16[(size_t)gy * p.S + ix0] : (unsigned short)0;
    } else {
      const int x = 2 * k - 1, gx = ix0 + x;              // gx is even
      unsigned short lo16 = 0, hi16 = 0;
      if (rowin) {
        if (even && gx >= 0 && gx + 1 < p.S) {
          const uint32_t w = *reinterpret_cast<const uint32_t*>(img16 + (size_t)gy * p.S + gx);
          lo16 = (unsigned short)(w & 0xffffu);
          hi16 = (unsigned short)(w >> 16);
        } else {
          if (gx >= 0 && gx < p.S) lo16 = img16[(size_t)gy * p.S + gx];
          if (gx + 1 >= 0 && gx + 1 < p.S) hi16 = img16[(size_t)gy * p.S + gx + 1];
        }
      }
      sIn16[y * kITP + x] = lo16;
      if (x + 1 < kIT) sIn16[y * kITP + x + 1] = hi16;
    }
  }
  __syncthreads();
  // 2b. im2col rows.  K index = r*8 + s (s = 7 is a zero column, r = 7 a zero chunk): chunk c of row m is the 7 taps of
  //     filter row c, i.e. 8 consecutive bf16 of the patch starting at an even column -> four aligned 32-bit loads.
  {
    const int c = tid & 7;                      // chunk (= filter row) is fixed per thread; rows advance by 32 per step
    int m = tid >> 3;
    int cy = m / kCT, cx = m - cy * kCT;
    const uint32_t cmask = (c < 7) ? 0xffffffffu : 0u;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (m < kConvPix) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(sIn + (2 * cy + (c < 7 ? c : 0)) * kITP + 2 * cx);
        v = make_uint4(src[0] & cmask, src[1] & cmask, src[2] & cmask, src[3] & 0x0000ffffu & cmask);
      }
      *reinterpret_cast<uint4*>(sA + m * 128 + ((c ^ (m & 7)) << 4)) = v;
      m += 32; cy += 2; cx += 2;
      if (cx >= kCT) { cx -= kCT; ++cy; }
    }
  }
  fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // 3. MMAs
  if (warp == 0 && elect_one()) {
    constexpr uint32_t idesc = make_idesc(64);
    const uint32_t a0 = base, w0 = base + 256 * 128;
    mbar_wait(bar_w, 0);
    tc_fence_after();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16(tmem_base + mt * 64, make_smem_desc(a0 + mt * kABytes + k * 32), make_smem_desc(w0 + k * 32), idesc, k != 0 ? 1u : 0u);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();

  // 4. epilogue: conv tile (bias, ReLU, bf16) -> shared memory over the A matrix; out-of-image conv pixels -> 0
  {
    const int quad = warp & 3, mt = warp >> 2;
    const int m = mt * 128 + quad * 32 + lane;
    const int cy = m / kCT, cx = m - cy * kCT;
    const int gy = cy0 + cy, gx = cx0 + cx;
    const bool inside = m < kConvPix && gy >= 0 && gy < p.H1 && gx >= 0 && gx < p.H1;
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * 64 + c0), v);
      if (m < kConvPix) {
        const uint32_t keep = inside ? 0xffffffffu : 0u;       // out-of-image conv pixels must not win the max-pool
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + i * 8));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + i * 8 + 4));
          uint4 o;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
          h[0] = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[i * 8 + 0]) + b0.x, 0.f), fmaxf(__uint_as_float(v[i * 8 + 1]) + b0.y, 0.f));
          h[1] = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[i * 8 + 2]) + b0.z, 0.f), fmaxf(__uint_as_float(v[i * 8 + 3]) + b0.w, 0.f));
          h[2] = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[i * 8 + 4]) + b1.x, 0.f), fmaxf(__uint_as_float(v[i * 8 + 5]) + b1.y, 0.f));
          h[3] = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[i * 8 + 6]) + b1.z, 0.f), fmaxf(__uint_as_float(v[i * 8 + 7]) + b1.w, 0.f));
          o.x &= keep; o.y &= keep; o.z &= keep; o.w &= keep;
          const int chunk = (c0 >> 3) + i;
          *reinterpret_cast<uint4*>(sA + m * 128 + ((chunk ^ (m & 7)) << 4)) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);

  // 5. 3x3 stride-2 max-pool over the conv tile (values are >= 0 after ReLU, so padding/out-of-image = 0 is neutral)
  for (int i = tid; i < kTP * kTP * 8; i += 256) {
    const int c = i & 7, px = i >> 3;
    const int pp = px / kTP, pq = px - pp * kTP;
    const int gp = tp0 + pp, gq = tq0 + pq;
    if (gp >= p.P || gq >= p.P) continue;
    __nv_bfloat162 mx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) mx[j] = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int m = (2 * pp + dy) * kCT + 2 * pq + dx;
        const uint4 v = *reinterpret_cast<const uint4*>(sA + m * 128 + ((c ^ (m & 7)) << 4));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) mx[j] = __hmax2(mx[j], h[j]);
      }
    uint4 o;
    __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) ho[j] = mx[j];
    *reinterpret_cast<uint4*>(p.out + (((size_t)n * p.P + gp) * p.P + gq) * 64 + c * 8) = o;
  }
}

constexpr int kStemSmem = 256 * 128 + 64 * 128 + ((kIT * kITP * 2 + 15) & ~15) + 48 + 1024;

int launch_stem_fused(const pdf_op& op, const TensorMapBlob& tmap_w, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    PDF_CHECK_CUDA(cudaFuncSetAttribute(stem_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemSmem));
    configured = true;
  }
  StemParams p;
  p.in = reinterpret_cast<const __nv_bfloat16*>(op.d_in);
  p.w = reinterpret_cast<const __nv_bfloat16*>(op.d_weight);
  p.bias = op.d_bias;
  p.out = reinterpret_cast<__nv_bfloat16*>(op.d_out);
  p.S = op.h;
  p.H1 = (op.h + 6 - 7) / 2 + 1;
  p.P = op.ho;
  p.tiles = ceil_div(p.P, kTP);
  dim3 grid(p.tiles * p.tiles, op.n);
  stem_fused_kernel<<<grid, 256, kStemSmem, s>>>(*reinterpret_cast<const CUtensorMap*>(&tmap_w), p);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

}  // namespace pdf
