// K5 (tensor path): convolution weight gradient on the 5th-gen tensor cores.
//
//   dW[k][r][s][c] += sum_{m = (n,p,q)} dY[m, k] * X[n, p*stride + r - pad, q*stride + s - pad, c]          bf16 x bf16 -> f32 (TMEM)
//
// As a GEMM the reduction runs over PIXELS, and both operands are stored pixel-major (dY is [M, Cout], an NHWC activation is
// [pixels, Cin]): each is the TRANSPOSE of a K-major operand.  tcgen05 takes them as they are -- "MN-major" operands (instruction
// descriptor bits 15/16), whose canonical SWIZZLE_128B layout ((64 MN elements = 128 bytes) x (8 K rows), next 64 MN elements at
// LBO, next 8 K rows at SBO = 1024) is exactly what a TMA tile [64 pixels x 64 channels] lands in shared memory.  So the forward
// pass's tensor maps are reused unchanged: dY through a 2-D map, X through the im2col-mode map (padding / stride / image wrap
// in hardware) with the tap's (s, r) offsets; no transposed copy of anything is ever made.
//
// One CTA = one (128-Cout block, tap group, Cin block, pixel slab): it streams its slab in k-blocks of 64 pixels through a TMA
// ring, accumulates [128 x T*N] in TMEM (T taps of one filter row share the dY tile), and adds its partial sum into dW with
// red.global.add.f32 at the end -- the pixel range is split over as many CTAs as fill the machine.
//   warp 0: TMA producer | warp 1: tcgen05.mma issuer | warps 2..5: epilogue (once, after the slab)
//
// Replaces the weight-gradient half of `loss.backward()` through torchvision's convolutions (models/mil_attention_finetune.py:225).
#include <algorithm>

#include "tc_common.cuh"
#include "ops.cuh"

namespace pdf {

constexpr int kWgPix = 64;                  // pixels (GEMM K) per k-block
constexpr int kWgMaxStages = 6;             // ring depth is a launch parameter: as many stages as fit next to the tile shape

static int g_wg_rowtile = 1;     // pdf_debug_set_wgrad_rowtile: 0 = im2col-mode loads for every filter > 1x1
static int g_wg_waves = 1;       // measured (profiles/r02_wgrad_waves.txt): 1 wave 26.2 ms per fine-tune step, 2: 26.8, 3: 27.3, 4: 27.8

struct WgParams {
  int M, Cout, C, R, S, Ho, Wo, stride, pad;
  int T;            // taps per CTA (consecutive s of one filter row, or 1)
  int N;            // Cin columns per tap in this CTA (64 | 128 | 256)
  int tap_groups, cin_blocks, slabs, slab_kb;   // grid decomposition; slab_kb = k-blocks per slab
  int stages;       // TMA ring depth (2..kWgMaxStages)
  int flat;         // 1x1 stride-1 convolution: X is the plain [M, C] matrix, loaded through a 2-D tiled map (no im2col traversal)
  int px;           // pixels (GEMM K) per k-block: kWgPix, or rows * W in row-tiled mode
  int rows;         // row-tiled mode (> 0): a k-block = `rows` full image rows; X through a 4-D TILED map whose box is (64 ch, W, rows)
  int bpi;          //   k-blocks per image = ceil(H / rows)
  int H, W;
  float* dw;        // [Cout][R][S][C] f32
};

// kind::f16 instruction descriptor with BOTH operands MN-major: D = f32, A = B = bf16
__host__ __device__ constexpr uint32_t make_idesc_mn(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
}
// MN-major SWIZZLE_128B descriptor: LBO = byte distance between 64-element blocks along M/N, SBO = 1024 (8 K rows)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__global__ void __launch_bounds__(192, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const int nsub = p.N / 64;                                   // 64-channel sub-tiles per tap
  const uint32_t sub = (uint32_t)p.px * 128u;                  // one [px x 64 ch] swizzled sub-tile
  const uint32_t stage_bytes = (2u + (uint32_t)(p.T * nsub)) * sub;
  const int kWgStages = p.stages;
  const uint32_t bar_full = base + kWgStages * stage_bytes;
  const uint32_t bar_empty = bar_full + 8 * kWgMaxStages;
  const uint32_t bar_done = bar_empty + 8 * kWgMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kWgStages * stage_bytes + 8 * (2 * kWgMaxStages + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // unit decomposition
  int u = blockIdx.x;
  const int slab = u % p.slabs; u /= p.slabs;
  const int cb = u % p.cin_blocks; u /= p.cin_blocks;
  const int tg = u % p.tap_groups; u /= p.tap_groups;
  const int kb0 = u;                                            // Cout block
  const int taps_total = p.R * p.S;
  const int tap0 = tg * p.T;
  const int n_taps = min(p.T, taps_total - tap0);
  const int total_kb = p.rows ? (p.M / (p.H * p.W)) * p.bpi : (p.M + kWgPix - 1) / kWgPix;
  const int kb_lo = slab * p.slab_kb, kb_hi = min(total_kb, kb_lo + p.slab_kb);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_dy); prefetch_tmap(&tmap_x);
    for (int s = 0; s < kWgStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  const uint32_t tmem_cols = 512;
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (kb_lo < kb_hi) {
    if (warp == 0) {
      if (elect_one()) {
        const int hw = p.Ho * p.Wo;
        for (int kb = kb_lo, g = 0; kb < kb_hi; ++kb, ++g) {
          const uint32_t stage = g % kWgStages;
          mbar_wait(bar_empty + 8 * stage, (((uint32_t)(g / kWgStages)) & 1u) ^ 1u);
          mbar_expect_tx(bar_full + 8 * stage, (2u + (uint32_t)(n_taps * nsub)) * sub);
          const uint32_t sa = base + stage * stage_bytes;
          int m0, n_img, pp, qq;
          if (p.rows) {                                          // `rows` image rows of image n_img starting at row pp
            n_img = kb / p.bpi;
            pp = (kb - n_img * p.bpi) * p.rows; qq = 0;
            m0 = (n_img * p.H + pp) * p.W;
          } else {
            m0 = kb * kWgPix;
            n_img = m0 / hw;
            const int rem = m0 - n_img * hw;
            pp = rem / p.Wo; qq = rem - pp * p.Wo;
          }
          if (p.rows) {      // dY through the same kind of 4-D box: output rows past the image are ZERO (an in-bounds X row would
                             // otherwise meet the next image's dY through the taps above the centre)
            tma_load_4d(sa, &tmap_dy, bar_full + 8 * stage, kb0 * 128, 0, pp, n_img);
            tma_load_4d(sa + sub, &tmap_dy, bar_full + 8 * stage, kb0 * 128 + 64, 0, pp, n_img);
          } else {
            tma_load_2d(sa, &tmap_dy, bar_full + 8 * stage, kb0 * 128, m0);               // [px x 64 couts] x 2 (columns past Cout: zero fill)
            tma_load_2d(sa + sub, &tmap_dy, bar_full + 8 * stage, kb0 * 128 + 64, m0);
          }
          const int w0 = qq * p.stride - p.pad, h0 = pp * p.stride - p.pad;
          for (int t = 0; t < n_taps; ++t) {
            const int tap = tap0 + t;
            const int r = tap / p.S, s = tap - r * p.S;
            for (int j = 0; j < nsub; ++j) {
              const uint32_t dst = sa + 2 * sub + (uint32_t)(t * nsub + j) * sub;
              // row-tiled: the box (64 ch, W, rows) starting at column s - pad, row pp + r - pad IS the tap-shifted pixel block, the
              // zero padding comes from the out-of-bounds fill
              if (p.rows) tma_load_4d(dst, &tmap_x, bar_full + 8 * stage, cb * p.N + j * 64, s - p.pad, pp + r - p.pad, n_img);
              else if (p.flat) tma_load_2d(dst, &tmap_x, bar_full + 8 * stage, cb * p.N + j * 64, m0);
              else tma_load_im2col_4d(dst, &tmap_x, bar_full + 8 * stage, cb * p.N + j * 64, w0, h0, n_img, (uint16_t)s, (uint16_t)r);
            }
          }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        const uint32_t idesc = make_idesc_mn(p.N);
        for (int kb = kb_lo, g = 0; kb < kb_hi; ++kb, ++g) {
          const uint32_t stage = g % kWgStages;
          mbar_wait(bar_full + 8 * stage, ((uint32_t)(g / kWgStages)) & 1u);
          tc_fence_after();
          const uint32_t sa = base + stage * stage_bytes;
          const int ksteps = p.px / 16;
#pragma unroll 4
          for (int k = 0; k < ksteps; ++k) {                      // 16 pixels per instruction = 2 swizzle atoms = 2048 bytes of rows
            const uint64_t da = make_desc_mn(sa + k * 2048, sub);
            for (int t = 0; t < n_taps; ++t) {
              const uint64_t db = make_desc_mn(sa + 2 * sub + (uint32_t)(t * nsub) * sub + k * 2048, sub);
              umma_f16(tmem_base + (uint32_t)(t * p.N), da, db, idesc, (g | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(bar_empty + 8 * stage);
        }
        umma_commit(bar_done);
      }
    } else {
      const int quad = warp & 3;
      const int row = quad * 32 + lane;                          // Cout row inside the block
      const int k = kb0 * 128 + row;
      mbar_wait(bar_done, 0);
      tc_fence_after();
      for (int t = 0; t < n_taps; ++t) {
        const int tap = tap0 + t;
        for (int c0 = 0; c0 < p.N; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * p.N + c0), v);
          if (k < p.Cout) {
            float* dst = p.dw + ((size_t)k * taps_total + tap) * p.C + cb * p.N + c0;
#pragma unroll
            for (int i = 0; i < 32; i += 4)                          // 16-byte vector reductions: a quarter of the L2 atomic operations
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(__uint_as_float(v[i])),
                           "f"(__uint_as_float(v[i + 1])), "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3])) : "memory");
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace pdf

using namespace pdf;

// NHWC activation as a 4-D tiled map whose box is `rows` full image rows (W columns) x 64 channels of one image
static int encode_rows_4d(TensorMapBlob* out, const void* ptr, int n, int h, int w, int c, int rows) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PDF_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    PDF_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  const cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)w, (cuuint32_t)rows, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box,
                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PDF_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(4d rows) failed (%d) nhwc=%d,%d,%d,%d rows=%d", (int)r, n, h, w, c, rows);
  return PDF_OK;
}

extern "C" int pdf_debug_set_wgrad_rowtile(int enable) {
  pdf::g_wg_rowtile = enable != 0;
  return PDF_OK;
}

/* tuning / A-B hook: waves of CTAs the pixel range of a weight gradient is split into (default 1) */
extern "C" int pdf_debug_set_wgrad_waves(int waves) {
  if (waves < 1 || waves > 8) { pdf::set_error("pdf_debug_set_wgrad_waves: 1..8"); return PDF_ERR_ARG; }
  pdf::g_wg_waves = waves;
  return PDF_OK;
}

/* Weight gradient of a bf16 NHWC convolution on the tensor cores.  geometry from `op`; d_x [n,h,w,c] bf16, d_dy [n,ho,wo,k] bf16,
 * d_dw [k][r][s][c] f32 (ACCUMULATED into: zero it first).  Needs c % 64 == 0 and k % 64 == 0. */
extern "C" int pdf_conv_wgrad_bf16(const pdf_op* op, const void* d_x, const void* d_dy, float* d_dw, pdf_stream_t stream) {
  PDF_REQUIRE(op && d_x && d_dy && d_dw, "pdf_conv_wgrad_bf16: null pointer");
  PDF_REQUIRE(op->n > 0 && op->c % 64 == 0 && op->k % 64 == 0 && op->r == op->s && op->stride >= 1 && op->stride <= 8,
              "pdf_conv_wgrad_bf16: needs Cin %% 64 == 0, Cout %% 64 == 0, square filter (c=%d k=%d)", op->c, op->k);
  PDF_REQUIRE(op->ho == (op->h + 2 * op->pad - op->r) / op->stride + 1 && op->wo == (op->w + 2 * op->pad - op->s) / op->stride + 1,
              "pdf_conv_wgrad_bf16: inconsistent output size");
  PDF_REQUIRE((reinterpret_cast<uintptr_t>(d_x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_dy) & 15) == 0, "pdf_conv_wgrad_bf16: pointers must be 16-byte aligned");
  if (int rc = load_driver_entry_points()) return rc;
  const int M = op->n * op->ho * op->wo;
  TensorMapBlob tdy, tx;
  pdf_op xo = *op;
  xo.d_in = d_x;
  const bool flat = op->r == 1 && op->s == 1 && op->stride == 1 && op->pad == 0;
  // row-tiled mode (same-size stride-1 filters on maps whose rows pack into 16-pixel MMA steps: 56 -> 2 rows, 28 -> 4, 14 -> 8 = 112
  // pixels per k-block): X through a 4-D TILED map instead of the im2col traversal, which delivers a fraction of the tiled rate
  // (the stem's patch matrix went 752 -> 240 us when its 1x1 "convolution" switched to a 2-D tiled map)
  int rows = 0;
  // (Cin <= 128 only: at N = 256 a 112-pixel stage is 86 KB, two ring stages, and those layers run at 680 TFLOP/s on the im2col path)
  if (g_wg_rowtile && !flat && op->c <= 128 && op->stride == 1 && op->r % 2 == 1 && op->pad == op->r / 2 && op->ho == op->h && op->wo == op->w &&
      op->w <= 128) {
    for (int rr = 1; rr <= 16 && rr * op->w <= 128; ++rr)                          // the largest block of whole rows <= 128 pixels that
      if ((rr * op->w) % 16 == 0 && (double)((op->h + rr - 1) / rr) * rr <= 1.15 * op->h) rows = rr;   // wastes <= 15 % on an image's last block
  }
  const int px = rows ? rows * op->w : kWgPix;
  if (rows) {
    if (int rc = encode_rows_4d(&tdy, d_dy, op->n, op->ho, op->wo, op->k, rows)) return rc;
    if (int rc = encode_rows_4d(&tx, d_x, op->n, op->h, op->w, op->c, rows)) return rc;
  } else if (int rc = encode_2d(&tdy, d_dy, (uint64_t)M, (uint64_t)op->k, (uint32_t)px)) {
    return rc;
  } else if (flat) {
    if (int rc = encode_2d(&tx, d_x, (uint64_t)M, (uint64_t)op->c, kWgPix)) return rc;
  } else {
    if (int rc = encode_im2col(&tx, xo, 0, kWgPix)) return rc;
  }
  WgParams p;
  p.flat = flat ? 1 : 0;
  p.px = px; p.rows = rows; p.bpi = rows ? (op->h + rows - 1) / rows : 0; p.H = op->h; p.W = op->w;
  p.M = M; p.Cout = op->k; p.C = op->c; p.R = op->r; p.S = op->s; p.Ho = op->ho; p.Wo = op->wo; p.stride = op->stride; p.pad = op->pad;
  // Cin columns per CTA: the whole Cin when it fits an instruction (N <= 256, a multiple of 64) so that dY streams once
  p.N = op->c <= 256 ? op->c : (op->c % 256 == 0 ? 256 : (op->c % 128 == 0 ? 128 : 64));
  p.T = (op->s > 1 && p.N <= 128) ? min(op->s, 512 / p.N) : 1;         // the taps of one filter row share the dY tile (T * N <= 512 columns)
  p.tap_groups = op->r * ((op->s + p.T - 1) / p.T);
  if (p.T > 1 && op->s % p.T != 0) { p.T = 1; p.tap_groups = op->r * op->s; }   // (tap groups never straddle filter rows)
  p.cin_blocks = op->c / p.N;
  const int cout_blocks = (op->k + 127) / 128;
  const int units = cout_blocks * p.tap_groups * p.cin_blocks;
  const int total_kb = rows ? op->n * p.bpi : (M + kWgPix - 1) / kWgPix;
  // pixel slabs: g_wg_waves waves of CTAs over the machine (one CTA per SM: the ring takes most of the shared memory).  Fewer slabs =
  // fewer partial sums through the L2 atomics, more slabs = shorter tail
  // (rounded DOWN: units * slabs must not exceed waves * SMs -- rounding up left a last wave of 1..24 CTAs on most layers, a whole
  //  extra CTA time: 333 -> 222 us on layer 1's 3x3 convolutions, ncu "Waves Per SM 2.01")
  int slabs = max(1, min(total_kb, (g_wg_waves * num_sms()) / units));
  p.slab_kb = (total_kb + slabs - 1) / slabs;
  slabs = (total_kb + p.slab_kb - 1) / p.slab_kb;
  p.slabs = slabs;
  p.dw = d_dw;
  const int stage_bytes = (2 + p.T * (p.N / 64)) * px * 128;
  const int tail = 8 * (2 * kWgMaxStages + 1) + 16 + 1024;
  p.stages = std::max(2, std::min(kWgMaxStages, (227 * 1024 - tail) / stage_bytes));
  const int smem = p.stages * stage_bytes + tail;
  static int configured = 0;
  if (smem > configured) {
    PDF_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  PDF_CHECK_CUDA(launch_pdl(wgrad_tc_kernel, dim3(units * slabs), dim3(192), (size_t)smem, as_stream(stream),
                            *reinterpret_cast<const CUtensorMap*>(&tdy), *reinterpret_cast<const CUtensorMap*>(&tx), p));
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}
