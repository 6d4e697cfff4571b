// K2: 3x3 stride-1 pad-1 convolution with the three HORIZONTAL taps of a filter row served by ONE activation tile.
//
// The generic implicit-GEMM kernel (conv_tc.cu) fetches the im2col A tile once per filter tap: nine L2 -> shared-memory
// transfers of the same pixels.  For the 128-channel layers that feed, not the tensor pipe, is the bound: with MMAs and epilogue
// switched off the TMA ring alone takes 108-123 of their 149-176 us (profiles/r01_conv_probe.txt).
//
// Here the im2col tensor map traverses W + 2 base pixels per image row instead of W (bounding-box upper corner +1 instead of
// -1): output positions live in "padded-row" space m' = (n*H + p)*(W+2) + q', q' in [0, W+2), the last two of every row dead.
// In that space the pixel the tap (r, s+1) needs at position m' is the pixel the tap (r, s) needs at m'+1 -- also across row and
// image boundaries, where im2col-mode TMA zero-fills the out-of-bounds columns.  One load of 130 consecutive positions for
// (r, s = 0) therefore serves s = 0, 1, 2 through MMA descriptors advanced by 0, 1, 2 rows of 128 bytes (a descriptor may start at
// any 128-byte row of a 128B-swizzled tile, profiles/r01_umma_shift_probe.txt).  A is fetched 3x instead of 9x; the price is
// (W+2)/W more MMA work (7 % at W = 28).  Unlike the halo-resident layer-1 kernel (conv3x3_tc.cu) nothing is quantised to an
// image: tiles run through rows and images, so small feature maps do not waste tile slots.
//
// Per k-step (filter row r, 64-channel chunk cc) a stage holds MT x [130 x 64] activations and the three [BLOCK_N x 64] weight
// blocks of taps (r, 0..2): 82 KB at MT = 2, N = 128, two stages.  Persistent CTAs, two TMEM accumulator sets, the epilogue of
// conv_tc.cu (bias in shared memory, whole-tile residual prefetch, ReLU, 256-bit bf16 stores) with the padded-row index mapping.
// warp 0: TMA producer | warp 1: tcgen05.mma issuer (+TMEM alloc) | warps 2..5: epilogue
#include "tc_common.cuh"
#include "ops.cuh"

namespace pdf {

constexpr int kHsN = 128;                         // output channels per tile
constexpr int kHsMT = 2;                          // 128-position sub-tiles per tile (share the weight blocks)
constexpr int kHsStages = 2;
constexpr int kHsPix = kBlockM + 2;               // positions per activation load
constexpr int kHsABytes = (kHsPix * 128 + 1023) / 1024 * 1024;   // 17 KB
constexpr int kHsBBytes = kHsN * 128;             // one tap's [N x 64] weight block
constexpr int kHsStage = kHsMT * kHsABytes + 3 * kHsBBytes;
constexpr int kHsBarOff = kHsStages * kHsStage;
constexpr int kHsNumBars = 2 * kHsStages + 4;
constexpr int kHsBiasOff = (kHsBarOff + kHsNumBars * 8 + 16 + 15) & ~15;
constexpr int kHsBiasMax = 2048;
constexpr int kHsDynamic = kHsBiasOff + kHsBiasMax * 4 + 1024;

struct HsParams {
  int M_pad;            // N * H * (W + 2) positions
  int H, W, Wp, Cout, cchunks, relu;
  const float* bias;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out;
};

__global__ void __launch_bounds__(192, 1)
conv3x3_hs_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const HsParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bar_full = base + kHsBarOff;
  const uint32_t bar_empty = bar_full + kHsStages * 8;
  const uint32_t bar_accfull = bar_empty + kHsStages * 8;
  const uint32_t bar_accempty = bar_accfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kHsBarOff + kHsNumBars * 8);
  float* s_bias = reinterpret_cast<float*>(smem + kHsBiasOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M_pad + kBlockM * kHsMT - 1) / (kBlockM * kHsMT);
  const int total_tiles = m_tiles * (p.Cout / kHsN);
  const int num_ks = 3 * p.cchunks;               // k-steps per tile: filter row x channel chunk
  for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) s_bias[i] = p.bias ? __ldg(p.bias + i) : 0.f;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < kHsStages; ++s) { mbar_init(bar_full + s * 8, 1); mbar_init(bar_empty + s * 8, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_accfull + a * 8, 1); mbar_init(bar_accempty + a * 8, 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);        // 2 accumulator sets x (2 sub-tiles x 128 columns)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // see common.cuh
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      uint32_t g = 0;
      const int hwp = p.H * p.Wp;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile / m_tiles, mtile = tile - nt * m_tiles;
        const int m0 = mtile * (kBlockM * kHsMT), n0 = nt * kHsN;
        const int n_sub = min(kHsMT, (p.M_pad - m0 + kBlockM - 1) / kBlockM);
        int n_img[kHsMT], w0[kHsMT], h0[kHsMT];
#pragma unroll
        for (int mt = 0; mt < kHsMT; ++mt) {
          const int ms = m0 + mt * kBlockM;
          n_img[mt] = ms / hwp;
          const int rem = ms - n_img[mt] * hwp;
          const int pp = rem / p.Wp;
          w0[mt] = (rem - pp * p.Wp) - 1;          // base pixel of position q' (filter column 0): w = q' - 1
          h0[mt] = pp - 1;
        }
        const uint32_t tx_bytes = (uint32_t)(n_sub * kHsPix * 128 + 3 * kHsBBytes);
        int r = 0, cc = 0;
        for (int ks = 0; ks < num_ks; ++ks, ++g) {
          const uint32_t stage = g % kHsStages, phase = (g / kHsStages) & 1u;
          mbar_wait(bar_empty + stage * 8, phase ^ 1u);
          mbar_expect_tx(bar_full + stage * 8, tx_bytes);
          const uint32_t sa = base + stage * kHsStage, sb = sa + kHsMT * kHsABytes;
#pragma unroll
          for (int mt = 0; mt < kHsMT; ++mt)
            if (mt < n_sub)
              tma_load_im2col_4d(sa + mt * kHsABytes, &tmap_a, bar_full + stage * 8, cc * kBlockK, w0[mt], h0[mt], n_img[mt], (uint16_t)0,
                                 (uint16_t)r);
#pragma unroll
          for (int s = 0; s < 3; ++s)              // weights [Cout][R][S][Cin]: column of (r, s, cc)
            tma_load_2d(sb + s * kHsBBytes, &tmap_b, bar_full + stage * 8, ((r * 3 + s) * p.cchunks + cc) * kBlockK, n0);
          if (++cc == p.cchunks) { cc = 0; ++r; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(kHsN);
      uint32_t g = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int mtile = tile % m_tiles;
        const int m0 = mtile * (kBlockM * kHsMT);
        const int n_sub = min(kHsMT, (p.M_pad - m0 + kBlockM - 1) / kBlockM);
        const int acc = it & 1;
        mbar_wait(bar_accempty + acc * 8, ((uint32_t)(it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(acc * kHsMT * kHsN);
        for (int ks = 0; ks < num_ks; ++ks, ++g) {
          const uint32_t stage = g % kHsStages, phase = (g / kHsStages) & 1u;
          mbar_wait(bar_full + stage * 8, phase);
          tc_fence_after();
          const uint32_t a_lo = smem_desc_lo(base + stage * kHsStage), b_lo = a_lo + (uint32_t)(kHsMT * kHsABytes / 16);
#pragma unroll
          for (int mt = 0; mt < kHsMT; ++mt) {
            if (mt < n_sub) {
#pragma unroll
              for (int s = 0; s < 3; ++s) {
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)   // tap s: the same tile, s rows (8 x 16-byte units each) further down
                  umma_f16_lo(d0 + mt * kHsN, a_lo + (uint32_t)(mt * (kHsABytes / 16) + s * 8 + k * 2),
                              b_lo + (uint32_t)(s * (kHsBBytes / 16) + k * 2), idesc, (ks | s | k) != 0 ? 1u : 0u);
              }
            }
          }
          umma_commit(bar_empty + stage * 8);
        }
        umma_commit(bar_accfull + acc * 8);
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int hwp = p.H * p.Wp;
    constexpr int kCPT = kHsN / 32;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int nt = tile / m_tiles, mtile = tile - nt * m_tiles;
      const int m0 = mtile * (kBlockM * kHsMT), n0 = nt * kHsN;
      const int n_sub = min(kHsMT, (p.M_pad - m0 + kBlockM - 1) / kBlockM);
      const int acc = it & 1;
      // position -> output pixel: live when q' < W
      size_t obase[kHsMT];
      bool live[kHsMT];
#pragma unroll
      for (int mt = 0; mt < kHsMT; ++mt) {
        const int m = m0 + mt * kBlockM + row;
        const int n = m / hwp;
        const int rem = m - n * hwp;
        const int pp = rem / p.Wp, qq = rem - pp * p.Wp;
        live[mt] = mt < n_sub && m < p.M_pad && qq < p.W;
        obase[mt] = (((size_t)n * p.H + pp) * p.W + qq) * p.Cout + n0;
      }
      uint32_t res[kHsMT * kCPT][2][8];
      if (p.residual) {
#pragma unroll
        for (int ch = 0; ch < kHsMT * kCPT; ++ch) {
          const int mt = ch / kCPT, c0 = (ch - mt * kCPT) * 32;
          if (live[mt]) {
            const __nv_bfloat16* rp = p.residual + obase[mt] + c0;
            ldg256_nc(rp, res[ch][0]);
            ldg256_nc(rp + 16, res[ch][1]);
          }
        }
      }
      mbar_wait(bar_accfull + acc * 8, (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < kHsMT * kCPT; ++ch) {
        const int mt = ch / kCPT, c0 = (ch - mt * kCPT) * 32;
        if (mt >= n_sub) continue;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kHsMT * kHsN + mt * kHsN + c0), v);
        if (live[mt]) {
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b = *reinterpret_cast<const float4*>(s_bias + n0 + c0 + i);
            f[i] = __uint_as_float(v[i]) + b.x; f[i + 1] = __uint_as_float(v[i + 1]) + b.y;
            f[i + 2] = __uint_as_float(v[i + 2]) + b.z; f[i + 3] = __uint_as_float(v[i + 3]) + b.w;
          }
          if (p.residual) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
              for (int j = 0; j < 8; ++j) { f[i * 16 + j * 2] += bf16_lo(res[ch][i][j]); f[i * 16 + j * 2 + 1] += bf16_hi(res[ch][i][j]); }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
          }
          __nv_bfloat16* op = p.out + obase[mt] + c0;
#pragma unroll
          for (int i = 0; i < 2; ++i)
            stg256(op + i * 16, pack_bf16x2(f[i * 16 + 0], f[i * 16 + 1]), pack_bf16x2(f[i * 16 + 2], f[i * 16 + 3]),
                   pack_bf16x2(f[i * 16 + 4], f[i * 16 + 5]), pack_bf16x2(f[i * 16 + 6], f[i * 16 + 7]),
                   pack_bf16x2(f[i * 16 + 8], f[i * 16 + 9]), pack_bf16x2(f[i * 16 + 10], f[i * 16 + 11]),
                   pack_bf16x2(f[i * 16 + 12], f[i * 16 + 13]), pack_bf16x2(f[i * 16 + 14], f[i * 16 + 15]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_accempty + acc * 8) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static int g_hs_mode = 1;    // 0 off, 1 on for eligible layers with Cout == 128 (one n-tile), 2 on for every eligible layer, 3 = 2 regardless of size

bool hs_eligible(const pdf_op& op) {
  if (!g_hs_mode) return false;
  if (!(op.r == 3 && op.s == 3 && op.stride == 1 && op.pad == 1 && op.h == op.ho && op.w == op.wo)) return false;
  if (op.c % kBlockK != 0 || op.k % kHsN != 0 || op.k > kHsBiasMax || op.out_f32 || op.d_weight2) return false;
  if (g_hs_mode == 1 && op.k != kHsN) return false;
  const long tiles = (long)ceil_div((long long)op.n * op.h * (op.w + 2), kBlockM * kHsMT) * (op.k / kHsN);
  if (g_hs_mode == 3) return op.w >= 4;            // (tests: small problems too)
  return op.w >= 4 && tiles >= 2L * num_sms();     // enough work for every SM, else the generic kernel's smaller tiles win
}

int launch_conv3x3_hs(const TcConv& tc, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    PDF_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_hs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHsDynamic));
    configured = true;
  }
  HsParams p;
  p.H = tc.Ho; p.W = tc.Wo; p.Wp = tc.Wo + 2; p.Cout = tc.Cout; p.cchunks = tc.cchunks; p.relu = tc.relu;
  p.M_pad = tc.n_images * tc.Ho * p.Wp;
  p.bias = tc.bias; p.residual = reinterpret_cast<const __nv_bfloat16*>(tc.residual); p.out = reinterpret_cast<__nv_bfloat16*>(tc.out);
  const int total_tiles = ceil_div(p.M_pad, kBlockM * kHsMT) * (tc.Cout / kHsN);
  const int grid = max(1, min(total_tiles, num_sms()));
  PDF_CHECK_CUDA(launch_pdl(conv3x3_hs_kernel, dim3(grid), dim3(192), (size_t)kHsDynamic, s, *reinterpret_cast<const CUtensorMap*>(&tc.tmap_a),
                            *reinterpret_cast<const CUtensorMap*>(&tc.tmap_b), p));
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

}  // namespace pdf

/* tuning / test hook for the horizontally-shared 3x3 kernel (conv3x3_hs.cu): 0 = off, 1 (default) = eligible 3x3 stride-1 layers with
 * Cout == 128, 2 = every eligible layer (Cout % 128 == 0).  Read when a plan is created. */
extern "C" int pdf_debug_set_hs_mode(int mode) {
  pdf::g_hs_mode = mode;
  return PDF_OK;
}
