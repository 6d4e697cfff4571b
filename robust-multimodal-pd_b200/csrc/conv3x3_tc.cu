// K2: 3x3 stride-1 pad-1 convolution, 64 -> 64 channels (ResNet layer 1), halo-resident variant.
//
// The generic implicit-GEMM kernel re-reads the activation once per filter tap (9x) through im2col TMA and is
// bound by the L2 -> shared-memory feed.  Here each 256-pixel tile loads its input HALO ONCE:
//
//   * output pixels are indexed in "padded-row" space  m' = p*(W+2) + q  (q >= W are two dead positions per row);
//   * one 4-D TMA box {64 ch, W+2 columns starting at w=-1, NR rows starting at h=p_lo-1, 1 image} lands the halo as
//     consecutive 128-byte rows (zero fill = the convolution's padding), 128B-swizzled;
//   * for tap (r,s) the A operand of the MMA is THE SAME shared-memory tile, read through a descriptor whose start
//     address is advanced by (r*(W+2)+s) rows -- the swizzle is a function of the absolute address, so any 128-byte
//     row is a legal operand start (profiles/r01_umma_shift_probe.txt);
//   * all 9 weight taps (72 KB) stay resident in shared memory; CTAs are persistent over tiles; two TMEM accumulator
//     sets let the epilogue of tile t overlap the MMAs of tile t+1.
//
// warp 0: TMA producer | warp 1: tcgen05.mma issuer (+TMEM alloc) | warps 2..5: epilogue (bias, residual, ReLU, bf16 NHWC)
#include "tc_common.cuh"
#include "ops.cuh"

namespace pdf {

constexpr int kHaloBM = 256;       // output positions (padded-row space) per tile
constexpr int kHaloStages = 2;
constexpr int kPrefetchAhead = 2;   // tiles warmed in L2 beyond the ones resident in shared memory

struct HaloParams {
  int H, W, Wp, NR, tiles_per_image, total_tiles, relu;
  const float* bias;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out;
};

__global__ void __launch_bounds__(192, 1)
conv3x3_c64_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t a_bytes = ((uint32_t)(p.NR * p.Wp * 128) + 1023u) & ~1023u;   // one halo stage
  const uint32_t s_w = base;                                   // 9 taps x [64 cout x 64 cin] bf16 = 9 x 8 KB
  const uint32_t s_a = base + 9 * 8192;
  const uint32_t bar0 = s_a + kHaloStages * a_bytes;
  const uint32_t bar_w = bar0, bar_afull = bar0 + 8, bar_aempty = bar_afull + 8 * kHaloStages;
  const uint32_t bar_accfull = bar_aempty + 8 * kHaloStages, bar_accempty = bar_accfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (bar_accempty + 16 - base));
  float4* s_bias4 = reinterpret_cast<float4*>(smem + (bar0 + 96 - base));               // 64 f32, broadcast reads in the epilogue
  if (threadIdx.x < 16) s_bias4[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(p.bias) + threadIdx.x);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
    mbar_init(bar_w, 1);
    for (int s = 0; s < kHaloStages; ++s) { mbar_init(bar_afull + 8 * s, 1); mbar_init(bar_aempty + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_accfull + 8 * a, 1); mbar_init(bar_accempty + 8 * a, 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 256);        // 2 accumulator sets x (2 halves x 64 columns)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // see common.cuh: the successor's prologue overlaps this kernel's tail,
  pdl_wait();                // activations are touched only after the predecessor has completed

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_w, 9 * 8192);
      for (int t = 0; t < 9; ++t) tma_load_2d(s_w + t * 8192, &tmap_w, bar_w, t * 64, 0);
      // Warm L2 with the halo rows of the tile kPrefetchAhead iterations ahead: the shared-memory ring is only two stages deep.
      // (Prefetching the residual rows the same way was measured to DOUBLE their DRAM traffic (1.1 GB vs 0.62 GB per launch in an ncu --set full capture): removed.)
      auto prefetch_tile = [&](int tile) {
        if (tile >= p.total_tiles) return;
        const int n = tile / p.tiles_per_image;
        const int m0 = (tile - n * p.tiles_per_image) * kHaloBM;
        const int p_lo = m0 / p.Wp;
        asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                     ::"l"(reinterpret_cast<uint64_t>(&tmap_a)), "r"(0), "r"(-1), "r"(p_lo - 1), "r"(n) : "memory");
      };
      for (int a = kHaloStages; a < kHaloStages + kPrefetchAhead; ++a) prefetch_tile(blockIdx.x + a * gridDim.x);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int stage = it % kHaloStages;
        const uint32_t phase = (uint32_t)(it / kHaloStages) & 1u;
        const int n = tile / p.tiles_per_image;
        const int m0 = (tile - n * p.tiles_per_image) * kHaloBM;
        const int p_lo = m0 / p.Wp;
        prefetch_tile(tile + (kHaloStages + kPrefetchAhead) * gridDim.x);
        mbar_wait(bar_aempty + 8 * stage, phase ^ 1u);
        mbar_expect_tx(bar_afull + 8 * stage, (uint32_t)(p.NR * p.Wp * 128));
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(s_a + stage * a_bytes), "l"(reinterpret_cast<uint64_t>(&tmap_a)), "r"(bar_afull + 8 * stage), "r"(0), "r"(-1),
              "r"(p_lo - 1), "r"(n)
            : "memory");
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(64);
      // one thread issues 72 MMAs of only 32 tensor-cycles each per tile: descriptor arithmetic must stay at ~1 add per MMA
      uint32_t tap_off[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) tap_off[t] = (uint32_t)((t / 3) * p.Wp + (t % 3)) * 8u;   // rows -> 16-byte units
      const uint32_t w_lo = smem_desc_lo(s_w);
      mbar_wait(bar_w, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int stage = it % kHaloStages, acc = it & 1;
        const uint32_t phase = (uint32_t)(it / kHaloStages) & 1u, accphase = (uint32_t)(it >> 1) & 1u;
        const int n = tile / p.tiles_per_image;
        const int m0 = (tile - n * p.tiles_per_image) * kHaloBM;
        const int off0 = m0 - (m0 / p.Wp) * p.Wp;
        mbar_wait(bar_accempty + 8 * acc, accphase ^ 1u);      // epilogue has drained this accumulator set
        mbar_wait(bar_afull + 8 * stage, phase);
        tc_fence_after();
        const uint32_t a_lo = smem_desc_lo(s_a + stage * a_bytes) + (uint32_t)off0 * 8u;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t d = tmem_base + (uint32_t)(acc * 128 + half * 64);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_lo(d, a_lo + tap_off[t] + (uint32_t)(half * 128 * 8 + k * 2), w_lo + (uint32_t)(t * 512 + k * 2), idesc,
                          (t | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(bar_aempty + 8 * stage);
        umma_commit(bar_accfull + 8 * acc);
      }
    }
  } else {
    const int quad = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t accphase = (uint32_t)(it >> 1) & 1u;
      const int n = tile / p.tiles_per_image;
      const int m0 = (tile - n * p.tiles_per_image) * kHaloBM;
      // residual rows of both halves are fetched BEFORE waiting on the accumulator: their latency hides behind the MMAs.
      // Each thread owns one output row (128 bytes): 256-bit accesses move it as four full sectors.
      uint32_t res[2][4][8];
      bool valid[2];
      size_t obase[2];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int m = m0 + half * 128 + quad * 32 + lane;
        const int pp = m / p.Wp, qq = m - pp * p.Wp;
        valid[half] = pp < p.H && qq < p.W;
        obase[half] = (((size_t)n * p.H + pp) * p.W + qq) * 64;
        if (p.residual && valid[half]) {
#pragma unroll
          for (int i = 0; i < 4; ++i) ldg256_nc(p.residual + obase[half] + i * 16, res[half][i]);
        }
      }
      mbar_wait(bar_accfull + 8 * acc, accphase);
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 128 + half * 64 + c0), v);
          if (valid[half]) {
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b = s_bias4[(c0 + i) >> 2];
              f[i] = __uint_as_float(v[i]) + b.x; f[i + 1] = __uint_as_float(v[i + 1]) + b.y;
              f[i + 2] = __uint_as_float(v[i + 2]) + b.z; f[i + 3] = __uint_as_float(v[i + 3]) + b.w;
            }
            if (p.residual) {
#pragma unroll
              for (int i = 0; i < 2; ++i) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const uint32_t w = res[half][(c0 >> 4) + i][j];
                  f[i * 16 + j * 2] += bf16_lo(w);
                  f[i * 16 + j * 2 + 1] += bf16_hi(w);
                }
              }
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            __nv_bfloat16* op = p.out + obase[half] + c0;
#pragma unroll
            for (int i = 0; i < 2; ++i)
              stg256(op + i * 16, pack_bf16x2(f[i * 16 + 0], f[i * 16 + 1]), pack_bf16x2(f[i * 16 + 2], f[i * 16 + 3]),
                     pack_bf16x2(f[i * 16 + 4], f[i * 16 + 5]), pack_bf16x2(f[i * 16 + 6], f[i * 16 + 7]),
                     pack_bf16x2(f[i * 16 + 8], f[i * 16 + 9]), pack_bf16x2(f[i * 16 + 10], f[i * 16 + 11]),
                     pack_bf16x2(f[i * 16 + 12], f[i * 16 + 13]), pack_bf16x2(f[i * 16 + 14], f[i * 16 + 15]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_accempty + 8 * acc) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

int halo_smem_bytes(int NR, int Wp) {
  const int a_bytes = (NR * Wp * 128 + 1023) & ~1023;
  return 9 * 8192 + kHaloStages * a_bytes + 8 * (1 + 2 * kHaloStages + 4) + 32 + 256 + 1024;
}

bool halo_eligible(const pdf_op& op) {
  if (!(op.r == 3 && op.s == 3 && op.stride == 1 && op.pad == 1 && op.c == 64 && op.k == 64 && op.h == op.ho && op.w == op.wo)) return false;
  const int Wp = op.w + 2;
  const int NR = (Wp - 1 + kHaloBM - 1) / Wp + 3;
  return op.w >= 8 && Wp <= 256 && NR <= 256 && halo_smem_bytes(NR, Wp) <= 225 * 1024;
}

int launch_conv3x3_halo(const TcConv& tc, cudaStream_t s) {
  static int configured = 0;
  const int smem = halo_smem_bytes(tc.halo_nr, tc.halo_wp);
  if (smem > configured) {
    PDF_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_c64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  HaloParams p;
  p.H = tc.Ho; p.W = tc.Wo; p.Wp = tc.halo_wp; p.NR = tc.halo_nr;
  p.tiles_per_image = ceil_div(tc.Ho * tc.halo_wp, kHaloBM);
  p.total_tiles = tc.n_images * p.tiles_per_image;
  p.relu = tc.relu; p.bias = tc.bias; p.residual = reinterpret_cast<const __nv_bfloat16*>(tc.residual);
  p.out = reinterpret_cast<__nv_bfloat16*>(tc.out);
  const int grid = min(p.total_tiles, num_sms());
  PDF_CHECK_CUDA(launch_pdl(conv3x3_c64_kernel, dim3(grid), dim3(192), (size_t)smem, s,
                            *reinterpret_cast<const CUtensorMap*>(&tc.tmap_a), *reinterpret_cast<const CUtensorMap*>(&tc.tmap_b), p));
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

}  // namespace pdf
