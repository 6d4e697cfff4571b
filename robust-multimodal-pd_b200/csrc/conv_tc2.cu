// K2 (bf16 path), CTA-PAIR variant of the implicit-GEMM convolution: tcgen05.mma.cta_group::2.
//
// Two CTAs of a cluster (two SMs of one TPC) work on ONE 256 x BLOCK_N output tile:
//   * CTA r loads ITS 128 rows of A (im2col TMA / 2-D tile) and ITS HALF of the weight tile (BLOCK_N/2 rows of B);
//   * the leader CTA's elected thread issues one M=256 MMA per K=16 step; the tensor cores of both SMs read both B halves,
//     so each SM's shared memory serves 4 KB of A + 1/2 of B per MMA instead of 4 KB + all of B -- the operand feed that
//     bounds the single-CTA kernel (profiles/r01_umma_rate.txt);
//   * every stage is 16 KB + BLOCK_N*64 bytes per CTA, so the ring is deeper (6-8 stages) in the same shared memory;
//   * TMA completions of both CTAs land on the LEADER's "full" barrier (.cta_group::2, peer bit cleared); tcgen05.commit
//     multicasts "stage free" and "accumulator ready" to both CTAs; the peer's epilogue warps arrive remotely on the
//     leader's "accumulator drained" barrier.
//
// Same warp roles as conv_tc_kernel: warp 0 TMA producer | warp 1 MMA issuer (leader only) + TMEM alloc | warps 2..5 epilogue.
#include "tc_common.cuh"
#include "ops.cuh"

namespace pdf {

struct Tc2Params {
  int M_total, Cout, Ho, Wo, stride, pad, S, cchunks, num_kb, relu, im2col, out_f32;
  const float* bias;
  const __nv_bfloat16* residual;
  void* out;
};

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the even CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma2_load_im2col_4d(uint32_t dst, const void* tmap, uint32_t bar, int c, int w, int h, int n,
                                                    uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma2_f16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi) : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {   // arrive on `bar` (same offset) in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// kind::f16 instruction descriptor for the pair: D=f32, A=B=bf16, K-major, M=256, N
__host__ __device__ constexpr uint32_t make_idesc_m256(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

constexpr int kBiasMax2 = 2048;

constexpr int kResidentKb = 18;     // BRES: k-blocks of weights kept in shared memory (3x3 x 128 input channels)

// BRES: the layer has ONE n-tile (Cout == BLOCK_N), so every tile multiplies by the same weights: each CTA of the pair keeps its
// half of them resident (kResidentKb x (BLOCK_N/2) x 128 B = 144 KB at N = 128) and only the activations stream through the
// ring.  The generic kernels re-fetch the weights for every tile, a third of the L2 -> shared-memory traffic that bounds the
// 128-channel layers (profiles/r01_conv_probe.txt).  MEASURED SLOWER (profiles/r01_op_times_resident_weights.txt: 187-192 us against
// 152-173 us for the single-CTA kernel with two M sub-tiles): next to 144 KB of weights only 64 KB of activations are in flight
// per CTA.  Kept as pair mode 3 (opt-in), parity-tested.
template <int BLOCK_N, int STAGES, bool BRES = false>
struct Smem2 {
  static constexpr int kBHalf = (BLOCK_N / 2) * kBlockK * 2;
  static constexpr int kStage = kABytes + (BRES ? 0 : kBHalf);
  static constexpr int kResOff = STAGES * kStage;                 // resident weights (BRES)
  static constexpr int kBarOff = kResOff + (BRES ? kResidentKb * kBHalf : 0);
  static constexpr int kNumBars = 2 * STAGES + 5;                 // full/empty ring, acc_full[2], acc_empty[2], weights
  static constexpr int kBiasOff = (kBarOff + kNumBars * 8 + 16 + 15) & ~15;   // f32 bias of every output channel (Cout <= kBiasMax2), staged once
  static constexpr int kDynamic = kBiasOff + kBiasMax2 * 4 + 1024;
  static_assert(2 * BLOCK_N <= 512, "two accumulator sets must fit the 512 TMEM columns");
};

template <int BLOCK_N, int STAGES, bool BRES = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Tc2Params p) {
  using L = Smem2<BLOCK_N, STAGES, BRES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bar_full = base + L::kBarOff;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  const uint32_t bar_accfull = bar_empty + STAGES * 8;
  const uint32_t bar_accempty = bar_accfull + 16;
  const uint32_t bar_w = bar_accempty + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kBarOff + L::kNumBars * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int m_tiles = (p.M_total + 2 * kBlockM - 1) / (2 * kBlockM);
  const int total_tiles = m_tiles * (p.Cout / BLOCK_N);
  float* s_bias = reinterpret_cast<float*>(smem + L::kBiasOff);    // as in conv_tc_kernel: no global load per chunk in the epilogue
  const bool bias_staged = p.bias != nullptr && p.Cout <= kBiasMax2;
  if (bias_staged)
    for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) s_bias[i] = __ldg(p.bias + i);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + s * 8, 1); mbar_init(bar_empty + s * 8, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_accfull + a * 8, 1); mbar_init(bar_accempty + a * 8, 8); }   // 4 epilogue warps x 2 CTAs
    mbar_init(bar_w, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(smem_u32(tmem_slot), 2 * BLOCK_N);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                     // barriers of BOTH CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      if (BRES) {   // this CTA's half of every weight k-block, once; all bytes of the pair are counted on the leader's barrier
        const uint32_t w_leader = bar_w & kPeerBitMask;
        if (leader) mbar_expect_tx(bar_w, 2u * (uint32_t)(p.num_kb * L::kBHalf));
        for (int kb = 0; kb < p.num_kb; ++kb)
          tma2_load_2d(base + L::kResOff + kb * L::kBHalf, &tmap_b, w_leader, kb * kBlockK, (int)rank * (BLOCK_N / 2));
      }
      uint32_t g = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        const int nt = tile / m_tiles, mtile = tile - nt * m_tiles;
        const int ms = mtile * (2 * kBlockM) + (int)rank * kBlockM;     // first output row of THIS CTA
        const int n0 = nt * BLOCK_N + (int)rank * (BLOCK_N / 2);        // first weight row of THIS CTA's B half
        int n_img = 0, w0 = 0, h0 = 0;
        if (p.im2col) {
          const int hw = p.Ho * p.Wo;
          n_img = ms / hw;
          const int rem = ms - n_img * hw;
          const int pp = rem / p.Wo, qq = rem - pp * p.Wo;
          w0 = qq * p.stride - p.pad;
          h0 = pp * p.stride - p.pad;
        }
        int cc = 0, r = 0, s = 0;
        for (int kb = 0; kb < p.num_kb; ++kb, ++g) {
          const uint32_t stage = g % STAGES, phase = (g / STAGES) & 1u;
          mbar_wait(bar_empty + stage * 8, phase ^ 1u);
          const uint32_t full_leader = (bar_full + stage * 8) & kPeerBitMask;
          if (leader) mbar_expect_tx(bar_full + stage * 8, 2u * (uint32_t)L::kStage);   // bytes of both CTAs land on this barrier
          const uint32_t sa = base + stage * L::kStage, sb = sa + kABytes;
          if (p.im2col) tma2_load_im2col_4d(sa, &tmap_a, full_leader, cc * kBlockK, w0, h0, n_img, (uint16_t)s, (uint16_t)r);
          else tma2_load_2d(sa, &tmap_a, full_leader, kb * kBlockK, ms);
          if (!BRES) tma2_load_2d(sb, &tmap_b, full_leader, kb * kBlockK, n0);
          if (++cc == p.cchunks) { cc = 0; if (++s == p.S) { s = 0; ++r; } }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_m256(BLOCK_N);
      uint32_t g = 0;
      int it = 0;
      if (BRES) mbar_wait(bar_w, 0);               // both halves of the resident weights have landed
      const uint32_t w_lo = smem_desc_lo(base + L::kResOff);
      for (int tile = pair; tile < total_tiles; tile += npairs, ++it) {
        const int acc = it & 1;
        mbar_wait(bar_accempty + acc * 8, ((uint32_t)(it >> 1) & 1u) ^ 1u);   // both CTAs' epilogues drained this accumulator set
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = 0; kb < p.num_kb; ++kb, ++g) {
          const uint32_t stage = g % STAGES, phase = (g / STAGES) & 1u;
          mbar_wait(bar_full + stage * 8, phase);
          tc_fence_after();
          const uint32_t a_lo = smem_desc_lo(base + stage * L::kStage);
          const uint32_t b_lo = BRES ? w_lo + (uint32_t)(kb * (L::kBHalf / 16)) : a_lo + (uint32_t)(kABytes / 16);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma2_f16_lo(d0, a_lo + (uint32_t)(k * 2), b_lo + (uint32_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          umma2_commit_mc(bar_empty + stage * 8);    // frees this stage in BOTH CTAs once the MMAs have read it
        }
        umma2_commit_mc(bar_accfull + acc * 8);      // accumulators of this tile complete, both CTAs
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t accempty_leader = bar_accempty & kPeerBitMask;
    int it = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs, ++it) {
      const int nt = tile / m_tiles, mtile = tile - nt * m_tiles;
      const int m = mtile * (2 * kBlockM) + (int)rank * kBlockM + row;
      const int n0 = nt * BLOCK_N;
      const int acc = it & 1;
      const bool mvalid = m < p.M_total;
      // the whole row's residual is in registers before the wait for the accumulator (see conv_tc_kernel)
      constexpr int kChunks = BLOCK_N / 32;
      uint32_t res[kChunks][2][8];
      if (p.residual && mvalid) {
#pragma unroll
        for (int ch = 0; ch < kChunks; ++ch) {
          const __nv_bfloat16* rp = p.residual + (size_t)m * p.Cout + n0 + ch * 32;
          ldg256_nc(rp, res[ch][0]);
          ldg256_nc(rp + 16, res[ch][1]);
        }
      }
      mbar_wait(bar_accfull + acc * 8, (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < kChunks; ++ch) {
        const int c0 = ch * 32;
        const int col = n0 + c0;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0), v);
        if (mvalid) {
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          if (p.bias) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b = bias_staged ? *reinterpret_cast<const float4*>(s_bias + col + i)
                                           : __ldg(reinterpret_cast<const float4*>(p.bias + col + i));
              f[i] += b.x; f[i + 1] += b.y; f[i + 2] += b.z; f[i + 3] += b.w;
            }
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
          if (p.out_f32) {
            float* op = reinterpret_cast<float*>(p.out) + (size_t)m * p.Cout + col;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              stg256(op + i * 8, __float_as_uint(f[8 * i]), __float_as_uint(f[8 * i + 1]), __float_as_uint(f[8 * i + 2]),
                     __float_as_uint(f[8 * i + 3]), __float_as_uint(f[8 * i + 4]), __float_as_uint(f[8 * i + 5]),
                     __float_as_uint(f[8 * i + 6]), __float_as_uint(f[8 * i + 7]));
          } else {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)m * p.Cout + col;
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
      if (lane == 0) mbar_arrive_cluster(accempty_leader + acc * 8);     // leader's barrier, from either CTA
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                     // the peer must not free TMEM / exit while the leader's MMAs still use its half
  if (warp == 1) tmem_dealloc_2sm(tmem_base, 2 * BLOCK_N);
}

// Measured on B200 (C2, 768 slices).  One launch at a time under ncu (profiles/r01_launches_pair_kernel.csv) the 256-wide 3x3
// layers take the same time as the single-CTA kernel and the 128-wide layers are 25-40 % SLOWER than the single-CTA kernel with
// two M sub-tiles per B tile.  Back to back, however, the conv stack runs into the 1000 W power cap (SM clock 1.4-1.55 GHz,
// profiles/r01_stack_power.txt): there the pair kernel, which reads half of the B operand from shared memory per CTA, lets
// layers 3-4 clock ~60 MHz higher and finish 5 % sooner.  Default: pairs for the 256-wide 3x3 layers only.
static int g_pair_mode = 2;     // 0 off, 1 every eligible Cout >= 128 layer, 2 only 3x3 layers with 256-wide tiles, 3 = 2 + resident weights at N = 128

bool pair_eligible(const TcConv& tc) {
  if (!g_pair_mode || tc.halo) return false;
  if (tc.block_n != 256 && tc.block_n != 128) return false;
  if (g_pair_mode == 2 && (tc.block_n != 256 || tc.R * tc.S == 1)) return false;
  if (g_pair_mode == 3 && tc.block_n != 256 && !(tc.R * tc.S == 9 && tc.block_n == 128 && tc.Cout == 128 && tc.R * tc.S * tc.cchunks <= kResidentKb)) return false;
  if (g_pair_mode == 3 && tc.block_n == 256 && tc.R * tc.S == 1) return false;
  const long tiles = (long)ceil_div(tc.M_total, 2 * kBlockM) * (tc.Cout / tc.block_n);
  return tiles >= num_sms() / 2;          // at least one tile per CTA pair
}

template <int BLOCK_N, int STAGES, bool BRES = false>
static int launch_tc2(const TcConv& tc, cudaStream_t s) {
  using L = Smem2<BLOCK_N, STAGES, BRES>;
  static bool configured = false;
  if (!configured) {
    PDF_CHECK_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<BLOCK_N, STAGES, BRES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic));
    configured = true;
  }
  Tc2Params p;
  p.M_total = tc.M_total; p.Cout = tc.Cout; p.Ho = tc.Ho; p.Wo = tc.Wo; p.stride = tc.stride; p.pad = tc.pad; p.S = tc.S;
  p.cchunks = tc.cchunks; p.num_kb = tc.R * tc.S * tc.cchunks; p.relu = tc.relu; p.im2col = tc.im2col; p.out_f32 = tc.out_f32;
  p.bias = tc.bias; p.residual = reinterpret_cast<const __nv_bfloat16*>(tc.residual); p.out = tc.out;
  const int total_tiles = ceil_div(tc.M_total, 2 * kBlockM) * (tc.Cout / BLOCK_N);
  const int pairs = max(1, min(total_tiles, num_sms() / 2));
  conv_tc2_kernel<BLOCK_N, STAGES, BRES><<<2 * pairs, 192, L::kDynamic, s>>>(*reinterpret_cast<const CUtensorMap*>(&tc.tmap_a),
                                                                             *reinterpret_cast<const CUtensorMap*>(&tc.tmap_b2), p);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

// Probe: back-to-back tcgen05.mma.cta_group::2 (M = 256 over the pair, N, K = 16) from shared memory, the pair counterpart of
// umma_rate_kernel: each CTA holds its 128 A rows and N/2 B rows.  out[pair] = cycles (leader clock).
template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
umma2_rate_kernel(int iters, int mode, unsigned long long* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t sa = base, sb = base + 3 * kABytes, bar = sb + BLOCK_N * 64;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 3 * kABytes + BLOCK_N * 64 + 16);
  const int warp = threadIdx.x >> 5;
  const bool leader = cluster_ctarank() == 0;
  for (uint32_t i = threadIdx.x; i < (3 * kABytes + BLOCK_N * 64) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc_2sm(smem_u32(tmem_slot), BLOCK_N < 32 ? 32 : BLOCK_N);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (leader && warp == 0 && elect_one()) {
    constexpr uint32_t idesc = make_idesc_m256(BLOCK_N);
    const uint32_t a_lo = smem_desc_lo(sa), b_lo = smem_desc_lo(sb);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const uint32_t shift = (mode & 1) ? (uint32_t)((t / 3) * 58 + (t % 3)) * 8u : 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma2_f16_lo(tmem_base, a_lo + shift + k * 2, b_lo + k * 2, idesc, 1u);
      }
    }
    umma2_commit_mc(bar);
    mbar_wait(bar, 0);
    out[blockIdx.x >> 1] = (unsigned long long)(clock64() - t0);
  }
  if (!leader && warp == 0 && elect_one()) mbar_wait(bar, 0);      // the multicast commit arrives here too
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm(tmem_base, BLOCK_N < 32 ? 32 : BLOCK_N);
}

int launch_umma2_rate(int N, int iters, int mode, int pairs, unsigned long long* d_cycles, cudaStream_t s) {
  const int smem = 3 * kABytes + N * 64 + 64 + 1024;
  if (N == 64) {
    PDF_CHECK_CUDA(cudaFuncSetAttribute(umma2_rate_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma2_rate_kernel<64><<<2 * pairs, 128, smem, s>>>(iters, mode, d_cycles);
  } else if (N == 128) {
    PDF_CHECK_CUDA(cudaFuncSetAttribute(umma2_rate_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma2_rate_kernel<128><<<2 * pairs, 128, smem, s>>>(iters, mode, d_cycles);
  } else {
    PDF_CHECK_CUDA(cudaFuncSetAttribute(umma2_rate_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma2_rate_kernel<256><<<2 * pairs, 128, smem, s>>>(iters, mode, d_cycles);
  }
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

static bool resident_eligible(const TcConv& tc) {      // one n-tile, all weight k-blocks fit next to the activation ring
  return tc.block_n == 128 && tc.Cout == 128 && tc.R * tc.S * tc.cchunks <= kResidentKb;
}

int launch_conv_tc2(const TcConv& tc, cudaStream_t s) {
  if (tc.block_n == 256) return launch_tc2<256, 6>(tc, s);
  if (g_pair_mode == 3 && resident_eligible(tc)) return launch_tc2<128, 4, true>(tc, s);
  return launch_tc2<128, 8>(tc, s);
}

}  // namespace pdf

/* tuning / test hook: 1 = route eligible Cout >= 128 layers through the CTA-pair (cta_group::2) kernel, 2 = only the 3x3 layers with
 * 256-wide tiles, 0 = off */
extern "C" int pdf_debug_enable_pair(int enable) {
  pdf::g_pair_mode = enable;
  return PDF_OK;
}
