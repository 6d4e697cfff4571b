// K3 (tensor path): the two linear layers of the MIL attention head as tcgen05 GEMMs on the FP32 bags, no conversion pass.
//
//   C[M, N] = act( A[M, K] * B[N, K]^T + bias )       A, B float32 in HBM, tcgen05.mma kind::tf32 (f32 operands read from
//                                                     shared memory, 10-bit mantissa products, f32 accumulation in TMEM)
//   mode 0  C = relu(.) stored as f32 [M, N]           h = ReLU(instance(x))                     models/mil_attention.py:42
//   mode 1  no C: N = 2A (gated: [W_v; W_u]) or A;     s[m] = sum_a w[a] tanh(v_a) (sigmoid(u_a)) + b_w    :43-46
//           the attention scores come straight out of the epilogue (one thread owns one instance row)
//
// The bags [n_bags * L, D] are the only large operand (4*L*D bytes per subject): they are read ONCE from HBM by
// TMA; the weight matrix is re-streamed from L2 per row tile.  Warp roles as in conv_tc.cu: TMA producer, single-thread
// MMA issuer, 4 epilogue warps, persistent over row tiles.  Two tile shapes: MT = 1 (128 rows, two TMEM accumulator sets: the
// epilogue of tile i overlaps the MMAs of tile i+1), the default, and MT = 2 (256 rows against ONE pass of the weight k-blocks,
// all 512 TMEM columns as one accumulator set), an opt-in that halves the weight share of the shared-memory fill but measured
// no faster (pdf_debug_set_mil_mt).
#include "tc_common.cuh"
#include "ops.cuh"

namespace pdf {

constexpr int kTfK = 32;                       // f32 elements per 128-byte swizzle row

struct TfParams {
  int M, N, K, mode, gated, A;
  const float* bias;     // [N]
  const float* w_w;      // mode 1: [A]
  const float* b_w;      // mode 1: [1]
  float* out;            // mode 0: [M, N]; mode 1: [M]
};

__host__ __device__ constexpr uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
      : "memory");
}

// hardware transcendental approximations (MUFU.TANH / MUFU.EX2 / MUFU.RCP, relative error ~2^-11 / 2^-22): the products feeding them
// carry tf32's 10-bit mantissa already, and the score epilogue -- 128..256 tanh + sigmoid per instance row, one row per thread --
// was 4x the GEMM it follows with libm's tanhf / expf
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

template <int BN, int MT>
__global__ void __launch_bounds__(192, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const TfParams p) {
  constexpr int kStage = MT * kABytes + BN * 128;
  constexpr int kTfStages = (MT == 1) ? 4 : 3;
  constexpr int kAccSets = (2 * MT * BN <= 512) ? 2 : 1;
  constexpr int kTmemCols = (kAccSets * MT * BN <= 64) ? 64 : (kAccSets * MT * BN <= 128 ? 128 : (kAccSets * MT * BN <= 256 ? 256 : 512));
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bar_full = base + kTfStages * kStage;
  const uint32_t bar_empty = bar_full + 8 * kTfStages;
  const uint32_t bar_accfull = bar_empty + 8 * kTfStages;
  const uint32_t bar_accempty = bar_accfull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kTfStages * kStage + 8 * (2 * kTfStages + 4));
  float* s_bias = reinterpret_cast<float*>(smem + kTfStages * kStage + 8 * (2 * kTfStages + 4) + 16);   // [BN] bias | [BN] w_w (mode 1)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < BN; i += blockDim.x) {
    s_bias[i] = (p.bias && i < p.N) ? __ldg(p.bias + i) : 0.f;
    s_bias[BN + i] = (p.mode == 1 && i < p.A) ? __ldg(p.w_w + i) : 0.f;
  }
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b);
    for (int s = 0; s < kTfStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_accfull + 8 * a, 1); mbar_init(bar_accempty + 8 * a, 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const int m_tiles = (p.M + MT * kBlockM - 1) / (MT * kBlockM);
  const int num_kb = p.K / kTfK;
  if (warp == 0) {
    if (elect_one()) {
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < num_kb; ++kb, ++g) {
          const uint32_t stage = g % kTfStages;
          mbar_wait(bar_empty + 8 * stage, ((g / kTfStages) & 1u) ^ 1u);
          mbar_expect_tx(bar_full + 8 * stage, (uint32_t)kStage);
          const uint32_t sa = base + stage * kStage;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)                           // (rows past M: zero fill)
            tma_load_2d(sa + mt * kABytes, &tmap_a, bar_full + 8 * stage, kb * kTfK, (tile * MT + mt) * kBlockM);
          tma_load_2d(sa + MT * kABytes, &tmap_b, bar_full + 8 * stage, kb * kTfK, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_tf32(BN);
      uint32_t g = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
        const int acc = it % kAccSets;
        mbar_wait(bar_accempty + 8 * acc, (((uint32_t)(it / kAccSets)) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(acc * MT * BN);
        for (int kb = 0; kb < num_kb; ++kb, ++g) {
          const uint32_t stage = g % kTfStages;
          mbar_wait(bar_full + 8 * stage, (g / kTfStages) & 1u);
          tc_fence_after();
          const uint32_t a_lo = smem_desc_lo(base + stage * kStage), b_lo = a_lo + (uint32_t)(MT * kABytes / 16);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int k = 0; k < 4; ++k)    // K = 8 f32 = 32 bytes per instruction
              umma_tf32_lo(d0 + (uint32_t)(mt * BN), a_lo + (uint32_t)(mt * (kABytes / 16) + k * 2), b_lo + (uint32_t)(k * 2), idesc,
                           (kb | k) != 0 ? 1u : 0u);
          umma_commit(bar_empty + 8 * stage);
        }
        umma_commit(bar_accfull + 8 * acc);
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
      const int acc = it % kAccSets;
      mbar_wait(bar_accfull + 8 * acc, ((uint32_t)(it / kAccSets)) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
      const int m = (tile * MT + mt) * kBlockM + row;
      const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((acc * MT + mt) * BN);
      if (p.mode == 0) {
#pragma unroll 1
        for (int c0 = 0; c0 < p.N; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(t0 + (uint32_t)c0, v);
          if (m < p.M) {
            float* op = p.out + (size_t)m * p.N + c0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(__uint_as_float(v[i * 8 + j]) + s_bias[c0 + i * 8 + j], 0.f);
              stg256(op + i * 8, __float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]),
                     __float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7]));
            }
          }
        }
      } else {
        float sc = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < p.A; c0 += 32) {
          uint32_t v[32], u[32];
          tmem_ld32(t0 + (uint32_t)c0, v);
          if (p.gated) tmem_ld32(t0 + (uint32_t)(p.A + c0), u);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float t = tanh_fast(__uint_as_float(v[i]) + s_bias[c0 + i]);
            if (p.gated) t *= sigmoid_fast(__uint_as_float(u[i]) + s_bias[p.A + c0 + i]);
            sc = fmaf(s_bias[BN + c0 + i], t, sc);
          }
        }
        if (m < p.M) p.out[m] = sc + __ldg(p.b_w);
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_accempty + 8 * acc) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// [rows, cols] row-major f32 matrix, box = [box_rows x 32 cols] (one 128-byte swizzle row of K per matrix row)
static int encode_2d_f32(TensorMapBlob* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    PDF_CHECK_CUDA(cudaFree(0));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PDF_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    PDF_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kTfK, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box,
                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PDF_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(f32) failed (%d) rows=%llu cols=%llu box_rows=%u", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, box_rows);
  return PDF_OK;
}

bool gemm_tf32_supported(int N, int K) { return (N == 64 || N == 128 || N == 256) && K >= kTfK && K % kTfK == 0; }

static int g_tf32_mt = 0;      // pdf_debug_set_mil_mt: 0 = by batch size, 1 / 2 = forced

template <int BN, int MT>
static int launch_tf32(const TensorMapBlob& ta, const TensorMapBlob& tb, const TfParams& p, cudaStream_t s) {
  constexpr int kStages = (MT == 1) ? 4 : 3;
  constexpr int smem = kStages * (MT * kABytes + BN * 128) + 8 * (2 * kStages + 4) + 16 + 2 * BN * 4 + 1024;
  static bool configured = false;
  if (!configured) {
    PDF_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<BN, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int grid = max(1, min(ceil_div(p.M, MT * kBlockM), num_sms()));
  PDF_CHECK_CUDA(launch_pdl(gemm_tf32_kernel<BN, MT>, dim3(grid), dim3(192), (size_t)smem, s, *reinterpret_cast<const CUtensorMap*>(&ta),
                            *reinterpret_cast<const CUtensorMap*>(&tb), p));
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

// mode 0: out[M,N] = relu(A B^T + bias); mode 1: out[M] = attention score (N = 2*A_dim gated | A_dim)
int launch_gemm_tf32(const float* A, const float* B, int M, int N, int K, int mode, int gated, int A_dim, const float* bias,
                     const float* w_w, const float* b_w, float* out, cudaStream_t s) {
  PDF_REQUIRE(gemm_tf32_supported(N, K), "tf32 GEMM: N must be 64, 128 or 256 and K a multiple of 32 (N=%d K=%d)", N, K);
  PDF_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(out) & 31) == 0, "tf32 GEMM: pointers must be aligned");
  TensorMapBlob ta, tb;
  if (int rc = encode_2d_f32(&ta, A, (uint64_t)M, (uint64_t)K, kBlockM)) return rc;
  if (int rc = encode_2d_f32(&tb, B, (uint64_t)N, (uint64_t)K, (uint32_t)N)) return rc;
  TfParams p;
  p.M = M; p.N = N; p.K = K; p.mode = mode; p.gated = gated; p.A = A_dim; p.bias = bias; p.w_w = w_w; p.b_w = b_w; p.out = out;
  // measured (profiles/r02_mil_times.txt, 2048 bags): 256-row tiles are NOT faster (253 vs 248 us per sweep) -- what they save in
  // weight re-streaming they lose with the single TMEM accumulator set and the shallower ring; they stay an opt-in
  const bool two = g_tf32_mt == 2;
  switch (N) {
    case 64: return two ? launch_tf32<64, 2>(ta, tb, p, s) : launch_tf32<64, 1>(ta, tb, p, s);
    case 128: return two ? launch_tf32<128, 2>(ta, tb, p, s) : launch_tf32<128, 1>(ta, tb, p, s);
    default: return two ? launch_tf32<256, 2>(ta, tb, p, s) : launch_tf32<256, 1>(ta, tb, p, s);
  }
}

}  // namespace pdf

/* tuning / A-B hook: row tiles per weight pass of the MIL tf32 GEMMs (0 = chosen from the batch size, 1, 2) */
extern "C" int pdf_debug_set_mil_mt(int mt) {
  if (mt < 0 || mt > 2) { pdf::set_error("pdf_debug_set_mil_mt: 0, 1 or 2"); return PDF_ERR_ARG; }
  pdf::g_tf32_mt = mt;
  return PDF_OK;
}
