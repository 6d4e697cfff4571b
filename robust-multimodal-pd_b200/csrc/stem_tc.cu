// K2 stem, fused: 7x7 stride-2 pad-3 convolution (one folded input channel) + bias + ReLU + 3x3 stride-2 pad-1
// max-pool in ONE persistent, warp-specialised kernel.
//
// The GEMM is TRANSPOSED so that the max-pool needs no shared-memory round trip:
//
//   D[lane = (variant v, channel c), column = site n]  =  Wa[128 x 96]  *  B[192 x 96]^T       (tcgen05, f32 in TMEM)
//
//   * a tile is 8 x 7 pooled pixels  <-  17 x 15 conv pixels  <-  a 39 x 35 input patch;
//   * site n = ((i*3+u)*16 + cx), i = 0..3, u = 0..2, cx = 0..14: its B row holds the 11 patch rows (8 pixels each:
//     K = t*8 + s, t = 0..10, s = 0..7) that cover conv row 4i+u (variant 0, filter rows at t = 0..6) AND conv row
//     4i+u+2 (variant 1, the same filter shifted down by 4 patch rows, t = 4..10) of conv column cx;
//   * lane v*64+c therefore owns, as TMEM COLUMNS, every conv pixel of channel c that the pooled rows 2i+v need:
//     the 3x3/2 max-pool, ReLU and the bf16 conversion are plain register arithmetic on tcgen05.ld results, and all 128 lanes
//     do useful work.  The bias is added by the tensor core: K slots 88/89 of every B row hold 1, the weights carry hi + lo terms.
//
//   warps 0..E-1 epilogue (E = kEpiWarps, 4 or 8): tcgen05.ld -> 3x3 max -> ReLU -> bf16 NHWC store.  With E = 8 two warps share
//                a TMEM lane quarter (warp w and w+4 both own lanes 32*(w%4)..+31) and take half of the accumulator's columns each
//   warp  E      TMA load of the weights, tcgen05.mma issue (6 x M128 N192 K16 per tile), TMEM alloc
//   next 10      build the B matrix (im2col) in shared memory, 128B-swizzled K-major, from the patch; a 3-stage ring
//                against the MMAs; two TMEM accumulators overlap epilogue and MMA
//   last warp    TMA producer: the 39 x 40 input patch of tile t+4 lands in a 4-stage ring while tile t is built
//
// The input image is read from a zero-PADDED buffer (origin at row/column 5, see pdf_stem_padded_dims): the patch
// origin is then (4*pp0, 4*pq0) >= 0 and every 8-pixel chunk of the staged patch is a 4-byte aligned shared-memory read.
// HBM traffic per image: the padded bf16 image in, P*P*64 bf16 out.
// Replaces conv1/bn1/relu/maxpool of torchvision's ResNet (`model(batch)`, data/openneuro_features.py:260).
#include "tc_common.cuh"
#include "ops.cuh"

namespace pdf {

constexpr int kPR = 8, kPQ = 7;                  // pooled tile: rows x columns
constexpr int kPatchRows = 4 * kPR + 7;          // 39
constexpr int kConvCols = 2 * kPQ + 1;           // 15
constexpr int kSites = 12 * 16;                  // 192 B rows
constexpr int kStemBBlock = kSites * 128;        // one 64-wide K block of B
constexpr int kStemBStage = 2 * kStemBBlock;
constexpr int kStemABlock = 128 * 128;
constexpr int kStemABytes = 2 * kStemABlock;
constexpr int kBuilderWarps = 10;                // 585 chunks over 320 threads: two rounds
#ifndef STEM_EPI_WARPS
#define STEM_EPI_WARPS 8
#endif
constexpr int kEpiWarps = STEM_EPI_WARPS;
constexpr int kEpiGroups = 16 / kEpiWarps;        // row groups (of 4) per epilogue warp
constexpr int kStemWarps = kEpiWarps + 1 + kBuilderWarps + 1;
constexpr int kStemBuilders = kBuilderWarps * 32;
constexpr int kBStages = 3;                      // B ring (the two TMEM accumulators bound the epilogue side)
constexpr int kPatchCols = 40;                   // 35 needed (+4 when the box start is rounded down to 16 bytes); 80-byte rows
constexpr int kPatchStages = 4;
constexpr int kPatchStage = (kPatchRows * kPatchCols * 2 + 127) / 128 * 128;
constexpr int kStemSmem = kStemABytes + kBStages * kStemBStage + kPatchStages * kPatchStage + 256 + 1024;

struct StemParams {
  const __nv_bfloat16* in;    // padded [n, rows, pitch]
  const int* border;          // NULL, or the border-correction blob of pdf_op.d_scale (see include/pdfusion_b200.h)
  __nv_bfloat16* out;         // [n, P, P, 64]
  int pitch, rows, H1, P, tiles_x, tiles_y, total_tiles;
  unsigned long long* trace;  // debug: per-role clock64 stamps of CTA 0 (pdf_debug_set_trace), or NULL
};

// stamp slot: [it][event]; events 0-2 builder, 3-5 MMA, 6-8 epilogue, 9-13 inside the epilogue of warp 0
// (compiled in only with -DPDF_STEM_TRACE: `make EXTRA=-DPDF_STEM_TRACE`; the stamps cost ~25 instructions per tile and warp)
#ifdef PDF_STEM_TRACE
#define STEM_TRACE(ev) do { if (p.trace && blockIdx.x == 0 && it < 64 && (threadIdx.x & 31) == 0) p.trace[it * 16 + (ev)] = clock64(); } while (0)
#else
#define STEM_TRACE(ev) do { } while (0)
#endif

// (image, tile row, tile column) of the tiles blockIdx.x, blockIdx.x + gridDim.x, ... without a division per tile
struct TileIter {
  int n, ty, tx, dn, dty, dtx, tiles_x, tiles_y;
  __device__ TileIter(int first, int stride, int tiles_x_, int tiles_y_) : tiles_x(tiles_x_), tiles_y(tiles_y_) {
    const int per = tiles_x * tiles_y;
    n = first / per;
    int r = first - n * per;
    ty = r / tiles_x;
    tx = r - ty * tiles_x;
    dn = stride / per;
    r = stride - dn * per;
    dty = r / tiles_x;
    dtx = r - dty * tiles_x;
  }
  __device__ __forceinline__ void next() {
    tx += dtx; ty += dty; n += dn;
    if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
    if (ty >= tiles_y) { ty -= tiles_y; ++n; }
  }
};

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
// max(x, 0) rounded to bf16 in one instruction (cvt.rn.relu)
__device__ __forceinline__ __nv_bfloat16 relu_bf16(float x) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %1;" : "=r"(r) : "f"(x));
  return __ushort_as_bfloat16((unsigned short)(r & 0xffffu));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// FIX: the input was normalised with per-channel statistics, border conv pixels get a class-dependent bias correction
// (p.border).  A separate instantiation keeps the ~3000 instructions of that path out of the common kernel.
template <bool FIX>
__global__ void __launch_bounds__(kStemWarps * 32, 1)
stem_fused_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_in, const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t sA = base, sB = base + kStemABytes;
  const uint32_t sP = sB + kBStages * kStemBStage;
  const uint32_t bar0 = sP + kPatchStages * kPatchStage;
  const uint32_t bar_w = bar0, bar_bfull = bar0 + 8, bar_bempty = bar_bfull + 8 * kBStages, bar_accfull = bar_bempty + 8 * kBStages;
  const uint32_t bar_accempty = bar_accfull + 16, bar_pfull = bar_accempty + 16, bar_pempty = bar_pfull + 8 * kPatchStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (bar_pempty + 8 * kPatchStages - base));

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  if (warp == kEpiWarps) {
    if (elect_one()) {
      prefetch_tmap(&tmap_w);
      mbar_init(bar_w, 1);
      for (int s = 0; s < kBStages; ++s) {
        mbar_init(bar_bfull + 8 * s, kBuilderWarps);   // one arrive per builder warp
        mbar_init(bar_bempty + 8 * s, 1);              // tcgen05.commit
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(bar_accfull + 8 * s, 1);             // tcgen05.commit
        mbar_init(bar_accempty + 8 * s, kEpiWarps);    // one arrive per epilogue warp
      }
      for (int s = 0; s < kPatchStages; ++s) { mbar_init(bar_pfull + 8 * s, 1); mbar_init(bar_pempty + 8 * s, kBuilderWarps); }
      fence_barrier_init();
      mbar_expect_tx(bar_w, kStemABytes);
      tma_load_2d(sA, &tmap_w, bar_w, 0, 0);
      tma_load_2d(sA + kStemABlock, &tmap_w, bar_w, 64, 0);
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 512);
  }
  // B rows of the unused column slot (cx = 15) and the K padding (t = 11) are never written by the builders: they only
  // have to be finite, so both stages are zeroed once
  for (int i = tid; i < kBStages * kStemBStage / 16; i += kStemWarps * 32)
    reinterpret_cast<uint4*>(smem + kStemABytes)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  // K slots 88 and 89 (t = 11, s = 0, 1) of every B row hold the constant 1: the weight matrix carries the bias there (split
  // into two bf16 terms), so the tensor core adds it and the epilogue has no per-pixel bias add
  for (int i = tid; i < kBStages * kSites; i += kStemWarps * 32) {
    const int stage = i / kSites, site = i - stage * kSites;
    *reinterpret_cast<uint32_t*>(smem + kStemABytes + stage * kStemBStage + kStemBBlock + site * 128 + ((3 ^ (site & 7)) << 4)) = 0x3F803F80u;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // see common.cuh: the successor's prologue overlaps this kernel's tail,
  pdl_wait();                // the image is read (and the output written) only after the predecessor has completed

  if (warp < kEpiWarps) {
    // ------------------------------------------------------------------------------------------ epilogue
    const int quarter = warp & 3, half = warp >> 2;     // TMEM lane quarter (fixed by warp % 4), column half of the accumulator
    const int v = quarter >> 1, c = (quarter & 1) * 32 + lane;
    // per-channel input statistics: conv pixels whose 7x7 window leaves the image need a class-dependent bias correction
    const int nc = FIX ? p.border[0] : 0;
    const int* cls = p.border + 2;
    const float* delta = reinterpret_cast<const float*>(p.border + 2 + p.H1) + c;
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int it = 0;
    TileIter ti(blockIdx.x, gridDim.x, p.tiles_x, p.tiles_y);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it, ti.next()) {
      const int stage = it & 1;
      const uint32_t phase = (uint32_t)(it >> 1) & 1u;
      const int n = ti.n;
      const int pp0 = ti.ty * kPR, pq0 = ti.tx * kPQ;
      const int cy0 = 2 * pp0 - 1, cx0 = 2 * pq0 - 1;
      const bool col_edge = cx0 < 0 || cx0 + kConvCols > p.H1;
      bool fix = false;
      if (FIX) fix = (cls[max(cy0, 0)] | cls[min(cy0 + 4 * kPR / 2, p.H1 - 1)] | cls[max(cx0, 0)] | cls[min(cx0 + kConvCols - 1, p.H1 - 1)]) != 0;
      const size_t orow_stride = (size_t)p.P * 64;
      __nv_bfloat16* otile = p.out + (((size_t)n * p.P + pp0 + v) * p.P + pq0) * 64 + c;
      mbar_wait(bar_accfull + 8 * stage, phase);
      if (warp == 0) STEM_TRACE(6);
      tc_fence_after();
      // tcgen05.ld moves only ~64-100 B/clk per SM (gpurun_out/stem_trace.txt): reading the 96 KB accumulator is the longest
      // stage of this kernel; issuing the loads further ahead (a register double buffer) was measured slower.
#pragma unroll
      for (int ii = 0; ii < kEpiGroups; ++ii) {
        const int i = kEpiGroups * half + ii;
        uint32_t r0[16], r1[16], r2[16];
        const uint32_t col = (uint32_t)(stage * 256 + i * 48);
        tmem_ld16_nowait(tlane + col, r0);
        tmem_ld16_nowait(tlane + col + 16, r1);
        tmem_ld16_nowait(tlane + col + 32, r2);
        tmem_wait_ld();
        if (warp == 0) STEM_TRACE(9 + ii);
        if (ii == kEpiGroups - 1) {                    // this warp's share is read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cta(bar_accempty + 8 * stage);
          if (warp == 0) STEM_TRACE(7);
        }
        const int pr = pp0 + 2 * i + v;
        if (pr >= p.P) continue;
        const int cy = cy0 + 4 * i + 2 * v;            // conv rows cy, cy+1, cy+2; the middle one (2*pr) is always valid
        const bool top = cy >= 0, bot = cy + 2 < p.H1;
        if (FIX && fix) {                              // (uniform branch: border tiles of a per-channel-normalised input only)
          auto correct = [&](uint32_t (&r)[16], int row) {
            if (row < 0 || row >= p.H1) return;
            const int rc = cls[row];
#pragma unroll
            for (int k = 0; k < kConvCols; ++k) {
              const int cx = cx0 + k;
              if (cx < 0 || cx >= p.H1) continue;
              const int cc = cls[cx];
              if (rc | cc) r[k] = __float_as_uint(__uint_as_float(r[k]) + __ldg(delta + (rc * nc + cc) * 64));
            }
          };
          correct(r0, cy); correct(r1, cy + 1); correct(r2, cy + 2);
        }
        __nv_bfloat16* orow = otile + (size_t)(2 * i) * orow_stride;
        float m[16];
        if (top && bot && !col_edge) {                 // interior rows of interior tile columns: 3-input maxima, ReLU in the convert
#pragma unroll
          for (int k = 0; k < kConvCols; ++k) m[k] = fmaxf(fmaxf(__uint_as_float(r0[k]), __uint_as_float(r1[k])), __uint_as_float(r2[k]));
#pragma unroll
          for (int j = 0; j < kPQ; ++j)
            orow[j * 64] = relu_bf16(fmaxf(fmaxf(m[2 * j], m[2 * j + 1]), m[2 * j + 2]));
        } else {
#pragma unroll
          for (int k = 0; k < kConvCols; ++k) {
            float x = __uint_as_float(r1[k]);
            if (top) x = fmaxf(x, __uint_as_float(r0[k]));
            if (bot) x = fmaxf(x, __uint_as_float(r2[k]));
            m[k] = x;
          }
          if (col_edge) {
#pragma unroll
            for (int k = 0; k < kConvCols; ++k)
              if (cx0 + k < 0 || cx0 + k >= p.H1) m[k] = -INFINITY;
          }
#pragma unroll
          for (int j = 0; j < kPQ; ++j)
            if (pq0 + j < p.P) orow[j * 64] = relu_bf16(fmaxf(fmaxf(m[2 * j], m[2 * j + 1]), m[2 * j + 2]));
        }
        if (ii == 0 && warp == 0) STEM_TRACE(13);
      }
      if (warp == 0) STEM_TRACE(8);
    }
  } else if (warp == kEpiWarps) {
    // ------------------------------------------------------------------------------------------ MMA issue
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(kSites);
      const uint32_t a_lo = smem_desc_lo(sA);
      mbar_wait(bar_w, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1, bs = it % kBStages;
        const uint32_t phase = (uint32_t)(it >> 1) & 1u, bphase = (uint32_t)(it / kBStages) & 1u;
        mbar_wait(bar_accempty + 8 * acc, phase ^ 1u);
        STEM_TRACE(3);
        mbar_wait(bar_bfull + 8 * bs, bphase);
        STEM_TRACE(4);
        tc_fence_after();
        const uint32_t b_lo = smem_desc_lo(sB + bs * kStemBStage);
        const uint32_t d = tmem_base + (uint32_t)(acc * 256);
#pragma unroll
        for (int ks = 0; ks < 6; ++ks)
          umma_f16_lo(d, a_lo + (uint32_t)((ks >> 2) * (kStemABlock / 16) + (ks & 3) * 2),
                      b_lo + (uint32_t)((ks >> 2) * (kStemBBlock / 16) + (ks & 3) * 2), idesc, ks != 0 ? 1u : 0u);
        umma_commit(bar_bempty + 8 * bs);
        umma_commit(bar_accfull + 8 * acc);
        STEM_TRACE(5);
      }
    }
  } else if (warp < kEpiWarps + 1 + kBuilderWarps) {
    // ------------------------------------------------------------------------------------------ B builders
    // The (patch row, conv column) chunks a thread copies, and the B rows each one lands in, are the same for every tile:
    // the offsets are computed once and live in registers.
    const int bt = tid - (kEpiWarps + 1) * 32;
    constexpr int kRounds = (kPatchRows * kConvCols + kStemBuilders - 1) / kStemBuilders;
    int ld_off[kRounds];
    uint32_t st_off[kRounds][6];
#pragma unroll
    for (int k = 0; k < kRounds; ++k) {
      const int idx = bt + k * kStemBuilders;
      const bool have = idx < kPatchRows * kConvCols;
      const int y = idx / kConvCols, cx = idx - y * kConvCols;
      ld_off[k] = have ? y * (kPatchCols * 2) + cx * 4 : -1;
      const int q0 = y >> 1, tpar = y & 1;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int q = q0 - j, t = tpar + 2 * j;      // conv row q (sites exist for q % 4 != 3) sees patch row y as its t-th row
        const bool ok = have && q >= 0 && q <= 14 && (q & 3) != 3 && t <= 10;
        const int site = ((q >> 2) * 3 + (q & 3)) * 16 + cx;
        st_off[k][j] = ok ? (uint32_t)((t >> 3) * kStemBBlock + site * 128 + (((t & 7) ^ (cx & 7)) << 4)) : 0xffffffffu;
      }
    }
    int it = 0;
    TileIter ti(blockIdx.x, gridDim.x, p.tiles_x, p.tiles_y);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it, ti.next()) {
      const int stage = it % kBStages, ps = it % kPatchStages;
      const uint32_t phase = (uint32_t)(it / kBStages) & 1u, pphase = (uint32_t)(it / kPatchStages) & 1u;
      const uint32_t pbase = sP + (uint32_t)ps * kPatchStage + (uint32_t)(((4 * ti.tx * kPQ) & 7) * 2);   // see the TMA producer
      const uint32_t bbase = sB + (uint32_t)stage * kStemBStage;
      mbar_wait(bar_pfull + 8 * ps, pphase);
      if (warp == kEpiWarps + 1) STEM_TRACE(0);
      mbar_wait(bar_bempty + 8 * stage, phase ^ 1u);
      if (warp == kEpiWarps + 1) STEM_TRACE(1);
#pragma unroll
      for (int k = 0; k < kRounds; ++k) {
        if (ld_off[k] >= 0) {
          const uint32_t a = pbase + (uint32_t)ld_off[k];
          uint32_t v0, v1, v2, v3;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v0) : "r"(a));
          asm volatile("ld.shared.b32 %0, [%1+4];" : "=r"(v1) : "r"(a));
          asm volatile("ld.shared.b32 %0, [%1+8];" : "=r"(v2) : "r"(a));
          asm volatile("ld.shared.b32 %0, [%1+12];" : "=r"(v3) : "r"(a));
#pragma unroll
          for (int j = 0; j < 6; ++j)
            if (st_off[k][j] != 0xffffffffu)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(bbase + st_off[k][j]), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
        }
      }
      fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cta(bar_bfull + 8 * stage);
        mbar_arrive_cta(bar_pempty + 8 * ps);
      }
      if (warp == kEpiWarps + 1) STEM_TRACE(2);
    }
  } else {
    // ------------------------------------------------------------------------------------------ patch TMA producer
    if (elect_one()) {
      prefetch_tmap(&tmap_in);
      int it = 0;
      TileIter ti(blockIdx.x, gridDim.x, p.tiles_x, p.tiles_y);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it, ti.next()) {
        const int ps = it % kPatchStages;
        const uint32_t pphase = (uint32_t)(it / kPatchStages) & 1u;
        const int n = ti.n, ty = ti.ty, tx = ti.tx;
        mbar_wait(bar_pempty + 8 * ps, pphase ^ 1u);
        mbar_expect_tx(bar_pfull + 8 * ps, kPatchRows * kPatchCols * 2);
        // padded coordinates of the patch origin: image row 4*pp0-5 -> padded row 4*pp0, likewise for columns.  A TMA box must
        // start on a 16-byte boundary of the innermost dimension (profiles/r01_tma_probe.txt): the column is rounded down to
        // a multiple of 8 pixels and the builders skip the 0 or 4 extra pixels
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(sP + (uint32_t)ps * kPatchStage), "l"(reinterpret_cast<uint64_t>(&tmap_in)), "r"(bar_pfull + 8 * ps),
                       "r"((4 * tx * kPQ) & ~7), "r"(4 * ty * kPR), "r"(n) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) tmem_dealloc(tmem_base, 512);
}

}  // namespace pdf

/* geometry of the zero-padded one-channel stem input for an S x S image (include/pdfusion_b200.h) */
extern "C" int pdf_stem_padded_dims(int S, int* pitch, int* rows) {
  using namespace pdf;
  PDF_REQUIRE(S >= 8 && pitch && rows, "pdf_stem_padded_dims: bad arguments");
  const int H1 = (S + 6 - 7) / 2 + 1, P = (H1 + 2 - 3) / 2 + 1;
  const int pq0max = kPQ * (ceil_div(P, kPQ) - 1), pp0max = kPR * (ceil_div(P, kPR) - 1);
  const int cols = max(S + PDF_STEM_PAD_LO, 4 * pq0max + 2 * (kConvCols - 1) + 8);
  *pitch = (cols + 7) / 8 * 8;
  *rows = max(S + PDF_STEM_PAD_LO, 4 * pp0max + kPatchRows);
  return PDF_OK;
}

static unsigned long long* g_stem_trace = nullptr;
/* debug hook: CTA 0 of the next fused-stem launches writes clock64 stamps [64 tiles][16 events] into d_buf (NULL = off) */
extern "C" int pdf_debug_set_trace(unsigned long long* d_buf) {
  g_stem_trace = d_buf;
  return PDF_OK;
}

namespace pdf {

int launch_stem_fused(const pdf_op& op, const TensorMapBlob& tmap_w, const TensorMapBlob& tmap_in, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    PDF_CHECK_CUDA(cudaFuncSetAttribute(stem_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemSmem));
    PDF_CHECK_CUDA(cudaFuncSetAttribute(stem_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemSmem));
    configured = true;
  }
  StemParams p;
  p.in = reinterpret_cast<const __nv_bfloat16*>(op.d_in);
  p.border = reinterpret_cast<const int*>(op.d_scale);
  p.out = reinterpret_cast<__nv_bfloat16*>(op.d_out);
  if (int rc = pdf_stem_padded_dims(op.h, &p.pitch, &p.rows)) return rc;
  p.H1 = (op.h + 6 - 7) / 2 + 1;
  p.P = op.ho;
  p.tiles_x = ceil_div(p.P, kPQ);
  p.tiles_y = ceil_div(p.P, kPR);
  p.total_tiles = p.tiles_x * p.tiles_y * op.n;
  p.trace = g_stem_trace;
  const int grid = max(1, min(p.total_tiles, num_sms()));
  const CUtensorMap& tw = *reinterpret_cast<const CUtensorMap*>(&tmap_w);
  const CUtensorMap& ti = *reinterpret_cast<const CUtensorMap*>(&tmap_in);
  if (p.border) PDF_CHECK_CUDA(launch_pdl(stem_fused_kernel<true>, dim3(grid), dim3(kStemWarps * 32), (size_t)kStemSmem, s, tw, ti, p));
  else PDF_CHECK_CUDA(launch_pdl(stem_fused_kernel<false>, dim3(grid), dim3(kStemWarps * 32), (size_t)kStemSmem, s, tw, ti, p));
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}

}  // namespace pdf
