// f4 ("simple" feature mode): intensity statistics of a resampled volume, one thread block per subject.
//
// Replaces `_compute_simple_features` (data/openneuro_features.py:34-73 of the reference): over vals = volume[volume > 0] (all
// voxels when none is positive) -- mean, std, min, max, np.median, np.percentile 10 / 90 / 1 / 99, a 10-bin np.histogram of the
// values clipped to [p1, p99], and (extra_stats) the central moments behind scipy.stats.skew / kurtosis.  The 8^3 grid means of
// the same function are a second trilinear zoom and run through K1a's resample kernel (preproc.cu), bit-exact like `_load_volume`.
//
// Order statistics are EXACT: a three-level radix select (12 + 10 + 10 bits of an order-preserving key of the float bits) for
// all 10 ranks at once -- the two neighbours of each numpy 'linear' virtual index, which numpy 2.x evaluates in float32
// (SURVEY.md A.2), and np.median's one or two middle ranks -- followed by numpy's `_lerp` in float32 round-to-nearest intrinsics.
// The histogram counts follow np.histogram's uniform-bin rule: float32 edges built as np.linspace builds them (arange * step +
// start, last edge = stop), a value belongs to bin i when edges[i] <= v < edges[i+1], the last bin closed.  Sums run in float64
// (numpy's pairwise float32 sums are reproduced to ~1e-7 relative, not bit for bit: the tests state that tolerance).
//
// The volume (96^3 = 3.5 MB by default) is read five times by its block and stays in L2; throughput is set by K1a's resample.
#include "common.cuh"

namespace pdf {

constexpr int kSfThreads = 1024;
constexpr int kSfQueries = 10;       // median pair, p10, p90, p1, p99 neighbour pairs
constexpr int kSfMaxBins = 64;

__device__ __forceinline__ uint32_t sf_key(float v) {            // order-preserving: a < b  <=>  key(a) < key(b)
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float sf_unkey(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ double sf_block_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < kSfThreads / 32; ++i) t += red[i];
  return t;
}

// numpy 'linear' percentile of n sorted values for q32 = float32(q) / float32(100): ranks of the two neighbours and gamma
__device__ void sf_virtual(unsigned n, float q32, unsigned* r0, unsigned* r1, float* gamma) {
  const float virt = __fmul_rn((float)(n - 1), q32);
  const float prev = floorf(virt);
  *gamma = __fsub_rn(virt, prev);
  if (virt >= (float)(n - 1)) { *r0 = *r1 = n - 1; }
  else if (virt < 0.f) { *r0 = *r1 = 0; }
  else { *r0 = (unsigned)prev; *r1 = *r0 + 1; }
}
__device__ float sf_lerp(float a, float b, float g) {            // numpy _lerp in float32
  const float d = __fsub_rn(b, a);
  if (g >= 0.5f) return __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.0f, g)));
  return __fadd_rn(a, __fmul_rn(d, g));
}

// out (doubles per subject, stride kSfOutStride): 0 n | 1 sum | 2 min | 3 max | 4 m2 | 5 m3 | 6 m4 (sums of (v-mean)^k) | 7 median |
// 8 p10 | 9 p90 | 10 p1 | 11 p99 | 12 all-voxel fallback flag | 16.. counts[bins] | 16+bins.. edges[bins+1]
constexpr int kSfOutStride = 16 + 2 * kSfMaxBins + 1;

__global__ void __launch_bounds__(kSfThreads, 1)
simple_stats_kernel(const float* __restrict__ vols, size_t voxels, int bins, double* __restrict__ out_all) {
  __shared__ uint32_t hist[kSfQueries * 1024];                   // level 0 uses the first 4096 entries
  __shared__ double red[kSfThreads / 32];
  __shared__ unsigned s_rank[kSfQueries], s_prefix[kSfQueries];   // residual rank / key prefix found so far
  __shared__ float s_gamma[5];
  __shared__ unsigned s_scan[kSfThreads];
  __shared__ float s_edges[kSfMaxBins + 1];
  __shared__ unsigned s_cnt[kSfMaxBins];
  __shared__ double s_mean;
  __shared__ float s_lo, s_hi;
  const float* vol = vols + (size_t)blockIdx.x * voxels;
  double* out = out_all + (size_t)blockIdx.x * kSfOutStride;
  const int tid = threadIdx.x;

  // pass A: how many positive voxels?
  double cnt = 0.0;
  for (size_t i = tid; i < voxels; i += kSfThreads) cnt += vol[i] > 0.f ? 1.0 : 0.0;
  const double n_pos = sf_block_sum(cnt, red);
  const bool use_all = n_pos == 0.0;                             // the reference falls back to every voxel
  const unsigned n = use_all ? (unsigned)voxels : (unsigned)n_pos;

  // sum / min / max
  double s = 0.0;
  float mn = INFINITY, mx = -INFINITY;
  for (size_t i = tid; i < voxels; i += kSfThreads) {
    const float v = vol[i];
    if (use_all || v > 0.f) { s += (double)v; mn = fminf(mn, v); mx = fmaxf(mx, v); }
  }
  const double sum = sf_block_sum(s, red);
  const double dmn = -sf_block_sum(0.0, red) ;                    // (keeps the barrier pattern uniform)
  (void)dmn;
  for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  __syncthreads();
  if ((tid & 31) == 0) { red[tid >> 5] = (double)mn; }
  __syncthreads();
  float gmn = INFINITY;
  for (int i = 0; i < kSfThreads / 32; ++i) gmn = fminf(gmn, (float)red[i]);
  __syncthreads();
  if ((tid & 31) == 0) { red[tid >> 5] = (double)mx; }
  __syncthreads();
  float gmx = -INFINITY;
  for (int i = 0; i < kSfThreads / 32; ++i) gmx = fmaxf(gmx, (float)red[i]);
  __syncthreads();

  if (tid == 0) {
    // ranks: np.median's middle pair, then the neighbours of the virtual indices of q = 10, 90, 1, 99
    s_rank[0] = (n % 2 == 0) ? n / 2 - 1 : n / 2;
    s_rank[1] = n / 2;
    const float qs[4] = {__fdiv_rn(10.0f, 100.0f), __fdiv_rn(90.0f, 100.0f), __fdiv_rn(1.0f, 100.0f), __fdiv_rn(99.0f, 100.0f)};
    for (int j = 0; j < 4; ++j) sf_virtual(n, qs[j], &s_rank[2 + 2 * j], &s_rank[3 + 2 * j], &s_gamma[1 + j]);
    for (int q = 0; q < kSfQueries; ++q) s_prefix[q] = 0;
    s_mean = sum / (double)n;
  }

  // three radix levels: 12, 10, 10 bits
  for (int level = 0; level < 3; ++level) {
    const int shift = level == 0 ? 20 : (level == 1 ? 10 : 0);
    const int nb = level == 0 ? 4096 : 1024;
    const int rows = level == 0 ? 1 : kSfQueries;
    for (int i = tid; i < rows * nb; i += kSfThreads) hist[i] = 0;
    __syncthreads();
    unsigned pre[kSfQueries];
    if (level > 0)
      for (int q = 0; q < kSfQueries; ++q) pre[q] = s_prefix[q];
    for (size_t i = tid; i < voxels; i += kSfThreads) {
      const float v = vol[i];
      if (!(use_all || v > 0.f)) continue;
      const uint32_t k = sf_key(v);
      if (level == 0) atomicAdd(&hist[k >> 20], 1u);
      else {
        const uint32_t hi_bits = k >> (shift + 10);
        const uint32_t b = (k >> shift) & 1023u;
        for (int q = 0; q < kSfQueries; ++q)
          if (hi_bits == pre[q]) atomicAdd(&hist[q * 1024 + b], 1u);
      }
    }
    __syncthreads();
    // per row: find the bin that holds the residual rank (block-wide scan over nb bins, nb / 1024 bins per thread)
    for (int row = 0; row < rows; ++row) {
      const uint32_t* h = hist + row * nb;
      const int per = nb / kSfThreads;                            // 4 or 1
      unsigned local = 0;
      for (int j = 0; j < per; ++j) local += h[tid * per + j];
      s_scan[tid] = local;
      __syncthreads();
      for (int o = 1; o < kSfThreads; o <<= 1) {                  // inclusive Hillis-Steele scan
        const unsigned add = tid >= o ? s_scan[tid - o] : 0;
        __syncthreads();
        s_scan[tid] += add;
        __syncthreads();
      }
      const unsigned before = s_scan[tid] - local;                // elements in bins below this thread's
      const int q_lo = level == 0 ? 0 : row, q_hi = level == 0 ? kSfQueries : row + 1;
      for (int q = q_lo; q < q_hi; ++q) {
        const unsigned r = s_rank[q];
        if (r >= before && r < before + local) {                  // exactly one thread per query
          unsigned acc = before;
          for (int j = 0; j < per; ++j) {
            const unsigned c = h[tid * per + j];
            if (r < acc + c) {
              s_rank[q] = r - acc;
              s_prefix[q] = (level == 0 ? 0u : (s_prefix[q] << 10)) | (unsigned)(tid * per + j);
              break;
            }
            acc += c;
          }
        }
      }
      __syncthreads();
    }
  }

  if (tid == 0) {
    float os[kSfQueries];
    for (int q = 0; q < kSfQueries; ++q) os[q] = sf_unkey(s_prefix[q]);
    const float median = (n % 2 == 0) ? __fmul_rn(__fadd_rn(os[0], os[1]), 0.5f) : os[1];     // np.mean of the middle pair in float32
    const float p10 = sf_lerp(os[2], os[3], s_gamma[1]), p90 = sf_lerp(os[4], os[5], s_gamma[2]);
    float lo = sf_lerp(os[6], os[7], s_gamma[3]), hi = sf_lerp(os[8], os[9], s_gamma[4]);
    out[0] = (double)n; out[1] = sum; out[2] = (double)gmn; out[3] = (double)gmx;
    out[7] = (double)median; out[8] = (double)p10; out[9] = (double)p90; out[10] = (double)lo; out[11] = (double)hi;
    out[12] = use_all ? 1.0 : 0.0;
    s_lo = lo; s_hi = hi;
    // np.histogram(range=(lo, hi)): equal outer edges are widened by 0.5; float32 edges as np.linspace builds them
    float first = lo, last = hi;
    if (first == last) { first = __fsub_rn(first, 0.5f); last = __fadd_rn(last, 0.5f); }
    const float step = __fdiv_rn(__fsub_rn(last, first), (float)bins);
    for (int i = 0; i < bins; ++i) s_edges[i] = __fadd_rn(__fmul_rn((float)i, step), first);
    s_edges[bins] = last;
    for (int i = 0; i <= bins; ++i) out[16 + bins + i] = (double)s_edges[i];
  }
  if (tid < kSfMaxBins) s_cnt[tid] = 0;
  __syncthreads();

  // pass E: central moments and the histogram of the clipped values
  const double mu = s_mean;
  const float lo = s_lo, hi = s_hi;
  double m2 = 0.0, m3 = 0.0, m4 = 0.0;
  for (size_t i = tid; i < voxels; i += kSfThreads) {
    const float v = vol[i];
    if (!(use_all || v > 0.f)) continue;
    const double d = (double)v - mu, d2 = d * d;
    m2 += d2; m3 += d2 * d; m4 += d2 * d2;
    const float c = fminf(fmaxf(v, lo), hi);
    if (c >= s_edges[0] && c <= s_edges[bins]) {
      int b = 0;
      while (b < bins - 1 && c >= s_edges[b + 1]) ++b;
      atomicAdd(&s_cnt[b], 1u);
    }
  }
  const double t2 = sf_block_sum(m2, red), t3 = sf_block_sum(m3, red), t4 = sf_block_sum(m4, red);
  __syncthreads();
  if (tid == 0) { out[4] = t2; out[5] = t3; out[6] = t4; }
  if (tid < bins) out[16 + tid] = (double)s_cnt[tid];
}

}  // namespace pdf

using namespace pdf;

/* Intensity statistics of `batch` volumes [batch, voxels] f32 for the "simple" feature mode (reference `_compute_simple_features`,
 * data/openneuro_features.py:34-73).  d_out: batch * pdf_simple_stats_stride() doubles, per subject
 * 0 n | 1 sum | 2 min | 3 max | 4..6 sums of (v-mean)^2,3,4 | 7 median | 8 p10 | 9 p90 | 10 p1 | 11 p99 | 12 all-voxel fallback |
 * 16.. counts[hist_bins] | 16+hist_bins.. edges[hist_bins+1]. */
extern "C" int pdf_simple_stats_stride(void) { return kSfOutStride; }

extern "C" int pdf_simple_stats(int batch, size_t voxels, int hist_bins, const float* d_vol, double* d_out, pdf_stream_t stream) {
  PDF_REQUIRE(batch > 0 && voxels > 0 && voxels < (1ull << 31) && hist_bins >= 1 && hist_bins <= kSfMaxBins && d_vol && d_out,
              "pdf_simple_stats: bad arguments (1 <= hist_bins <= %d)", kSfMaxBins);
  simple_stats_kernel<<<batch, kSfThreads, 0, as_stream(stream)>>>(d_vol, voxels, hist_bins, d_out);
  PDF_CHECK_LAUNCH();
  return PDF_OK;
}
