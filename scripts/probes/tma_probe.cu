// Stand-alone probe: TMA tiled load of a [box_rows x box_cols] bf16 box from a [rows, pitch] tensor, no swizzle.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, int c0, int c1, int bytes, uint16_t* out, int dyn_off) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem) + dyn_off;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(c0), "r"(c1) : "memory");
    uint32_t ok = 0; int spins = 0;
    while (!ok && spins++ < (1 << 22))
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b) : "memory");
    out[0] = ok ? 1 : 0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[1 + i] = reinterpret_cast<uint16_t*>(smem + dyn_off)[i];
}
int main() {
  cudaFree(0);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  struct Cfg { int pitch, rows, bc, br, c0, c1, off; } cfgs[] = {
    {232, 693, 40, 39, 0, 0, 0}, {232, 693, 40, 39, 28, 32, 0}, {232, 693, 64, 32, 0, 0, 0}, {256, 693, 64, 64, 0, 0, 0},
    {232, 693, 40, 39, 0, 0, 131072}, {232, 693, 40, 39, 28, 32, 131072 + 3200}, {128, 128, 64, 64, 0, 0, 0}};
  for (auto& c : cfgs) {
    size_t n = (size_t)c.pitch * c.rows;
    std::vector<uint16_t> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (uint16_t)(i % 65521);
    uint16_t *d, *o; cudaMalloc(&d, n * 2); cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice);
    int bytes = c.bc * c.br * 2;
    cudaMalloc(&o, bytes + 16); cudaMemset(o, 0, bytes + 16);
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)c.pitch, (cuuint64_t)c.rows}; cuuint64_t str[1] = {(cuuint64_t)c.pitch * 2};
    cuuint32_t box[2] = {(cuuint32_t)c.bc, (cuuint32_t)c.br}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int smem = c.off + bytes + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<<<1, 128, smem>>>(tm, c.c0, c.c1, bytes, o, c.off);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint16_t> ho(bytes / 2 + 1); cudaMemcpy(ho.data(), o, bytes + 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < c.br && e == cudaSuccess; ++y) for (int x = 0; x < c.bc; ++x) {
      size_t gi = (size_t)(c.c1 + y) * c.pitch + c.c0 + x;
      uint16_t want = (c.c0 + x < c.pitch && c.c1 + y < c.rows) ? h[gi] : 0;
      if (ho[1 + y * c.bc + x] != want) ++bad;
    }
    printf("pitch %d rows %d box %dx%d at (%d,%d) smem_off %d: encode %d run %s done %d mismatches %d\n", c.pitch, c.rows, c.bc, c.br, c.c0, c.c1,
           c.off, (int)r, cudaGetErrorString(e), (int)ho[0], bad);
    if (e != cudaSuccess) { cudaDeviceReset(); cudaFree(0); }
  }
  return 0;
}
