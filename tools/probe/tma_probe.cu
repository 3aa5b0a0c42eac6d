// TMA probe (development aid, not product code): what do cp.async.bulk.tensor 2D tile loads and tile::gather4 loads
// with CU_TENSOR_MAP_SWIZZLE_64B put where in shared memory?  Rows of 8 doubles; row r word k holds r*8+k.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tg, const __grid_constant__ CUtensorMap tt, double *out, int4 rows, int trow, unsigned *info)
{
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  double *s = reinterpret_cast<double *>(smem);
  if (threadIdx.x == 0) {
    info[0] = s32(smem);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) s[i] = -1.0;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(256u + 1024u) : "memory");
    // gather4: rows.x..w -> smem bytes [1024, 1280)
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(s32(smem + 1024 + 256)), "l"(&tg), "r"(0), "r"(rows.x), "r"(rows.y), "r"(rows.z), "r"(rows.w), "r"(s32(&bar)) : "memory");
    // tile load: 16 rows from trow -> smem bytes [0, 1024)
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(s32(smem)), "l"(&tt), "r"(0), "r"(trow), "r"(s32(&bar)) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(s32(&bar)), "r"(0u) : "memory");
  for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = s[i];
}
typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main()
{
  const int R = 4096;
  std::vector<double> h((size_t)R * 8);
  for (int r = 0; r < R; r++) for (int k = 0; k < 8; k++) h[(size_t)r * 8 + k] = r * 8 + k;
  double *d, *out; unsigned *info;
  CK(cudaMalloc(&d, h.size() * 8)); CK(cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&out, 512 * 8)); CK(cudaMalloc(&info, 64));
  void *fp = nullptr; cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
  if (!fp) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  encode_fn enc = (encode_fn)fp;
  CUtensorMap tg, tt;
  cuuint64_t gdim[2] = {8, (cuuint64_t)R}, gstr[1] = {64};
  cuuint32_t box_g[2] = {8, 1}, box_t[2] = {8, 16}, es[2] = {1, 1};
  for (int sw = 0; sw < 2; sw++) {
    CUtensorMapSwizzle mode = sw ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r1 = enc(&tg, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, gdim, gstr, box_g, es, CU_TENSOR_MAP_INTERLEAVE_NONE, mode, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&tt, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, gdim, gstr, box_t, es, CU_TENSOR_MAP_INTERLEAVE_NONE, mode, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("swizzle %d: encode gather map rc=%d, tile map rc=%d\n", sw, (int)r1, (int)r2);
    if (r1 || r2) continue;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192));
    probe<<<1, 128, 4096>>>(tg, tt, out, make_int4(100, 7, 2049, 33), 160, info);
    CK(cudaDeviceSynchronize());
    std::vector<double> o(512); unsigned hi[4];
    CK(cudaMemcpy(o.data(), out, 512 * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hi, info, 16, cudaMemcpyDeviceToHost));
    printf("dynamic smem base 0x%x\n", hi[0]);
    printf("tile rows 160..175 -> smem rows 0..15 (each entry: source row:word)\n");
    for (int r = 0; r < 16; r++) { printf("  smem row %2d:", r); for (int k = 0; k < 8; k++) { long v = (long)o[r * 8 + k]; printf(" %ld:%ld", v / 8, v % 8); } printf("\n"); }
    printf("gather4 rows {100,7,2049,33} -> smem bytes 1280.. (rows 20..23)\n");
    for (int r = 16; r < 26; r++) { printf("  smem row %2d:", r); for (int k = 0; k < 8; k++) { long v = (long)o[r * 8 + k]; printf(" %ld:%ld", v / 8, v % 8); } printf("\n"); }
  }
  return 0;
}
