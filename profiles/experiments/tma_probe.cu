// tma_probe.cu -- stand-alone probe: one 3-D cp.async.bulk.tensor load of a BX x BY x BZ box out of a 32^3
// f32 tensor, descriptor in global memory or as a __grid_constant__ parameter.  Prints what lands in smem.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
__device__ __forceinline__ unsigned saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int BX, int BY, int BZ>
__global__ void probe(const void* tmap_g, const __grid_constant__ CUtensorMap tmap_p, int use_param, int x, int y, int z,
                      float* out)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  float* buf = reinterpret_cast<float*>(smem);
  if (threadIdx.x == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(saddr(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    const void* d = use_param ? (const void*)&tmap_p : tmap_g;
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(saddr(&bar)),
                 "r"(BX * BY * BZ * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(saddr(buf)), "l"(d), "r"(x), "r"(y), "r"(z), "r"(saddr(&bar)) : "memory");
  }
  unsigned ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(saddr(&bar)) : "memory");
  for (int i = threadIdx.x; i < BX * BY * BZ; i += blockDim.x) out[i] = buf[i];
  if (threadIdx.x == 0) out[BX * BY * BZ] = (float)(saddr(buf) & 1023u);
}
static int g_x = 3, g_y = 5;
template <int BX, int BY, int BZ>
int run(int use_param, int cz = -1)
{
  const int N = 32;
  std::vector<float> h(N * N * N);
  for (int i = 0; i < N * N * N; ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&o, BX * BY * BZ * 4 + 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  typedef CUresult (*fn_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  fn_t fn = (fn_t)sym;
  alignas(64) CUtensorMap m;
  cuuint64_t dims[3] = { N, N, N }, strides[2] = { N * 4, N * N * 4 };
  cuuint32_t box[3] = { BX, BY, BZ }, es[3] = { 1, 1, 1 };
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("x=%d y=%d box %dx%dx%d param=%d cz=%d q=%d fn=%p encode=%d ", g_x, g_y, BX, BY, BZ, use_param, cz, (int)q, sym, (int)r);
  { const unsigned long long* w = (const unsigned long long*)&m; printf("desc[0..3]=%llx %llx %llx %llx ", w[0], w[1], w[2], w[3]); }
  void* mg;
  cudaMalloc(&mg, 128);
  cudaMemcpy(mg, &m, 128, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe<BX, BY, BZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, BX * BY * BZ * 4);
  probe<BX, BY, BZ><<<1, 128, BX * BY * BZ * 4>>>(mg, m, use_param, g_x, g_y, cz, o);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run=%s ", cudaGetErrorString(e));
  if (e == cudaSuccess)
  {
    std::vector<float> g(BX * BY * BZ);
    cudaMemcpy(g.data(), o, g.size() * 4, cudaMemcpyDeviceToHost);
    // expect element (x=3+i, y=5+j, z=-1+k): z=-1 plane zero-filled
    printf("g[0]=%g (want 0) g[plane]=%g (want %g) g[plane+row+1]=%g (want %g)", g[0], g[BX * BY], (float)(5 * N + 3),
           g[BX * BY + BX + 1], (float)(6 * N + 4));
  }
  printf("\n");
  return e != cudaSuccess;
}
int main(int argc, char** argv)
{
  int which = argc > 1 ? atoi(argv[1]) : 0;
  if (argc > 4) { g_x = atoi(argv[2]); g_y = atoi(argv[3]); return run<12, 12, 12>(0, atoi(argv[4])); }
  if (which == 0) return run<12, 12, 12>(0);
  if (which == 1) return run<12, 12, 12>(1);
  if (which == 2) return run<16, 12, 12>(0);
  if (which == 3) return run<16, 16, 8>(0);
  if (which == 4) return run<8, 8, 8>(0);
  if (which == 5) return run<8, 8, 8>(0, 0);
  if (which == 6) return run<16, 16, 8>(1, 0);
  return 0;
}
