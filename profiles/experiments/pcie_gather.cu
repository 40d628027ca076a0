// pcie_gather.cu -- how fast can a kernel pull a SPARSE subset of 128-byte lines out of mapped pinned host
// memory, against cudaMemcpyAsync of the whole buffer?  Decides whether a ray-guided sparse publish
// (mark touched lines, fetch only those) can beat the dense upload of a 512^3 f32 block.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pcie_gather pcie_gather.cu && ./pcie_gather
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// one warp per group of 32 lines; lane i owns the flag of line base+i; for every wanted line the whole warp
// copies its 128 bytes (one 4-byte element per lane), UNROLL lines in flight per warp
template <int UNROLL>
__global__ void fetch_lines(const unsigned char* __restrict__ want, const float* __restrict__ host,
                            float* __restrict__ dev, size_t n_lines)
{
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t base = warp * 32; base < n_lines; base += n_warps * 32)
  {
    const bool w = (base + lane < n_lines) && want[base + lane];
    unsigned m = __ballot_sync(0xffffffffu, w);
    while (m)
    {
      float v[UNROLL];
      size_t at[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
      {
        at[u] = ~(size_t)0;
        if (m)
        {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          at[u] = (base + b) * 32 + lane;
          v[u] = __ldcs(host + at[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (at[u] != ~(size_t)0) dev[at[u]] = v[u];
    }
  }
}

// 16 bytes per lane: a warp moves 4 consecutive lines (512 B) per instruction when all four are wanted
__global__ void fetch_lines_v4(const unsigned char* __restrict__ want, const float4* __restrict__ host,
                               float4* __restrict__ dev, size_t n_lines)
{
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  // a warp owns 128 lines (16 KiB) per iteration: 32 quads of 4 lines, lane q*? -> 8 lanes per line
  for (size_t base = warp * 128; base < n_lines; base += n_warps * 128)
  {
    float4 v[4];
    bool w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      // iteration u covers lines base + u*32 + (lane>>3)*... : 4 lines per instruction, 8 lanes each
#pragma unroll 1
      for (int dummy = 0; dummy < 1; ++dummy) {}
      w[u] = false;
    }
    for (int it = 0; it < 32; it += 4)
    {
#pragma unroll
      for (int u = 0; u < 4; ++u)
      {
        const size_t line = base + (size_t)(it + u) * 4 + (lane >> 3);
        w[u] = line < n_lines && want[line];
        if (w[u]) v[u] = __ldcs(host + line * 8 + (lane & 7));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
      {
        const size_t line = base + (size_t)(it + u) * 4 + (lane >> 3);
        if (w[u]) dev[line * 8 + (lane & 7)] = v[u];
      }
    }
  }
}

// 64-byte granules: flag per granule, half a warp (16 lanes x 4 B) per granule, UNROLL pairs in flight
template <int UNROLL>
__global__ void fetch_half_lines(const unsigned char* __restrict__ want, const float* __restrict__ host,
                                 float* __restrict__ dev, size_t n_gran)
{
  const int lane = threadIdx.x & 31;
  const int half = lane >> 4, sub = lane & 15;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t base = warp * 64; base < n_gran; base += n_warps * 64)
  {
    // half h of the warp owns granules base + 32 h .. base + 32 h + 31; lane flags: 2 per lane
    const bool w0 = (base + lane < n_gran) && want[base + lane];
    const bool w1 = (base + 32 + lane < n_gran) && want[base + 32 + lane];
    const unsigned m0 = __ballot_sync(0xffffffffu, w0), m1 = __ballot_sync(0xffffffffu, w1);
    unsigned m = half ? m1 : m0;
    const size_t hb = base + 32 * half;
    // both halves iterate max(popc) times (warp-synchronous loads are not required here)
    while (__any_sync(0xffffffffu, m != 0))
    {
      float v[UNROLL];
      size_t at[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
      {
        at[u] = ~(size_t)0;
        if (m)
        {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          at[u] = (hb + b) * 16 + sub;
          v[u] = __ldcs(host + at[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (at[u] != ~(size_t)0) dev[at[u]] = v[u];
    }
  }
}

int main()
{
  const size_t n = (size_t)512 * 512 * 512;
  const size_t n_lines = n / 32;
  float* host = nullptr;
  CK(cudaHostAlloc(&host, n * 4, cudaHostAllocMapped));
  for (size_t i = 0; i < n; i += 1024) host[i] = (float)i;
  float* hdev = nullptr;
  CK(cudaHostGetDevicePointer(&hdev, host, 0));
  float* dev = nullptr;
  CK(cudaMalloc(&dev, n * 4));
  unsigned char* want = nullptr;
  CK(cudaMalloc(&want, n_lines));
  std::vector<unsigned char> hw(n_lines);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float ms;
  // dense copy baseline
  for (int r = 0; r < 3; ++r)
  {
    cudaEventRecord(a);
    CK(cudaMemcpyAsync(dev, host, n * 4, cudaMemcpyHostToDevice));
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    cudaEventElapsedTime(&ms, a, b);
  }
  printf("cudaMemcpyAsync dense 537 MB: %.3f ms  %.1f GB/s\n", ms, n * 4 / ms / 1e6);
  // patterns: density d of lines wanted; "rows": whole x-rows (2 KiB = 16 lines) of every k-th z-slice pair
  const double dens[] = { 1.0, 0.5, 0.26, 0.1 };
  for (int pat = 0; pat < 2; ++pat)
    for (double d : dens)
    {
      size_t cnt = 0;
      srand(1);
      if (pat == 0)
        for (size_t i = 0; i < n_lines; ++i) { hw[i] = (rand() / (double)RAND_MAX) < d; cnt += hw[i]; }
      else
      {
        // slabs: 2 of every `period` z-slices (a slice = 512*512*4/128 = 8192 lines)
        const int period = (int)(2.0 / d + 0.5);
        for (size_t i = 0; i < n_lines; ++i) { const size_t z = i / 8192; hw[i] = (z % period) < 2; cnt += hw[i]; }
      }
      CK(cudaMemcpy(want, hw.data(), n_lines, cudaMemcpyHostToDevice));
      for (int variant = 0; variant < 4; ++variant)
      {
        float best = 1e9f;
        for (int r = 0; r < 3; ++r)
        {
          cudaEventRecord(a);
          if (variant == 0) fetch_lines<1><<<148 * 8, 256>>>(want, hdev, dev, n_lines);
          if (variant == 1) fetch_lines<4><<<148 * 8, 256>>>(want, hdev, dev, n_lines);
          if (variant == 2) fetch_lines<8><<<148 * 8, 256>>>(want, hdev, dev, n_lines);
          if (variant == 3) fetch_lines_v4<<<148 * 8, 256>>>(want, (const float4*)hdev, (float4*)dev, n_lines);
          cudaEventRecord(b);
          CK(cudaEventSynchronize(b));
          cudaEventElapsedTime(&ms, a, b);
          best = ms < best ? ms : best;
        }
        printf("pattern %s density %.2f (%zu lines, %.1f MB) variant %d: %.3f ms  %.1f GB/s\n", pat ? "slabs " : "random",
               d, cnt, cnt * 128 / 1e6, variant, best, cnt * 128 / best / 1e6);
      }
    }
  // 64-byte granules at the same densities (random pattern)
  {
    const size_t n_gran = n / 16;
    unsigned char* want64 = nullptr;
    CK(cudaMalloc(&want64, n_gran));
    std::vector<unsigned char> hw64(n_gran);
    for (double d : dens)
    {
      size_t cnt = 0;
      srand(2);
      for (size_t i = 0; i < n_gran; ++i) { hw64[i] = (rand() / (double)RAND_MAX) < d; cnt += hw64[i]; }
      CK(cudaMemcpy(want64, hw64.data(), n_gran, cudaMemcpyHostToDevice));
      float best = 1e9f;
      for (int r = 0; r < 3; ++r)
      {
        cudaEventRecord(a);
        fetch_half_lines<4><<<148 * 8, 256>>>(want64, hdev, dev, n_gran);
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
      }
      printf("64-byte granules random density %.2f (%zu granules, %.1f MB): %.3f ms  %.1f GB/s\n", d, cnt, cnt * 64 / 1e6,
             best, cnt * 64 / best / 1e6);
    }
  }
  CK(cudaGetLastError());
  return 0;
}
