// stage.cu -- demand staging of a host-resident field (VR_HOST_STAGED).
//
// In situ, the simulation's field lives in host memory and changes every cycle, so every publish pays
// for moving it to the GPU: 537 MB for a 512^3 f32 block, ~10 ms over PCIe Gen5 -- two orders of
// magnitude more than the ~0.1 ms the sampler then spends on it.  But at the reference's default
// sampling (samples = 100, a step of ~9 voxels) the rays only ever touch about a quarter of the
// field's 128-byte lines.  So instead of copying the block, the publish just registers it, and each
// trace is preceded by
//   1. a pre-pass of the sampler itself (sampler.cu, mode 4) that walks the same rays through the same
//      cells -- same code, same rounding -- and flags the lines the gathers will read, and
//   2. this kernel, which pulls the flagged lines that are not on the device yet out of the mapped host
//      array (one coalesced 128-byte read per line per warp, several in flight) into the staging
//      buffer the sampler then reads.
// Measured on B200 (profiles/r1_v5_pcie_gather.txt): the sparse pull runs at the same ~51 GB/s as the
// dense copy at any density, so the transfer time shrinks with the touched fraction.  Lines stay
// resident until the next publish: later views (cinema orbits) only fetch what they add, and once
// most of the block has been pulled the rest follows in one sweep and the pre-passes stop.
#include "vr_internal.h"

namespace vr
{
namespace
{
constexpr int kUnroll = 4;

// ALL: fetch every line that is not resident yet (the block is about to be fully staged)
template <bool ALL>
__global__ void __launch_bounds__(256) fetch_lines_kernel(unsigned char* __restrict__ want,
                                                          unsigned char* __restrict__ have,
                                                          const unsigned int* __restrict__ src,
                                                          unsigned int* __restrict__ dst, size_t n_lines,
                                                          size_t n_words, unsigned long long* __restrict__ n_have)
{
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t base = warp * 32; base < n_lines; base += n_warps * 32)
  {
    // lane i owns the flags of line base + i
    bool need = false;
    if (base + lane < n_lines)
    {
      const bool w = ALL || want[base + lane] != 0;
      if (w)
      {
        if (!ALL) want[base + lane] = 0; // re-armed for the next pre-pass
        need = have[base + lane] == 0;
        if (need) have[base + lane] = 1;
      }
    }
    unsigned m = __ballot_sync(0xffffffffu, need);
    if (m && lane == 0) atomicAdd(n_have, (unsigned long long)__popc(m));
    while (m)
    {
      unsigned int v[kUnroll];
      size_t at[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
      {
        at[u] = ~(size_t)0;
        if (m)
        {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          const size_t w = (base + b) * 32 + lane;
          if (w < n_words)
          {
            at[u] = w;
            v[u] = __ldcs(src + w);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        if (at[u] != ~(size_t)0) dst[at[u]] = v[u];
    }
  }
}
__global__ void count_lines_kernel(const unsigned char* __restrict__ have, size_t n, unsigned long long* out)
{
  unsigned long long c = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    c += have[i] != 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
// dense[i] = strided[offset + i * stride]: one component of an interleaved Blueprint mcarray, or a padded array
// (what ascent_vtkh_data_adapter.cpp:1836-1887 wraps in a vtkm::cont::ArrayHandleStride)
template <typename T>
__global__ void gather_strided_kernel(const T* __restrict__ src, size_t stride, size_t n, T* __restrict__ dst)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i * stride];
}
} // namespace

void preload_stage_kernels()
{
  preload_kernel(fetch_lines_kernel<true>);
  preload_kernel(fetch_lines_kernel<false>);
  preload_kernel(count_lines_kernel);
  preload_kernel(gather_strided_kernel<unsigned int>);
  preload_kernel(gather_strided_kernel<unsigned long long>);
}

cudaError_t launch_gather_strided(const void* src, int elem_bytes, size_t stride, size_t n, void* dst, int sm_count,
                                  cudaStream_t s)
{
  if (n == 0) return cudaSuccess;
  size_t grid = (size_t)sm_count * 8;
  const size_t need = (n + 255) / 256;
  if (grid > need) grid = need;
  if (elem_bytes == 4)
    gather_strided_kernel<unsigned int><<<(unsigned)grid, 256, 0, s>>>(static_cast<const unsigned int*>(src), stride, n,
                                                                      static_cast<unsigned int*>(dst));
  else
    gather_strided_kernel<unsigned long long><<<(unsigned)grid, 256, 0, s>>>(static_cast<const unsigned long long*>(src),
                                                                            stride, n, static_cast<unsigned long long*>(dst));
  return cudaGetLastError();
}

cudaError_t launch_count_lines(const unsigned char* have, size_t n_lines, unsigned long long* out, cudaStream_t s)
{
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  if (n_lines) count_lines_kernel<<<256, 256, 0, s>>>(have, n_lines, out);
  return cudaGetLastError();
}

cudaError_t launch_fetch_lines(unsigned char* want, unsigned char* have, const void* src, void* dst,
                               size_t n_lines, size_t n_bytes, bool all, unsigned long long* n_have,
                               int sm_count, cudaStream_t s)
{
  if (n_lines == 0) return cudaSuccess;
  size_t grid = (size_t)sm_count * 8;
  const size_t need = (n_lines + 255) / 256;
  if (grid > need) grid = need;
  if (all)
    fetch_lines_kernel<true><<<(unsigned)grid, 256, 0, s>>>(want, have, static_cast<const unsigned int*>(src),
                                                           static_cast<unsigned int*>(dst), n_lines, n_bytes / 4, n_have);
  else
    fetch_lines_kernel<false><<<(unsigned)grid, 256, 0, s>>>(want, have, static_cast<const unsigned int*>(src),
                                                            static_cast<unsigned int*>(dst), n_lines, n_bytes / 4, n_have);
  return cudaGetLastError();
}

} // namespace vr
