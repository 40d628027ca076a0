// png.cu -- Render::Save's PNG encode on the device.
//
// The reference turns each finished canvas into a PNG on the host: PNGEncoder::Encode
// (src/libs/png_utils/ascent_png_encoder.cpp:258-303) converts the float canvas to RGBA8, flips the rows and
// calls lodepng with `btype = 2, use_lz77 = 0` ("use less aggressive compression": Huffman coding only) --
// tens of milliseconds of one host core per frame, against ~0.1 ms for the frame itself here.  This unit
// produces the complete PNG byte stream on the GPU, so that only the compressed file crosses PCIe:
//
//   png_rows_kernel      one CTA per scanline: Sub filter (type 1, bpp 4), run-length tokens (a literal, then
//                        length/distance-1 matches -- cleared background and flat colour collapse to a few
//                        bytes per row), fixed-Huffman bit packing (deflate BTYPE = 01) by 256 threads at once:
//                        every thread counts the bits of its segment's tokens, a block scan gives its bit
//                        offset, then it ORs its codes into the row's bit buffer in shared memory.  Each row
//                        is its own deflate block, closed by an empty stored block (the Z_SYNC_FLUSH marker
//                        00 00 FF FF) so that rows end on byte boundaries and can be concatenated.  The CTA also
//                        leaves the row's Adler-32 sums and the raw CRC-32 of its compressed bytes.
//   png_finish_kernel    one CTA: scan of the row sizes, Adler-32 and CRC-32 of the whole stream by COMBINING
//                        the rows' partial values (both checksums are linear: CRC over GF(2)[x]/P, Adler over
//                        Z/65521), signature + IHDR + IDAT header, final empty block, checksums, IEND.
//   png_compact_kernel   rows copied from their fixed-size slots to their final offsets.
//
// The output is a valid PNG (RGBA, 8 bits) whose decoded pixels equal vr_canvas_download_rgba8's bytes; it is
// NOT byte-identical to lodepng's file (a different but equally lossless deflate stream) -- the reference's own
// image comparison decodes both sides (ascent_png_compare.cpp).  tests/png_model.py restates this exact stream
// layout on the CPU; the GPU test compares byte for byte against it and decodes the file with PIL.
#include <cstdint>

#include "vr_internal.h"

namespace vr
{
namespace
{
constexpr int kPngThreads = 256;
constexpr unsigned kCrcPoly = 0xEDB88320u; // CRC-32 (reflected)
constexpr unsigned kAdlerMod = 65521u;

// ---- GF(2) arithmetic of the CRC, reflected representation: bit 31 = x^0 (zlib's multmodp / x2nmodp)
__device__ __forceinline__ unsigned gf_mul(unsigned a, unsigned b)
{
  unsigned p = 0;
#pragma unroll 1
  for (unsigned m = 0x80000000u; m; m >>= 1)
  {
    if (a & m) p ^= b;
    b = (b & 1u) ? (b >> 1) ^ kCrcPoly : (b >> 1);
  }
  return p;
}
// x^(8 n) mod P
__device__ __forceinline__ unsigned gf_xpow8(unsigned long long n)
{
  unsigned p = 0x80000000u;  // x^0
  unsigned sq = 0x00800000u; // x^8  (bit 31 - 8)
  while (n)
  {
    if (n & 1ull) p = gf_mul(p, sq);
    sq = gf_mul(sq, sq);
    n >>= 1;
  }
  return p;
}
// the CRC register after one more byte (no table: used for a handful of header bytes only)
__device__ __forceinline__ unsigned crc_byte(unsigned reg, unsigned char b)
{
  reg ^= b;
#pragma unroll
  for (int k = 0; k < 8; ++k) reg = (reg & 1u) ? (reg >> 1) ^ kCrcPoly : (reg >> 1);
  return reg;
}

__device__ __forceinline__ unsigned rev_bits(unsigned v, int n) { return __brev(v) >> (32 - n); }

// fixed Huffman codes (RFC 1951, 3.2.6), already bit-reversed for the LSB-first stream
__device__ __forceinline__ void literal_code(unsigned v, unsigned& code, int& n)
{
  if (v < 144u) { code = rev_bits(0x30u + v, 8); n = 8; }
  else { code = rev_bits(0x190u + (v - 144u), 9); n = 9; }
}
// a match of length L (3..258) at distance 1: length symbol + extra bits + the 5-bit distance code 0
__device__ __forceinline__ void match_code(unsigned L, unsigned& code, int& n)
{
  unsigned k, eb, base;
  const unsigned d = L - 3u;
  if (L == 258u) { k = 28u; eb = 0u; base = 258u; }
  else if (d < 8u) { k = d; eb = 0u; base = L; }
  else
  {
    eb = (31u - (unsigned)__clz(d)) - 2u;
    k = 4u * eb + 4u + ((d >> eb) & 3u);
    base = 3u + ((4u + (k & 3u)) << eb);
  }
  const unsigned sym = 257u + k;
  if (sym < 280u) { code = rev_bits(sym - 256u, 7); n = 7; }
  else { code = rev_bits(0xC0u + (sym - 280u), 8); n = 8; }
  code |= (L - base) << n;
  n += (int)eb;
  n += 5; // distance code 0 (distance 1): five zero bits
}

// tokens of the filtered bytes f[begin, end): for every run of equal bytes one literal, then matches of up to 258
// at distance 1, then 0-2 literals.  WRITE = false only counts bits.
template <bool WRITE>
__device__ __forceinline__ unsigned emit_segment(const unsigned char* f, int begin, int end, unsigned* words, unsigned bitpos)
{
  unsigned pos = bitpos;
  auto put = [&](unsigned code, int n) {
    if (WRITE)
    {
      const unsigned w = pos >> 5, o = pos & 31u;
      atomicOr(&words[w], code << o);
      if (o + (unsigned)n > 32u) atomicOr(&words[w + 1], code >> (32u - o));
    }
    pos += (unsigned)n;
  };
  int i = begin;
  while (i < end)
  {
    const unsigned char b = f[i];
    int L = 1;
    while (i + L < end && f[i + L] == b) ++L;
    unsigned code;
    int n;
    literal_code(b, code, n);
    put(code, n);
    int rem = L - 1;
    while (rem >= 3)
    {
      const int m = rem < 258 ? rem : 258;
      unsigned mc;
      int mn;
      match_code((unsigned)m, mc, mn);
      put(mc, mn);
      rem -= m;
    }
    for (; rem > 0; --rem) put(code, n);
    i += L;
  }
  return pos - bitpos;
}

// dynamic shared memory: [filtered bytes, padded to 4][bit buffer words]
__global__ void __launch_bounds__(kPngThreads) png_rows_kernel(const uchar4* __restrict__ rgba, int W, int H, unsigned char* slots,
                                                               unsigned slot_stride, unsigned* __restrict__ sizes,
                                                               unsigned* __restrict__ adler_a, unsigned* __restrict__ adler_b,
                                                               unsigned* __restrict__ crc_raw, unsigned* __restrict__ crc_pow)
{
  extern __shared__ __align__(16) unsigned char s_png[];
  __shared__ unsigned s_crc_table[256];
  __shared__ unsigned s_scan[kPngThreads / 32];
  __shared__ unsigned long long s_red[2][kPngThreads / 32];
  const int n = 1 + 4 * W;                   // filter byte + pixels
  const int n_pad = (n + 3) & ~3;
  unsigned char* f = s_png;
  unsigned* words = reinterpret_cast<unsigned*>(s_png + n_pad);
  const int n_words = (int)(slot_stride / 4u);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    unsigned c = (unsigned)tid;
#pragma unroll
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ kCrcPoly : (c >> 1);
    s_crc_table[tid] = c;
  }
  for (int row = blockIdx.x; row < H; row += gridDim.x)
  {
    __syncthreads(); // (the previous row's buffers are free)
    // ---- Sub filter, bpp = 4: byte - the same channel of the pixel to the left
    const uchar4* src = rgba + (size_t)row * W;
    if (tid == 0) f[0] = 1;
    for (int x = tid; x < W; x += kPngThreads)
    {
      const uchar4 p = src[x];
      const uchar4 q = x > 0 ? src[x - 1] : make_uchar4(0, 0, 0, 0);
      f[1 + 4 * x + 0] = (unsigned char)(p.x - q.x);
      f[1 + 4 * x + 1] = (unsigned char)(p.y - q.y);
      f[1 + 4 * x + 2] = (unsigned char)(p.z - q.z);
      f[1 + 4 * x + 3] = (unsigned char)(p.w - q.w);
    }
    for (int k = tid; k < n_words; k += kPngThreads) words[k] = 0u;
    __syncthreads();
    // ---- bits of every thread's segment, exclusive scan
    const int seg = (n + kPngThreads - 1) / kPngThreads;
    const int begin = min(tid * seg, n), end = min(begin + seg, n);
    const unsigned my_bits = emit_segment<false>(f, begin, end, nullptr, 0u);
    unsigned incl = my_bits;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_scan[warp] = incl;
    // ---- Adler-32 sums of the row: a = sum d_i, b = sum (n - i) d_i
    unsigned long long sa = 0, sb = 0;
    for (int i = begin; i < end; ++i)
    {
      sa += f[i];
      sb += (unsigned long long)(n - i) * f[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      sa += __shfl_xor_sync(0xffffffffu, sa, o);
      sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    if (lane == 0) { s_red[0][warp] = sa; s_red[1][warp] = sb; }
    __syncthreads();
    unsigned warp_off = 0, total_bits = 0;
#pragma unroll
    for (int w = 0; w < kPngThreads / 32; ++w)
    {
      if (w < warp) warp_off += s_scan[w];
      total_bits += s_scan[w];
    }
    // block header: BFINAL = 0, BTYPE = 01 -> bits 0,1,0 (LSB first) = value 2 in three bits
    const unsigned my_pos = 3u + warp_off + incl - my_bits;
    if (tid == 0) atomicOr(&words[0], 2u);
    emit_segment<true>(f, begin, end, words, my_pos);
    // end-of-block (seven zero bits) + the empty stored block's header (three zero bits): nothing to OR in
    const unsigned body_bytes = (3u + total_bits + 7u + 3u + 7u) / 8u;
    const unsigned row_bytes = body_bytes + 4u; // + LEN = 0000, NLEN = FFFF
    __syncthreads();
    if (tid == 0)
    {
      unsigned char* wb = reinterpret_cast<unsigned char*>(words);
      wb[body_bytes + 0] = 0x00; wb[body_bytes + 1] = 0x00; wb[body_bytes + 2] = 0xFF; wb[body_bytes + 3] = 0xFF;
      unsigned long long a = 0, b = 0;
      for (int w = 0; w < kPngThreads / 32; ++w) { a += s_red[0][w]; b += s_red[1][w]; }
      adler_a[row] = (unsigned)(a % kAdlerMod);
      adler_b[row] = (unsigned)(b % kAdlerMod);
      sizes[row] = row_bytes;
    }
    __syncthreads();
    // ---- raw CRC (register starts at 0, no final xor) of the row's compressed bytes: 32 lanes take a
    // contiguous share each, lane 0 chains the shares with the x^(8 len) shift operator
    if (warp == 0)
    {
      const unsigned char* wb = reinterpret_cast<const unsigned char*>(words);
      const unsigned share = (row_bytes + 31u) / 32u;
      const unsigned b0 = min((unsigned)lane * share, row_bytes), b1 = min(b0 + share, row_bytes);
      unsigned reg = 0;
      for (unsigned i = b0; i < b1; ++i) reg = s_crc_table[(reg ^ wb[i]) & 0xffu] ^ (reg >> 8);
      const unsigned pw = gf_xpow8(b1 - b0);
      unsigned acc = 0;
      for (int l = 0; l < 32; ++l)
      {
        const unsigned r_l = __shfl_sync(0xffffffffu, reg, l);
        const unsigned p_l = __shfl_sync(0xffffffffu, pw, l);
        if (lane == 0) acc = gf_mul(acc, p_l) ^ r_l;
      }
      if (lane == 0)
      {
        crc_raw[row] = acc;
        crc_pow[row] = gf_xpow8(row_bytes);
      }
    }
    // ---- the row's bytes to its slot (whole words; the tail beyond row_bytes is never read)
    unsigned* dst = reinterpret_cast<unsigned*>(slots + (size_t)row * slot_stride);
    const int used_words = (int)((row_bytes + 3u) / 4u);
    for (int k = tid; k < used_words; k += kPngThreads) dst[k] = words[k];
  }
}

constexpr int kFinishThreads = 1024;
constexpr unsigned kPngHeaderBytes = 8u + 25u + 8u + 2u; // signature, IHDR chunk, IDAT length + type, zlib header

// offsets[r] = first byte of row r inside the file; out[...] = everything except the rows; *total = file size
__global__ void __launch_bounds__(kFinishThreads) png_finish_kernel(int W, int H, const unsigned* __restrict__ sizes,
                                                                     const unsigned* __restrict__ adler_a,
                                                                     const unsigned* __restrict__ adler_b,
                                                                     const unsigned* __restrict__ crc_raw,
                                                                     const unsigned* __restrict__ crc_pow,
                                                                     unsigned long long* __restrict__ offsets,
                                                                     unsigned char* __restrict__ out, unsigned long long capacity,
                                                                     unsigned long long* __restrict__ total_out)
{
  __shared__ unsigned long long s_sum[kFinishThreads];
  __shared__ unsigned s_c[kFinishThreads], s_p[kFinishThreads];
  __shared__ unsigned long long s_a[kFinishThreads], s_b[kFinishThreads];
  const int tid = threadIdx.x;
  // rows [r0, r1) of this thread (contiguous, in order)
  const int per = (H + kFinishThreads - 1) / kFinishThreads;
  const int r0 = min(tid * per, H), r1 = min(r0 + per, H);
  const unsigned long long n = 1ull + 4ull * (unsigned long long)W;
  unsigned long long bytes = 0, a = 0, b = 0;
  unsigned c = 0, p = 0x80000000u; // (CRC, shift) of an empty string: identity of the combine below
  for (int r = r0; r < r1; ++r)
  {
    bytes += sizes[r];
    c = gf_mul(c, crc_pow[r]) ^ crc_raw[r];
    p = gf_mul(p, crc_pow[r]);
    a += adler_a[r];
    // b_total = sum_r (b_r + (H - 1 - r) n a_r): the rows after r add their lengths to every weight of row r
    b += adler_b[r] + (((unsigned long long)(H - 1 - r) * n) % kAdlerMod) * adler_a[r] % kAdlerMod;
  }
  s_sum[tid] = bytes; s_c[tid] = c; s_p[tid] = p; s_a[tid] = a % kAdlerMod; s_b[tid] = b % kAdlerMod;
  __syncthreads();
  // exclusive scan of the byte counts (Hillis-Steele on 1024 entries) for the row offsets
  unsigned long long incl = bytes;
  for (int o = 1; o < kFinishThreads; o <<= 1)
  {
    const unsigned long long v = tid >= o ? s_sum[tid - o] : 0ull;
    __syncthreads();
    incl += v;
    s_sum[tid] = incl;
    __syncthreads();
  }
  unsigned long long off = kPngHeaderBytes + incl - bytes;
  for (int r = r0; r < r1; ++r) { offsets[r] = off; off += sizes[r]; }
  // ordered tree reduction of (crc, shift) pairs: left o right = (crc_l * shift_r ^ crc_r, shift_l * shift_r)
  for (int o = 1; o < kFinishThreads; o <<= 1)
  {
    if ((tid & (2 * o - 1)) == 0)
    {
      const unsigned cr = s_c[tid + o], pr = s_p[tid + o];
      s_c[tid] = gf_mul(s_c[tid], pr) ^ cr;
      s_p[tid] = gf_mul(s_p[tid], pr);
      s_a[tid] = (s_a[tid] + s_a[tid + o]) % kAdlerMod;
      s_b[tid] = (s_b[tid] + s_b[tid + o]) % kAdlerMod;
    }
    __syncthreads();
  }
  if (tid != 0) return;
  const unsigned long long rows_bytes = s_sum[kFinishThreads - 1];
  const unsigned long long zlen = 2ull + rows_bytes + 5ull + 4ull; // zlib header, rows, final empty block, Adler-32
  const unsigned long long total = 8ull + 25ull + 12ull + zlen + 12ull;
  *total_out = total;
  if (total > capacity) return; // (the host reports the overflow)
  auto be32 = [](unsigned char* q, unsigned v) { q[0] = (unsigned char)(v >> 24); q[1] = (unsigned char)(v >> 16); q[2] = (unsigned char)(v >> 8); q[3] = (unsigned char)v; };
  unsigned char* q = out;
  const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
  for (int k = 0; k < 8; ++k) q[k] = sig[k];
  q += 8;
  // IHDR: width, height, bit depth 8, colour type 6 (RGBA), deflate, adaptive filtering, no interlace
  be32(q, 13u);
  q[4] = 'I'; q[5] = 'H'; q[6] = 'D'; q[7] = 'R';
  be32(q + 8, (unsigned)W);
  be32(q + 12, (unsigned)H);
  q[16] = 8; q[17] = 6; q[18] = 0; q[19] = 0; q[20] = 0;
  unsigned reg = 0xFFFFFFFFu;
  for (int k = 4; k < 21; ++k) reg = crc_byte(reg, q[k]);
  be32(q + 21, reg ^ 0xFFFFFFFFu);
  q += 25;
  // IDAT
  be32(q, (unsigned)zlen);
  q[4] = 'I'; q[5] = 'D'; q[6] = 'A'; q[7] = 'T';
  q[8] = 0x78; q[9] = 0x01; // zlib: deflate, 32 KiB window, no preset dictionary, fastest
  reg = 0xFFFFFFFFu;
  for (int k = 4; k < 10; ++k) reg = crc_byte(reg, q[k]);
  reg = gf_mul(reg, s_p[0]) ^ s_c[0]; // ... all rows at once
  unsigned char* tail = out + kPngHeaderBytes + rows_bytes;
  tail[0] = 0x01; tail[1] = 0x00; tail[2] = 0x00; tail[3] = 0xFF; tail[4] = 0xFF; // BFINAL = 1, stored, empty
  // Adler-32 with the initial a = 1: a = 1 + sum d, b = N + sum (N - i) d_i, N = H n
  const unsigned long long N = (unsigned long long)H * n;
  const unsigned ad_a = (unsigned)((1ull + s_a[0]) % kAdlerMod);
  const unsigned ad_b = (unsigned)((N % kAdlerMod + s_b[0]) % kAdlerMod);
  be32(tail + 5, (ad_b << 16) | ad_a);
  for (int k = 0; k < 9; ++k) reg = crc_byte(reg, tail[k]);
  be32(tail + 9, reg ^ 0xFFFFFFFFu);
  // IEND
  unsigned char* e = tail + 13;
  be32(e, 0u);
  e[4] = 'I'; e[5] = 'E'; e[6] = 'N'; e[7] = 'D';
  reg = 0xFFFFFFFFu;
  for (int k = 4; k < 8; ++k) reg = crc_byte(reg, e[k]);
  be32(e + 8, reg ^ 0xFFFFFFFFu);
}

__global__ void __launch_bounds__(256) png_compact_kernel(const unsigned char* __restrict__ slots, unsigned slot_stride,
                                                          const unsigned* __restrict__ sizes,
                                                          const unsigned long long* __restrict__ offsets, int H,
                                                          unsigned char* __restrict__ out, unsigned long long capacity,
                                                          const unsigned long long* __restrict__ total)
{
  if (*total > capacity) return;
  for (int row = blockIdx.x; row < H; row += gridDim.x)
  {
    const unsigned char* src = slots + (size_t)row * slot_stride;
    unsigned char* dst = out + offsets[row];
    const unsigned nb = sizes[row];
    for (unsigned k = threadIdx.x; k < nb; k += blockDim.x) dst[k] = src[k];
  }
}
} // namespace

// worst case of a row: every byte a 9-bit literal, + block header, end-of-block, stored header, 4 marker bytes
unsigned png_slot_stride(int W)
{
  const unsigned long long n = 1ull + 4ull * (unsigned long long)W;
  const unsigned long long bytes = (3ull + 9ull * n + 10ull + 7ull) / 8ull + 4ull;
  return (unsigned)((bytes + 8ull + 15ull) & ~15ull); // (+ a spare word: the bit writer may touch word w + 1)
}
size_t png_capacity(int W, int H) { return (size_t)png_slot_stride(W) * (size_t)H + 128; }
size_t png_rows_smem(int W) { return (size_t)((1 + 4 * W + 3) & ~3) + png_slot_stride(W); }

void preload_png_kernels()
{
  preload_kernel(png_rows_kernel);
  preload_kernel(png_finish_kernel);
  preload_kernel(png_compact_kernel);
}

// rgba: W x H RGBA8 on the device, row 0 = the PNG's first (top) scanline.  scratch: png_scratch_bytes(W, H).
// out: device buffer of `capacity` bytes; total_dev receives the file size (also when it exceeds the capacity).
cudaError_t launch_png_encode(const uchar4* rgba, int W, int H, unsigned char* scratch, unsigned char* out,
                              unsigned long long capacity, unsigned long long* total_dev, int sm_count, cudaStream_t s)
{
  const unsigned stride = png_slot_stride(W);
  unsigned char* slots = scratch;
  unsigned* sizes = reinterpret_cast<unsigned*>(scratch + (size_t)stride * H);
  unsigned* adler_a = sizes + H;
  unsigned* adler_b = adler_a + H;
  unsigned* crc_raw = adler_b + H;
  unsigned* crc_pow = crc_raw + H;
  unsigned long long* offsets = reinterpret_cast<unsigned long long*>(
    (reinterpret_cast<uintptr_t>(crc_pow + H) + 7) & ~(uintptr_t)7);
  const size_t smem = png_rows_smem(W);
  cudaError_t e = cudaFuncSetAttribute(png_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, png_rows_kernel, kPngThreads, smem);
  if (per_sm < 1) per_sm = 1;
  int grid = sm_count * per_sm;
  if (grid > H) grid = H;
  png_rows_kernel<<<grid, kPngThreads, smem, s>>>(rgba, W, H, slots, stride, sizes, adler_a, adler_b, crc_raw, crc_pow);
  png_finish_kernel<<<1, kFinishThreads, 0, s>>>(W, H, sizes, adler_a, adler_b, crc_raw, crc_pow, offsets, out, capacity,
                                                  total_dev);
  int cgrid = sm_count * 4;
  if (cgrid > H) cgrid = H;
  png_compact_kernel<<<cgrid, 256, 0, s>>>(slots, stride, sizes, offsets, H, out, capacity, total_dev);
  return cudaGetLastError();
}
size_t png_scratch_bytes(int W, int H)
{
  return (size_t)png_slot_stride(W) * (size_t)H + (size_t)H * 5 * sizeof(unsigned) + 8 + (size_t)H * sizeof(unsigned long long);
}
} // namespace vr
