// Launchers for the upsweep (all-pass histogram + bin scan) and the multi-GPU splitter-rank kernel.
#include "configs.h"
#include "histogram.cuh"

namespace b200rs
{

template <class U, int PASSES>
static cudaError_t launch_hist_p(
  const void* keys, unsigned long long n, unsigned long long* bins, int begin_bit, int end_bit, const KeyXform& xf,
  unsigned grid, cudaStream_t stream, uint32_t* zero_flag)
{
  using L = HistLayout<int(sizeof(U))>;
  // identity / integer / floating-point transform (histogram.cuh); 8-byte keys have one body
  const int mode = xf.float_mask != 0 ? 2 : ((xf.sign_mask != 0 || xf.desc_mask != 0) ? 1 : 0);
  auto kernel    = sizeof(U) > 4 || mode == 2 ? histogram_kernel<U, PASSES, 2>
                 : mode == 1                  ? histogram_kernel<U, PASSES, 1>
                                              : histogram_kernel<U, PASSES, 0>;
  // keys of at most 32 bits: one table more, the kernel aligns its counters to the size of one pass's table (histogram.cuh)
  const size_t smem = size_t(PASSES + (sizeof(U) <= 4 && PASSES <= 4 ? 1 : 0)) * RADIX * L::REPLICAS * 4;
  if (smem > 48 * 1024)
  {
    // per-device attribute; cheap and idempotent, legal during stream capture
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess)
    {
      return e;
    }
  }
  kernel<<<grid, HIST_THREADS, smem, stream>>>(static_cast<const U*>(keys), n, bins, begin_bit, end_bit, xf, zero_flag);
  return cudaPeekAtLastError();
}

template <class U>
static cudaError_t launch_hist_t(
  const void* keys, unsigned long long n, unsigned long long* bins, int passes, int begin_bit, int end_bit,
  const KeyXform& xf, int sm_count, cudaStream_t stream, uint32_t* zero_flag)
{
  constexpr unsigned long long VEC = 16 / sizeof(U);
  // persistent grid: one 1024-thread CTA per SM (it owns up to 128 KB of replicated counters), never more CTAs than work
  unsigned long long want = (n / VEC + HIST_THREADS * HIST_UNROLL - 1) / (HIST_THREADS * HIST_UNROLL);
  unsigned grid           = unsigned(sm_count);
  if (want < grid)
  {
    grid = want < 1 ? 1u : unsigned(want);
  }
  switch (passes)
  {
    case 1: return launch_hist_p<U, 1>(keys, n, bins, begin_bit, end_bit, xf, grid, stream, zero_flag);
    case 2: return launch_hist_p<U, (sizeof(U) >= 2 ? 2 : 1)>(keys, n, bins, begin_bit, end_bit, xf, grid, stream, zero_flag);
    case 3: return launch_hist_p<U, (sizeof(U) >= 4 ? 3 : 1)>(keys, n, bins, begin_bit, end_bit, xf, grid, stream, zero_flag);
    case 4: return launch_hist_p<U, (sizeof(U) >= 4 ? 4 : 1)>(keys, n, bins, begin_bit, end_bit, xf, grid, stream, zero_flag);
    case 5: return launch_hist_p<U, (sizeof(U) >= 8 ? 5 : 1)>(keys, n, bins, begin_bit, end_bit, xf, grid, stream, zero_flag);
    case 6: return launch_hist_p<U, (sizeof(U) >= 8 ? 6 : 1)>(keys, n, bins, begin_bit, end_bit, xf, grid, stream, zero_flag);
    case 7: return launch_hist_p<U, (sizeof(U) >= 8 ? 7 : 1)>(keys, n, bins, begin_bit, end_bit, xf, grid, stream, zero_flag);
    case 8: return launch_hist_p<U, (sizeof(U) >= 8 ? 8 : 1)>(keys, n, bins, begin_bit, end_bit, xf, grid, stream, zero_flag);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_histogram(
  const void* keys, unsigned long long n, int key_bytes, unsigned long long* bins, int passes, int begin_bit,
  int end_bit, const KeyXform& xf, int sm_count, cudaStream_t stream, uint32_t* zero_flag)
{
  switch (key_bytes)
  {
    case 1:
      return launch_hist_t<uint8_t>(keys, n, bins, passes, begin_bit, end_bit, xf, sm_count, stream, zero_flag);
    case 2:
      return launch_hist_t<uint16_t>(keys, n, bins, passes, begin_bit, end_bit, xf, sm_count, stream, zero_flag);
    case 4:
      return launch_hist_t<uint32_t>(keys, n, bins, passes, begin_bit, end_bit, xf, sm_count, stream, zero_flag);
    case 8:
      return launch_hist_t<uint64_t>(keys, n, bins, passes, begin_bit, end_bit, xf, sm_count, stream, zero_flag);
    default:
      return cudaErrorNotSupported;
  }
}

cudaError_t launch_scan_bins(unsigned long long* bins, int passes, cudaStream_t stream)
{
  scan_bins_kernel<<<passes, RADIX, 0, stream>>>(bins);
  return cudaPeekAtLastError();
}

// One thread per splitter: lower/upper bound in the locally sorted keys under the sort's own order.
template <class U>
__global__ void splitter_ranks_kernel(
  const U* __restrict__ sorted, unsigned long long n, const KeyXform kx, const U* __restrict__ splitters, int m,
  unsigned long long* lt, unsigned long long* eq)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m)
  {
    return;
  }
  const XformT<U> xf(kx);
  const U s = digit_view(twiddle_in(splitters[i], xf), xf);
  unsigned long long lo = 0, hi = n;
  while (lo < hi) // first position whose key is not before s
  {
    const unsigned long long mid = lo + (hi - lo) / 2;
    const U k                    = digit_view(twiddle_in(sorted[mid], xf), xf);
    if (k < s)
    {
      lo = mid + 1;
    }
    else
    {
      hi = mid;
    }
  }
  const unsigned long long lower = lo;
  hi                             = n;
  while (lo < hi) // first position whose key is after s
  {
    const unsigned long long mid = lo + (hi - lo) / 2;
    const U k                    = digit_view(twiddle_in(sorted[mid], xf), xf);
    if (k <= s)
    {
      lo = mid + 1;
    }
    else
    {
      hi = mid;
    }
  }
  lt[i] = lower;
  eq[i] = lo - lower;
}

template <class U>
static cudaError_t launch_split_t(
  const void* sorted, unsigned long long n, const KeyXform& xf, const void* splitters, int m, unsigned long long* lt,
  unsigned long long* eq, cudaStream_t stream)
{
  if (m <= 0)
  {
    return cudaSuccess;
  }
  splitter_ranks_kernel<U><<<(m + 127) / 128, 128, 0, stream>>>(
    static_cast<const U*>(sorted), n, xf, static_cast<const U*>(splitters), m, lt, eq);
  return cudaPeekAtLastError();
}

cudaError_t launch_splitter_ranks(
  const void* sorted_keys, unsigned long long n, int key_bytes, const KeyXform& xf, const void* splitters,
  int num_splitters, unsigned long long* lt, unsigned long long* eq, cudaStream_t stream)
{
  switch (key_bytes)
  {
    case 1:
      return launch_split_t<uint8_t>(sorted_keys, n, xf, splitters, num_splitters, lt, eq, stream);
    case 2:
      return launch_split_t<uint16_t>(sorted_keys, n, xf, splitters, num_splitters, lt, eq, stream);
    case 4:
      return launch_split_t<uint32_t>(sorted_keys, n, xf, splitters, num_splitters, lt, eq, stream);
    case 8:
      return launch_split_t<uint64_t>(sorted_keys, n, xf, splitters, num_splitters, lt, eq, stream);
    default:
      return cudaErrorNotSupported;
  }
}

} // namespace b200rs
