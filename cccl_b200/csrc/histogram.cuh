// Upsweep: ONE read of the keys produces the digit histogram of every 8-bit pass, then a tiny kernel
// turns each pass's 256 counts into exclusive offsets.
//
// Replaces DeviceRadixSortHistogramKernel / AgentRadixSortHistogram and DeviceRadixSortExclusiveSumKernel
// (/root/reference/cub/cub/device/dispatch/kernels/kernel_radix_sort.cuh:447-473, :563-606,
//  cub/cub/agent/agent_radix_sort_histogram.cuh:118-121, :184-279).  Differences: 128-bit loads with the key
// transform fused, a persistent grid sized from the SM count, COPIES lane-interleaved sub-histograms per pass so
// that low-entropy inputs (all lanes hitting one bin) serialise COPIES times less, 64-bit global bins.
#pragma once

#include "common.cuh"

namespace b200rs
{

constexpr int HIST_THREADS = 512;
constexpr int HIST_COPIES  = 4;
constexpr int HIST_UNROLL  = 4; // 16-byte loads in flight per thread

template <class U, int MAXP>
__device__ __forceinline__ void hist_accumulate(
  U bits, const XformT<U>& xf, uint32_t* h_lane, int passes, int begin_bit, int end_bit)
{
  const U view = digit_view(twiddle_in(bits, xf), xf);
#pragma unroll
  for (int p = 0; p < MAXP; ++p)
  {
    if (p < passes)
    {
      const int bit       = begin_bit + p * RADIX_BITS;
      const int nbits     = min(RADIX_BITS, end_bit - bit);
      const uint32_t d    = uint32_t(view >> bit) & ((1u << nbits) - 1u);
      atomicAdd(h_lane + p * (RADIX * HIST_COPIES) + d * HIST_COPIES, 1u);
    }
  }
}

// bins: [passes][256] uint64, zero on entry; accumulated with global atomics.
template <class U>
__global__ void __launch_bounds__(HIST_THREADS)
histogram_kernel(const U* __restrict__ keys, unsigned long long n, unsigned long long* bins, int passes, int begin_bit,
                 int end_bit, const KeyXform kx)
{
  constexpr int MAXP = int(sizeof(U));          // at most one pass per key byte
  constexpr int VEC  = 16 / int(sizeof(U));     // keys per 128-bit load
  __shared__ uint32_t h[MAXP * RADIX * HIST_COPIES];

  const XformT<U> xf(kx);
  for (int i = threadIdx.x; i < MAXP * RADIX * HIST_COPIES; i += HIST_THREADS)
  {
    h[i] = 0;
  }
  __syncthreads();
  uint32_t* h_lane = h + (threadIdx.x & (HIST_COPIES - 1));

  // split [0,n) into a scalar head up to 16-byte alignment, a vector body and a scalar tail
  const unsigned long long addr = reinterpret_cast<unsigned long long>(keys);
  unsigned long long head       = ((16 - (addr & 15)) & 15) / sizeof(U);
  if (head > n)
  {
    head = n;
  }
  const unsigned long long nvec = (n - head) / VEC;
  const unsigned long long tail = head + nvec * VEC;

  if (blockIdx.x == 0)
  {
    for (unsigned long long i = threadIdx.x; i < head; i += HIST_THREADS)
    {
      hist_accumulate<U, MAXP>(keys[i], xf, h_lane, passes, begin_bit, end_bit);
    }
    for (unsigned long long i = tail + threadIdx.x; i < n; i += HIST_THREADS)
    {
      hist_accumulate<U, MAXP>(keys[i], xf, h_lane, passes, begin_bit, end_bit);
    }
  }

  const uint4* vkeys = reinterpret_cast<const uint4*>(keys + head);
  const unsigned long long stride = (unsigned long long) gridDim.x * HIST_THREADS;
  unsigned long long v            = (unsigned long long) blockIdx.x * HIST_THREADS + threadIdx.x;
  // main loop: HIST_UNROLL independent 16-byte loads per thread, each warp instruction covers 512 contiguous bytes
  for (; v + stride * (HIST_UNROLL - 1) < nvec; v += stride * HIST_UNROLL)
  {
    uint4 q[HIST_UNROLL];
#pragma unroll
    for (int u = 0; u < HIST_UNROLL; ++u)
    {
      q[u] = __ldg(vkeys + v + stride * u);
    }
#pragma unroll
    for (int u = 0; u < HIST_UNROLL; ++u)
    {
      const U* e = reinterpret_cast<const U*>(&q[u]);
#pragma unroll
      for (int j = 0; j < VEC; ++j)
      {
        hist_accumulate<U, MAXP>(e[j], xf, h_lane, passes, begin_bit, end_bit);
      }
    }
  }
  for (; v < nvec; v += stride)
  {
    const uint4 q = __ldg(vkeys + v);
    const U* e    = reinterpret_cast<const U*>(&q);
#pragma unroll
    for (int j = 0; j < VEC; ++j)
    {
      hist_accumulate<U, MAXP>(e[j], xf, h_lane, passes, begin_bit, end_bit);
    }
  }
  __syncthreads();

  for (int i = threadIdx.x; i < passes * RADIX; i += HIST_THREADS)
  {
    unsigned long long c = 0;
#pragma unroll
    for (int k = 0; k < HIST_COPIES; ++k)
    {
      c += h[i * HIST_COPIES + k];
    }
    if (c != 0)
    {
      atomicAdd(bins + i, c);
    }
  }
}

// In-place exclusive scan of each pass's 256 bins; one block of 256 threads per pass.
__global__ void __launch_bounds__(RADIX) scan_bins_kernel(unsigned long long* bins)
{
  __shared__ unsigned long long wsum[RADIX / 32];
  unsigned long long* b = bins + size_t(blockIdx.x) * RADIX;
  const uint32_t lane   = threadIdx.x & 31;
  const uint32_t warp   = threadIdx.x >> 5;
  const unsigned long long c = b[threadIdx.x];
  unsigned long long incl    = c;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1)
  {
    const unsigned long long n = __shfl_up_sync(0xffffffffu, incl, s);
    if (lane >= uint32_t(s))
    {
      incl += n;
    }
  }
  if (lane == 31)
  {
    wsum[warp] = incl;
  }
  __syncthreads();
  unsigned long long base = 0;
  for (uint32_t w = 0; w < warp; ++w)
  {
    base += wsum[w];
  }
  b[threadIdx.x] = base + incl - c;
}

} // namespace b200rs
