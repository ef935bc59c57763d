// Multi-GPU support kernels for the partition-first protocol (cccl_b200/multi_gpu.py):
//
//   select_histogram_kernel   one round of the exact MSD radix select over the UNSORTED shard: for every candidate
//                             prefix (the high digits chosen so far for one splitter) the 256-bin histogram of the
//                             next digit among the keys that carry that prefix;
//   bucket_ids_kernel         destination bucket of every key against the final splitter values:
//                             id = 2 * #{splitters below the key} + [key equals a splitter], so keys strictly between
//                             two splitters and keys tied with a splitter land in separate buckets (ties are later
//                             split by (source rank, position) on the host side of the protocol).
//
// Both work in the sort's own key domain: bit-ordered value after twiddle_in (descending included) with -0.0 viewed as
// +0.0 (common.cuh digit_view), so "equal" means exactly what the final stable sort treats as equal.
//
// They restate, for radix keys and exact splitters, the counting rounds of the reference's multi-GPU sort
// (/root/reference/cudax/include/cuda/experimental/__multi_gpu/algorithm/sort/hss/histogramming.h:522-610).
#include "../../include/b200rs.h"
#include "common.cuh"

namespace b200rs
{

constexpr int MAX_SPLITTERS = 15;

struct SplitterSet
{
  unsigned long long v[MAX_SPLITTERS + 1];
  int count;
};

constexpr int PART_THREADS = 512;

template <class U>
__global__ void __launch_bounds__(PART_THREADS) select_histogram_kernel(
  const U* __restrict__ keys, unsigned long long n, const KeyXform kx, const SplitterSet prefixes, int hi_shift,
  int lo_shift, unsigned long long* __restrict__ hist)
{
  __shared__ unsigned int sh[MAX_SPLITTERS * RADIX];
  const int np = prefixes.count;
  for (int i = threadIdx.x; i < np * RADIX; i += PART_THREADS)
  {
    sh[i] = 0;
  }
  __syncthreads();
  const XformT<U> xf(kx);
  constexpr int UNROLL = 4; // independent loads in flight per thread
  const unsigned long long stride = (unsigned long long) gridDim.x * PART_THREADS * UNROLL;
  // whole blocks iterate together (the trip count only depends on blockIdx) so the ballots below are convergent
  for (unsigned long long base = (unsigned long long) blockIdx.x * PART_THREADS * UNROLL; base < n; base += stride)
  {
    U raw[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
    {
      const unsigned long long i = base + (unsigned long long) u * PART_THREADS + threadIdx.x;
      raw[u]                     = i < n ? keys[i] : U(0);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
    {
      const unsigned long long i = base + (unsigned long long) u * PART_THREADS + threadIdx.x;
      const bool in              = i < n;
      const U v                  = digit_view(twiddle_in(raw[u], xf), xf);
      // hi_shift == key bits on the first round: every key carries the (empty) prefix
      const unsigned long long hi = hi_shift >= int(sizeof(U) * 8) ? 0ull : (unsigned long long) (v >> hi_shift);
      const unsigned int bin      = (unsigned int) (v >> lo_shift) & (RADIX - 1);
      for (int p = 0; p < np; ++p)
      {
        const bool hit        = in && hi == prefixes.v[p];
        const unsigned int hm = __ballot_sync(0xffffffffu, hit);
        if (hit)
        {
          // warp-aggregated: one shared atomic per distinct bin, so all-equal keys do not serialise
          const unsigned int peers = __match_any_sync(hm, bin);
          if ((peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0)
          {
            atomicAdd(&sh[p * RADIX + bin], (unsigned int) __popc(peers));
          }
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < np * RADIX; i += PART_THREADS)
  {
    const unsigned int c = sh[i];
    if (c != 0)
    {
      atomicAdd(&hist[i], (unsigned long long) c);
    }
  }
}

// Four consecutive keys per thread, their ids stored as one 32-bit word (ids must be 4-byte aligned: checked by
// the launcher, which falls back to VEC = 1).
template <class U, int VEC>
__global__ void __launch_bounds__(PART_THREADS) bucket_ids_kernel(
  const U* __restrict__ keys, unsigned long long n, const KeyXform kx, const SplitterSet splitters,
  unsigned char* __restrict__ ids)
{
  const XformT<U> xf(kx);
  const int m = splitters.count;
  const unsigned long long stride = (unsigned long long) gridDim.x * PART_THREADS * VEC;
  for (unsigned long long i = ((unsigned long long) blockIdx.x * PART_THREADS + threadIdx.x) * VEC; i < n; i += stride)
  {
    U raw[VEC];
#pragma unroll
    for (int u = 0; u < VEC; ++u)
    {
      raw[u] = i + u < n ? keys[i + u] : U(0);
    }
    unsigned int packed = 0;
#pragma unroll
    for (int u = 0; u < VEC; ++u)
    {
      const U v       = digit_view(twiddle_in(raw[u], xf), xf);
      unsigned int id = 0;
      for (int j = 0; j < m; ++j)
      {
        const U s = U(splitters.v[j]);
        id += (v > s ? 1u : 0u) + (v >= s ? 1u : 0u);
      }
      packed |= id << (8 * u);
    }
    if (VEC == 4 && i + 3 < n)
    {
      *reinterpret_cast<unsigned int*>(ids + i) = packed;
    }
    else
    {
#pragma unroll
      for (int u = 0; u < VEC; ++u)
      {
        if (i + u < n)
        {
          ids[i + u] = (unsigned char) (packed >> (8 * u));
        }
      }
    }
  }
}

template <class U>
static cudaError_t launch_select_t(
  const void* keys, unsigned long long n, const KeyXform& xf, const SplitterSet& pf, int round, unsigned long long* hist,
  int sms, cudaStream_t stream)
{
  const int bits     = int(sizeof(U) * 8);
  const int lo_shift = bits - RADIX_BITS * (round + 1);
  const int hi_shift = lo_shift + RADIX_BITS;
  unsigned long long want = (n + PART_THREADS * 4 - 1) / (PART_THREADS * 4);
  unsigned grid           = unsigned(sms) * 4;
  grid                    = want < grid ? unsigned(want) : grid;
  select_histogram_kernel<U><<<grid, PART_THREADS, 0, stream>>>(
    static_cast<const U*>(keys), n, xf, pf, hi_shift, lo_shift, hist);
  return cudaPeekAtLastError();
}

template <class U>
static cudaError_t launch_ids_t(
  const void* keys, unsigned long long n, const KeyXform& xf, const SplitterSet& sp, unsigned char* ids, int sms,
  cudaStream_t stream)
{
  if (reinterpret_cast<size_t>(ids) % 4 == 0)
  {
    unsigned long long want = (n + PART_THREADS * 4 - 1) / (PART_THREADS * 4);
    unsigned grid           = unsigned(sms) * 8;
    grid                    = want < grid ? unsigned(want) : grid;
    bucket_ids_kernel<U, 4><<<grid, PART_THREADS, 0, stream>>>(static_cast<const U*>(keys), n, xf, sp, ids);
  }
  else
  {
    unsigned long long want = (n + PART_THREADS - 1) / PART_THREADS;
    unsigned grid           = unsigned(sms) * 8;
    grid                    = want < grid ? unsigned(want) : grid;
    bucket_ids_kernel<U, 1><<<grid, PART_THREADS, 0, stream>>>(static_cast<const U*>(keys), n, xf, sp, ids);
  }
  return cudaPeekAtLastError();
}

static int sm_count(int* out)
{
  int dev       = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
  {
    return int(e);
  }
  e = cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev);
  return int(e);
}

} // namespace b200rs

using namespace b200rs;

extern "C" {

int b200rs_select_histogram(
  const void* d_keys,
  uint64_t num_items,
  int key_kind,
  int key_bytes,
  int descending,
  const uint64_t* h_prefixes,
  int num_prefixes,
  int round,
  uint64_t* d_hist,
  b200rs_stream_t stream_)
{
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (key_kind < 0 || key_kind > 2 || (key_bytes != 1 && key_bytes != 2 && key_bytes != 4 && key_bytes != 8)
      || (key_kind == 2 && key_bytes < 4) || num_prefixes < 0 || num_prefixes > MAX_SPLITTERS || round < 0
      || round >= key_bytes || d_hist == nullptr || (num_prefixes > 0 && h_prefixes == nullptr))
  {
    return int(cudaErrorInvalidValue);
  }
  if (num_prefixes == 0)
  {
    return 0;
  }
  cudaError_t e = cudaMemsetAsync(d_hist, 0, size_t(num_prefixes) * RADIX * sizeof(uint64_t), stream);
  if (e != cudaSuccess || num_items == 0)
  {
    return int(e);
  }
  int sms = 0;
  if (int rc = sm_count(&sms))
  {
    return rc;
  }
  SplitterSet pf;
  pf.count = num_prefixes;
  for (int i = 0; i < num_prefixes; ++i)
  {
    pf.v[i] = h_prefixes[i];
  }
  const KeyXform xf         = make_xform(key_kind, key_bytes, descending);
  unsigned long long* hist = reinterpret_cast<unsigned long long*>(d_hist);
  switch (key_bytes)
  {
    case 1: return int(launch_select_t<uint8_t>(d_keys, num_items, xf, pf, round, hist, sms, stream));
    case 2: return int(launch_select_t<uint16_t>(d_keys, num_items, xf, pf, round, hist, sms, stream));
    case 4: return int(launch_select_t<uint32_t>(d_keys, num_items, xf, pf, round, hist, sms, stream));
    default: return int(launch_select_t<uint64_t>(d_keys, num_items, xf, pf, round, hist, sms, stream));
  }
}

int b200rs_bucket_ids(
  const void* d_keys,
  uint64_t num_items,
  int key_kind,
  int key_bytes,
  int descending,
  const uint64_t* h_splitters,
  int num_splitters,
  uint8_t* d_ids,
  b200rs_stream_t stream_)
{
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (key_kind < 0 || key_kind > 2 || (key_bytes != 1 && key_bytes != 2 && key_bytes != 4 && key_bytes != 8)
      || (key_kind == 2 && key_bytes < 4) || num_splitters < 0 || num_splitters > MAX_SPLITTERS
      || (num_splitters > 0 && h_splitters == nullptr) || (num_items > 0 && d_ids == nullptr))
  {
    return int(cudaErrorInvalidValue);
  }
  if (num_items == 0)
  {
    return 0;
  }
  int sms = 0;
  if (int rc = sm_count(&sms))
  {
    return rc;
  }
  SplitterSet sp;
  sp.count = num_splitters;
  for (int i = 0; i < num_splitters; ++i)
  {
    sp.v[i] = h_splitters[i];
  }
  const KeyXform xf = make_xform(key_kind, key_bytes, descending);
  switch (key_bytes)
  {
    case 1: return int(launch_ids_t<uint8_t>(d_keys, num_items, xf, sp, d_ids, sms, stream));
    case 2: return int(launch_ids_t<uint16_t>(d_keys, num_items, xf, sp, d_ids, sms, stream));
    case 4: return int(launch_ids_t<uint32_t>(d_keys, num_items, xf, sp, d_ids, sms, stream));
    default: return int(launch_ids_t<uint64_t>(d_keys, num_items, xf, sp, d_ids, sms, stream));
  }
}

} // extern "C"
