// Top-K selection on top of the radix-select kernels.
//
// Replaces cub::DeviceTopK::{Max,Min}{Keys,Pairs} (/root/reference/cub/cub/device/device_topk.cuh:297,775,1238 ...,
// dispatch/dispatch_topk.cuh: AIR top-k, a radix select over 11-bit digits followed by a filter).  Contract kept: the K
// best keys (and their values) are written in NO particular order, and which of several keys tied with the K-th one are
// returned is unspecified (the reference requires `determinism::not_guaranteed, output_ordering::unsorted`).
//
// Here: the K-th key is found EXACTLY by the same MSD radix select the multi-GPU sort uses for its splitters --
// round 0 = the upsweep's top-digit histogram, round 1 = a full scan that also compacts the keys that can still matter,
// later rounds scan only those candidates; after every histogram a one-warp kernel picks the bin holding rank K -- then ONE
// filter pass writes every key strictly better than the K-th and as many of its ties as are still needed.  Three reads of
// the keys, no sort, no host wait.
#include <cuda_runtime.h>

#include "../../include/b200rs.h"
#include "common.cuh"

#include <atomic>
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace b200rs
{

struct TopkState
{
  unsigned long long prefix;   // high digits of the K-th key chosen so far (bit-ordered domain of the requested order)
  unsigned long long below;    // keys strictly better than every key carrying that prefix
  unsigned long long k;        // K, capped to num_items
  unsigned long long need_eq;  // ties of the K-th key still to emit (after the last round)
  unsigned long long out_count; // items written so far by the filter
  unsigned long long eq_taken;  // ties handed out so far
};

__global__ void topk_init_kernel(TopkState* st, unsigned long long k)
{
  st->prefix = st->below = st->need_eq = st->out_count = st->eq_taken = 0;
  st->k = k;
}

// one warp: first bin whose running count reaches the remaining rank
__global__ void __launch_bounds__(32) topk_pick_kernel(const unsigned long long* hist, TopkState* st, int last)
{
  const uint32_t lane = threadIdx.x;
  unsigned long long g[8], sum = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j)
  {
    g[j] = hist[lane * 8 + j];
    sum += g[j];
  }
  unsigned long long incl = sum;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1)
  {
    const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, s);
    incl += lane >= uint32_t(s) ? up : 0ull;
  }
  const unsigned long long want = st->k - st->below; // >= 1
  unsigned long long c = incl - sum, before = 0, eq = 0;
  uint32_t bin = 0;
  bool found   = false;
#pragma unroll
  for (int j = 0; j < 8; ++j)
  {
    if (!found && c + g[j] >= want)
    {
      found  = true;
      bin    = lane * 8 + j;
      before = c;
      eq     = g[j];
    }
    c += g[j];
  }
  const uint32_t who = __ballot_sync(0xffffffffu, found);
  const int src      = who != 0 ? __ffs(who) - 1 : 31;
  if (who == 0 && lane == 31)
  {
    bin    = RADIX - 1;
    before = c - g[7];
    eq     = g[7];
  }
  bin    = __shfl_sync(0xffffffffu, bin, src);
  before = __shfl_sync(0xffffffffu, before, src);
  eq     = __shfl_sync(0xffffffffu, eq, src);
  if (lane == 0)
  {
    st->below += before;
    st->prefix = st->prefix * RADIX + bin;
    if (last)
    {
      st->need_eq = st->k - st->below; // <= eq
      (void) eq;
    }
  }
}

constexpr int TOPK_THREADS = 256;
constexpr int TOPK_ITEMS   = 8; // keys per thread and chunk

// block-wide exclusive prefix of one count per thread (TOPK_THREADS threads); total returned to every thread
__device__ __forceinline__ uint32_t topk_block_scan(uint32_t v, uint32_t* warp_sums, uint32_t& total)
{
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl       = v;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1)
  {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, s);
    incl += lane >= uint32_t(s) ? up : 0u;
  }
  __syncthreads(); // warp_sums may still be read by the previous scan
  if (lane == 31)
  {
    warp_sums[warp] = incl;
  }
  __syncthreads();
  uint32_t before = 0, all = 0;
#pragma unroll
  for (int w = 0; w < TOPK_THREADS / 32; ++w)
  {
    const uint32_t s = warp_sums[w];
    before += uint32_t(w) < warp ? s : 0u;
    all += s;
  }
  total = all;
  return before + incl - v;
}

// Every key strictly better than the K-th, and the first need_eq ties that ask (any subset of the ties is a valid
// answer).  A CTA handles chunks of TOPK_THREADS * TOPK_ITEMS keys: per chunk ONE atomic reserves the output range of
// all its selected items (and one more hands out tie tickets, only while ties are still needed and present), so the two
// global counters see n / 2048 atomics instead of one per warp row (k = 2^23 of 2^28: 4.2 -> ms).
template <class U, int VBYTES>
__device__ __forceinline__ void
topk_filter_body(const U* keys, U* keys_out, const typename value_of<VBYTES>::type* vals,
                 typename value_of<VBYTES>::type* vals_out, unsigned long long n, const KeyXform& kx, TopkState* st,
                 const U kth, const unsigned long long need_eq, uint32_t* warp_sums, unsigned long long* s_base)
{
  const XformT<U> xf(kx);
  constexpr unsigned long long CHUNK = (unsigned long long) TOPK_THREADS * TOPK_ITEMS;
  for (unsigned long long base = (unsigned long long) blockIdx.x * CHUNK; base < n;
       base += (unsigned long long) gridDim.x * CHUNK)
  {
    U raw[TOPK_ITEMS];
    uint32_t better = 0, tie = 0; // bit j: item j of this thread
#pragma unroll
    for (int j = 0; j < TOPK_ITEMS; ++j)
    {
      const unsigned long long i = base + (unsigned long long) j * TOPK_THREADS + threadIdx.x;
      raw[j]                     = i < n ? keys[i] : U(0);
    }
#pragma unroll
    for (int j = 0; j < TOPK_ITEMS; ++j)
    {
      const unsigned long long i = base + (unsigned long long) j * TOPK_THREADS + threadIdx.x;
      const U t                  = digit_view(twiddle_in(raw[j], xf), xf);
      better |= (i < n && t < kth) ? (1u << j) : 0u;
      tie |= (i < n && t == kth) ? (1u << j) : 0u;
    }
    // almost every chunk selects nothing when K << n: one barrier decides, the scans below are skipped
    if (__syncthreads_or((better | (need_eq != 0 ? tie : 0u)) != 0) == 0)
    {
      continue;
    }
    uint32_t take = better;
    if (need_eq != 0 && __syncthreads_or(tie != 0))
    {
      uint32_t ties_total;
      const uint32_t ties_before = topk_block_scan(uint32_t(__popc(tie)), warp_sums, ties_total);
      if (threadIdx.x == 0)
      {
        s_base[1] = atomicAdd(&st->eq_taken, (unsigned long long) ties_total);
      }
      __syncthreads();
      unsigned long long ticket = s_base[1] + ties_before;
#pragma unroll
      for (int j = 0; j < TOPK_ITEMS; ++j)
      {
        if (tie & (1u << j))
        {
          take |= ticket < need_eq ? (1u << j) : 0u;
          ++ticket;
        }
      }
    }
    uint32_t takes_total;
    const uint32_t takes_before = topk_block_scan(uint32_t(__popc(take)), warp_sums, takes_total);
    if (takes_total != 0) // uniform over the CTA
    {
      if (threadIdx.x == 0)
      {
        s_base[0] = atomicAdd(&st->out_count, (unsigned long long) takes_total);
      }
      __syncthreads();
      unsigned long long pos = s_base[0] + takes_before;
#pragma unroll
      for (int j = 0; j < TOPK_ITEMS; ++j)
      {
        if (take & (1u << j))
        {
          keys_out[pos] = raw[j];
          if (VBYTES > 0)
          {
            vals_out[pos] = vals[base + (unsigned long long) j * TOPK_THREADS + threadIdx.x];
          }
          ++pos;
        }
      }
    }
  }
}

template <class U, int VBYTES>
__global__ void __launch_bounds__(TOPK_THREADS)
topk_filter_kernel(const U* keys, U* keys_out, const typename value_of<VBYTES>::type* vals,
                   typename value_of<VBYTES>::type* vals_out, unsigned long long n, const KeyXform kx, TopkState* st)
{
  __shared__ uint32_t warp_sums[TOPK_THREADS / 32];
  __shared__ unsigned long long s_base[2];
  topk_filter_body<U, VBYTES>(keys, keys_out, vals, vals_out, n, kx, st, U(st->prefix), st->need_eq, warp_sums, s_base);
}

// Small inputs: the whole selection in ONE cooperative launch.  Below a few million keys every kernel of the select is
// shorter than the gap between two launches (11 stream operations ~ 60 us); here the same rounds run between grid-wide
// barriers: zero | per digit, most significant first: every CTA histograms the next digit of the keys that carry the
// prefix chosen so far (all keys, they are L2-resident at this size) and, after the barrier, picks the bin holding rank
// K from the grid-wide counts itself (every CTA computes the same answer: no second barrier) | filter.
struct TopkSmallArgs
{
  const void* keys;
  void* keys_out;
  const void* vals;
  void* vals_out;
  unsigned long long n, k;
  KeyXform xf;
  TopkState* st;
  unsigned int* hist; // [key bytes][256], zeroed by the kernel
};

template <class U, int VBYTES>
__global__ void __launch_bounds__(TOPK_THREADS) topk_small_kernel(const TopkSmallArgs a)
{
  using V = typename value_of<VBYTES>::type;
  __shared__ uint32_t sh[RADIX];
  __shared__ uint32_t warp_sums[TOPK_THREADS / 32];
  __shared__ unsigned long long s_base[2];
  __shared__ unsigned long long s_pick[2]; // chosen bin, keys before it
  cg::grid_group grid = cg::this_grid();
  constexpr int ROUNDS = int(sizeof(U));
  constexpr int BITS   = ROUNDS * 8;
  const XformT<U> xf(a.xf);
  const U* keys              = static_cast<const U*>(a.keys);
  const unsigned long long n = a.n;
  const uint32_t gtid = blockIdx.x * TOPK_THREADS + threadIdx.x, gsize = gridDim.x * TOPK_THREADS;
  for (uint32_t i = gtid; i < uint32_t(ROUNDS) * RADIX; i += gsize)
  {
    a.hist[i] = 0;
  }
  if (gtid == 0)
  {
    a.st->out_count = 0;
    a.st->eq_taken  = 0;
  }
  grid.sync();
  unsigned long long prefix = 0, below = 0;
  for (int r = 0; r < ROUNDS; ++r)
  {
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int lo_shift = BITS - 8 * (r + 1);
    // four independent loads in flight per thread (eight, and twice the CTAs: measured slower, r4q) (the keys are L2-resident: the scan is latency-bound otherwise; a
    // 32-way replicated histogram was measured too -- slower: zeroing and folding 32 KB per round costs more than the
    // bank conflicts of round 0, profiles/r4o_topk_small_sweep.jsonl)
    for (unsigned long long i0 = gtid; i0 < n; i0 += 4ull * gsize)
    {
      U raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
      {
        const unsigned long long i = i0 + (unsigned long long) u * gsize;
        raw[u]                     = i < n ? keys[i] : U(0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
      {
        const unsigned long long i = i0 + (unsigned long long) u * gsize;
        const U t                  = digit_view(twiddle_in(raw[u], xf), xf);
        // r == 0: every key carries the (empty) prefix; a shift by the full width is avoided
        if (i < n && (r == 0 || (unsigned long long) (t >> (lo_shift + 8)) == prefix))
        {
          atomicAdd(&sh[(unsigned int) (t >> lo_shift) & (RADIX - 1)], 1u);
        }
      }
    }
    __syncthreads();
    if (sh[threadIdx.x] != 0)
    {
      atomicAdd(&a.hist[r * RADIX + threadIdx.x], sh[threadIdx.x]);
    }
    grid.sync();
    // pick: first bin whose running count reaches the remaining rank (256 threads: one bin each)
    {
      const uint32_t c = a.hist[r * RADIX + threadIdx.x];
      uint32_t total;
      const uint32_t before = topk_block_scan(c, warp_sums, total);
      const unsigned long long want = a.k - below; // >= 1
      if (before < want && want <= (unsigned long long) before + c)
      {
        s_pick[0] = threadIdx.x;
        s_pick[1] = before;
      }
      __syncthreads();
      prefix = prefix * RADIX + s_pick[0];
      below += s_pick[1];
      __syncthreads();
    }
  }
  topk_filter_body<U, VBYTES>(keys, static_cast<U*>(a.keys_out), static_cast<const V*>(a.vals), static_cast<V*>(a.vals_out),
                              n, a.xf, a.st, U(prefix), a.k - below, warp_sums, s_base);
}

template <class U, int VB>
static cudaError_t launch_topk_small(const TopkSmallArgs& a, int sms, cudaStream_t stream)
{
  auto kernel = topk_small_kernel<U, VB>;
  int per_sm  = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TOPK_THREADS, 0);
  if (e != cudaSuccess)
  {
    return e;
  }
  // every CTA must be resident; few CTAs make the barriers cheap: about 4096 keys per CTA and round
  unsigned long long want = (a.n + 4095) / 4096;
  const unsigned long long cap = (unsigned long long) sms * (per_sm > 4 ? 4 : (per_sm < 1 ? 1 : per_sm));
  const unsigned grid          = unsigned(want < 1 ? 1 : (want < cap ? want : cap));
  void* params[]               = {const_cast<TopkSmallArgs*>(&a)};
  return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), dim3(grid), dim3(TOPK_THREADS), params, 0, stream);
}

template <class U>
static cudaError_t launch_topk_small_v(int vb, const TopkSmallArgs& a, int sms, cudaStream_t stream)
{
  switch (vb)
  {
    case 0: return launch_topk_small<U, 0>(a, sms, stream);
    case 1: return launch_topk_small<U, 1>(a, sms, stream);
    case 2: return launch_topk_small<U, 2>(a, sms, stream);
    case 4: return launch_topk_small<U, 4>(a, sms, stream);
    case 8: return launch_topk_small<U, 8>(a, sms, stream);
    case 16: return launch_topk_small<U, 16>(a, sms, stream);
    default: return cudaErrorNotSupported;
  }
}

template <class U, int VB>
static cudaError_t launch_filter(const void* keys, void* keys_out, const void* vals, void* vals_out, unsigned long long n,
                                 const KeyXform& xf, TopkState* st, int sms, cudaStream_t stream)
{
  using V                       = typename value_of<VB>::type;
  const unsigned long long chunk = (unsigned long long) TOPK_THREADS * TOPK_ITEMS;
  const unsigned long long want  = (n + chunk - 1) / chunk;
  const unsigned grid            = unsigned(want < (unsigned long long) sms * 8 ? want : (unsigned long long) sms * 8);
  topk_filter_kernel<U, VB><<<grid, TOPK_THREADS, 0, stream>>>(static_cast<const U*>(keys), static_cast<U*>(keys_out),
                                                              static_cast<const V*>(vals), static_cast<V*>(vals_out), n, xf,
                                                              st);
  return cudaPeekAtLastError();
}

template <class U>
static cudaError_t launch_filter_v(int vb, const void* keys, void* keys_out, const void* vals, void* vals_out,
                                   unsigned long long n, const KeyXform& xf, TopkState* st, int sms, cudaStream_t stream)
{
  switch (vb)
  {
    case 0: return launch_filter<U, 0>(keys, keys_out, vals, vals_out, n, xf, st, sms, stream);
    case 1: return launch_filter<U, 1>(keys, keys_out, vals, vals_out, n, xf, st, sms, stream);
    case 2: return launch_filter<U, 2>(keys, keys_out, vals, vals_out, n, xf, st, sms, stream);
    case 4: return launch_filter<U, 4>(keys, keys_out, vals, vals_out, n, xf, st, sms, stream);
    case 8: return launch_filter<U, 8>(keys, keys_out, vals, vals_out, n, xf, st, sms, stream);
    case 16: return launch_filter<U, 16>(keys, keys_out, vals, vals_out, n, xf, st, sms, stream);
    default: return cudaErrorNotSupported;
  }
}

} // namespace b200rs

using namespace b200rs;

// inputs of at most this many key bytes (2^23 4-byte keys) take the one-launch kernel (b200rs_set_topk_small_max
// overrides, 0 = never)
static std::atomic<unsigned long long> g_topk_small_max_bytes{32ull << 20}; // measured crossover, profiles/r4p_*

extern "C" int b200rs_set_topk_small_max(unsigned long long key_bytes_total)
{
  g_topk_small_max_bytes.store(key_bytes_total, std::memory_order_relaxed);
  return 0;
}

static size_t t_align(size_t x)
{
  return (x + 255) / 256 * 256;
}

extern "C" int b200rs_topk(
  void* d_temp_storage,
  size_t* temp_storage_bytes,
  const void* d_keys_in,
  void* d_keys_out,
  const void* d_values_in,
  void* d_values_out,
  uint64_t num_items,
  uint64_t k,
  int key_kind,
  int key_bytes,
  int value_bytes,
  int largest,
  b200rs_stream_t stream_)
{
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (temp_storage_bytes == nullptr || key_kind < 0 || key_kind > 2)
  {
    return int(cudaErrorInvalidValue);
  }
  if ((key_bytes != 1 && key_bytes != 2 && key_bytes != 4 && key_bytes != 8)
      || (value_bytes != 0 && value_bytes != 1 && value_bytes != 2 && value_bytes != 4 && value_bytes != 8
          && value_bytes != 16)
      || (key_kind == 2 && key_bytes < 2))
  {
    return int(cudaErrorNotSupported);
  }
  k = k < num_items ? k : num_items; // capped (device_topk.cuh:281-283)
  const uint64_t cand_cap = num_items / 8 + (uint64_t(1) << 16);
  const size_t off_state = 0, off_hist = 256, off_cstate = off_hist + t_align(RADIX * 8);
  const size_t off_cand = off_cstate + t_align(1026 * 8);
  const size_t total    = off_cand + t_align(size_t(cand_cap) * key_bytes) + 255;
  if (d_temp_storage == nullptr)
  {
    *temp_storage_bytes = (num_items == 0 || k == 0) ? 1 : total;
    return 0;
  }
  if (num_items == 0 || k == 0)
  {
    return 0;
  }
  if (*temp_storage_bytes < total)
  {
    return int(cudaErrorInvalidValue);
  }
  if (d_keys_in == nullptr || d_keys_out == nullptr || (value_bytes > 0 && (d_values_in == nullptr || d_values_out == nullptr)))
  {
    return int(cudaErrorInvalidValue);
  }
  cudaError_t e = cudaSuccess;
  if (k == num_items) // everything is selected: a copy
  {
    e = cudaMemcpyAsync(d_keys_out, d_keys_in, size_t(num_items) * key_bytes, cudaMemcpyDeviceToDevice, stream);
    if (e == cudaSuccess && value_bytes > 0)
    {
      e = cudaMemcpyAsync(d_values_out, d_values_in, size_t(num_items) * value_bytes, cudaMemcpyDeviceToDevice, stream);
    }
    return int(e);
  }
  unsigned char* base = reinterpret_cast<unsigned char*>(t_align(reinterpret_cast<size_t>(d_temp_storage)));
  TopkState* st       = reinterpret_cast<TopkState*>(base + off_state);
  uint64_t* hist      = reinterpret_cast<uint64_t*>(base + off_hist);
  if (num_items * uint64_t(key_bytes) <= g_topk_small_max_bytes.load(std::memory_order_relaxed))
  {
    // one cooperative launch (the 256 * 8-byte histogram area of the general path holds key_bytes * 256 u32 counters)
    int dev0 = 0, sms0 = 0;
    e = cudaGetDevice(&dev0);
    if (e == cudaSuccess)
    {
      e = cudaDeviceGetAttribute(&sms0, cudaDevAttrMultiProcessorCount, dev0);
    }
    if (e != cudaSuccess)
    {
      return int(e);
    }
    TopkSmallArgs sa;
    sa.keys     = d_keys_in;
    sa.keys_out = d_keys_out;
    sa.vals     = d_values_in;
    sa.vals_out = d_values_out;
    sa.n        = num_items;
    sa.k        = k;
    sa.xf       = make_xform(key_kind, key_bytes, largest ? 1 : 0);
    sa.st       = st;
    sa.hist     = reinterpret_cast<unsigned int*>(hist);
    switch (key_bytes)
    {
      case 1: return int(launch_topk_small_v<uint8_t>(value_bytes, sa, sms0, stream));
      case 2: return int(launch_topk_small_v<uint16_t>(value_bytes, sa, sms0, stream));
      case 4: return int(launch_topk_small_v<uint32_t>(value_bytes, sa, sms0, stream));
      default: return int(launch_topk_small_v<uint64_t>(value_bytes, sa, sms0, stream));
    }
  }
  uint64_t* cstate    = reinterpret_cast<uint64_t*>(base + off_cstate);
  void* cand          = base + off_cand;
  // "largest" = the first K of the DESCENDING order: the kernels' own descending transform
  const int descending = largest ? 1 : 0;
  const int bits       = key_bytes * 8;
  topk_init_kernel<<<1, 1, 0, stream>>>(st, k);
  for (int rnd = 0; rnd < key_bytes; ++rnd)
  {
    int rc = 0;
    if (rnd == 0)
    {
      rc = b200rs_digit_histogram(d_keys_in, num_items, key_kind, key_bytes, bits - 8, bits, descending, hist, stream_);
    }
    else
    {
      const bool emit = rnd == 1 && key_bytes > 2, use = rnd > 1 && key_bytes > 2;
      rc = b200rs_select_histogram(d_keys_in, num_items, key_kind, key_bytes, descending,
                                   reinterpret_cast<const uint64_t*>(&st->prefix), 1, rnd, hist, use ? cand : nullptr,
                                   use ? cstate : nullptr, emit ? cand : nullptr, emit ? cstate : nullptr, cand_cap, stream_);
    }
    if (rc != 0)
    {
      return rc;
    }
    topk_pick_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<const unsigned long long*>(hist), st,
                                           rnd == key_bytes - 1 ? 1 : 0);
    if ((e = cudaPeekAtLastError()) != cudaSuccess)
    {
      return int(e);
    }
  }
  int dev = 0, sms = 0;
  e = cudaGetDevice(&dev);
  if (e == cudaSuccess)
  {
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (e != cudaSuccess)
  {
    return int(e);
  }
  const KeyXform xf = make_xform(key_kind, key_bytes, descending);
  switch (key_bytes)
  {
    case 1: return int(launch_filter_v<uint8_t>(value_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out, num_items, xf, st, sms, stream));
    case 2: return int(launch_filter_v<uint16_t>(value_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out, num_items, xf, st, sms, stream));
    case 4: return int(launch_filter_v<uint32_t>(value_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out, num_items, xf, st, sms, stream));
    default: return int(launch_filter_v<uint64_t>(value_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out, num_items, xf, st, sms, stream));
  }
}
