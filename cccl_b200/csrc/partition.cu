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
constexpr int SELECT_STATE_WORDS = 2 + 1024; // candidate state: {unused, overflow flag, per-CTA counts}

// Candidate keys: the first full round (round 1) appends every key that carries ANY of the prefixes to a compact
// buffer; later rounds only scan that buffer (about n * prefixes / 256 keys for uniform data).  Every CTA owns one
// slice of the buffer (capacity / gridDim.x keys) and reserves space with a shared-memory counter, so there is no
// global atomic on the way; state[1] != 0 when some slice overflowed (heavily duplicated keys) -- then later rounds
// scan all keys -- and state[2 + b] = number of candidates in the slice of CTA b.  Emitting and consuming launches use
// the same grid (SELECT_STATE_WORDS bounds it).
//
// VEC keys per thread and load (16-byte loads when the array is 16-byte aligned, else VEC = 1), two loads in flight.
// Duplicate prefixes share one histogram row while counting (row of the first occurrence) and get copies at the end,
// so a key belongs to at most one row: one compare loop, one warp-aggregated shared atomic.
template <class U, int VEC>
__global__ void __launch_bounds__(PART_THREADS) select_histogram_kernel(
  const U* __restrict__ keys, unsigned long long n, const KeyXform kx, const unsigned long long* __restrict__ d_prefixes,
  int np, int hi_shift, int lo_shift, unsigned long long* __restrict__ hist, const U* __restrict__ cand_in,
  const unsigned long long* __restrict__ cand_state_in, U* __restrict__ cand_out,
  unsigned long long* __restrict__ cand_state_out, unsigned long long cand_capacity)
{
  __shared__ unsigned int sh[MAX_SPLITTERS * RADIX];
  __shared__ unsigned long long s_prefix[MAX_SPLITTERS + 1];
  __shared__ int s_first[MAX_SPLITTERS + 1];
  __shared__ signed char s_row[RADIX]; // round 1 (one-byte prefixes): histogram row of a top byte, or -1
  __shared__ unsigned int s_emitted;
  const unsigned long long slice_cap = (cand_capacity / gridDim.x) & ~15ull; // slices stay 16-byte aligned
  if (threadIdx.x == 0)
  {
    s_emitted = 0;
  }
  const bool one_byte_prefix = hi_shift == int(sizeof(U) * 8) - RADIX_BITS;
  if (threadIdx.x < RADIX)
  {
    s_row[threadIdx.x] = -1;
  }
  for (int i = threadIdx.x; i < np * RADIX; i += PART_THREADS)
  {
    sh[i] = 0;
  }
  __syncthreads();
  if (threadIdx.x < np)
  {
    const unsigned long long pf = d_prefixes[threadIdx.x];
    int first                   = threadIdx.x;
    for (int q = int(threadIdx.x) - 1; q >= 0; --q)
    {
      first = d_prefixes[q] == pf ? q : first;
    }
    s_prefix[threadIdx.x] = pf;
    s_first[threadIdx.x]  = first;
    if (one_byte_prefix && first == int(threadIdx.x))
    {
      // integer keys: the table is indexed by the RAW top byte (sign / descending flips folded into the index), so the
      // scan below tests a key with one shift and one byte load; float keys go through the transform (the -0.0 rule
      // and the sign-dependent flip need the whole key)
      const unsigned int flips = (unsigned int) ((kx.sign_mask ^ kx.desc_mask) >> (sizeof(U) * 8 - RADIX_BITS));
      const unsigned int idx   = kx.float_mask != 0 ? (unsigned int) pf : ((unsigned int) pf ^ flips);
      s_row[idx & (RADIX - 1)] = (signed char) threadIdx.x;
    }
  }
  __syncthreads();
  // source: this CTA's slice of the candidate buffer when there is one and no slice overflowed
  const bool from_slice = cand_state_in != nullptr && cand_state_in[1] == 0;
  if (from_slice)
  {
    keys = cand_in + (unsigned long long) blockIdx.x * slice_cap;
    n    = cand_state_in[2 + blockIdx.x];
  }
  U* my_out = cand_out != nullptr ? cand_out + (unsigned long long) blockIdx.x * slice_cap : nullptr;
  const XformT<U> xf(kx);
  const unsigned int lane = threadIdx.x & 31;
  constexpr int LOADS     = 2; // independent vector loads in flight per thread
  struct alignas(sizeof(U) * VEC) Vec
  {
    U k[VEC];
  };
  const unsigned long long per_iter = (unsigned long long) PART_THREADS * VEC * LOADS;
  const unsigned long long stride   = from_slice ? per_iter : (unsigned long long) gridDim.x * per_iter;
  const unsigned long long start    = from_slice ? 0ull : (unsigned long long) blockIdx.x * per_iter;
  // whole blocks iterate together (the trip count only depends on blockIdx) so the votes below are convergent
  const bool is_float = kx.float_mask != 0;
  auto body           = [&](auto check_tag, unsigned long long base) {
    constexpr bool CHECK = decltype(check_tag)::value; // only the last iteration of a block can run past n
    Vec raw[LOADS];
#pragma unroll
    for (int l = 0; l < LOADS; ++l)
    {
      const unsigned long long i = base + ((unsigned long long) l * PART_THREADS + threadIdx.x) * VEC;
      if (!CHECK || i + VEC <= n)
      {
        raw[l] = *reinterpret_cast<const Vec*>(keys + i);
      }
      else
      {
#pragma unroll
        for (int j = 0; j < VEC; ++j)
        {
          raw[l].k[j] = i + j < n ? keys[i + j] : U(0);
        }
      }
    }
    if (one_byte_prefix)
    {
      // The only round that scans every key.  Each thread tests its LOADS * VEC keys with one table look-up each, then
      // the warp handles ALL its hits of the batch together: one vote to leave early, one warp scan + one shared atomic
      // to reserve candidate space (instead of a ballot, a branch and an atomic per key).  Dense hits (duplicated keys)
      // fall through to the per-key path below, which aggregates equal (row, bin) pairs.
      int rows[LOADS * VEC];
      unsigned int nh = 0;
#pragma unroll
      for (int l = 0; l < LOADS; ++l)
      {
#pragma unroll
        for (int j = 0; j < VEC; ++j)
        {
          const unsigned long long i = base + ((unsigned long long) l * PART_THREADS + threadIdx.x) * VEC + j;
          U v                        = raw[l].k[j];
          if (is_float)
          {
            v = digit_view(twiddle_in(v, xf), xf);
          }
          const int row = (!CHECK || i < n) ? int(s_row[(unsigned int) (v >> (sizeof(U) * 8 - RADIX_BITS)) & (RADIX - 1)]) : -1;
          rows[l * VEC + j] = row;
          nh += row >= 0 ? 1u : 0u;
        }
      }
      const unsigned int tot = __reduce_add_sync(0xffffffffu, nh);
      if (tot == 0)
      {
        return;
      }
      if (tot < 64)
      {
        unsigned int incl = nh;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
          const unsigned int up = __shfl_up_sync(0xffffffffu, incl, d);
          incl += lane >= (unsigned int) d ? up : 0u;
        }
        unsigned int pos = 0;
        bool room        = false;
        if (my_out != nullptr)
        {
          if (lane == 0)
          {
            pos = atomicAdd(&s_emitted, tot);
          }
          pos  = __shfl_sync(0xffffffffu, pos, 0);
          room = (unsigned long long) pos + tot <= slice_cap;
          if (!room && lane == 0)
          {
            cand_state_out[1] = 1; // overflow: later rounds scan every key
          }
          pos += incl - nh;
        }
#pragma unroll
        for (int q = 0; q < LOADS * VEC; ++q)
        {
          if (rows[q] >= 0)
          {
            const U k = raw[q / VEC].k[q % VEC];
            const U v = digit_view(twiddle_in(k, xf), xf);
            atomicAdd(&sh[(unsigned int) rows[q] * RADIX + ((unsigned int) (v >> lo_shift) & (RADIX - 1))], 1u);
            if (room)
            {
              my_out[pos++] = k;
            }
          }
        }
        return;
      }
    }
#pragma unroll
    for (int l = 0; l < LOADS; ++l)
    {
#pragma unroll
      for (int j = 0; j < VEC; ++j)
      {
        const unsigned long long i = base + ((unsigned long long) l * PART_THREADS + threadIdx.x) * VEC + j;
        const bool inside          = !CHECK || i < n;
        int row                    = -1;
        U v                        = raw[l].k[j];
        if (one_byte_prefix)
        {
          // the only round that scans every key: one table look-up per key (integer keys: on the raw top byte)
          if (is_float)
          {
            v = digit_view(twiddle_in(v, xf), xf);
          }
          row = inside ? int(s_row[(unsigned int) (v >> (sizeof(U) * 8 - RADIX_BITS)) & (RADIX - 1)]) : -1;
          if (!__any_sync(0xffffffffu, row >= 0))
          {
            continue;
          }
          if (!is_float)
          {
            v = digit_view(twiddle_in(v, xf), xf);
          }
        }
        else
        {
          if (!__any_sync(0xffffffffu, inside))
          {
            continue;
          }
          v = digit_view(twiddle_in(v, xf), xf);
          // hi_shift == key bits on the first round: every key carries the (empty) prefix
          const unsigned long long hi = hi_shift >= int(sizeof(U) * 8) ? 0ull : (unsigned long long) (v >> hi_shift);
          if (inside)
          {
            for (int p = 0; p < np; ++p)
            {
              row = (hi == s_prefix[p] && s_first[p] == p) ? p : row;
            }
          }
        }
        const unsigned int bin = (unsigned int) (v >> lo_shift) & (RADIX - 1);
        const bool hit         = row >= 0;
        const unsigned int hm  = __ballot_sync(0xffffffffu, hit);
        if (hit)
        {
          const unsigned int slot = (unsigned int) row * RADIX + bin;
          if (__popc(hm) >= 8)
          {
            // many hits in one warp means duplicated keys: aggregate, one shared atomic per distinct (row, bin), so
            // all-equal keys do not serialise (MATCH.ANY itself is slow -- about one per 40 cycles per SM -- which is
            // why the common case of a few scattered hits below does not use it)
            const unsigned int peers = __match_any_sync(hm, slot);
            if ((peers & ((1u << lane) - 1u)) == 0)
            {
              atomicAdd(&sh[slot], (unsigned int) __popc(peers));
            }
          }
          else
          {
            atomicAdd(&sh[slot], 1u);
          }
        }
        if (my_out != nullptr && hm != 0)
        {
          unsigned int pos = 0;
          if (lane == 0)
          {
            pos = atomicAdd(&s_emitted, (unsigned int) __popc(hm));
          }
          pos = __shfl_sync(0xffffffffu, pos, 0);
          if ((unsigned long long) pos + __popc(hm) <= slice_cap)
          {
            if (hit)
            {
              my_out[pos + __popc(hm & ((1u << lane) - 1u))] = raw[l].k[j];
            }
          }
          else if (lane == 0)
          {
            cand_state_out[1] = 1; // overflow: later rounds scan every key
          }
        }
      }
    }
  };
  for (unsigned long long base = start; base < n; base += stride)
  {
    if (base + per_iter <= n)
    {
      body(std::false_type{}, base);
    }
    else
    {
      body(std::true_type{}, base);
    }
  }
  __syncthreads();
  if (my_out != nullptr && threadIdx.x == 0)
  {
    cand_state_out[2 + blockIdx.x] = s_emitted; // (beyond slice_cap only together with the overflow flag)
  }
  for (int i = threadIdx.x; i < np * RADIX; i += PART_THREADS)
  {
    const unsigned int c = sh[s_first[i / RADIX] * RADIX + (i % RADIX)];
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
  const void* keys, unsigned long long n, const KeyXform& xf, const unsigned long long* d_prefixes, int np, int round,
  unsigned long long* hist, const void* cand_in, const unsigned long long* cand_state_in, void* cand_out,
  unsigned long long* cand_state_out, unsigned long long cand_capacity, int sms, cudaStream_t stream)
{
  const int bits     = int(sizeof(U) * 8);
  const int lo_shift = bits - RADIX_BITS * (round + 1);
  const int hi_shift = lo_shift + RADIX_BITS;
  constexpr int VEC       = 16 / int(sizeof(U));
  const bool aligned      = (reinterpret_cast<size_t>(keys) | reinterpret_cast<size_t>(cand_in)) % 16 == 0;
  const unsigned long long per = (unsigned long long) PART_THREADS * 2 * (aligned ? VEC : 1);
  unsigned long long want = (n + per - 1) / per;
  unsigned grid           = unsigned(sms) * 4;
  grid                    = grid > unsigned(SELECT_STATE_WORDS - 2) ? unsigned(SELECT_STATE_WORDS - 2) : grid;
  grid                    = want < grid ? unsigned(want) : grid;
  if (aligned)
  {
    select_histogram_kernel<U, VEC><<<grid, PART_THREADS, 0, stream>>>(
      static_cast<const U*>(keys), n, xf, d_prefixes, np, hi_shift, lo_shift, hist, static_cast<const U*>(cand_in),
      cand_state_in, static_cast<U*>(cand_out), cand_state_out, cand_capacity);
  }
  else
  {
    select_histogram_kernel<U, 1><<<grid, PART_THREADS, 0, stream>>>(
      static_cast<const U*>(keys), n, xf, d_prefixes, np, hi_shift, lo_shift, hist, static_cast<const U*>(cand_in),
      cand_state_in, static_cast<U*>(cand_out), cand_state_out, cand_capacity);
  }
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
  const uint64_t* d_prefixes,
  int num_prefixes,
  int round,
  uint64_t* d_hist,
  const void* d_candidates_in,
  const uint64_t* d_candidate_state_in,
  void* d_candidates_out,
  uint64_t* d_candidate_state_out,
  uint64_t candidate_capacity,
  b200rs_stream_t stream_)
{
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if ((d_candidate_state_in != nullptr && d_candidates_in == nullptr)
      || (d_candidates_out != nullptr && d_candidate_state_out == nullptr))
  {
    return int(cudaErrorInvalidValue);
  }
  if (key_kind < 0 || key_kind > 2 || (key_bytes != 1 && key_bytes != 2 && key_bytes != 4 && key_bytes != 8)
      || (key_kind == 2 && key_bytes < 2) || num_prefixes < 0 || num_prefixes > MAX_SPLITTERS || round < 0
      || round >= key_bytes || d_hist == nullptr || (num_prefixes > 0 && d_prefixes == nullptr))
  {
    return int(cudaErrorInvalidValue);
  }
  if (num_prefixes == 0)
  {
    return 0;
  }
  cudaError_t e = cudaMemsetAsync(d_hist, 0, size_t(num_prefixes) * RADIX * sizeof(uint64_t), stream);
  if (e == cudaSuccess && d_candidates_out != nullptr)
  {
    e = cudaMemsetAsync(d_candidate_state_out, 0, SELECT_STATE_WORDS * sizeof(uint64_t), stream);
  }
  if (e != cudaSuccess || num_items == 0)
  {
    return int(e);
  }
  int sms = 0;
  if (int rc = sm_count(&sms))
  {
    return rc;
  }
  const unsigned long long* pf  = reinterpret_cast<const unsigned long long*>(d_prefixes);
  const unsigned long long* csi = reinterpret_cast<const unsigned long long*>(d_candidate_state_in);
  unsigned long long* cso       = reinterpret_cast<unsigned long long*>(d_candidate_state_out);
  const KeyXform xf         = make_xform(key_kind, key_bytes, descending);
  unsigned long long* hist = reinterpret_cast<unsigned long long*>(d_hist);
  switch (key_bytes)
  {
    case 1: return int(launch_select_t<uint8_t>(d_keys, num_items, xf, pf, num_prefixes, round, hist, d_candidates_in, csi,
                                                 d_candidates_out, cso, candidate_capacity, sms, stream));
    case 2: return int(launch_select_t<uint16_t>(d_keys, num_items, xf, pf, num_prefixes, round, hist, d_candidates_in, csi,
                                                 d_candidates_out, cso, candidate_capacity, sms, stream));
    case 4: return int(launch_select_t<uint32_t>(d_keys, num_items, xf, pf, num_prefixes, round, hist, d_candidates_in, csi,
                                                 d_candidates_out, cso, candidate_capacity, sms, stream));
    default: return int(launch_select_t<uint64_t>(d_keys, num_items, xf, pf, num_prefixes, round, hist, d_candidates_in, csi,
                                                 d_candidates_out, cso, candidate_capacity, sms, stream));
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
      || (key_kind == 2 && key_bytes < 2) || num_splitters < 0 || num_splitters > MAX_SPLITTERS
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
