// The whole radix sort for mid-size inputs in ONE cooperative launch (latency path).
//
// The general path enqueues 3 + passes stream operations (memset, upsweep histogram, bin scan, one onesweep launch
// per digit); below ~2^21 items every one of them is shorter than the gap between two launches, so the sort time is
// launch latency.  The reference attacks the same regime with programmatic dependent launch
// (/root/reference/cub/cub/device/dispatch/dispatch_radix_sort.cuh:1755-1756) and benchmarks it from 2^16 items
// (cub/benchmarks/bench/radix_sort/keys.cu:60-64).  Here the phases are the same device code as the general path
// (onesweep_tile of onesweep.cuh does every digit pass, chained scan included) inside one kernel whose CTAs are all
// resident (cooperative launch) and meet at grid-wide barriers:
//
//   zero bins + first look-back array | per-pass digit histograms of all keys | exclusive scan of the bins |
//   pass 0 | pass 1 | ... (each pass also zeroes the next pass's look-back array)
//
// Tiles are small (256 threads x 8 keys) so that 2^16 keys already spread over 32 SMs; a CTA takes tiles
// blockIdx, blockIdx + grid, ... in increasing order, which keeps the look-back deadlock-free without a ticket counter.
#include <cooperative_groups.h>

#include "configs.h"
#include "onesweep.cuh"

namespace cg = cooperative_groups;

namespace b200rs
{

constexpr int SMALL_NT      = 256;
constexpr int SMALL_IPT     = 8;
constexpr int SMALL_OPT     = OPT_FMA_NOT | OPT_LB_WINDOW | OPT_CTR16;
constexpr int SMALL_MAXPASS = 8;

struct SmallArgs
{
  PassArgs pass[SMALL_MAXPASS]; // exactly what the general path would launch, one entry per digit pass
  unsigned long long* bins;     // [passes][256]
  int passes;
  uint32_t tiles;
};

template <class U, int VB, bool FLOATK>
__global__ void __launch_bounds__(SMALL_NT) small_sort_kernel(const SmallArgs s)
{
  using L = OnesweepSmem<U, VB, SMALL_NT, SMALL_IPT, SMALL_OPT>;
  constexpr uint32_t TILE = L::TILE;
  extern __shared__ __align__(1024) unsigned char smem[];
  cg::grid_group grid  = cg::this_grid();
  const uint32_t sbase = uint32_t(__cvta_generic_to_shared(smem));
  const uint32_t tid   = threadIdx.x;
  const uint32_t gtid  = blockIdx.x * SMALL_NT + tid;
  const uint32_t gsize = gridDim.x * SMALL_NT;
  const int passes     = s.passes;
  const uint32_t n     = s.pass[0].num_items;

  // ---- phase 0: zero the bins and the first pass's look-back words
  for (uint32_t i = gtid; i < uint32_t(passes) * RADIX; i += gsize)
  {
    s.bins[i] = 0;
  }
  for (uint32_t i = gtid; i < s.tiles * RADIX; i += gsize)
  {
    s.pass[0].lookback[i] = 0;
  }
  // ---- phase 1: digit histograms of every pass from one read of the keys (shared-memory counters per CTA)
  uint32_t* sh = reinterpret_cast<uint32_t*>(smem);
  for (uint32_t i = tid; i < uint32_t(passes) * RADIX; i += SMALL_NT)
  {
    sh[i] = 0;
  }
  __syncthreads();
  {
    const XformT<U> xf(s.pass[0].xf);
    const U* kin = static_cast<const U*>(s.pass[0].keys_in);
    // whole warps iterate together so that the vote below is convergent
    for (uint32_t base = blockIdx.x * SMALL_NT + (tid & ~31u); base < n; base += gsize)
    {
      const uint32_t i  = base + (tid & 31u);
      const bool inside = i < n;
      const U v         = inside ? digit_view(twiddle_in(kin[i], xf), xf) : U(0);
      for (int p = 0; p < passes; ++p)
      {
        const uint32_t d  = uint32_t(v >> s.pass[p].shift) & s.pass[p].mask;
        const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
        // a warp whose keys share the digit (all-equal / few-unique inputs) adds once instead of 32 serialised atomics
        if (__all_sync(0xffffffffu, inside && d == d0))
        {
          if ((tid & 31u) == 0)
          {
            atomicAdd(&sh[p * RADIX + d0], 32u);
          }
        }
        else if (inside)
        {
          atomicAdd(&sh[p * RADIX + d], 1u);
        }
      }
    }
  }
  __syncthreads();
  grid.sync(); // the bins are zero everywhere before anybody adds to them
  for (uint32_t i = tid; i < uint32_t(passes) * RADIX; i += SMALL_NT)
  {
    if (sh[i] != 0)
    {
      atomicAdd(&s.bins[i], (unsigned long long) sh[i]);
    }
  }
  grid.sync();
  // ---- phase 2: counts -> exclusive offsets, one CTA per pass
  for (int p = blockIdx.x; p < passes; p += gridDim.x)
  {
    unsigned long long* b        = s.bins + p * RADIX;
    const unsigned long long c   = b[tid];
    const uint32_t lane          = tid & 31, warp = tid >> 5;
    unsigned long long incl      = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= uint32_t(d))
      {
        incl += o;
      }
    }
    __shared__ unsigned long long wsum[SMALL_NT / 32];
    __syncthreads();
    if (lane == 31)
    {
      wsum[warp] = incl;
    }
    __syncthreads();
    unsigned long long before = 0;
    for (uint32_t w = 0; w < warp; ++w)
    {
      before += wsum[w];
    }
    b[tid] = before + incl - c;
  }
  grid.sync();
  // ---- the digit passes
  for (int p = 0; p < passes; ++p)
  {
    const PassArgs& a = s.pass[p];
    for (uint32_t tile = blockIdx.x; tile < s.tiles; tile += gridDim.x)
    {
      __syncthreads(); // the previous tile's (or phase's) shared memory is dead
      {
        constexpr int WORDS = L::NW * RADIX * L::CTR_BYTES / 4;
#pragma unroll
        for (int j = 0; j < WORDS / SMALL_NT; ++j)
        {
          sts32(sbase + L::OFF_WARP + (j * SMALL_NT + tid) * 4, 0);
        }
        if (tid == 0)
        {
          sts32(sbase + L::OFF_MISC + 44, 0);
        }
      }
      __syncthreads();
      const uint32_t tile_base = tile * TILE;
      const uint32_t valid     = min(TILE, n - tile_base);
      if (valid == TILE)
      {
        onesweep_tile<U, VB, SMALL_NT, SMALL_IPT, RANK_BALLOT, SMALL_OPT, FLOATK, false, true>(a, sbase, tile, tile_base,
                                                                                                 valid);
      }
      else
      {
        onesweep_tile<U, VB, SMALL_NT, SMALL_IPT, RANK_BALLOT, SMALL_OPT, FLOATK, false, false>(a, sbase, tile, tile_base,
                                                                                                  valid);
      }
    }
    if (a.lookback_next != nullptr)
    {
      for (uint32_t i = gtid; i < a.lookback_next_tiles * RADIX; i += gsize)
      {
        a.lookback_next[i] = 0;
      }
    }
    if (p + 1 < passes)
    {
      grid.sync();
    }
  }
}

template <class U, int VB>
static cudaError_t launch_small_t(const SmallArgs& s, int sms, cudaStream_t stream, int* grid_out)
{
  using L                  = OnesweepSmem<U, VB, SMALL_NT, SMALL_IPT, SMALL_OPT>;
  constexpr bool CAN_FLOAT = sizeof(U) >= 2;
  const bool flt           = CAN_FLOAT && s.pass[0].xf.float_mask != 0;
  auto kernel              = flt ? small_sort_kernel<U, VB, CAN_FLOAT> : small_sort_kernel<U, VB, false>;
  size_t smem              = L::BYTES;
  const size_t hist_bytes  = size_t(SMALL_MAXPASS) * RADIX * 4;
  smem                     = smem < hist_bytes ? hist_bytes : smem;
  int per_sm               = 0;
  cudaError_t e            = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SMALL_NT, smem);
  if (e != cudaSuccess)
  {
    return e;
  }
  if (per_sm < 1)
  {
    return cudaErrorLaunchOutOfResources;
  }
  // every CTA must be resident; small grids make the barriers cheaper, so never more CTAs than tiles
  unsigned grid = unsigned(sms) * unsigned(per_sm > 4 ? 4 : per_sm);
  grid          = s.tiles < grid ? s.tiles : grid;
  grid          = grid < unsigned(s.passes) ? unsigned(s.passes) : grid; // one CTA per pass scans the bins
  if (grid_out != nullptr)
  {
    *grid_out = int(grid);
  }
  void* params[] = {const_cast<SmallArgs*>(&s)};
  return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), dim3(grid), dim3(SMALL_NT), params, smem,
                                     stream);
}

bool small_sort_supported(int key_bytes, int value_bytes)
{
  return (key_bytes == 4 || key_bytes == 8) && (value_bytes == 0 || value_bytes == 4 || value_bytes == 8);
}

unsigned long long small_sort_tile_items()
{
  return (unsigned long long) SMALL_NT * SMALL_IPT;
}

cudaError_t launch_small_sort(const PassArgs* passes, int num_passes, unsigned long long* bins, unsigned tiles,
                              int key_bytes, int value_bytes, int sms, cudaStream_t stream)
{
  if (num_passes < 1 || num_passes > SMALL_MAXPASS || !small_sort_supported(key_bytes, value_bytes))
  {
    return cudaErrorNotSupported;
  }
  SmallArgs s;
  for (int p = 0; p < num_passes; ++p)
  {
    s.pass[p] = passes[p];
  }
  for (int p = num_passes; p < SMALL_MAXPASS; ++p)
  {
    s.pass[p] = passes[num_passes - 1];
  }
  s.bins   = bins;
  s.passes = num_passes;
  s.tiles  = tiles;
#define B200RS_SMALL(KB, KT, VBV)               \
  if (key_bytes == KB && value_bytes == VBV)    \
  {                                             \
    return launch_small_t<KT, VBV>(s, sms, stream, nullptr); \
  }
  B200RS_SMALL(4, uint32_t, 0)
  B200RS_SMALL(4, uint32_t, 4)
  B200RS_SMALL(4, uint32_t, 8)
  B200RS_SMALL(8, uint64_t, 0)
  B200RS_SMALL(8, uint64_t, 4)
  B200RS_SMALL(8, uint64_t, 8)
#undef B200RS_SMALL
  return cudaErrorNotSupported;
}

} // namespace b200rs
