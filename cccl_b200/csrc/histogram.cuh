// Upsweep: ONE read of the keys produces the digit histogram of every 8-bit pass, then a tiny kernel
// turns each pass's 256 counts into exclusive offsets.
//
// Replaces DeviceRadixSortHistogramKernel / AgentRadixSortHistogram and DeviceRadixSortExclusiveSumKernel
// (/root/reference/cub/cub/device/dispatch/kernels/kernel_radix_sort.cuh:447-473, :563-606,
//  cub/cub/agent/agent_radix_sort_histogram.cuh:118-121, :184-279).
//
// What bounds this kernel on B200 is the shared-memory atomic pipe, not HBM: ncu on the first version (4
// interleaved sub-histograms, profiles/r1_histogram_*) showed the LSU wavefront pipe at 92 % with 13 wavefronts per
// 32 keys -- 4 atomics per key, each ~3.3-way bank-conflicted because random digits of 32 lanes collide in the 32
// banks -- and 43 instructions per key.  So:
//   * REPLICAS sub-histograms per pass laid out bin-major, replica-minor: counter (pass, bin, replica) lives at word
//     (pass*256 + bin)*REPLICAS + replica.  With REPLICAS == 32 and replica == lane, lane l only ever touches bank l:
//     every warp-wide atomic is ONE conflict-free wavefront whatever the digits are (all-equal keys included).
//     4-byte keys: 4 passes x 256 x 32 x 4 B = 128 KB; 8-byte keys use 16 replicas (2-way conflicts) to fit 8 passes;
//   * one persistent CTA of 1024 threads per SM owns that shared memory;
//   * no-return shared atomics (red.shared.add) on 32-bit shared-window addresses, 128-bit key loads, the key
//     transform fused, the pass count a template parameter (no per-key branches);
//   * 64-bit global bins, one global atomic per non-zero (pass, bin) per CTA.
#pragma once

#include "common.cuh"

namespace b200rs
{

constexpr int HIST_THREADS = 1024;
constexpr int HIST_UNROLL  = 2; // 16-byte loads in flight per thread

template <int KEY_BYTES>
struct HistLayout
{
  static constexpr int MAXP     = KEY_BYTES;                    // at most one pass per key byte
  static constexpr int REPLICAS = KEY_BYTES <= 4 ? 32 : 16;     // sub-histograms per pass
  static constexpr size_t BYTES = size_t(MAXP) * RADIX * REPLICAS * 4;
};

__device__ __forceinline__ void red_shared_add(uint32_t addr, uint32_t v)
{
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// `mine` = shared-window address of this thread's replica of (pass 0, bin 0)
template <class U, int PASSES, int REPLICAS>
__device__ __forceinline__ void
hist_accumulate(U bits, const XformT<U>& xf, uint32_t mine, int begin_bit, int end_bit, bool& saw_zero_alias)
{
  const U t    = twiddle_in(bits, xf);
  const U view = digit_view(t, xf);
  saw_zero_alias |= t == xf.neg_zero; // only meaningful (and only consumed) for floating-point keys
#pragma unroll
  for (int p = 0; p < PASSES; ++p)
  {
    const int bit    = begin_bit + p * RADIX_BITS;
    uint32_t d       = uint32_t(view >> bit);
    if (p == PASSES - 1)
    {
      d &= (1u << min(RADIX_BITS, end_bit - bit)) - 1u; // the last pass may be narrower than 8 bits
    }
    else
    {
      d &= uint32_t(RADIX - 1);
    }
    red_shared_add(mine + (p * RADIX + d) * (REPLICAS * 4), 1u);
  }
}

// Keys of at most 32 bits: the digit is extracted already scaled to the counter stride and merged with the replica's base
// (one rotate + one LOP3 per digit, the pass's table selected by an immediate offset of the RED), instead of shift, mask,
// multiply-add: 26.6 -> warp instructions per 32 keys for four digits (the kernel is ALU-bound, section 3.1 of DESIGN.md).
// Needs the counters aligned to one pass's table (256 bins * REPLICAS * 4 bytes); the kernel aligns them itself.
template <int OFFSET>
__device__ __forceinline__ void red_shared_inc_at(uint32_t addr)
{
  asm volatile("red.shared.add.u32 [%0+%1], %2;" ::"r"(addr), "n"(OFFSET), "r"(1u) : "memory");
}

struct HistFold
{
  uint32_t rot[4], msk[4]; // per pass: rotate-right amount and (digit mask << log2(stride))
};

template <int PASSES, int REPLICAS, int P = 0>
__device__ __forceinline__ void hist_accumulate_fold(uint32_t view, uint32_t mine, const HistFold& f)
{
  if constexpr (P < PASSES)
  {
    red_shared_inc_at<P * RADIX * REPLICAS * 4>((__funnelshift_r(view, view, f.rot[P]) & f.msk[P]) | mine);
    hist_accumulate_fold<PASSES, REPLICAS, P + 1>(view, mine, f);
  }
}

// bins: [PASSES][256] uint64, zero on entry; accumulated with global atomics.
// MODE: 0 = the key transform is the identity (unsigned keys, ascending): no per-key transform at all; 1 = integer
// transform (sign flip and / or descending inversion); 2 = floating point (sign-dependent flip, -0.0 viewed as +0.0, and
// the aliased-zero flag for the passes).
template <class U, int PASSES, int MODE>
__global__ void __launch_bounds__(HIST_THREADS, 1)
histogram_kernel(const U* __restrict__ keys, unsigned long long n, unsigned long long* bins, int begin_bit, int end_bit,
                 const KeyXform kx, uint32_t* zero_flag)
{
  using L                = HistLayout<int(sizeof(U))>;
  constexpr int REPLICAS = L::REPLICAS;
  constexpr int VEC      = 16 / int(sizeof(U)); // keys per 128-bit load
  constexpr bool FOLD         = sizeof(U) <= 4 && PASSES <= 4;
  constexpr uint32_t PASS_TBL = RADIX * REPLICAS * 4; // bytes of one pass's replicated counters
  extern __shared__ __align__(16) unsigned char hsmem[];
  // FOLD: the tables start at the next multiple of PASS_TBL (the launcher reserves the slack)
  const uint32_t sraw  = uint32_t(__cvta_generic_to_shared(hsmem));
  const uint32_t sbase = FOLD ? (sraw + PASS_TBL - 1u) & ~(PASS_TBL - 1u) : sraw;
  uint32_t* h          = reinterpret_cast<uint32_t*>(hsmem + (sbase - sraw));
  HistFold fold;
  if constexpr (FOLD)
  {
    constexpr int LOG_STRIDE = REPLICAS == 32 ? 7 : 6;
#pragma unroll
    for (int p = 0; p < PASSES; ++p)
    {
      const int bit   = begin_bit + p * RADIX_BITS;
      const int nbits = min(RADIX_BITS, end_bit - bit);
      fold.rot[p]     = uint32_t(bit - LOG_STRIDE) & 31u;
      fold.msk[p]     = ((1u << nbits) - 1u) << LOG_STRIDE;
    }
  }

  const XformT<U> xf(kx);
  for (int i = threadIdx.x; i < PASSES * RADIX * REPLICAS; i += HIST_THREADS)
  {
    h[i] = 0;
  }
  __syncthreads();
  const uint32_t mine = sbase + (threadIdx.x & (REPLICAS - 1)) * 4;
  bool saw            = false; // a key whose pattern is ranked as the other zero (floats: -0.0 / +0.0) was read
  auto accumulate = [&](U bits) {
    if constexpr (FOLD)
    {
      U t = bits;
      if constexpr (MODE >= 1)
      {
        t = twiddle_in(bits, xf);
      }
      if constexpr (MODE == 2)
      {
        saw |= t == xf.neg_zero;
        t = digit_view(t, xf);
      }
      hist_accumulate_fold<PASSES, REPLICAS>(uint32_t(t), mine, fold);
    }
    else
    {
      hist_accumulate<U, PASSES, REPLICAS>(bits, xf, mine, begin_bit, end_bit, saw);
    }
  };

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
      accumulate(keys[i]);
    }
    for (unsigned long long i = tail + threadIdx.x; i < n; i += HIST_THREADS)
    {
      accumulate(keys[i]);
    }
  }

  const uint4* vkeys              = reinterpret_cast<const uint4*>(keys + head);
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
        accumulate(e[j]);
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
      accumulate(e[j]);
    }
  }
  // the passes skip the per-key zero test while this word stays 0 (PassArgs::zero_flag)
  if (__syncthreads_or(saw ? 1 : 0) != 0 && threadIdx.x == 0 && zero_flag != nullptr)
  {
    *zero_flag = 1u;
  }

  // fold the replicas: consecutive threads read consecutive words (conflict-free), REPLICAS-wide segmented sum by shuffles
  for (int i = threadIdx.x; i < PASSES * RADIX * REPLICAS; i += HIST_THREADS)
  {
    unsigned long long c = h[i];
#pragma unroll
    for (int s = REPLICAS / 2; s > 0; s >>= 1)
    {
      c += __shfl_down_sync(0xffffffffu, c, s, REPLICAS);
    }
    if ((threadIdx.x & (REPLICAS - 1)) == 0 && c != 0)
    {
      atomicAdd(bins + i / REPLICAS, c);
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
