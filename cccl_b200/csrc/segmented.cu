// Segmented radix sort: every segment [begin[s], end[s]) of one key (and value) array sorted independently.
//
// Replaces cub::DeviceSegmentedRadixSort (/root/reference/cub/cub/device/device_segmented_radix_sort.cuh:86,234,
// dispatch/dispatch_segmented_radix_sort.cuh, kernel DeviceSegmentedRadixSortKernel in
// dispatch/kernels/kernel_segmented_radix_sort.cuh:113-300).  The reference launches one kernel per 6-bit digit pass,
// each CTA doing upsweep + scan + downsweep of one segment, every pass a round trip through HBM whatever the segment
// size.  Here ONE launch sorts everything, one CTA per segment:
//   * a segment of at most one tile (5120 / 2560 / 1280 items for a dominant item size of <= 4 / 8 / 16 bytes) is read
//     once, sorted entirely in shared memory (single_tile_sort: all 8-bit passes) and written once;
//   * a longer segment runs its 8-bit passes inside the same CTA: the digit histograms of every pass from one read of
//     the segment, then per pass a scan and the tiles in order, each ranked with the same warp-ballot code and scattered behind running per-digit offsets
//     (the chained scan of the unsegmented kernel degenerates to these running offsets when one CTA owns all tiles).
//     Passes ping-pong between the output array and a scratch copy so that the input is never written and the last
//     pass lands in the output (same rule as the unsegmented pointer API).
// Positions outside every segment are neither read nor written (device_segmented_radix_sort.cuh:59).
//
// Key order: the reference's segmented kernel never inverts keys for descending sorts (it reverses the digit instead,
// kernel_segmented_radix_sort.cuh:213-272), so the -0.0 == +0.0 rule applies in the un-inverted domain at EVERY segment
// size -- the rule make_xform calls single_tile_rule (pinned by tests/golden/segmented/cubseg_f32_*).
#include "../../include/b200rs.h"
#include "segmented_long.h"

#include <atomic>
#include "single_tile.cuh"

namespace b200rs
{

struct SegArgs
{
  const void* keys_in;
  void* keys_out;
  void* keys_tmp;
  const void* vals_in;
  void* vals_out;
  void* vals_tmp;
  const void* begin_offsets;
  const void* end_offsets;
  long long num_segments;
  int offset_bytes; // 4 or 8 (signed or unsigned: segment lengths are differences)
  int begin_bit;
  int end_bit;
  uint32_t all_ones;
  KeyXform xf;
  const SegLongCtl* long_ctl; // != nullptr: segments longer than long_min are sorted by segmented_long.cu
  uint32_t long_min;
  uint32_t tiny_max; // segments of at most this many items are sorted by segmented_tiny_kernel (0 = none)
  uint32_t batch;    // segments a CTA of segmented_sort_kernel examines at a time (1 .. ST_THREADS)
};

__device__ __forceinline__ long long load_offset(const void* p, long long i, int bytes)
{
  return bytes == 8 ? static_cast<const long long*>(p)[i] : (long long) static_cast<const unsigned int*>(p)[i];
}

// One 8-bit pass of one long segment by the calling CTA (src -> dst, both already offset to the segment).
template <class U, int VBYTES, int IPT, bool FLOATK>
__device__ __forceinline__ void segment_pass(
  const U* src, U* dst, const typename value_of<VBYTES>::type* vsrc, typename value_of<VBYTES>::type* vdst,
  const uint32_t len, const int bit, const uint32_t mask, const bool first, const bool last, const SegArgs& a,
  const uint32_t sbase, const uint32_t* counts, uint32_t* s_off, uint32_t* s_excl)
{
  using L = SingleTileSmem<U, VBYTES, IPT>;
  using V = typename value_of<VBYTES>::type;
  constexpr int NW        = L::NW;
  constexpr uint32_t TILE = L::TILE;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t s_cnt  = sbase + L::OFF_CNT;
  const uint32_t s_misc = sbase + L::OFF_MISC;
  const uint32_t s_keys = sbase + L::OFF_KEYS;
  const uint32_t s_vals = sbase + L::OFF_VALS;
  const uint32_t s_mine = s_cnt + warp * (RADIX * 2);
  const XformT<U> xf(a.xf);
  const U neg_zero = U(a.xf.neg_zero), pos_zero = U(a.xf.pos_zero);

  // ---- this pass's digit counts (from the all-pass histogram of the segment) -> exclusive offsets
  s_off[tid] = counts[tid];
  __syncthreads();
  {
    const uint32_t c = s_off[tid];
    uint32_t incl    = c;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1)
    {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, s);
      if (lane >= uint32_t(s))
      {
        incl += up;
      }
    }
    if (lane == 31)
    {
      sts32(s_misc + warp * 4, incl);
    }
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w)
    {
      const uint32_t ws = lds32(s_misc + w * 4);
      before += (uint32_t(w) < warp) ? ws : 0u;
    }
    s_off[tid] = before + incl - c;
  }
  __syncthreads();

  // ---- the tiles, in order
  const uint32_t lt_mask = lanemask_lt(), gt_mask = lanemask_gt();
  const uint32_t chunk = warp * 32 * IPT + lane;
  for (uint32_t t0 = 0; t0 < len; t0 += TILE)
  {
    const uint32_t valid = (len - t0) < TILE ? (len - t0) : TILE;
    U key[IPT];
    V val[VBYTES > 0 ? IPT : 1];
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      const uint32_t p = chunk + i * 32;
      U k              = p < valid ? src[t0 + p] : U(~U(0)); // padding: largest digit, last in tile order
      if (first && p < valid)
      {
        k = twiddle_in(k, xf);
      }
      key[i] = k;
      if (VBYTES > 0 && p < valid)
      {
        val[i] = vsrc[t0 + p];
      }
    }
#pragma unroll
    for (int j = 0; j < NW * RADIX * 2 / 4 / ST_THREADS; ++j)
    {
      sts32(s_cnt + (j * ST_THREADS + tid) * 4, 0);
    }
    __syncthreads(); // also: the previous tile's staged items have been scattered
    uint32_t rank[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      const uint32_t d = pass_digit<FLOATK>(key[i], bit, mask, neg_zero, pos_zero);
      uint32_t b, c;
      match_digit_ballot_fma(d, a.all_ones, b, c);
      const uint32_t before = __popc(b & c & lt_mask);
      const uint32_t ctr    = s_mine + d * 2;
      const uint32_t next   = ctr_ld<true>(ctr) + before + 1;
      if ((b & c & gt_mask) == 0)
      {
        ctr_st<true>(ctr, next);
      }
      __syncwarp();
      rank[i] = next - 1;
    }
    __syncthreads();
    // per-digit totals of the tile, exclusive scan over digits, per-warp bases
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w)
    {
      total += ctr_ld<true>(s_cnt + (w * RADIX + tid) * 2);
    }
    uint32_t incl = total;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1)
    {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, s);
      if (lane >= uint32_t(s))
      {
        incl += up;
      }
    }
    if (lane == 31)
    {
      sts32(s_misc + warp * 4, incl);
    }
    __syncthreads();
    uint32_t excl = incl - total;
#pragma unroll
    for (int w = 0; w < NW; ++w)
    {
      const uint32_t ws = lds32(s_misc + w * 4);
      excl += (uint32_t(w) < warp) ? ws : 0u;
    }
    s_excl[tid] = excl;
    {
      uint32_t run = excl;
#pragma unroll
      for (int w = 0; w < NW; ++w)
      {
        const uint32_t addr = s_cnt + (w * RADIX + tid) * 2;
        const uint32_t c    = ctr_ld<true>(addr);
        ctr_st<true>(addr, run);
        run += c;
      }
    }
    __syncthreads();
    // stage in digit order
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      const uint32_t d = pass_digit<FLOATK>(key[i], bit, mask, neg_zero, pos_zero);
      const uint32_t r = rank[i] + ctr_ld<true>(s_mine + d * 2);
      sts_t<U>(s_keys + r * uint32_t(sizeof(U)), key[i]);
      if (VBYTES > 0)
      {
        sts_t<V>(s_vals + r * uint32_t(sizeof(V)), val[i]);
      }
    }
    __syncthreads();
    // scatter behind the running offsets: staged position p of digit d goes to s_off[d] + (p - excl[d])
    for (uint32_t p = tid; p < valid; p += ST_THREADS)
    {
      const U k        = lds_t<U>(s_keys + p * uint32_t(sizeof(U)));
      const uint32_t d = pass_digit<FLOATK>(k, bit, mask, neg_zero, pos_zero);
      const uint32_t o = s_off[d] + (p - s_excl[d]);
      dst[o]           = last ? twiddle_out(k, xf) : k;
      if (VBYTES > 0)
      {
        vdst[o] = lds_t<V>(s_vals + p * uint32_t(sizeof(V)));
      }
    }
    __syncthreads();
    // padding keys carry the largest digit and rank last: they are not part of its count
    s_off[tid] += total - ((tid == mask) ? (TILE - valid) : 0u);
    // (the barrier at the top of the next tile orders this update before the next scatter)
  }
  __syncthreads(); // the CTA's global stores of this pass are visible to its loads of the next
}

template <class U, int VBYTES, int IPT, bool FLOATK>
__global__ void __launch_bounds__(ST_THREADS, 3) segmented_sort_kernel(const SegArgs a)
{
  using L = SingleTileSmem<U, VBYTES, IPT>;
  using V = typename value_of<VBYTES>::type;
  constexpr int MAXPASS = int(sizeof(U)); // 8-bit digits
  extern __shared__ __align__(16) unsigned char seg_smem[];
  __shared__ uint32_t s_off[RADIX];
  __shared__ uint32_t s_excl[RADIX];
  __shared__ uint32_t s_hist[MAXPASS][RADIX]; // long segments: digit counts of every pass, from ONE read of the segment
  const uint32_t sbase = uint32_t(__cvta_generic_to_shared(seg_smem));
  // A CTA looks at a.batch segments at a time, one per thread, and keeps the ones that are its business (not empty, not
  // tiny, not long): with 168 K tiny segments the kernel used to spend 0.22 ms on every thread of every CTA loading the
  // offsets of segments it then skipped.  The host sizes the batch so that there are still >= 16 batches per SM.
  __shared__ uint32_t s_todo[ST_THREADS];
  __shared__ uint32_t s_ntodo;
  const bool long_on = a.long_ctl != nullptr && a.long_ctl->overflow == 0;
  for (long long batch = (long long) blockIdx.x * a.batch; batch < a.num_segments; batch += (long long) gridDim.x * a.batch)
  {
    __syncthreads(); // the previous batch's list is dead
    if (threadIdx.x == 0)
    {
      s_ntodo = 0;
    }
    __syncthreads();
    {
      const long long seg = batch + threadIdx.x;
      if (threadIdx.x < a.batch && seg < a.num_segments)
      {
        const long long b0   = load_offset(a.begin_offsets, seg, a.offset_bytes);
        const long long e0   = load_offset(a.end_offsets, seg, a.offset_bytes);
        const long long len0 = a.offset_bytes == 8 ? e0 - b0 : (long long) (int) (uint32_t(e0) - uint32_t(b0));
        // long: whole-grid passes (segmented_long.cu); tiny: one warp per segment (segmented_tiny_kernel)
        if (len0 > (long long) a.tiny_max && !(long_on && len0 > (long long) a.long_min))
        {
          s_todo[atomicAdd(&s_ntodo, 1u)] = threadIdx.x;
        }
      }
    }
    __syncthreads();
    const uint32_t ntodo = s_ntodo;
  for (uint32_t q = 0; q < ntodo; ++q)
  {
    const long long seg = batch + s_todo[q];
    const long long b   = load_offset(a.begin_offsets, seg, a.offset_bytes);
    const long long e   = load_offset(a.end_offsets, seg, a.offset_bytes);
    const long long len = a.offset_bytes == 8 ? e - b : (long long) (int) (uint32_t(e) - uint32_t(b));
    const U* kin = static_cast<const U*>(a.keys_in) + b;
    U* kout      = static_cast<U*>(a.keys_out) + b;
    const V* vin = VBYTES > 0 ? static_cast<const V*>(a.vals_in) + b : nullptr;
    V* vout      = VBYTES > 0 ? static_cast<V*>(a.vals_out) + b : nullptr;
    __syncthreads(); // the previous segment's shared memory is dead
    if (len <= (long long) L::TILE)
    {
      SingleTileArgs st;
      st.keys_in   = kin;
      st.keys_out  = kout;
      st.vals_in   = vin;
      st.vals_out  = vout;
      st.num_items = uint32_t(len);
      st.begin_bit = a.begin_bit;
      st.end_bit   = a.end_bit;
      st.all_ones  = a.all_ones;
      st.xf        = a.xf;
      single_tile_sort<U, VBYTES, IPT, FLOATK>(st, sbase);
      continue;
    }
    U* ktmp      = static_cast<U*>(a.keys_tmp) + b;
    V* vtmp      = VBYTES > 0 ? static_cast<V*>(a.vals_tmp) + b : nullptr;
    const int passes = (a.end_bit - a.begin_bit + RADIX_BITS - 1) / RADIX_BITS;
    if (passes == 0) // empty bit range: the segment is copied (dispatch_radix_sort.cuh:1966-1977)
    {
      for (long long i = threadIdx.x; i < len; i += ST_THREADS)
      {
        kout[i] = kin[i];
        if (VBYTES > 0)
        {
          vout[i] = vin[i];
        }
      }
      continue;
    }
    // digit histograms of every pass from one read of the segment (they do not depend on the item order)
    for (int i = threadIdx.x; i < passes * RADIX; i += ST_THREADS)
    {
      (&s_hist[0][0])[i] = 0;
    }
    __syncthreads();
    {
      const XformT<U> xf(a.xf);
      const U neg_zero = U(a.xf.neg_zero), pos_zero = U(a.xf.pos_zero);
      const uint32_t lane = threadIdx.x & 31;
      for (uint32_t base = threadIdx.x & ~31u; base < uint32_t(len); base += ST_THREADS)
      {
        const uint32_t i  = base + lane;
        const bool inside = i < uint32_t(len);
        const U k         = inside ? twiddle_in(kin[i], xf) : U(0);
        for (int p = 0; p < passes; ++p)
        {
          const int bit       = a.begin_bit + p * RADIX_BITS;
          const int nbits     = (a.end_bit - bit) < RADIX_BITS ? (a.end_bit - bit) : RADIX_BITS;
          const uint32_t d    = pass_digit<FLOATK>(k, bit, (1u << nbits) - 1u, neg_zero, pos_zero);
          const uint32_t d0   = __shfl_sync(0xffffffffu, d, 0);
          if (__all_sync(0xffffffffu, inside && d == d0)) // a warp of equal digits adds once
          {
            if (lane == 0)
            {
              atomicAdd(&s_hist[p][d0], 32u);
            }
          }
          else if (inside)
          {
            atomicAdd(&s_hist[p][d], 1u);
          }
        }
      }
    }
    __syncthreads();
    const U* src  = kin;
    const V* vsrc = vin;
    for (int p = 0; p < passes; ++p)
    {
      // never write the input; the last pass lands in the output (dispatch_radix_sort.cuh:1850-1857)
      const bool to_out = ((passes - 1 - p) % 2) == 0;
      U* dst            = to_out ? kout : ktmp;
      V* vdst           = to_out ? vout : vtmp;
      const int bit     = a.begin_bit + p * RADIX_BITS;
      const int nbits   = (a.end_bit - bit) < RADIX_BITS ? (a.end_bit - bit) : RADIX_BITS;
      segment_pass<U, VBYTES, IPT, FLOATK>(src, dst, vsrc, vdst, uint32_t(len), bit, (1u << nbits) - 1u, p == 0,
                                           p == passes - 1, a, sbase, s_hist[p], s_off, s_excl);
      src  = dst;
      vsrc = vdst;
    }
  }
  }
}

// Tiny segments: ONE WARP per segment of at most TINY_ROWS * 32 items.  A 256-thread CTA per 100-item segment is mostly
// overhead (barriers, 8 warps' counter tables, a 256-digit block scan per pass); here a warp keeps the segment in
// registers (item j * 32 + lane in slot j), ranks every 8-bit pass with the same match-by-ballot code against ONE
// 256-entry counter table, scans the table itself (8 digits per lane) and permutes through a private staging area.
// No block-wide barrier anywhere; items past the end never take part (their lanes are masked out of the ballots).
constexpr int TINY_ROWS  = 8;
constexpr int TINY_WARPS = 4; // per CTA

template <class U, int VBYTES, bool FLOATK>
__global__ void __launch_bounds__(TINY_WARPS * 32) segmented_tiny_kernel(const SegArgs a)
{
  using V = typename value_of<VBYTES>::type;
  constexpr int ITEM = int(sizeof(U)) > VBYTES ? int(sizeof(U)) : VBYTES;
  __shared__ uint32_t s_cnt[TINY_WARPS][RADIX];
  __shared__ __align__(16) unsigned char s_stage[TINY_WARPS][TINY_ROWS * 32 * ITEM];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* cnt       = s_cnt[warp];
  U* stage_k          = reinterpret_cast<U*>(s_stage[warp]);
  V* stage_v          = reinterpret_cast<V*>(s_stage[warp]);
  const XformT<U> xf(a.xf);
  const U neg_zero = U(a.xf.neg_zero), pos_zero = U(a.xf.pos_zero);
  const uint32_t lt_mask = lanemask_lt();
  const int passes       = (a.end_bit - a.begin_bit + RADIX_BITS - 1) / RADIX_BITS;
  const long long nwarps = (long long) gridDim.x * TINY_WARPS;
  for (long long seg = (long long) blockIdx.x * TINY_WARPS + warp; seg < a.num_segments; seg += nwarps)
  {
    const long long b   = load_offset(a.begin_offsets, seg, a.offset_bytes);
    const long long e   = load_offset(a.end_offsets, seg, a.offset_bytes);
    const long long len = a.offset_bytes == 8 ? e - b : (long long) (int) (uint32_t(e) - uint32_t(b));
    if (len <= 0 || len > (long long) a.tiny_max)
    {
      continue;
    }
    const uint32_t n    = uint32_t(len);
    const uint32_t rows = (n + 31) / 32; // warp-uniform
    const U* kin = static_cast<const U*>(a.keys_in) + b;
    U* kout      = static_cast<U*>(a.keys_out) + b;
    U key[TINY_ROWS];
    V val[VBYTES > 0 ? TINY_ROWS : 1];
#pragma unroll
    for (int j = 0; j < TINY_ROWS; ++j)
    {
      const uint32_t i = j * 32 + lane;
      if (i < n)
      {
        key[j] = twiddle_in(kin[i], xf);
        if (VBYTES > 0)
        {
          val[j] = static_cast<const V*>(a.vals_in)[b + i];
        }
      }
    }
    for (int p = 0; p < passes; ++p)
    {
      const int bit       = a.begin_bit + p * RADIX_BITS;
      const int nbits     = (a.end_bit - bit) < RADIX_BITS ? (a.end_bit - bit) : RADIX_BITS;
      const uint32_t mask = (1u << nbits) - 1u;
#pragma unroll
      for (int q = 0; q < RADIX / 32; ++q)
      {
        cnt[q * 32 + lane] = 0;
      }
      __syncwarp();
      uint32_t rank[TINY_ROWS];
#pragma unroll
      for (int j = 0; j < TINY_ROWS; ++j)
      {
        if (uint32_t(j) < rows)
        {
          const bool valid    = uint32_t(j) * 32 + lane < n;
          const uint32_t d    = valid ? pass_digit<FLOATK>(key[j], bit, mask, neg_zero, pos_zero) : 0u;
          const uint32_t live = __ballot_sync(0xffffffffu, valid);
          uint32_t mb, mc;
          match_digit_ballot(d, mb, mc);
          const uint32_t peers = mb & mc & live;
          if (valid)
          {
            const uint32_t old = cnt[d];
            rank[j]            = old + uint32_t(__popc(peers & lt_mask));
            if ((peers >> lane) >> 1 == 0) // highest peer lane: the new running count
            {
              cnt[d] = old + uint32_t(__popc(peers));
            }
          }
          __syncwarp();
        }
      }
      // exclusive scan of the 256 counts: lane l owns digits 8l .. 8l + 7
      {
        uint32_t c[8], sum = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
        {
          c[q] = cnt[lane * 8 + q];
          sum += c[q];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1)
        {
          const uint32_t up = __shfl_up_sync(0xffffffffu, incl, sft);
          incl += lane >= uint32_t(sft) ? up : 0u;
        }
        uint32_t run = incl - sum;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q)
        {
          cnt[lane * 8 + q] = run;
          run += c[q];
        }
      }
      __syncwarp();
      // permute: keys (then values) through the staging area
#pragma unroll
      for (int j = 0; j < TINY_ROWS; ++j)
      {
        if (uint32_t(j) * 32 + lane < n)
        {
          rank[j] += cnt[pass_digit<FLOATK>(key[j], bit, mask, neg_zero, pos_zero)];
          stage_k[rank[j]] = key[j];
        }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < TINY_ROWS; ++j)
      {
        if (uint32_t(j) * 32 + lane < n)
        {
          key[j] = stage_k[j * 32 + lane];
        }
      }
      if (VBYTES > 0)
      {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < TINY_ROWS; ++j)
        {
          if (uint32_t(j) * 32 + lane < n)
          {
            stage_v[rank[j]] = val[j];
          }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < TINY_ROWS; ++j)
        {
          if (uint32_t(j) * 32 + lane < n)
          {
            val[j] = stage_v[j * 32 + lane];
          }
        }
      }
      __syncwarp();
    }
#pragma unroll
    for (int j = 0; j < TINY_ROWS; ++j)
    {
      const uint32_t i = j * 32 + lane;
      if (i < n)
      {
        kout[i] = twiddle_out(key[j], xf);
        if (VBYTES > 0)
        {
          static_cast<V*>(a.vals_out)[b + i] = val[j];
        }
      }
    }
  }
}

template <class U, int VB>
static cudaError_t launch_seg_tiny(const SegArgs& a, int sms, cudaStream_t stream)
{
  constexpr bool CAN_FLOAT = sizeof(U) >= 2;
  auto kernel              = segmented_tiny_kernel<U, VB, false>;
  if (CAN_FLOAT && a.xf.float_mask != 0)
  {
    kernel = segmented_tiny_kernel<U, VB, CAN_FLOAT>;
  }
  const long long want = (a.num_segments + TINY_WARPS - 1) / TINY_WARPS;
  const long long cap  = (long long) sms * 32;
  kernel<<<unsigned(want < cap ? want : cap), TINY_WARPS * 32, 0, stream>>>(a);
  return cudaPeekAtLastError();
}

template <class U, int VB>
static cudaError_t launch_seg(const SegArgs& a, int sms, cudaStream_t stream)
{
  constexpr int IPT        = SingleTileShape<U, VB>::IPT;
  using L                  = SingleTileSmem<U, VB, IPT>;
  constexpr bool CAN_FLOAT = sizeof(U) >= 2;
  auto kernel              = segmented_sort_kernel<U, VB, IPT, false>;
  if (CAN_FLOAT && a.xf.float_mask != 0)
  {
    kernel = segmented_sort_kernel<U, VB, IPT, CAN_FLOAT>;
  }
  if (L::BYTES > 32 * 1024) // static shared memory (offsets + all-pass histograms) comes on top
  {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(L::BYTES));
    if (e != cudaSuccess)
    {
      return e;
    }
  }
  if (a.tiny_max != 0)
  {
    cudaError_t e = launch_seg_tiny<U, VB>(a, sms, stream);
    if (e != cudaSuccess)
    {
      return e;
    }
  }
  // one CTA per batch of segments (taken round-robin); a CTA sorts the short segments of its batches one by one.  Batch
  // size: as large as leaves >= 16 batches per SM, at most one segment per thread
  SegArgs b               = a;
  const long long per     = a.num_segments / ((long long) sms * 16);
  b.batch                 = uint32_t(per < 1 ? 1 : (per > ST_THREADS ? ST_THREADS : per));
  const long long cap     = (long long) sms * 64;
  const long long batches = (a.num_segments + b.batch - 1) / b.batch;
  const unsigned grid     = unsigned(batches < cap ? batches : cap);
  kernel<<<grid, ST_THREADS, L::BYTES, stream>>>(b);
  return cudaPeekAtLastError();
}

template <class U>
static cudaError_t launch_seg_v(int value_bytes, const SegArgs& a, int sms, cudaStream_t stream)
{
  switch (value_bytes)
  {
    case 0: return launch_seg<U, 0>(a, sms, stream);
    case 1: return launch_seg<U, 1>(a, sms, stream);
    case 2: return launch_seg<U, 2>(a, sms, stream);
    case 4: return launch_seg<U, 4>(a, sms, stream);
    case 8: return launch_seg<U, 8>(a, sms, stream);
    case 16: return launch_seg<U, 16>(a, sms, stream);
    default: return cudaErrorNotSupported;
  }
}

} // namespace b200rs

using namespace b200rs;

static cudaError_t launch_seg_k(int key_bytes, int value_bytes, const SegArgs& a, int sms, cudaStream_t stream)
{
  switch (key_bytes)
  {
    case 1: return launch_seg_v<uint8_t>(value_bytes, a, sms, stream);
    case 2: return launch_seg_v<uint16_t>(value_bytes, a, sms, stream);
    case 4: return launch_seg_v<uint32_t>(value_bytes, a, sms, stream);
    default: return launch_seg_v<uint64_t>(value_bytes, a, sms, stream);
  }
}

// side stream + fork / join events of the calling host thread, one set per device (created on first use, kept)
struct SideStream
{
  cudaStream_t stream;
  cudaEvent_t fork, join;
  bool made;
};
static SideStream* side_stream_of(int dev)
{
  static thread_local SideStream t_side[32] = {};
  if (dev < 0 || dev >= 32)
  {
    return nullptr;
  }
  SideStream& s = t_side[dev];
  if (!s.made)
  {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess
        || cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess
        || cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess)
    {
      (void) cudaGetLastError();
      return nullptr;
    }
    s.made = true;
  }
  return &s;
}

// segments longer than this many items take the whole-grid passes (0 = never); a segment of at most one tile of the
// per-segment kernel is always sorted in shared memory by one CTA whatever the value
static std::atomic<uint32_t> g_seg_long_min{SEG_LONG_MIN_DEFAULT};

// segments of at most this many items are sorted by one warp each (at most TINY_ROWS * 32; 0 = never)
static std::atomic<uint32_t> g_seg_tiny_max{TINY_ROWS * 32};

extern "C" int b200rs_set_segmented_tiny_max(unsigned long long items)
{
  g_seg_tiny_max.store(items > TINY_ROWS * 32 ? uint32_t(TINY_ROWS * 32) : uint32_t(items), std::memory_order_relaxed);
  return 0;
}

extern "C" int b200rs_set_segmented_long_min(unsigned long long items)
{
  g_seg_long_min.store(items > 0xffffffffull ? 0xffffffffu : uint32_t(items), std::memory_order_relaxed);
  return 0;
}

static size_t seg_align(size_t x)
{
  return (x + 255) / 256 * 256;
}

extern "C" int b200rs_segmented_sort(
  void* d_temp_storage,
  size_t* temp_storage_bytes,
  const void* d_keys_in,
  void* d_keys_out,
  const void* d_values_in,
  void* d_values_out,
  uint64_t num_items,
  uint64_t num_segments,
  const void* d_begin_offsets,
  const void* d_end_offsets,
  int offset_bytes,
  int key_kind,
  int key_bytes,
  int value_bytes,
  int begin_bit,
  int end_bit,
  int descending,
  b200rs_stream_t stream_)
{
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (temp_storage_bytes == nullptr || key_kind < 0 || key_kind > 2 || (offset_bytes != 4 && offset_bytes != 8))
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
  if (begin_bit < 0 || end_bit < begin_bit || end_bit > key_bytes * 8)
  {
    return int(cudaErrorInvalidValue);
  }
  const int passes    = (end_bit - begin_bit + RADIX_BITS - 1) / RADIX_BITS;
  const bool need_tmp = passes > 1 && num_items > single_tile_capacity(key_bytes, value_bytes);
  // scratch copies for the passes of segments longer than one tile (the size query cannot know the segment sizes)
  const size_t off_keys = 0;
  const size_t off_vals = need_tmp ? seg_align(size_t(num_items) * key_bytes) : 0;
  size_t total          = need_tmp ? off_vals + seg_align(size_t(num_items) * value_bytes) + 255 : 1;
  // whole-grid passes for segments longer than long_min (segmented_long.cu): a table of at most
  // num_items / long_min disjoint long segments, their bins and the chained-scan rows of all their tiles
  uint32_t long_min = g_seg_long_min.load(std::memory_order_relaxed);
  if (long_min != 0 && long_min < single_tile_capacity(key_bytes, value_bytes))
  {
    long_min = uint32_t(single_tile_capacity(key_bytes, value_bytes)); // one tile is sorted in shared memory, always
  }
  const bool use_long = long_min != 0 && passes > 0 && num_items > long_min && num_items < (1ull << 32)
                     && seg_long_supported(key_bytes, value_bytes);
  const uint32_t max_long    = use_long ? uint32_t(num_items / long_min) + 1u : 0u;
  const uint32_t long_tile   = use_long ? seg_long_tile_items(key_bytes, value_bytes) : 1u;
  const uint32_t tiles_bound = use_long ? uint32_t(num_items / long_tile) + max_long + 1u : 0u;
  const size_t off_ctl  = seg_align(total);
  const size_t off_tab  = off_ctl + seg_align(sizeof(SegLongCtl));
  const size_t off_bins = off_tab + seg_align(size_t(max_long) * sizeof(SegLong));
  const size_t off_lb   = off_bins + seg_align(size_t(max_long) * passes * RADIX * sizeof(unsigned long long));
  const size_t lb_bytes = seg_align(size_t(tiles_bound) * RADIX * sizeof(uint32_t));
  if (use_long)
  {
    total = off_lb + 2 * lb_bytes + 255;
  }
  if (d_temp_storage == nullptr)
  {
    *temp_storage_bytes = total;
    return 0;
  }
  if (*temp_storage_bytes < total)
  {
    return int(cudaErrorInvalidValue);
  }
  if (num_items == 0 || num_segments == 0)
  {
    return 0;
  }
  if (d_keys_in == nullptr || d_keys_out == nullptr || d_begin_offsets == nullptr || d_end_offsets == nullptr
      || (value_bytes > 0 && (d_values_in == nullptr || d_values_out == nullptr)))
  {
    return int(cudaErrorInvalidValue);
  }
  unsigned char* base = reinterpret_cast<unsigned char*>(seg_align(reinterpret_cast<size_t>(d_temp_storage)));
  SegArgs a;
  a.keys_in       = d_keys_in;
  a.keys_out      = d_keys_out;
  a.keys_tmp      = need_tmp ? base + off_keys : nullptr;
  a.vals_in       = value_bytes > 0 ? d_values_in : nullptr;
  a.vals_out      = value_bytes > 0 ? d_values_out : nullptr;
  a.vals_tmp      = (need_tmp && value_bytes > 0) ? base + off_vals : nullptr;
  a.begin_offsets = d_begin_offsets;
  a.end_offsets   = d_end_offsets;
  a.num_segments  = (long long) num_segments;
  a.offset_bytes  = offset_bytes;
  a.begin_bit     = begin_bit;
  a.end_bit       = end_bit;
  a.all_ones      = 0xffffffffu;
  a.xf            = make_xform(key_kind, key_bytes, descending, /*single_tile_rule=*/true);
  a.long_ctl      = use_long ? reinterpret_cast<const SegLongCtl*>(base + off_ctl) : nullptr;
  a.long_min      = long_min;
  a.tiny_max      = g_seg_tiny_max.load(std::memory_order_relaxed);
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess)
  {
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (e != cudaSuccess)
  {
    return int(e);
  }
  if (use_long)
  {
    // the long segments first: their table and overflow flag must exist before the per-segment kernel reads them
    SegLongPlan p;
    p.keys_in       = a.keys_in;
    p.keys_out      = a.keys_out;
    p.keys_tmp      = a.keys_tmp;
    p.vals_in       = a.vals_in;
    p.vals_out      = a.vals_out;
    p.vals_tmp      = a.vals_tmp;
    p.begin_offsets = a.begin_offsets;
    p.end_offsets   = a.end_offsets;
    p.num_segments  = a.num_segments;
    p.offset_bytes  = a.offset_bytes;
    p.begin_bit     = begin_bit;
    p.end_bit       = end_bit;
    p.passes        = passes;
    p.xf            = a.xf;
    p.ctl           = reinterpret_cast<SegLongCtl*>(base + off_ctl);
    p.table         = reinterpret_cast<SegLong*>(base + off_tab);
    p.bins          = reinterpret_cast<unsigned long long*>(base + off_bins);
    p.lookback[0]   = reinterpret_cast<uint32_t*>(base + off_lb);
    p.lookback[1]   = reinterpret_cast<uint32_t*>(base + off_lb + lb_bytes);
    p.zero_bytes    = off_lb + lb_bytes - off_bins;
    p.max_long      = max_long;
    p.tiles_bound   = tiles_bound;
    p.long_min      = long_min;
    p.sms           = sms;
    // The table of the long segments first (the per-segment kernel reads its overflow flag), then the two halves run
    // side by side: the whole-grid passes on a side stream forked from the caller's stream, the per-segment kernel on
    // the caller's stream, joined before returning.  With no long segment the side stream's (empty) launches hide behind
    // the per-segment kernel; with both kinds the short segments fill the tails of the long passes.  Legal under capture.
    e = seg_long_sort(p, 0, key_bytes, value_bytes, stream);
    SideStream* side = e == cudaSuccess ? side_stream_of(dev) : nullptr;
    if (e == cudaSuccess && side == nullptr)
    {
      e = seg_long_sort(p, 1, key_bytes, value_bytes, stream); // no side stream: everything in order on one stream
    }
    else if (e == cudaSuccess)
    {
      if ((e = cudaEventRecord(side->fork, stream)) == cudaSuccess
          && (e = cudaStreamWaitEvent(side->stream, side->fork, 0)) == cudaSuccess)
      {
        e = seg_long_sort(p, 1, key_bytes, value_bytes, side->stream);
      }
      const cudaError_t e2 = launch_seg_k(key_bytes, value_bytes, a, sms, stream);
      cudaError_t e3       = cudaEventRecord(side->join, side->stream);
      if (e3 == cudaSuccess)
      {
        e3 = cudaStreamWaitEvent(stream, side->join, 0);
      }
      return int(e != cudaSuccess ? e : (e2 != cudaSuccess ? e2 : e3));
    }
    if (e != cudaSuccess)
    {
      return int(e);
    }
  }
  return int(launch_seg_k(key_bytes, value_bytes, a, sms, stream));
}
