// Segmented radix sort, LONG segments: whole-grid passes.
//
// segmented.cu gives every segment to one CTA for all its passes -- ideal up to a few tiles, but a handful of very long
// segments would leave most of the GPU idle.  Segments longer than a threshold (b200rs_set_segmented_long_min) take this path instead: the
// device-wide onesweep pass (onesweep.cuh, the same device code as b200rs_sort) run over ALL long segments at once.
// Tiles never straddle segments; every segment has its own digit offsets and its own chained-scan rows, so a launch is
// many independent onesweep problems sharing one grid:
//   classify (segment table of the long segments, built on the device: the host never sees the offsets) -> tile
//   prefix -> per-segment all-pass histograms -> per-segment bin scans -> one launch per 8-bit digit.
// Same buffers and ping-pong rule as segmented.cu (the input is never written, the last pass lands in the output), same
// key transform (the segmented kernel's -0.0 rule).  Reference: one CTA per segment and pass
// (/root/reference/cub/cub/device/dispatch/kernels/kernel_segmented_radix_sort.cuh:113-300).
#include "../../include/b200rs.h"
#include "onesweep.cuh"
#include "segmented_long.h"

namespace b200rs
{

__device__ __forceinline__ long long seg_load_offset(const void* p, long long i, int bytes)
{
  return bytes == 8 ? static_cast<const long long*>(p)[i] : (long long) static_cast<const unsigned int*>(p)[i];
}

__global__ void seg_long_classify_kernel(
  const void* begin_offsets, const void* end_offsets, long long num_segments, int offset_bytes, SegLongCtl* ctl,
  SegLong* table, uint32_t max_long, uint32_t long_min)
{
  const long long seg = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (seg >= num_segments)
  {
    return;
  }
  const long long b   = seg_load_offset(begin_offsets, seg, offset_bytes);
  const long long e   = seg_load_offset(end_offsets, seg, offset_bytes);
  const long long len = offset_bytes == 8 ? e - b : (long long) (int) (uint32_t(e) - uint32_t(b));
  if (len > (long long) long_min)
  {
    const uint32_t idx = atomicAdd(&ctl->n_long, 1u);
    if (idx < max_long && len < (1ll << 30))
    {
      table[idx].begin      = (unsigned long long) b;
      table[idx].len        = uint32_t(len);
      table[idx].first_tile = 0;
    }
    else
    {
      ctl->overflow = 1; // more long segments than disjoint segments can produce, or a segment of 2^30 items and more:
                         // segmented.cu sorts everything
    }
  }
}

// one block: first_tile = exclusive prefix of the segments' tile counts, and the total
__global__ void __launch_bounds__(1024) seg_long_prefix_kernel(SegLongCtl* ctl, SegLong* table, uint32_t tile_items)
{
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  const uint32_t n    = ctl->n_long;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0)
  {
    carry = 0;
  }
  __syncthreads();
  if (ctl->overflow != 0)
  {
    return;
  }
  for (uint32_t base = 0; base < n; base += 1024)
  {
    const uint32_t i     = base + threadIdx.x;
    const uint32_t tiles = i < n ? (table[i].len + tile_items - 1) / tile_items : 0u;
    uint32_t incl        = tiles;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1)
    {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, s);
      incl += lane >= uint32_t(s) ? up : 0u;
    }
    if (lane == 31)
    {
      warp_sums[warp] = incl;
    }
    __syncthreads();
    uint32_t before = carry;
    for (uint32_t w = 0; w < warp; ++w)
    {
      before += warp_sums[w];
    }
    if (i < n)
    {
      table[i].first_tile = before + incl - tiles;
    }
    __syncthreads();
    if (threadIdx.x == 1023)
    {
      carry = before + incl;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0)
  {
    ctl->total_tiles = carry;
  }
}

// segment of a global tile id: last entry whose first_tile <= tile
__device__ __forceinline__ uint32_t seg_of_tile(const SegLong* table, uint32_t n, uint32_t tile)
{
  uint32_t lo = 0, hi = n; // first_tile is non-decreasing
  while (hi - lo > 1)
  {
    const uint32_t mid = (lo + hi) >> 1;
    if (table[mid].first_tile <= tile)
    {
      lo = mid;
    }
    else
    {
      hi = mid;
    }
  }
  return lo;
}

// all-pass digit histograms of the long segments, one tile of `tile_items` keys per CTA iteration
template <class U, bool FLOATK>
__global__ void __launch_bounds__(256) seg_long_hist_kernel(
  const U* keys, const SegLongCtl* ctl, const SegLong* table, unsigned long long* bins, int passes, int begin_bit,
  int end_bit, const KeyXform kx, uint32_t tile_items)
{
  __shared__ uint32_t h[8 * RADIX];
  __shared__ uint32_t s_seg;
  if (ctl->overflow != 0)
  {
    return;
  }
  const uint32_t total = ctl->total_tiles, n_long = ctl->n_long;
  const XformT<U> xf(kx);
  const U neg_zero = U(kx.neg_zero), pos_zero = U(kx.pos_zero);
  for (uint32_t tile = blockIdx.x; tile < total; tile += gridDim.x)
  {
    for (int i = threadIdx.x; i < passes * RADIX; i += 256)
    {
      h[i] = 0;
    }
    if (threadIdx.x == 0)
    {
      s_seg = seg_of_tile(table, n_long, tile);
    }
    __syncthreads();
    const uint32_t s    = s_seg;
    const SegLong sg    = table[s];
    const uint32_t off  = (tile - sg.first_tile) * tile_items;
    const uint32_t cnt  = sg.len - off < tile_items ? sg.len - off : tile_items;
    const U* k          = keys + sg.begin + off;
    for (uint32_t i = threadIdx.x; i < cnt; i += 256)
    {
      const U t = twiddle_in(k[i], xf);
      for (int p = 0; p < passes; ++p)
      {
        const int bit   = begin_bit + p * RADIX_BITS;
        const int nbits = (end_bit - bit) < RADIX_BITS ? (end_bit - bit) : RADIX_BITS;
        atomicAdd(&h[p * RADIX + pass_digit<FLOATK>(t, bit, (1u << nbits) - 1u, neg_zero, pos_zero)], 1u);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RADIX; i += 256)
    {
      if (h[i] != 0)
      {
        atomicAdd(&bins[size_t(s) * passes * RADIX + i], (unsigned long long) h[i]);
      }
    }
    __syncthreads();
  }
}

// in-place exclusive scan of one (segment, pass) row of 256 bins per block
__global__ void __launch_bounds__(RADIX) seg_long_scan_kernel(const SegLongCtl* ctl, unsigned long long* bins, int passes)
{
  __shared__ unsigned long long wsum[RADIX / 32];
  if (blockIdx.x / passes >= ctl->n_long || ctl->overflow != 0)
  {
    return;
  }
  unsigned long long* b      = bins + size_t(blockIdx.x) * RADIX;
  const uint32_t lane        = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long c = b[threadIdx.x];
  unsigned long long incl    = c;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1)
  {
    const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, s);
    incl += lane >= uint32_t(s) ? up : 0ull;
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

// One 8-bit digit pass over every long segment: a CTA takes the next global tile, finds its segment and runs the
// unsegmented tile code on that segment's arrays, offsets and chained-scan rows.
template <class U, int VB, int NT, int IPT, int MINB, int OPT, bool FLOATK>
__global__ void __launch_bounds__(NT, MINB) seg_onesweep_kernel(
  const PassArgs a, const SegLongCtl* ctl, const SegLong* table, int pass, int passes)
{
  using L = OnesweepSmem<U, VB, NT, IPT, OPT>;
  using V = typename value_of<VB>::type;
  constexpr int TILE = L::TILE;
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = uint32_t(__cvta_generic_to_shared(smem));
  const uint32_t tid   = threadIdx.x;
  if (ctl->overflow != 0)
  {
    return;
  }
  const uint32_t total = ctl->total_tiles;
  if (blockIdx.x >= total)
  {
    return; // the grid is sized for the worst case; exactly `total` CTAs stay and take the tile ids 0 .. total - 1
  }
  if (tid == 0)
  {
    // dynamic tile ids: a tile only starts after all its predecessors (of every segment) started
    const uint32_t t = atomicAdd(a.tile_counter, 1u);
    sts32(sbase + L::OFF_MISC + 32, t);
    sts32(sbase + L::OFF_MISC + 44, 0);
    sts32(sbase + L::OFF_MISC + 36, t < total ? seg_of_tile(table, ctl->n_long, t) : 0u);
  }
  {
    constexpr int WORDS = L::NW * RADIX * L::CTR_BYTES / 4;
#pragma unroll
    for (int j = 0; j < WORDS / NT; ++j)
    {
      sts32(sbase + L::OFF_WARP + (j * NT + tid) * 4, 0);
    }
  }
  __syncthreads();
  const uint32_t tile = lds32(sbase + L::OFF_MISC + 32);
  // the chained-scan rows of the NEXT pass are zeroed by this launch (every CTA, also the ones without a tile)
  if (tid < RADIX && a.lookback_next != nullptr)
  {
    for (uint32_t t = tile; t < total; t += gridDim.x)
    {
      a.lookback_next[size_t(t) * RADIX + tid] = 0;
    }
  }
  if (tile >= total)
  {
    return;
  }
  const uint32_t s = lds32(sbase + L::OFF_MISC + 36);
  const SegLong sg = table[s];
  PassArgs b       = a;
  b.keys_in        = static_cast<const U*>(a.keys_in) + sg.begin;
  b.keys_out       = static_cast<U*>(a.keys_out) + sg.begin;
  if (VB > 0)
  {
    b.vals_in  = static_cast<const V*>(a.vals_in) + sg.begin;
    b.vals_out = static_cast<V*>(a.vals_out) + sg.begin;
  }
  b.lookback      = a.lookback + size_t(sg.first_tile) * RADIX;
  b.lookback_next = nullptr;
  b.bins          = a.bins + (size_t(s) * passes + pass) * RADIX;
  b.bins_next     = nullptr;
  b.num_items     = sg.len;
  const uint32_t t_in      = tile - sg.first_tile;
  const uint32_t tile_base = t_in * uint32_t(TILE);
  const uint32_t valid     = min(uint32_t(TILE), sg.len - tile_base);
  if (valid == uint32_t(TILE))
  {
    onesweep_tile<U, VB, NT, IPT, RANK_BALLOT, OPT, FLOATK, false, true>(b, sbase, t_in, tile_base, valid);
  }
  else
  {
    onesweep_tile<U, VB, NT, IPT, RANK_BALLOT, OPT, FLOATK, false, false>(b, sbase, t_in, tile_base, valid);
  }
}

constexpr int SEG_BASE = OPT_FMA_NOT | OPT_LB_WINDOW | OPT_CTR16;
// tile shape per (key bytes, value bytes): the defaults of the unsegmented tables (inst_k*.cu)
template <class U, int VB>
struct SegLongShape;
template <>
struct SegLongShape<uint32_t, 0>
{
  static constexpr int IPT = 44, MINB = 3, OPT = SEG_BASE | OPT_SHORT_WARP | OPT_FOLD | OPT_FOLD_PTR;
};
template <>
struct SegLongShape<uint32_t, 4>
{
  static constexpr int IPT = 32, MINB = 3, OPT = SEG_BASE | OPT_SHORT_WARP | OPT_FOLD | OPT_FOLD_PTR;
};
template <>
struct SegLongShape<uint64_t, 0>
{
  static constexpr int IPT = 20, MINB = 4, OPT = SEG_BASE;
};
template <>
struct SegLongShape<uint64_t, 4>
{
  static constexpr int IPT = 20, MINB = 3, OPT = SEG_BASE | OPT_SHORT_WARP;
};

template <>
struct SegLongShape<uint32_t, 8>
{
  static constexpr int IPT = 20, MINB = 3, OPT = SEG_BASE;
};
template <>
struct SegLongShape<uint64_t, 8>
{
  static constexpr int IPT = 16, MINB = 3, OPT = SEG_BASE;
};

// phase 0: the table of the long segments (the per-segment kernel needs its overflow flag); phase 1: everything else
template <class U, int VB>
static cudaError_t seg_long_run(const SegLongPlan& p, int phase, cudaStream_t stream)
{
  using S            = SegLongShape<U, VB>;
  constexpr int NT   = 256;
  using L            = OnesweepSmem<U, VB, NT, S::IPT, S::OPT>;
  const bool flt     = p.xf.float_mask != 0;
  const int passes   = p.passes;
  cudaError_t e = cudaSuccess;
  if (phase == 0)
  {
    if ((e = cudaMemsetAsync(p.ctl, 0, sizeof(SegLongCtl), stream)) != cudaSuccess)
    {
      return e;
    }
    seg_long_classify_kernel<<<unsigned((p.num_segments + 255) / 256), 256, 0, stream>>>(
      p.begin_offsets, p.end_offsets, p.num_segments, p.offset_bytes, p.ctl, p.table, p.max_long, p.long_min);
    return cudaPeekAtLastError();
  }
  // one memset clears every bin and the first pass's chained-scan rows (contiguous: bins | lookback[0])
  if ((e = cudaMemsetAsync(p.bins, 0, p.zero_bytes, stream)) != cudaSuccess)
  {
    return e;
  }
  seg_long_prefix_kernel<<<1, 1024, 0, stream>>>(p.ctl, p.table, uint32_t(L::TILE));
  const unsigned hist_grid = p.tiles_bound < unsigned(p.sms) * 8u ? p.tiles_bound : unsigned(p.sms) * 8u;
  if (flt)
  {
    seg_long_hist_kernel<U, true><<<hist_grid, 256, 0, stream>>>(
      static_cast<const U*>(p.keys_in), p.ctl, p.table, p.bins, passes, p.begin_bit, p.end_bit, p.xf, uint32_t(L::TILE));
  }
  else
  {
    seg_long_hist_kernel<U, false><<<hist_grid, 256, 0, stream>>>(
      static_cast<const U*>(p.keys_in), p.ctl, p.table, p.bins, passes, p.begin_bit, p.end_bit, p.xf, uint32_t(L::TILE));
  }
  seg_long_scan_kernel<<<p.max_long * unsigned(passes), RADIX, 0, stream>>>(p.ctl, p.bins, passes);
  if ((e = cudaPeekAtLastError()) != cudaSuccess)
  {
    return e;
  }
  auto kernel = flt ? seg_onesweep_kernel<U, VB, NT, S::IPT, S::MINB, S::OPT, true>
                    : seg_onesweep_kernel<U, VB, NT, S::IPT, S::MINB, S::OPT, false>;
  if (L::BYTES > 48 * 1024)
  {
    if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(L::BYTES))) != cudaSuccess)
    {
      return e;
    }
  }
  const bool identity = p.xf.float_mask == 0 && p.xf.sign_mask == 0 && p.xf.desc_mask == 0;
  const void* src_k   = p.keys_in;
  const void* src_v   = p.vals_in;
  for (int pass = 0; pass < passes; ++pass)
  {
    // never write the input; the last pass lands in the output (dispatch_radix_sort.cuh:1850-1857)
    const bool to_out = ((passes - 1 - pass) % 2) == 0;
    void* dst_k       = to_out ? p.keys_out : p.keys_tmp;
    void* dst_v       = to_out ? p.vals_out : p.vals_tmp;
    const int bit     = p.begin_bit + pass * RADIX_BITS;
    const int nbits   = (p.end_bit - bit) < RADIX_BITS ? (p.end_bit - bit) : RADIX_BITS;
    PassArgs a;
    a.keys_in             = src_k;
    a.keys_out            = dst_k;
    a.vals_in             = VB > 0 ? src_v : nullptr;
    a.vals_out            = VB > 0 ? dst_v : nullptr;
    // chained-scan rows ping-pong between two arrays: this launch zeroes the rows of the next one
    a.lookback            = p.lookback[pass & 1];
    a.lookback_next       = pass + 1 < passes ? p.lookback[(pass + 1) & 1] : nullptr;
    a.lookback_next_tiles = 0;
    a.tile_counter        = &p.ctl->tile_counter[pass];
    a.bins                = p.bins;
    a.bins_next           = nullptr;
    a.num_items           = 0;
    a.num_tiles           = p.tiles_bound;
    a.all_ones            = 0xffffffffu;
    a.shift               = bit;
    a.mask                = (1u << nbits) - 1u;
    a.first_pass          = pass == 0 && !identity;
    a.last_pass           = pass == passes - 1 && !identity;
    a.big                 = 0;
    a.xf                  = p.xf;
    a.num_splitters       = 0;
    a.peer                = nullptr;
    a.plan                = nullptr;
    a.sm_count            = p.sms;
    a.zero_flag           = nullptr;
    kernel<<<p.tiles_bound, NT, L::BYTES, stream>>>(a, p.ctl, p.table, pass, passes);
    if ((e = cudaPeekAtLastError()) != cudaSuccess)
    {
      return e;
    }
    src_k = dst_k;
    src_v = dst_v;
  }
  return cudaSuccess;
}

bool seg_long_supported(int key_bytes, int value_bytes)
{
  return (key_bytes == 4 || key_bytes == 8) && (value_bytes == 0 || value_bytes == 4 || value_bytes == 8);
}

template <class U>
static uint32_t tile_items_of(int value_bytes)
{
  return 256u * uint32_t(value_bytes == 0   ? SegLongShape<U, 0>::IPT
                         : value_bytes == 4 ? SegLongShape<U, 4>::IPT
                                            : SegLongShape<U, 8>::IPT);
}

uint32_t seg_long_tile_items(int key_bytes, int value_bytes)
{
  return key_bytes == 4 ? tile_items_of<uint32_t>(value_bytes) : tile_items_of<uint64_t>(value_bytes);
}

template <class U>
static cudaError_t seg_long_sort_v(const SegLongPlan& p, int phase, int value_bytes, cudaStream_t stream)
{
  return value_bytes == 0   ? seg_long_run<U, 0>(p, phase, stream)
         : value_bytes == 4 ? seg_long_run<U, 4>(p, phase, stream)
                            : seg_long_run<U, 8>(p, phase, stream);
}

cudaError_t seg_long_sort(const SegLongPlan& p, int phase, int key_bytes, int value_bytes, cudaStream_t stream)
{
  return key_bytes == 4 ? seg_long_sort_v<uint32_t>(p, phase, value_bytes, stream)
                        : seg_long_sort_v<uint64_t>(p, phase, value_bytes, stream);
}

} // namespace b200rs
