// The one-CTA sort of at most one tile as a device function: used by single_tile.cu (whole problem) and segmented.cu
// (one segment).  See single_tile.cu for the description and the reference citations.
#pragma once

#include "configs.h"
#include "onesweep.cuh"

namespace b200rs
{

constexpr int ST_THREADS = 256;

template <class U, int VBYTES, int IPT>
struct SingleTileSmem
{
  static constexpr int NW   = ST_THREADS / 32;
  static constexpr int TILE = ST_THREADS * IPT;
  static constexpr uint32_t OFF_CNT  = 0;                              // u16 [NW][256]
  static constexpr uint32_t OFF_MISC = OFF_CNT + NW * RADIX * 2;       // u32 [16]
  static constexpr uint32_t OFF_KEYS = OFF_MISC + 64;
  static constexpr uint32_t OFF_VALS = OFF_KEYS + ((TILE * uint32_t(sizeof(U)) + 15) / 16) * 16;
  static constexpr size_t BYTES      = size_t(OFF_VALS) + size_t(TILE) * VBYTES;
};

struct SingleTileArgs
{
  const void* keys_in;
  void* keys_out;
  const void* vals_in;
  void* vals_out;
  uint32_t num_items;
  int begin_bit;
  int end_bit;
  uint32_t all_ones;
  KeyXform xf;
};

// The whole sort of one tile (a.num_items <= TILE) by the calling CTA; sbase = shared-window address of L::BYTES bytes.
template <class U, int VBYTES, int IPT, bool FLOATK>
__device__ __forceinline__ void single_tile_sort(const SingleTileArgs& a, const uint32_t sbase)
{
  using L = SingleTileSmem<U, VBYTES, IPT>;
  using V = typename value_of<VBYTES>::type;
  constexpr int NW = L::NW;
  const uint32_t tid    = threadIdx.x;
  const uint32_t lane   = tid & 31;
  const uint32_t warp   = tid >> 5;
  const uint32_t s_cnt  = sbase + L::OFF_CNT;
  const uint32_t s_misc = sbase + L::OFF_MISC;
  const uint32_t s_keys = sbase + L::OFF_KEYS;
  const uint32_t s_vals = sbase + L::OFF_VALS;
  const uint32_t s_mine = s_cnt + warp * (RADIX * 2);
  const uint32_t n      = a.num_items;
  const XformT<U> xf(a.xf);
  const U neg_zero = U(a.xf.neg_zero);
  const U pos_zero = U(a.xf.pos_zero);

  U key[IPT];
  V val[VBYTES > 0 ? IPT : 1];
  // A partial tile is spread evenly over the warps: `rows` = ceil(n / 256) rows of 32 items per warp instead of IPT, so a
  // 1 000-item segment costs every warp 4 ranked rows -- not warp 0 twenty and six warps twenty rows of padding.  Warp w
  // still owns a contiguous chunk, so ranks follow (warp, row, lane) == input order.  Rows >= `rows` are skipped whole.
  const uint32_t rows  = (n + ST_THREADS - 1) / ST_THREADS; // <= IPT, warp-uniform
  const uint32_t chunk = warp * 32 * rows + lane;
#pragma unroll
  for (int i = 0; i < IPT; ++i)
  {
    const uint32_t p = chunk + i * 32;
    key[i]           = U(~U(0)); // padding sorts last
    if (uint32_t(i) < rows && p < n)
    {
      key[i] = twiddle_in(static_cast<const U*>(a.keys_in)[p], xf);
      if (VBYTES > 0)
      {
        val[i] = static_cast<const V*>(a.vals_in)[p];
      }
    }
  }

  const uint32_t lt_mask = lanemask_lt();
  const uint32_t gt_mask = lanemask_gt();
  for (int bit = a.begin_bit; bit < a.end_bit; bit += RADIX_BITS)
  {
    const int nbits     = (a.end_bit - bit) < RADIX_BITS ? (a.end_bit - bit) : RADIX_BITS;
    const uint32_t mask = (1u << nbits) - 1u;
#pragma unroll
    for (int j = 0; j < NW * RADIX * 2 / 4 / ST_THREADS; ++j)
    {
      sts32(s_cnt + (j * ST_THREADS + tid) * 4, 0);
    }
    __syncthreads(); // also: everybody has read back the previous pass's staged items

    uint32_t rank[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      if (uint32_t(i) < rows)
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
        __syncwarp(); // the next row's loads of this counter come after the leader's store (memory model, racecheck)
        rank[i] = next - 1;
      }
    }
    __syncthreads();

    // digit totals, exclusive scan over digits, per-warp bases
    uint32_t total = 0, incl = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w)
    {
      total += ctr_ld<true>(s_cnt + (w * RADIX + tid) * 2);
    }
    incl = total;
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
    uint32_t run = incl - total;
#pragma unroll
    for (int w = 0; w < NW; ++w)
    {
      const uint32_t ws = lds32(s_misc + w * 4);
      run += (uint32_t(w) < warp) ? ws : 0u;
    }
#pragma unroll
    for (int w = 0; w < NW; ++w)
    {
      const uint32_t addr = s_cnt + (w * RADIX + tid) * 2;
      const uint32_t c    = ctr_ld<true>(addr);
      ctr_st<true>(addr, run);
      run += c;
    }
    __syncthreads();

    // stage in digit order, read back in tile order
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      if (uint32_t(i) < rows)
      {
        const uint32_t d = pass_digit<FLOATK>(key[i], bit, mask, neg_zero, pos_zero);
        const uint32_t r = rank[i] + ctr_ld<true>(s_mine + d * 2);
        sts_t<U>(s_keys + r * uint32_t(sizeof(U)), key[i]);
        if (VBYTES > 0)
        {
          sts_t<V>(s_vals + r * uint32_t(sizeof(V)), val[i]);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      if (uint32_t(i) < rows)
      {
        const uint32_t p = chunk + i * 32;
        key[i]           = lds_t<U>(s_keys + p * uint32_t(sizeof(U)));
        if (VBYTES > 0)
        {
          val[i] = lds_t<V>(s_vals + p * uint32_t(sizeof(V)));
        }
      }
    }
  }

#pragma unroll
  for (int i = 0; i < IPT; ++i)
  {
    const uint32_t p = chunk + i * 32;
    if (uint32_t(i) < rows && p < n)
    {
      static_cast<U*>(a.keys_out)[p] = twiddle_out(key[i], xf);
      if (VBYTES > 0)
      {
        static_cast<V*>(a.vals_out)[p] = val[i];
      }
    }
  }
}

template <class U, int VBYTES, int IPT, bool FLOATK>
__global__ void __launch_bounds__(ST_THREADS) single_tile_kernel(const SingleTileArgs a)
{
  extern __shared__ __align__(16) unsigned char st_smem[];
  single_tile_sort<U, VBYTES, IPT, FLOATK>(a, uint32_t(__cvta_generic_to_shared(st_smem)));
}

// items per thread by dominant item size: 20 (<= 4 bytes), 10 (8 bytes), 5 (16 bytes)
template <class U, int VB>
struct SingleTileShape
{
  static constexpr int DOM = int(sizeof(U)) > VB ? int(sizeof(U)) : VB;
  static constexpr int IPT = DOM <= 4 ? 20 : (DOM <= 8 ? 10 : 5);
};

} // namespace b200rs
