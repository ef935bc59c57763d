// One stable 8-bit digit pass ("onesweep") for sm_100a.
//
// Replaces DeviceRadixSortOnesweepKernel / AgentRadixSortOnesweep
// (/root/reference/cub/cub/device/dispatch/kernels/kernel_radix_sort.cuh:498-558,
//  cub/cub/agent/agent_radix_sort_onesweep.cuh:152-739) and the ranking it calls
// (cub/cub/block/block_radix_rank.cuh:913-1213).  Same contract seen from outside: keys_out/vals_out
// receive the items of keys_in/vals_in stably partitioned by the digit (key >> shift) & mask, starting at
// the per-digit global offsets in `bins`.  The inside is a different design:
//
//   * no "early counts" histogram pre-pass over the tile: the warp-private running offsets that ranking
//     maintains ARE the per-warp digit histograms once the last item is ranked, so every key costs one
//     match + one shared load + (leader only) one shared store;
//   * ranking with MATCH.ANY (one instruction) instead of an 8-ballot loop (template switch keeps the
//     ballot variant for measurement);
//   * the chained scan status words of the NEXT pass are zeroed by this pass (no memset between passes);
//   * keys stay bit-ordered in HBM between passes (transform fused into first load / last store);
//   * 64-bit per-digit output bases in shared memory, so one launch addresses arrays beyond 2^32 items.
//
// Stable order inside a tile: warp w owns the contiguous chunk [w*32*IPT, (w+1)*32*IPT); its item i of
// lane l is element i*32+l of the chunk (each load instruction covers one contiguous 32-key run, fully
// coalesced).  Ranks follow (warp, item, lane) == input position order.
#pragma once

#include "common.cuh"

namespace b200rs
{

enum RankAlgo
{
  RANK_MATCH  = 0,
  RANK_BALLOT = 1
};

template <class U, int VBYTES, int NT, int IPT>
struct OnesweepSmem
{
  static constexpr int NW          = NT / 32;
  static constexpr int TILE        = NT * IPT;
  static constexpr int ITEM_BYTES  = int(sizeof(U)) > VBYTES ? int(sizeof(U)) : VBYTES;
  static constexpr size_t OFF_WARP = 0;                                        // u32 [NW][256]
  static constexpr size_t OFF_GOFF = OFF_WARP + size_t(NW) * RADIX * 4;        // u64 [256]
  static constexpr size_t OFF_MISC = OFF_GOFF + size_t(RADIX) * 8;             // u32 [16]
  static constexpr size_t OFF_DATA = OFF_MISC + 64;                            // staged tile
  static constexpr size_t BYTES    = OFF_DATA + size_t(TILE) * ITEM_BYTES;
};

template <class U, int VBYTES, int NT, int IPT, int RANK, int MINB>
__global__ void __launch_bounds__(NT, MINB) onesweep_kernel(const PassArgs a)
{
  using L = OnesweepSmem<U, VBYTES, NT, IPT>;
  using V = typename value_of<VBYTES>::type;
  constexpr int NW   = L::NW;
  constexpr int TILE = L::TILE;
  static_assert(NT >= RADIX && NT % 32 == 0, "one thread per digit is required");

  extern __shared__ __align__(16) unsigned char smem[];
  uint32_t* warp_off        = reinterpret_cast<uint32_t*>(smem + L::OFF_WARP);
  unsigned long long* goff  = reinterpret_cast<unsigned long long*>(smem + L::OFF_GOFF);
  uint32_t* misc            = reinterpret_cast<uint32_t*>(smem + L::OFF_MISC);
  U* skeys                  = reinterpret_cast<U*>(smem + L::OFF_DATA);
  V* svals                  = reinterpret_cast<V*>(smem + L::OFF_DATA);

  const XformT<U> xf(a.xf);
  const uint32_t tid  = threadIdx.x;
  const uint32_t lane = tid & 31;
  const uint32_t warp = tid >> 5;
  const int shift     = a.shift;
  const uint32_t dmask = a.mask;

  // ---- dynamic tile id: a tile only starts after all its predecessors started (look-back cannot deadlock)
  if (tid == 0)
  {
    misc[8] = atomicAdd(a.tile_counter, 1u);
  }
  uint32_t* my_off = warp_off + warp * RADIX;
#pragma unroll
  for (int j = 0; j < RADIX / 32; ++j)
  {
    my_off[j * 32 + lane] = 0;
  }
  __syncthreads();
  const uint32_t tile      = misc[8];
  const uint32_t tile_base = tile * uint32_t(TILE);
  const uint32_t valid     = min(uint32_t(TILE), a.num_items - tile_base);
  const bool full          = valid == uint32_t(TILE);

  // ---- load keys, warp-striped
  U key[IPT];
  const uint32_t chunk = warp * 32 * IPT + lane;
  {
    const U* kin = static_cast<const U*>(a.keys_in) + tile_base + chunk;
    if (full)
    {
#pragma unroll
      for (int i = 0; i < IPT; ++i)
      {
        key[i] = kin[i * 32];
      }
      if (a.first_pass)
      {
#pragma unroll
        for (int i = 0; i < IPT; ++i)
        {
          key[i] = twiddle_in(key[i], xf);
        }
      }
    }
    else
    {
#pragma unroll
      for (int i = 0; i < IPT; ++i)
      {
        const bool ok = chunk + i * 32 < valid;
        U k           = ok ? kin[i * 32] : U(0);
        if (a.first_pass)
        {
          k = twiddle_in(k, xf);
        }
        key[i] = ok ? k : U(~U(0)); // padding ranks last: max digit, last in tile order
      }
    }
  }

  // ---- rank: warp-private running digit offsets
  uint32_t rank[IPT];
  const uint32_t lt_mask = lanemask_lt();
#pragma unroll
  for (int i = 0; i < IPT; ++i)
  {
    const uint32_t d = digit_of(digit_view(key[i], xf), shift, dmask);
    uint32_t peers;
    if (RANK == RANK_MATCH)
    {
      peers = __match_any_sync(0xffffffffu, d);
    }
    else
    {
      peers = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < RADIX_BITS; ++b)
      {
        const bool bit      = (d >> b) & 1;
        const uint32_t vote = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? vote : ~vote;
      }
    }
    const uint32_t before = __popc(peers & lt_mask);
    const uint32_t off    = my_off[d];
    __syncwarp();
    if (before == 0)
    {
      my_off[d] = off + __popc(peers);
    }
    __syncwarp();
    rank[i] = off + before;
  }
  __syncthreads();

  // ---- per-digit tile totals (one thread per digit), publish, block-wide exclusive scan over digits
  uint32_t total = 0, excl = 0;
  uint32_t wcount[NW];
  uint32_t* lb_word = a.lookback + size_t(tile) * RADIX + tid;
  if (tid < RADIX)
  {
#pragma unroll
    for (int w = 0; w < NW; ++w)
    {
      wcount[w] = warp_off[w * RADIX + tid];
      total += wcount[w];
    }
    st_relaxed_u32(lb_word, (tile == 0 ? LB_INCLUSIVE : LB_PARTIAL) | total);
    uint32_t incl = total;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1)
    {
      const uint32_t n = __shfl_up_sync(0xffffffffu, incl, s);
      if (lane >= uint32_t(s))
      {
        incl += n;
      }
    }
    if (lane == 31)
    {
      misc[warp] = incl;
    }
    excl = incl - total;
  }
  __syncthreads();
  if (tid < RADIX)
  {
#pragma unroll
    for (int w = 0; w < RADIX / 32; ++w)
    {
      excl += (uint32_t(w) < warp) ? misc[w] : 0u;
    }
    uint32_t run = excl;
#pragma unroll
    for (int w = 0; w < NW; ++w)
    {
      warp_off[w * RADIX + tid] = run;
      run += wcount[w];
    }
  }
  __syncthreads();

  // ---- stage keys in shared memory in digit order
#pragma unroll
  for (int i = 0; i < IPT; ++i)
  {
    const uint32_t d = digit_of(digit_view(key[i], xf), shift, dmask);
    rank[i] += my_off[d];
    skeys[rank[i]] = key[i];
  }

  // values are fetched now so their latency hides behind the look-back
  V val[VBYTES > 0 ? IPT : 1];
  if (VBYTES > 0)
  {
    const V* vin = static_cast<const V*>(a.vals_in) + tile_base + chunk;
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      if (full || chunk + i * 32 < valid)
      {
        val[i] = vin[i * 32];
      }
    }
  }

  // ---- decoupled look-back over predecessor tiles (one thread per digit)
  if (tid < RADIX)
  {
    uint32_t prefix = 0;
    if (tile > 0)
    {
      const uint32_t* w = lb_word - RADIX;
      while (true)
      {
        const uint32_t s = ld_relaxed_u32(w);
        if ((s & LB_FLAG_MASK) == 0)
        {
          continue; // predecessor has started (dynamic tile ids) but not published yet
        }
        prefix += s & LB_VALUE_MASK;
        if (s & LB_INCLUSIVE)
        {
          break;
        }
        w -= RADIX;
      }
      st_relaxed_u32(lb_word, LB_INCLUSIVE | (prefix + total));
    }
    const unsigned long long gbase = a.bins[tid] + prefix;
    goff[tid]                      = gbase - excl;
    if (a.bins_next != nullptr && tile_base + valid == a.num_items)
    {
      a.bins_next[tid] = gbase + total;
    }
    if (a.lookback_next != nullptr)
    {
      for (uint32_t t = tile; t < a.lookback_next_tiles; t += gridDim.x)
      {
        a.lookback_next[size_t(t) * RADIX + tid] = 0;
      }
    }
  }
  __syncthreads();

  // ---- coalesced scatter: consecutive threads write consecutive staged positions
  uint32_t digs[(IPT + 3) / 4];
  U* kout = static_cast<U*>(a.keys_out);
#pragma unroll
  for (int i = 0; i < IPT; ++i)
  {
    const uint32_t pos = i * NT + tid;
    uint32_t d         = 0;
    if (full || pos < valid)
    {
      const U k = skeys[pos];
      d         = digit_of(digit_view(k, xf), shift, dmask);
      kout[goff[d] + pos] = a.last_pass ? twiddle_out(k, xf) : k;
    }
    if (VBYTES > 0)
    {
      if ((i & 3) == 0)
      {
        digs[i / 4] = 0;
      }
      digs[i / 4] |= d << (8 * (i & 3));
    }
  }

  if (VBYTES > 0)
  {
    __syncthreads(); // staged keys are dead
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      if (full || chunk + i * 32 < valid)
      {
        svals[rank[i]] = val[i];
      }
    }
    __syncthreads();
    V* vout = static_cast<V*>(a.vals_out);
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      const uint32_t pos = i * NT + tid;
      if (full || pos < valid)
      {
        const uint32_t d    = (digs[i / 4] >> (8 * (i & 3))) & 0xffu;
        vout[goff[d] + pos] = svals[pos];
      }
    }
  }
}

} // namespace b200rs
