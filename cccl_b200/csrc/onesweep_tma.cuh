// One stable 8-bit digit pass for sm_100a whose scatter is done by the TMA engine ("onesweep, bulk-store variant").
//
// Same contract as onesweep.cuh (replaces DeviceRadixSortOnesweepKernel / AgentRadixSortOnesweep,
// /root/reference/cub/cub/device/dispatch/kernels/kernel_radix_sort.cuh:498-558,
// cub/cub/agent/agent_radix_sort_onesweep.cuh:152-739, ranking cub/cub/block/block_radix_rank.cuh:913-1213):
// keys_out/vals_out receive keys_in/vals_in stably partitioned by the digit (key >> shift) & mask, starting at
// the per-digit global offsets in `bins`.
//
// Why a second kernel: ncu (profiles/r1a_onesweep_ncu.txt) shows the pass is bound ON CHIP -- the half-rate integer
// ALU pipe (75 %) and shared-memory wavefronts (74 %) -- with DRAM at 32 %.  The per-key scatter loop of the classic
// design (LDS key, re-extract digit, LDS digit offset, 64-bit address, STG: 8 instructions and ~3.5 LSU wavefronts per
// key) is the part that needs no arithmetic at all: after staging, each digit's keys form ONE contiguous run in shared
// memory that goes to ONE contiguous range of the output.  So the runs are handed to the TMA engine:
//
//   * one thread per digit issues one `cp.async.bulk.global.shared::cta` for the 16-byte-aligned middle of its run and
//     stores the (< 16 bytes) head and tail itself; measured on B200 (tools/ubench/tma_scatter.cu): 128-byte runs
//     sustain 6.3 TB/s of writes, 256-byte runs 6.9 TB/s, i.e. the copy engine is never the limiter;
//   * bulk copies need source and destination 16-byte aligned, so run d is staged at a shared-memory position with the
//     same phase (mod 16 bytes) as its global destination.  The destination is only known after the decoupled
//     look-back, therefore the order is rank -> digit totals -> publish -> look-back -> padded scan -> stage -> store
//     (the classic kernel stages first and looks back afterwards);
//   * the look-back polls a window of LBW predecessors per round trip (independent loads in flight) instead of one;
//   * the per-lane complement in match-by-ballot is a predicated IMAD (x * -1 + -1 == ~x) so eight of the sixteen
//     logic operations per key move from the ALU pipe to the idle FMA pipe;
//   * values take the same route through their own staging buffer, so the key copies are still being read by the TMA
//     engine while values are staged;
//   * output offsets are per digit, 64-bit, held in the digit thread's registers: no "big" variant, no per-key offset
//     arithmetic.
//
// Requirements checked by the host (dispatch.cu): keys_out / vals_out 16-byte aligned; key size 4 or 8 bytes; value
// size 0, 4, 8 or 16 bytes.  Everything else runs on the classic kernel.
//
// Stable order inside a tile: warp w owns the contiguous chunk [w*32*IPT, (w+1)*32*IPT); its item i of lane l is
// element i*32+l of the chunk.  Ranks follow (warp, item, lane) == input position order.
#pragma once

#include "onesweep.cuh"

namespace b200rs
{

template <class U, int VBYTES, int NT, int IPT>
struct TmaSmem
{
  static constexpr int NW    = NT / 32;
  static constexpr int TILE  = NT * IPT;
  static constexpr int A_K   = 16 / int(sizeof(U));                    // keys per 16 bytes
  static constexpr int A_V   = VBYTES == 0 ? 1 : (VBYTES >= 16 ? 1 : 16 / VBYTES);
  static constexpr int AMAX  = A_K > A_V ? A_K : A_V;                  // staging phase granularity (items)
  static constexpr int SLOTS = TILE + RADIX * 2 * (AMAX - 1) + AMAX;   // every run may be padded front and back
  static constexpr uint32_t OFF_CNT  = 0;                              // u32 [NW][256] running offsets / bases
  static constexpr uint32_t OFF_MISC = OFF_CNT + NW * RADIX * 4;       // u32 [16]
  static constexpr uint32_t OFF_KEYS = OFF_MISC + 64;                  // U [SLOTS] (16-byte aligned)
  static constexpr uint32_t OFF_VALS = OFF_KEYS + ((SLOTS * uint32_t(sizeof(U)) + 15) / 16) * 16;
  static constexpr size_t BYTES      = size_t(OFF_VALS) + size_t(SLOTS) * VBYTES;
};

__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit()
{
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_all()
{
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Run d of this tile -> global memory: scalar head up to the first 16-byte boundary of the destination, one bulk copy
// for the aligned middle, scalar tail.  `src` is the shared-window address of the run's first item; by construction
// (src mod 16) == (dst mod 16).
template <class T>
__device__ __forceinline__ bool store_run(T* dst, uint32_t src, uint32_t n)
{
  constexpr uint32_t A = sizeof(T) >= 16 ? 1 : 16 / sizeof(T);
  uint32_t head        = 0;
  if (A > 1)
  {
    head = (A - (uint32_t(reinterpret_cast<size_t>(dst) / sizeof(T)) & (A - 1))) & (A - 1);
    head = min(head, n);
#pragma unroll
    for (uint32_t j = 0; j + 1 < A; ++j)
    {
      if (j < head)
      {
        dst[j] = lds_t<T>(src + j * uint32_t(sizeof(T)));
      }
    }
  }
  const uint32_t mid = (n - head) & ~(A - 1);
  if (mid > 0)
  {
    bulk_store(dst + head, src + head * uint32_t(sizeof(T)), mid * uint32_t(sizeof(T)));
  }
  if (A > 1)
  {
    const uint32_t done = head + mid;
#pragma unroll
    for (uint32_t j = 0; j + 1 < A; ++j)
    {
      if (done + j < n)
      {
        dst[done + j] = lds_t<T>(src + (done + j) * uint32_t(sizeof(T)));
      }
    }
  }
  return mid > 0;
}

template <class U, int VBYTES, int NT, int IPT, int LBW, bool FLOATK, bool FULL>
__device__ __forceinline__ void onesweep_tma_tile(
  const PassArgs& a, const uint32_t sbase, const uint32_t tile, const uint32_t tile_base, const uint32_t valid)
{
  using L = TmaSmem<U, VBYTES, NT, IPT>;
  using V = typename value_of<VBYTES>::type;
  constexpr int NW   = L::NW;
  constexpr int AMAX = L::AMAX;

  const uint32_t tid   = threadIdx.x;
  const uint32_t lane  = tid & 31;
  const uint32_t warp  = tid >> 5;
  const int shift      = a.shift;
  const uint32_t dmask = a.mask;
  const uint32_t ones  = a.all_ones;
  const U neg_zero     = U(a.xf.neg_zero);
  const U pos_zero     = U(a.xf.pos_zero);
  const uint32_t s_cnt  = sbase + L::OFF_CNT;
  const uint32_t s_misc = sbase + L::OFF_MISC;
  const uint32_t s_keys = sbase + L::OFF_KEYS;
  const uint32_t s_vals = sbase + L::OFF_VALS;
  const uint32_t s_mine = s_cnt + warp * (RADIX * 4);

  // ---- load keys, warp-striped
  U key[IPT];
  const uint32_t chunk = warp * 32 * IPT + lane;
  {
    const U* kin = static_cast<const U*>(a.keys_in) + tile_base + chunk;
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      key[i] = (FULL || chunk + i * 32 < valid) ? kin[i * 32] : U(0);
    }
    if (a.first_pass)
    {
      const XformT<U> xf(a.xf);
#pragma unroll
      for (int i = 0; i < IPT; ++i)
      {
        key[i] = twiddle_in(key[i], xf);
      }
    }
    if (!FULL)
    {
#pragma unroll
      for (int i = 0; i < IPT; ++i)
      {
        if (chunk + i * 32 >= valid)
        {
          key[i] = U(~U(0)); // padding ranks last: max digit, last in tile order
        }
      }
    }
  }

  // ---- rank inside the warp: rank2 holds (offset of the key among the warp's keys of its digit) + 1
  uint32_t rank2[(IPT + 1) / 2];
  const uint32_t lt_mask = lanemask_lt();
  const uint32_t gt_mask = lanemask_gt();
#pragma unroll
  for (int i = 0; i < IPT; ++i)
  {
    const uint32_t d = pass_digit<FLOATK>(key[i], shift, dmask, neg_zero, pos_zero);
    uint32_t b, c; // peers == b & c
    match_digit_ballot_fma(d, ones, b, c);
    const uint32_t before = __popc(b & c & lt_mask);
    const uint32_t ctr    = s_mine + d * 4;
    const uint32_t next   = lds32(ctr) + before + 1;
    if ((b & c & gt_mask) == 0) // highest peer lane: its position + 1 is the new running count
    {
      sts32(ctr, next);
    }
    __syncwarp(); // the next row's loads of this counter come after the leader's store (memory model, racecheck)
    put16(rank2, i, next);
  }
  __syncthreads();

  // ---- digit threads: tile totals, publish, look-back, staging layout
  uint32_t run_len = 0, start = 0;
  unsigned long long gbase = 0;
  if (tid < RADIX)
  {
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w)
    {
      total += lds32(s_cnt + (w * RADIX + tid) * 4);
    }
    uint32_t* lb_word = a.lookback + size_t(tile) * RADIX + tid;
    st_relaxed_u32(lb_word, (tile == 0 ? LB_INCLUSIVE : LB_PARTIAL) | total);

    // decoupled look-back, LBW predecessors per round trip
    uint32_t prefix = 0;
    if (tile > 0)
    {
      int t                = int(tile) - 1;
      const uint32_t* base = a.lookback + tid;
      while (true)
      {
        uint32_t w[LBW];
#pragma unroll
        for (int j = 0; j < LBW; ++j)
        {
          w[j] = (t - j >= 0) ? ld_relaxed_u32(base + size_t(t - j) * RADIX) : LB_INCLUSIVE;
        }
        int state = 0, used = 0; // 0 consuming, 1 hit a word that is not ready, 2 reached an inclusive prefix
#pragma unroll
        for (int j = 0; j < LBW; ++j)
        {
          if (state == 0)
          {
            if ((w[j] & LB_FLAG_MASK) == 0)
            {
              state = 1;
            }
            else
            {
              prefix += w[j] & LB_VALUE_MASK;
              ++used;
              if (w[j] & LB_INCLUSIVE)
              {
                state = 2;
              }
            }
          }
        }
        if (state == 2)
        {
          break;
        }
        if (used == 0)
        {
          __nanosleep(100); // the predecessor is still ranking: do not burn issue slots polling
        }
        t -= used;
      }
      st_relaxed_u32(lb_word, LB_INCLUSIVE | (prefix + total));
    }
    gbase = a.bins[tid] + prefix;
    if (a.bins_next != nullptr && tile_base + valid == a.num_items)
    {
      const uint32_t pad = (!FULL && tid == dmask) ? uint32_t(L::TILE) - valid : 0u;
      a.bins_next[tid]   = gbase + total - pad;
    }
    run_len = total - ((!FULL && tid == dmask) ? uint32_t(L::TILE) - valid : 0u);

    // staging layout: run d gets a slot of whole 16-byte groups and starts inside it at the phase of its destination
    const uint32_t phase = uint32_t(gbase) & uint32_t(AMAX - 1);
    const uint32_t slot  = total == 0 ? 0u : ((total + phase + uint32_t(AMAX - 1)) & ~uint32_t(AMAX - 1));
    uint32_t incl        = slot;
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
      sts32(s_misc + warp * 4, incl);
    }
    start = incl - slot + phase;
  }
  __syncthreads();
  if (tid < RADIX)
  {
#pragma unroll
    for (int w = 0; w < RADIX / 32; ++w)
    {
      const uint32_t ws = lds32(s_misc + w * 4);
      start += (uint32_t(w) < warp) ? ws : 0u;
    }
    uint32_t run = start - 1; // ranks are stored + 1
#pragma unroll
    for (int w = 0; w < NW; ++w)
    {
      const uint32_t addr = s_cnt + (w * RADIX + tid) * 4;
      const uint32_t c    = lds32(addr);
      sts32(addr, run);
      run += c;
    }
  }
  __syncthreads();

  // ---- stage keys in digit order (already in output format on the last pass)
  auto stage_keys = [&](auto last_tag) {
    constexpr bool LAST = decltype(last_tag)::value;
    const XformT<U> xf(a.xf);
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      const uint32_t d = pass_digit<FLOATK>(key[i], shift, dmask, neg_zero, pos_zero);
      const uint32_t r = get16(rank2, i) + lds32(s_mine + d * 4);
      if (VBYTES > 0)
      {
        update16(rank2, i, r);
      }
      sts_t<U>(s_keys + r * uint32_t(sizeof(U)), LAST ? twiddle_out(key[i], xf) : key[i]);
    }
  };
  if (a.last_pass)
  {
    stage_keys(std::true_type{});
  }
  else
  {
    stage_keys(std::false_type{});
  }

  // values are requested now; they arrive while the key runs are handed to the copy engine
  V val[VBYTES > 0 ? IPT : 1];
  if (VBYTES > 0)
  {
    const V* vin = static_cast<const V*>(a.vals_in) + tile_base + chunk;
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      if (FULL || chunk + i * 32 < valid)
      {
        val[i] = vin[i * 32];
      }
    }
  }
  fence_async_smem();
  __syncthreads();

  // ---- one thread per digit hands its run to the TMA engine
  bool issued = false;
  if (tid < RADIX && run_len > 0)
  {
    issued = store_run<U>(static_cast<U*>(a.keys_out) + gbase, s_keys + start * uint32_t(sizeof(U)), run_len);
  }

  if (VBYTES > 0)
  {
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      if (FULL || chunk + i * 32 < valid)
      {
        sts_t<V>(s_vals + get16(rank2, i) * uint32_t(sizeof(V)), val[i]);
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid < RADIX && run_len > 0)
    {
      issued |= store_run<V>(static_cast<V*>(a.vals_out) + gbase, s_vals + start * uint32_t(sizeof(V)), run_len);
    }
  }
  if (issued)
  {
    bulk_commit();
  }

  // the next launch's look-back words are zeroed by this one
  if (tid < RADIX && a.lookback_next != nullptr)
  {
    for (uint32_t t = tile; t < a.lookback_next_tiles; t += gridDim.x)
    {
      a.lookback_next[size_t(t) * RADIX + tid] = 0;
    }
  }
  if (issued)
  {
    bulk_wait_read_all(); // shared memory must outlive the copies that read it
  }
}

template <class U, int VBYTES, int NT, int IPT, int MINB, int LBW, bool FLOATK>
__global__ void __launch_bounds__(NT, MINB) onesweep_tma_kernel(const PassArgs a)
{
  using L = TmaSmem<U, VBYTES, NT, IPT>;
  constexpr int TILE = L::TILE;
  static_assert(NT >= RADIX && NT % 32 == 0, "one thread per digit is required");
  static_assert(L::SLOTS < 65536, "staged positions are kept in 16 bits");

  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = uint32_t(__cvta_generic_to_shared(smem));
  const uint32_t tid   = threadIdx.x;

  // ---- dynamic tile id: a tile only starts after all its predecessors started (look-back cannot deadlock)
  if (tid == 0)
  {
    sts32(sbase + L::OFF_MISC + 32, atomicAdd(a.tile_counter, 1u));
  }
  {
    const uint32_t row = sbase + L::OFF_CNT + (tid >> 5) * (RADIX * 4) + (tid & 31) * 4;
#pragma unroll
    for (int j = 0; j < RADIX / 32; ++j)
    {
      sts32(row + j * 128, 0);
    }
  }
  __syncthreads();
  const uint32_t tile      = lds32(sbase + L::OFF_MISC + 32);
  const uint32_t tile_base = tile * uint32_t(TILE);
  const uint32_t valid     = min(uint32_t(TILE), a.num_items - tile_base);
  if (valid == uint32_t(TILE))
  {
    onesweep_tma_tile<U, VBYTES, NT, IPT, LBW, FLOATK, true>(a, sbase, tile, tile_base, valid);
  }
  else
  {
    onesweep_tma_tile<U, VBYTES, NT, IPT, LBW, FLOATK, false>(a, sbase, tile, tile_base, valid);
  }
}

} // namespace b200rs
