// One stable 8-bit digit pass ("onesweep") for sm_100a -- software-pipelined persistent kernel.
//
// Replaces DeviceRadixSortOnesweepKernel / AgentRadixSortOnesweep
// (/root/reference/cub/cub/device/dispatch/kernels/kernel_radix_sort.cuh:498-558,
//  cub/cub/agent/agent_radix_sort_onesweep.cuh:152-739) and the ranking it calls
// (cub/cub/block/block_radix_rank.cuh:913-1213).  Same contract seen from outside: keys_out/vals_out
// receive the items of keys_in/vals_in stably partitioned by the digit (key >> shift) & mask, starting at
// the per-digit global offsets in `bins`.  The inside is a different design, driven by the ncu finding
// (profiles/) that on B200 this kernel is bound by issue slots, the half-rate integer ALU pipe, shared-memory
// wavefronts and -- a quarter of the time -- by waiting for the decoupled look-back, not by HBM:
//
//   * persistent CTAs (grid = resident CTAs, dynamic tile ids) made of NCW "compute" warps and ONE "helper"
//     warp.  The helper warp owns everything that talks to other CTAs: the decoupled look-back over predecessor
//     tiles (8 digits per lane, 8 independent polls in flight), publishing the inclusive prefix, the per-digit
//     output offsets, zeroing the next launch's status words.  Compute warps never wait on global memory
//     round trips of the chained scan: tile j's offsets are only needed one iteration later;
//   * a two-stage software pipeline per CTA over double-buffered staging memory:
//         rank(j) | digit totals(j), publish | stage(j) -> smem[j&1] | prefetch keys(j+1) | scatter(j-1)
//     so the global loads of tile j+1 fly during the scatter of tile j-1, and the look-back of tile j during
//     scatter(j-1), rank(j+1), totals(j+1), stage(j+1);
//   * producer/consumer hand-over between the compute warps and the helper warp with named barriers
//     (bar.arrive / bar.sync, two ids per direction alternating with the buffer parity);
//   * no "early counts" histogram pre-pass over the tile: the warp-private running offsets that ranking
//     maintains ARE the per-warp digit histograms once the last item is ranked;
//   * match-by-ballot with the per-lane complement done as a predicated IMAD (x*-1 + -1 == ~x) so it runs on the
//     FMA pipe, and three-input LOP3 ANDs: 8 VOTE + 8 IMAD + 5 LOP3 per key instead of 8 VOTE + 16+ LOP3;
//   * the highest peer lane is the leader, so ONE POPC per key gives both the lane's rank among its peers
//     and (for the leader) the group size;
//   * shared memory is addressed with explicit 32-bit shared-window addresses;
//   * float -0.0 handling is compiled in only for floating-point keys; 64-bit output offsets only for arrays of
//     2^32 items and more;
//   * the chained-scan status words of the NEXT launch are zeroed by this one (no memset between passes);
//   * keys stay bit-ordered in HBM between passes (transform fused into first load / last store).
//
// Stable order inside a tile: compute warp w owns the contiguous chunk [w*32*IPT, (w+1)*32*IPT); its item i of
// lane l is element i*32+l of the chunk (each load instruction covers one contiguous 32-key run, fully
// coalesced).  Ranks follow (warp, item, lane) == input position order.
#pragma once

#include "common.cuh"

namespace b200rs
{

template <class U, int VBYTES, int NCW, int IPT, int NBUF = 3, int NHW = 4>
struct OnesweepSmem
{
  static constexpr int NB            = NBUF;                           // staging buffers (tiles in flight per CTA)
  static constexpr int NTC           = NCW * 32;                       // compute threads
  static constexpr int NT            = NTC + NHW * 32;                 // + the helper warps
  static constexpr int DPL           = RADIX / (NHW * 32);             // digits per helper lane
  static constexpr int TILE          = NTC * IPT;
  static constexpr uint32_t OFF_CNT  = 0;                              // u32 [NCW][256] running offsets / bases
  static constexpr uint32_t OFF_TOT  = OFF_CNT + NCW * RADIX * 4;      // u32 [NB][256] tile digit totals
  static constexpr uint32_t OFF_EXCL = OFF_TOT + NBUF * RADIX * 4;     // u32 [NB][256] exclusive digit prefix in tile
  static constexpr uint32_t OFF_GOFF = OFF_EXCL + NBUF * RADIX * 4;    // u64 [NB][256] per-digit output offsets
  static constexpr uint32_t OFF_MISC = OFF_GOFF + NBUF * RADIX * 8;    // u32 [16]: warp sums[8], tile0, next[NB]
  static constexpr uint32_t OFF_KEYS = OFF_MISC + 64;                  // U [NB][TILE] (16-byte aligned)
  static constexpr uint32_t OFF_VALS = OFF_KEYS + NBUF * TILE * uint32_t(sizeof(U)); // V [NB][TILE]
  static constexpr size_t BYTES      = size_t(OFF_VALS) + NBUF * size_t(TILE) * VBYTES;
  static_assert(NBUF >= 2 && NBUF <= 3, "two or three staging buffers");
};

__device__ __forceinline__ uint32_t lanemask_gt()
{
  uint32_t r;
  asm("mov.u32 %0, %%lanemask_gt;" : "=r"(r));
  return r;
}

// ---- shared memory through 32-bit shared-window addresses.  "memory" clobbers keep the compiler from moving
// these across each other; within a converged warp the LSU executes them in program order.
__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v)
{
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long lds64(uint32_t addr)
{
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, unsigned long long v)
{
  asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
template <class T>
__device__ __forceinline__ T lds_t(uint32_t addr)
{
  T v;
  if (sizeof(T) == 1)
  {
    uint32_t t;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t) : "r"(addr) : "memory");
    v = *reinterpret_cast<T*>(&t);
  }
  else if (sizeof(T) == 2)
  {
    uint16_t t;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(t) : "r"(addr) : "memory");
    v = *reinterpret_cast<T*>(&t);
  }
  else if (sizeof(T) == 4)
  {
    uint32_t t = lds32(addr);
    v          = *reinterpret_cast<T*>(&t);
  }
  else if (sizeof(T) == 8)
  {
    unsigned long long t = lds64(addr);
    v                    = *reinterpret_cast<T*>(&t);
  }
  else
  {
    unsigned long long t[2] = {lds64(addr), lds64(addr + 8)};
    v                       = *reinterpret_cast<T*>(t);
  }
  return v;
}
template <class T>
__device__ __forceinline__ void sts_t(uint32_t addr, T v)
{
  if (sizeof(T) == 1)
  {
    uint32_t t = *reinterpret_cast<uint8_t*>(&v);
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(t) : "memory");
  }
  else if (sizeof(T) == 2)
  {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(*reinterpret_cast<uint16_t*>(&v)) : "memory");
  }
  else if (sizeof(T) == 4)
  {
    sts32(addr, *reinterpret_cast<uint32_t*>(&v));
  }
  else if (sizeof(T) == 8)
  {
    sts64(addr, *reinterpret_cast<unsigned long long*>(&v));
  }
  else
  {
    const unsigned long long* t = reinterpret_cast<const unsigned long long*>(&v);
    sts64(addr, t[0]);
    sts64(addr + 8, t[1]);
  }
}

// ---- named barriers (id 0 is __syncthreads)
template <uint32_t ID, uint32_t COUNT>
__device__ __forceinline__ void bar_sync_imm()
{
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}
template <uint32_t ID, uint32_t COUNT>
__device__ __forceinline__ void bar_arrive_imm()
{
  asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}
// barrier id = BASE + buffer index, with all ids as immediates so ptxas reserves only the barriers in use
template <uint32_t BASE, uint32_t COUNT>
__device__ __forceinline__ void bar_sync2(uint32_t idx)
{
  if (idx == 0)
  {
    bar_sync_imm<BASE, COUNT>();
  }
  else if (idx == 1)
  {
    bar_sync_imm<BASE + 1, COUNT>();
  }
  else
  {
    bar_sync_imm<BASE + 2, COUNT>();
  }
}
template <uint32_t BASE, uint32_t COUNT>
__device__ __forceinline__ void bar_arrive2(uint32_t idx)
{
  if (idx == 0)
  {
    bar_arrive_imm<BASE, COUNT>();
  }
  else if (idx == 1)
  {
    bar_arrive_imm<BASE + 1, COUNT>();
  }
  else
  {
    bar_arrive_imm<BASE + 2, COUNT>();
  }
}
constexpr uint32_t BAR_COMPUTE = 1; // compute warps only
constexpr uint32_t BAR_TOTALS  = 2; // +buffer: compute (arrive) -> helper (sync): tile totals published
constexpr uint32_t BAR_GOFF    = 5; // +buffer: helper (arrive) -> compute (sync): output offsets ready

// Lanes of the warp whose 8-bit digit equals this lane's, returned as two words whose AND is the peer mask (the
// caller folds the final AND into its own three-input LOP3s).  Per bit: ballot, complement for lanes whose bit is
// clear -- a predicated multiply-add by `m1` (== 0xffffffff, opaque to the compiler so it stays an IMAD on the FMA
// pipe: x * -1 + -1 == ~x) -- then three-input ANDs.  Bits above the pass's digit width are zero in every lane.
__device__ __forceinline__ void match_digit_ballot(uint32_t d, uint32_t m1, uint32_t& b, uint32_t& c)
{
  asm volatile(
    "{\n"
    ".reg .pred p0, p1, p2, p3;\n"
    ".reg .b32 v0, v1, v2, v3, v4, v5, v6, v7, t, dh;\n"
    "shr.u32 dh, %2, 4;\n"
    "and.b32 t, %2, 1; setp.ne.u32 p0, t, 0;\n"
    "and.b32 t, %2, 2; setp.ne.u32 p1, t, 0;\n"
    "and.b32 t, %2, 4; setp.ne.u32 p2, t, 0;\n"
    "and.b32 t, %2, 8; setp.ne.u32 p3, t, 0;\n"
    "vote.sync.ballot.b32 v0, p0, 0xffffffff;\n"
    "vote.sync.ballot.b32 v1, p1, 0xffffffff;\n"
    "vote.sync.ballot.b32 v2, p2, 0xffffffff;\n"
    "vote.sync.ballot.b32 v3, p3, 0xffffffff;\n"
    "@!p0 not.b32 v0, v0;\n"
    "@!p1 not.b32 v1, v1;\n"
    "@!p2 not.b32 v2, v2;\n"
    "@!p3 not.b32 v3, v3;\n"
    "and.b32 t, dh, 1; setp.ne.u32 p0, t, 0;\n"
    "and.b32 t, dh, 2; setp.ne.u32 p1, t, 0;\n"
    "and.b32 t, dh, 4; setp.ne.u32 p2, t, 0;\n"
    "and.b32 t, dh, 8; setp.ne.u32 p3, t, 0;\n"
    "vote.sync.ballot.b32 v4, p0, 0xffffffff;\n"
    "vote.sync.ballot.b32 v5, p1, 0xffffffff;\n"
    "vote.sync.ballot.b32 v6, p2, 0xffffffff;\n"
    "vote.sync.ballot.b32 v7, p3, 0xffffffff;\n"
    "@!p0 not.b32 v4, v4;\n"
    "@!p1 not.b32 v5, v5;\n"
    "@!p2 not.b32 v6, v6;\n"
    "@!p3 not.b32 v7, v7;\n"
    "lop3.b32 t, v0, v1, v2, 0x80;\n"
    "lop3.b32 %0, v3, v4, v5, 0x80;\n"
    "lop3.b32 %1, v6, v7, t, 0x80;\n"
    "}\n"
    : "=r"(b), "=r"(c)
    : "r"(d), "r"(m1));
}

template <int N>
__device__ __forceinline__ void put16(uint32_t (&pk)[N], int i, uint32_t v)
{
  pk[i / 2] = (i & 1) ? __byte_perm(pk[i / 2], v, 0x5410) : v; // v < 65536; the even half initialises the pair
}
template <int N>
__device__ __forceinline__ void update16(uint32_t (&pk)[N], int i, uint32_t v)
{
  pk[i / 2] = __byte_perm(pk[i / 2], v, (i & 1) ? 0x5410 : 0x3254); // replace one half, keep the other
}
template <int N>
__device__ __forceinline__ uint32_t get16(const uint32_t (&pk)[N], int i)
{
  return (i & 1) ? (pk[i / 2] >> 16) : (pk[i / 2] & 0xffffu);
}

template <bool FLOATK, class U>
__device__ __forceinline__ uint32_t pass_digit(U key, int shift, uint32_t mask, U neg_zero, U pos_zero)
{
  if (FLOATK)
  {
    key = key == neg_zero ? pos_zero : key; // -0.0 ranks as +0.0; the stored bits are untouched
  }
  return uint32_t(key >> shift) & mask;
}

// ------------------------------------------------------------------------------------------------ helper warps
// NHW warps per CTA, each lane owning DPL = 256 / (32 NHW) consecutive digits: for every tile the CTA's compute warps
// have ranked, walk back over the predecessor tiles' status words, publish the inclusive prefix, and leave the
// per-digit output offsets in shared memory.
//
// Windowed decoupled look-back.  At B200 tile rates (one tile every ~20-40 ns chip-wide against an L2 round trip of
// several hundred ns) the nearest tile with an INCLUSIVE word is typically 10-25 tiles back, so walking one predecessor
// per round trip would make the look-back the bottleneck of the CTA.  Each round reads the next LB_WIN predecessor
// rows of the lane's digits at once (one DPL-wide vector load per row, LB_WIN independent loads in flight) and consumes
// them nearest-first, stopping at the first unpublished or INCLUSIVE word.
template <int DPL>
__device__ __forceinline__ void ld_status(const uint32_t* p, uint32_t (&v)[DPL])
{
  if (DPL == 1)
  {
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v[0]) : "l"(p) : "memory");
  }
  else if (DPL == 2)
  {
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[DPL > 1 ? 1 : 0]) : "l"(p) : "memory");
  }
  else
  {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[DPL > 1 ? 1 : 0]), "=r"(v[DPL > 2 ? 2 : 0]), "=r"(v[DPL > 3 ? 3 : 0])
                 : "l"(p)
                 : "memory");
  }
}

template <class L, bool BIG>
__device__ __forceinline__ void onesweep_helper(const PassArgs& a, const uint32_t sbase, uint32_t tile)
{
  constexpr uint32_t NT   = L::NT;
  constexpr uint32_t TILE = L::TILE;
  constexpr int DPL       = L::DPL;
  constexpr int LB_WIN    = DPL >= 4 ? 4 : 8;
  static_assert(DPL == 1 || DPL == 2 || DPL == 4, "1, 2 or 4 digits per helper lane");
  const uint32_t hlane     = threadIdx.x - L::NTC;     // 0 .. NHW*32-1
  const uint32_t d0        = hlane * DPL;              // first digit of this lane
  const uint32_t num_tiles = a.num_tiles;

  uint32_t buf = 0;
  while (true)
  {
    bar_sync2<BAR_TOTALS, NT>(buf); // totals / exclusive prefixes of `tile` are in shared memory, partials published
    const uint32_t next  = lds32(sbase + L::OFF_MISC + (9 + buf) * 4);
    const uint32_t s_tot = sbase + L::OFF_TOT + (buf * RADIX + d0) * 4;

    uint32_t prefix[DPL];
#pragma unroll
    for (int u = 0; u < DPL; ++u)
    {
      prefix[u] = 0;
    }
    if (tile > 0)
    {
      uint32_t k[DPL]; // next row to consume, per digit
#pragma unroll
      for (int u = 0; u < DPL; ++u)
      {
        k[u] = tile - 1;
      }
      uint32_t pending = (1u << DPL) - 1u;
      while (pending != 0)
      {
        // all pending digits of a lane are read from the rows below the FARTHEST-behind digit's next row; digits that
        // are further along simply skip rows they have already consumed
        uint32_t top = 0;
#pragma unroll
        for (int u = 0; u < DPL; ++u)
        {
          top = (pending & (1u << u)) ? max(top, k[u]) : top;
        }
        uint32_t s[LB_WIN][DPL];
#pragma unroll
        for (int w = 0; w < LB_WIN; ++w)
        {
#pragma unroll
          for (int u = 0; u < DPL; ++u)
          {
            s[w][u] = 0;
          }
          if (top >= uint32_t(w))
          {
            ld_status<DPL>(a.lookback + size_t(top - w) * RADIX + d0, s[w]);
          }
        }
#pragma unroll
        for (int u = 0; u < DPL; ++u)
        {
          bool open = (pending & (1u << u)) != 0;
#pragma unroll
          for (int w = 0; w < LB_WIN; ++w)
          {
            const bool mine = open && (top - w) == k[u]; // the row this digit wants next
            // zero flags: that predecessor has only started: stop and poll again
            if (mine && (s[w][u] & LB_FLAG_MASK) != 0)
            {
              prefix[u] += s[w][u] & LB_VALUE_MASK;
              k[u]--;
              if (s[w][u] & LB_INCLUSIVE)
              {
                pending &= ~(1u << u);
                open = false;
                // publish as early as possible: successors stop walking here
                st_relaxed_u32(a.lookback + size_t(tile) * RADIX + d0 + u,
                               LB_INCLUSIVE | (prefix[u] + lds32(s_tot + u * 4)));
              }
            }
            else if (mine)
            {
              open = false;
            }
          }
        }
      }
    }

    const uint32_t tile_base = tile * TILE;
    const bool last_tile     = a.bins_next != nullptr && (a.num_items - tile_base) <= TILE;
#pragma unroll
    for (int u = 0; u < DPL; ++u)
    {
      const uint32_t d               = d0 + u;
      const uint32_t excl            = lds32(sbase + L::OFF_EXCL + (buf * RADIX + d) * 4);
      const unsigned long long gbase = a.bins[d] + prefix[u]; // L1-resident after the CTA's first tile
      // element offset such that out[off + staged position] is the output slot (wraps consistently when negative)
      if (BIG)
      {
        sts64(sbase + L::OFF_GOFF + (buf * RADIX + d) * 8, gbase - excl);
      }
      else
      {
        sts32(sbase + L::OFF_GOFF + (buf * RADIX + d) * 4, uint32_t(gbase) - excl);
      }
      if (last_tile)
      {
        a.bins_next[d] = gbase + lds32(s_tot + u * 4);
      }
    }
    __threadfence_block();
    bar_arrive2<BAR_GOFF, NT>(buf);

    if (a.lookback_next != nullptr)
    {
      for (uint32_t t = tile; t < a.lookback_next_tiles; t += num_tiles)
      {
#pragma unroll
        for (int u = 0; u < DPL; ++u)
        {
          a.lookback_next[size_t(t) * RADIX + d0 + u] = 0;
        }
      }
    }
    tile = next;
    buf  = (buf + 1 == uint32_t(L::NB)) ? 0u : buf + 1;
    if (tile >= num_tiles)
    {
      break;
    }
  }
}

// ------------------------------------------------------------------------------------------------ compute warps
template <class U, int VBYTES, int NCW, int IPT, int NBUF, bool FLOATK, bool BIG>
__device__ __forceinline__ void onesweep_compute(const PassArgs& a, const uint32_t sbase, uint32_t tile)
{
  using L = OnesweepSmem<U, VBYTES, NCW, IPT, NBUF>;
  constexpr uint32_t LAG = NBUF - 1; // a tile is scattered LAG iterations after it was ranked
  using V = typename value_of<VBYTES>::type;
  constexpr uint32_t NTC  = L::NTC;
  constexpr uint32_t NT   = L::NT;
  constexpr uint32_t TILE = L::TILE;
  constexpr bool PAIRS    = VBYTES > 0;

  const uint32_t tid_      = threadIdx.x;
  const uint32_t lane_     = tid_ & 31;
  const uint32_t warp_     = tid_ >> 5;
  const int shift          = a.shift;
  const uint32_t dmask     = a.mask;
  const uint32_t m1        = a.all_ones;
  const U neg_zero         = U(a.xf.neg_zero);
  const U pos_zero         = U(a.xf.pos_zero);
  const uint32_t num_tiles = a.num_tiles;
  const uint32_t s_cnt     = sbase + L::OFF_CNT;
  const uint32_t s_misc    = sbase + L::OFF_MISC;
  const uint32_t s_mine    = s_cnt + warp_ * (RADIX * 4); // this warp's running offsets
  const uint32_t chunk     = warp_ * 32 * IPT + lane_;
  const uint32_t lt_mask   = lanemask_lt();
  const uint32_t gt_mask   = lanemask_gt();

  U key[IPT];
  V val[PAIRS ? IPT : 1];

  auto load_tile = [&](uint32_t t) {
    const uint32_t tile_base = t * TILE;
    const uint32_t valid     = min(TILE, a.num_items - tile_base);
    const U* kin             = static_cast<const U*>(a.keys_in) + tile_base + chunk;
    if (valid == TILE)
    {
#pragma unroll
      for (int i = 0; i < IPT; ++i)
      {
        key[i] = kin[i * 32];
      }
    }
    else
    {
#pragma unroll
      for (int i = 0; i < IPT; ++i)
      {
        key[i] = (chunk + i * 32 < valid) ? kin[i * 32] : U(0);
      }
    }
    if (PAIRS)
    {
      const V* vin = static_cast<const V*>(a.vals_in) + tile_base + chunk;
      if (valid == TILE)
      {
#pragma unroll
        for (int i = 0; i < IPT; ++i)
        {
          val[i] = vin[i * 32];
        }
      }
      else
      {
#pragma unroll
        for (int i = 0; i < IPT; ++i)
        {
          if (chunk + i * 32 < valid)
          {
            val[i] = vin[i * 32];
          }
        }
      }
    }
  };

  // coalesced scatter of a staged tile: consecutive threads write consecutive staged positions
  auto scatter_tile = [&](uint32_t t, uint32_t buf) {
    const uint32_t tile_base = t * TILE;
    const uint32_t valid     = min(TILE, a.num_items - tile_base);
    const uint32_t s_keys    = sbase + L::OFF_KEYS + buf * TILE * uint32_t(sizeof(U));
    const uint32_t s_vals    = sbase + L::OFF_VALS + buf * TILE * uint32_t(VBYTES);
    const uint32_t s_goff    = sbase + L::OFF_GOFF + buf * RADIX * (BIG ? 8 : 4);
    U* kout                  = static_cast<U*>(a.keys_out);
    V* vout                  = static_cast<V*>(a.vals_out);
    const XformT<U> xf(a.xf);
    const bool last = a.last_pass != 0;
    auto body       = [&](auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
      for (int i = 0; i < IPT; ++i)
      {
        const uint32_t pos = i * NTC + tid_;
        if (FULL || pos < valid)
        {
          const U k        = lds_t<U>(s_keys + pos * uint32_t(sizeof(U)));
          const uint32_t d = pass_digit<FLOATK>(k, shift, dmask, neg_zero, pos_zero);
          const U o        = last ? twiddle_out(k, xf) : k;
          if (BIG)
          {
            const unsigned long long off = lds64(s_goff + d * 8) + pos;
            kout[off]                    = o;
            if (PAIRS)
            {
              vout[off] = lds_t<V>(s_vals + pos * uint32_t(VBYTES));
            }
          }
          else
          {
            const uint32_t off = lds32(s_goff + d * 4) + pos;
            kout[off]          = o;
            if (PAIRS)
            {
              vout[off] = lds_t<V>(s_vals + pos * uint32_t(VBYTES));
            }
          }
        }
      }
    };
    if (valid == TILE)
    {
      body(std::true_type{});
    }
    else
    {
      body(std::false_type{});
    }
  };

  load_tile(tile);

  uint32_t buf   = 0; // staging buffer of the tile being ranked
  uint32_t obuf  = 0; // staging buffer of the oldest tile not scattered yet
  uint32_t pend_ = 0; // tiles ranked and staged but not scattered yet (<= LAG)
  uint32_t pt0 = 0, pt1 = 0; // their tile ids, oldest first
  while (true)
  {
    const uint32_t tile_base = tile * TILE;
    const uint32_t valid_    = min(TILE, a.num_items - tile_base);

    // ---- transform on first pass; padding of the ragged last tile ranks last (max digit, last in tile order)
    if (a.first_pass)
    {
      const XformT<U> xf(a.xf);
#pragma unroll
      for (int i = 0; i < IPT; ++i)
      {
        key[i] = twiddle_in(key[i], xf);
      }
    }
    if (valid_ != TILE)
    {
#pragma unroll
      for (int i = 0; i < IPT; ++i)
      {
        if (chunk + i * 32 >= valid_)
        {
          key[i] = U(~U(0));
        }
      }
    }

    // ---- rank: warp-private running digit offsets; afterwards the warp's row holds its digit histogram.
    // Warp-relative ranks are kept two 16-bit values per register.
    uint32_t rank2[(IPT + 1) / 2];
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      const uint32_t d = pass_digit<FLOATK>(key[i], shift, dmask, neg_zero, pos_zero);
      uint32_t b, c;
      match_digit_ballot(d, m1, b, c);
      const uint32_t before = __popc(b & c & lt_mask);
      const uint32_t ctr    = s_mine + d * 4;
      const uint32_t off    = lds32(ctr);
      // every peer stores the same new running offset (benign same-value race): no leader predicate, which would
      // have to stay live while ptxas packs the next item's digit-bit tests into one R2P (P0..P6 all at once)
      sts32(ctr, off + __popc(b & c));
      put16(rank2, i, off + before);
    }
    bar_sync_imm<BAR_COMPUTE, NTC>();

    // Loop-invariant conditions below (tid < 256, lane == 31, full tile, ...) must not be kept in predicate
    // registers across the ranking code: ptxas turns the eight digit-bit tests of every item into one R2P, which
    // needs P0..P6 at once and fails to allocate if anything else is live.  Laundering the inputs through an empty
    // asm makes the compiler recompute those predicates here.
    uint32_t tid = tid_, lane = lane_, warp = warp_, valid = valid_, pend = pend_;
    asm volatile("" : "+r"(tid), "+r"(lane), "+r"(warp), "+r"(valid), "+r"(pend));

    // Stagger: this CTA publishes the totals of its next tile only after the look-back of its previous tile has
    // completed.  Normally that happened long ago (the helper had a whole iteration), so this costs nothing; but if
    // the look-back ever falls behind (all CTAs publishing while nobody has an inclusive prefix yet, e.g. at kernel
    // start), running ahead would leave every helper a walk of ~#CTAs rows per tile, for ever.  Holding the publication
    // back lets the inclusive frontier catch up and keeps the walks a few rows long (the same self-organisation a
    // non-pipelined onesweep gets for free).
    if (pend > 0)
    {
      bar_sync2<BAR_GOFF, NT>(buf == 0 ? uint32_t(NBUF) - 1 : buf - 1);
    }

    // ---- per-digit tile totals (one thread per digit), publish, block-wide exclusive scan over digits
    uint32_t total = 0, excl = 0;
    if (tid < RADIX)
    {
#pragma unroll
      for (int w = 0; w < NCW; ++w)
      {
        total += lds32(s_cnt + (w * RADIX + tid) * 4);
      }
      st_relaxed_u32(a.lookback + size_t(tile) * RADIX + tid, (tile == 0 ? LB_INCLUSIVE : LB_PARTIAL) | total);
      sts32(sbase + L::OFF_TOT + (buf * RADIX + tid) * 4, total);
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
        sts32(s_misc + warp * 4, incl);
      }
      excl = incl - total;
    }
    else if (tid == RADIX)
    {
      // claim the next tile: a tile only starts after all its predecessors started (look-back cannot deadlock)
      sts32(s_misc + (9 + buf) * 4, atomicAdd(a.tile_counter, 1u));
    }
    bar_sync_imm<BAR_COMPUTE, NTC>();
    if (tid < RADIX)
    {
#pragma unroll
      for (int w = 0; w < RADIX / 32; ++w)
      {
        const uint32_t ws = lds32(s_misc + w * 4);
        excl += (uint32_t(w) < warp) ? ws : 0u;
      }
      sts32(sbase + L::OFF_EXCL + (buf * RADIX + tid) * 4, excl);
      uint32_t run = excl;
#pragma unroll
      for (int w = 0; w < NCW; ++w)
      {
        const uint32_t addr = s_cnt + (w * RADIX + tid) * 4;
        const uint32_t c    = lds32(addr);
        sts32(addr, run);
        run += c;
      }
    }
    __threadfence_block();
    bar_arrive2<BAR_TOTALS, NT>(buf); // the helper warp may start the look-back of this tile
    bar_sync_imm<BAR_COMPUTE, NTC>();

    // ---- stage keys (and values) in shared memory in digit order
    {
      const uint32_t s_keys = sbase + L::OFF_KEYS + buf * TILE * uint32_t(sizeof(U));
      const uint32_t s_vals = sbase + L::OFF_VALS + buf * TILE * uint32_t(VBYTES);
#pragma unroll
      for (int i = 0; i < IPT; ++i)
      {
        const uint32_t d = pass_digit<FLOATK>(key[i], shift, dmask, neg_zero, pos_zero);
        const uint32_t r = get16(rank2, i) + lds32(s_mine + d * 4);
        sts_t<U>(s_keys + r * uint32_t(sizeof(U)), key[i]);
        if (PAIRS)
        {
          if (valid == TILE || chunk + i * 32 < valid)
          {
            sts_t<V>(s_vals + r * uint32_t(VBYTES), val[i]);
          }
        }
      }
    }
    // this warp's counter row is dead: zero it for the next tile
    __syncwarp();
#pragma unroll
    for (int j = 0; j < RADIX / 32; ++j)
    {
      sts32(s_mine + (j * 32 + lane) * 4, 0);
    }

    // ---- prefetch the next tile while the previous one is scattered
    const uint32_t next = lds32(s_misc + (9 + buf) * 4);
    if (next < num_tiles)
    {
      load_tile(next);
    }
    if (pend == LAG)
    {
      scatter_tile(pt0, obuf); // its output offsets were awaited before the current tile was published
      obuf = (obuf + 1 == uint32_t(NBUF)) ? 0u : obuf + 1;
      pt0  = pt1;
      pend--;
    }
    if (pend == 0)
    {
      pt0 = tile;
    }
    else
    {
      pt1 = tile;
    }
    pend_ = pend + 1;
    tile  = next;
    buf   = (buf + 1 == uint32_t(NBUF)) ? 0u : buf + 1;
    if (tile >= num_tiles)
    {
      break;
    }
  }
  // drain: only the newest staged tile has not been awaited yet
  while (pend_ > 0)
  {
    if (pend_ == 1)
    {
      bar_sync2<BAR_GOFF, NT>(obuf);
    }
    scatter_tile(pt0, obuf);
    obuf = (obuf + 1 == uint32_t(NBUF)) ? 0u : obuf + 1;
    pt0  = pt1;
    pend_--;
  }
}

template <class U, int VBYTES, int NCW, int IPT, int NBUF, int MINB, bool FLOATK, bool BIG>
__global__ void __launch_bounds__(OnesweepSmem<U, VBYTES, NCW, IPT, NBUF>::NT, MINB) onesweep_kernel(const PassArgs a)
{
  using L = OnesweepSmem<U, VBYTES, NCW, IPT, NBUF>;
  static_assert(L::NTC >= RADIX + 32, "one compute thread per digit plus one to claim tiles");
  static_assert(L::TILE <= 65536, "staged positions are kept in 16 bits");

  extern __shared__ __align__(16) unsigned char smem[];
  const uint32_t sbase = uint32_t(__cvta_generic_to_shared(smem));
  const uint32_t tid   = threadIdx.x;

  // ---- first tile id; zero the running-offset rows
  if (tid == 0)
  {
    sts32(sbase + L::OFF_MISC + 8 * 4, atomicAdd(a.tile_counter, 1u));
  }
  if (tid < L::NTC)
  {
    const uint32_t row = sbase + L::OFF_CNT + (tid >> 5) * (RADIX * 4) + (tid & 31) * 4;
#pragma unroll
    for (int j = 0; j < RADIX / 32; ++j)
    {
      sts32(row + j * 128, 0);
    }
  }
  __syncthreads();
  const uint32_t tile = lds32(sbase + L::OFF_MISC + 8 * 4);
  if (tile >= a.num_tiles)
  {
    return; // more CTAs than tiles
  }
  if (tid >= L::NTC)
  {
    onesweep_helper<L, BIG>(a, sbase, tile);
    return; // exits here so that the compute path below is not inside a divergent region of the kernel
  }
  onesweep_compute<U, VBYTES, NCW, IPT, NBUF, FLOATK, BIG>(a, sbase, tile);
}

} // namespace b200rs
