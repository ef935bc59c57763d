// One stable 8-bit digit pass ("onesweep") for sm_100a.
//
// Replaces DeviceRadixSortOnesweepKernel / AgentRadixSortOnesweep
// (/root/reference/cub/cub/device/dispatch/kernels/kernel_radix_sort.cuh:498-558,
//  cub/cub/agent/agent_radix_sort_onesweep.cuh:152-739) and the ranking it calls
// (cub/cub/block/block_radix_rank.cuh:913-1213).  Same contract seen from outside: keys_out/vals_out
// receive the items of keys_in/vals_in stably partitioned by the digit (key >> shift) & mask, starting at
// the per-digit global offsets in `bins`.  The inside is a different design, driven by the ncu finding that
// on B200 this kernel is ISSUE-SLOT and shared-memory-wavefront bound, not HBM bound (profiles/r1_*):
// everything is about instructions and shared-memory transactions per key.
//
//   * no "early counts" histogram pre-pass over the tile: the warp-private running offsets that ranking
//     maintains ARE the per-warp digit histograms once the last item is ranked (saves one shared atomic
//     per key);
//   * match-by-ballot: the eight digit-bit predicates come from two R2P (four bits each), then per bit one VOTE and
//     one predicated NOT, and the eight complemented ballots are ANDed three at a time (3 LOP3); MATCH.ANY is kept
//     as a template switch for measurement only -- it runs at ~1 per 40 cycles per SM on B200 and loses by 1.6x;
//   * the highest peer lane is the leader, so ONE POPC per key gives both the lane's rank among its peers
//     and (for the leader) the group size;
//   * shared memory is addressed with explicit 32-bit shared-window addresses (one LEA per access);
//   * per-digit output offsets in shared memory are 32-bit element offsets (one LDS.32 + IADD3 + IMAD.WIDE +
//     STG per key); a BIG variant with 64-bit offsets serves arrays of 2^32 items and more;
//   * float -0.0 handling is compiled in only for floating-point keys;
//   * the chained scan status words of the NEXT launch are zeroed by this one (no memset between passes);
//   * keys stay bit-ordered in HBM between passes (transform fused into first load / last store).
//
// Stable order inside a tile: warp w owns the contiguous chunk [w*32*IPT, (w+1)*32*IPT); its item i of
// lane l is element i*32+l of the chunk (each load instruction covers one contiguous 32-key run, fully
// coalesced).  Ranks follow (warp, item, lane) == input position order.
#pragma once

#include <cstddef>

#include "common.cuh"

namespace b200rs
{

enum RankAlgo
{
  RANK_MATCH  = 0,
  RANK_BALLOT = 1
};

// Optimisation switches of the kernel (template parameter OPT), kept switchable so each one can be measured alone.
enum OnesweepOpt
{
  OPT_FMA_NOT   = 1, // per-lane ballot complement as a predicated IMAD (FMA pipe) instead of LOP3 (ALU pipe)
  OPT_LB_WINDOW = 2, // look-back polls 8 predecessors per round trip
  OPT_CTR16     = 4, // 16-bit warp counters: half the words per bank, fewer shared-memory bank conflicts
  OPT_BUCKET    = 8, // the digit is the key's destination bucket against PassArgs::splitters (multi-GPU partition pass)
  // (bits 16 and 32 were a hand-pipelined staging loop and an early look-back request: measured 3 % slower / neutral
  //  on B200, removed -- DESIGN.md section 6)
  // single-digit short circuits (reference: agent_radix_sort_onesweep.cuh:386-467)
  OPT_SHORT_WARP  = 64,  // a warp whose keys share one digit skips ranking
  OPT_SHORT_TILE  = 128, // a full tile whose keys share one digit skips staging and is copied straight to its run
  OPT_SHORT       = 64 + 128,
  // 4-byte keys, 16-bit counters: the digit is extracted already scaled and merged with the table base (one rotate +
  // one LOP3 give the shared-memory address), and the scatter folds the staged position into a per-thread pointer
  OPT_FOLD        = 256,
  OPT_FOLD_PTR    = 512, // with OPT_FOLD: scatter through a per-thread pointer + biased offsets instead of offset + position
  OPT_DH_FMA      = 1024 // with OPT_FOLD: the high nibble of the digit is brought down by a multiply-high (FMA pipe)
                         // instead of a shift (ALU pipe, the binding pipe of the rank loop)
};

template <class U, int VBYTES, int NT, int IPT, int OPT = 0>
struct OnesweepSmem
{
  static constexpr int NW          = NT / 32;
  static constexpr int TILE        = NT * IPT;
  static constexpr int ITEM_BYTES  = int(sizeof(U)) > VBYTES ? int(sizeof(U)) : VBYTES;
  static constexpr int CTR_BYTES   = (OPT & OPT_CTR16) ? 2 : 4;
  static constexpr uint32_t OFF_WARP = 0;                             // u32/u16 [NW][256] running offsets
  static constexpr uint32_t OFF_GOFF = OFF_WARP + NW * RADIX * CTR_BYTES; // u64 [256] per-digit output offsets
  static constexpr uint32_t OFF_END  = OFF_GOFF + RADIX * 8;          // u32 [256] end of each digit's staged run
                                                                      // (bucket mode only, else empty)
  static constexpr uint32_t OFF_PEER = OFF_END + ((OPT & 8) ? RADIX * 4 : 0); // PeerTable copy (bucket mode only)
  static constexpr uint32_t OFF_SPLIT = OFF_PEER + ((OPT & 8) ? uint32_t((sizeof(PeerTable) + 15) / 16 * 16) : 0u); // u64 [16] splitters + u32 count (bucket mode only)
  static constexpr uint32_t OFF_MISC = OFF_SPLIT + ((OPT & 8) ? 16u * 8u + 16u : 0u); // u32 [16]
  static constexpr uint32_t OFF_DATA = OFF_MISC + 64;                 // staged tile (16-byte aligned)
  static constexpr size_t BYTES      = size_t(OFF_DATA) + size_t(TILE) * ITEM_BYTES;
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

// lanes of the warp whose 8-bit digit equals this lane's: per bit {predicate, ballot, flip for lanes whose bit is
// clear, and}.  Bits above the pass's digit width are zero in every lane and cost nothing in correctness.
__device__ __forceinline__ void match_digit_ballot(uint32_t d, uint32_t& b, uint32_t& c)
{
  // Returned as two words whose AND is the peer mask: the caller folds that AND into its own three-input LOP3s.
  // The digit bits are tested four at a time so that ptxas packs each group into ONE R2P (register bits ->
  // predicates) and never needs more than four predicates at once; the complemented ballots are combined with
  // three-input ANDs (3 LOP3 instead of 7).
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
    : "r"(d));
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

// The digit of one key for this launch: bits [shift, shift + 8) of the bit-ordered key.  (Bucket mode -- the key's
// destination bucket against the splitters -- is evaluated once per tile, splitter-major, by bucket_ids_of_tile.)
template <bool FLOATK, bool BUCKET, class U>
__device__ __forceinline__ uint32_t tile_digit(const PassArgs& a, U key, int shift, uint32_t mask, U neg_zero, U pos_zero)
{
  static_assert(!BUCKET, "bucket ids come from bucket_ids_of_tile");
  return pass_digit<FLOATK>(key, shift, mask, neg_zero, pos_zero);
}

// Bucket mode: id = 2 * #{splitters below the key} + [key equals a splitter], for every key of the thread, packed four
// ids per register.  Splitter-major: each splitter is read ONCE per thread (a broadcast shared-memory load from the
// table the kernel prologue filled from the kernel arguments or from a device-side PartitionPlan), then compared with
// all IPT keys.  Monotone in the key, so the all-ones padding key gets the largest id.
template <bool FLOATK, int IPT, class U>
__device__ __forceinline__ void
bucket_ids_of_tile(uint32_t s_split, const U (&key)[IPT], U neg_zero, U pos_zero, uint32_t (&bid)[(IPT + 3) / 4])
{
#pragma unroll
  for (int q = 0; q < (IPT + 3) / 4; ++q)
  {
    bid[q] = 0;
  }
  uint32_t ns;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ns) : "r"(s_split + 128u));
#pragma unroll 1
  for (uint32_t j = 0; j < ns; ++j)
  {
    unsigned long long s64;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(s64) : "r"(s_split + j * 8u));
    const U s = U(s64);
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      U k = key[i];
      if (FLOATK)
      {
        k = k == neg_zero ? pos_zero : k;
      }
      bid[i / 4] += ((k > s ? 1u : 0u) + (k >= s ? 1u : 0u)) << (8 * (i & 3));
    }
  }
}

// Bucket mode with remote destinations: byte address of the item with partitioned index idx and bucket d.
// s_peer = shared-window address of the CTA's PeerTable copy.
__device__ __forceinline__ unsigned long long peer_address(
  uint32_t s_peer, uint32_t bucket_table_off, uint32_t rank_table_off, uint32_t d, uint32_t idx, uint32_t item_bytes)
{
  unsigned long long p = lds64(s_peer + bucket_table_off + d * 8);
  if (p == 0) // a segment boundary falls inside this bucket (keys tied with a splitter): find the rank by index
  {
    const uint32_t nd = lds32(s_peer);
    uint32_t r        = 0;
    for (uint32_t j = 0; j + 1 < nd; ++j)
    {
      r += idx >= lds32(s_peer + uint32_t(offsetof(PeerTable, seg_end)) + j * 4) ? 1u : 0u;
    }
    p = lds64(s_peer + rank_table_off + r * 8);
  }
  return p + (unsigned long long) idx * item_bytes;
}

// Bucket mode with remote destinations, the scatter: run by run (one run = the tile's items of one bucket, contiguous
// at the destination), every warp store covering ONE 128-byte-aligned line of the destination.  Measured on B200
// (tools/ubench/peer_store.cu): SM stores into a peer GPU reach 700 GB/s when each warp's 128 bytes are line-aligned,
// 410-420 GB/s when they straddle two lines -- which is what a position-order scatter does, runs starting anywhere.
// s_end[b] = end of bucket b's staged run, s_goff[b] + staged position = partitioned index.
template <class T, int NT, class F>
__device__ __forceinline__ void scatter_runs_to_peers(
  uint32_t s_end, uint32_t s_goff, uint32_t s_peer, uint32_t bucket_table_off, uint32_t rank_table_off, uint32_t s_data,
  uint32_t num_buckets, uint32_t valid, F on_store)
{
  constexpr uint32_t NW   = NT / 32;
  constexpr uint32_t LINE = 128 / sizeof(T) < 32 ? 32 : 128 / sizeof(T); // items per warp store (>= one line)
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t start = 0;
  for (uint32_t b = 0; b < num_buckets; ++b)
  {
    const uint32_t end  = lds32(s_end + b * 4);
    const uint32_t stop = end < valid ? end : valid;
    if (stop > start)
    {
      const uint32_t goff          = lds32(s_goff + b * 4);
      const unsigned long long ptr = lds64(s_peer + bucket_table_off + b * 8);
      if (ptr != 0)
      {
        // byte address of the run's first item; m = items between the previous line boundary and it
        const unsigned long long first = ptr + (unsigned long long) (goff + start) * sizeof(T);
        const uint32_t m               = uint32_t(first & 127u) / uint32_t(sizeof(T));
        const uint32_t count           = stop - start;
        const uint32_t lines           = (count + m + LINE - 1) / LINE;
        for (uint32_t l = warp; l < lines; l += NW)
        {
          const uint32_t v = l * LINE + lane;
          if (v >= m && v - m < count)
          {
            const T x = lds_t<T>(s_data + (start + v - m) * uint32_t(sizeof(T)));
            *reinterpret_cast<T*>(first + (unsigned long long) (v - m) * sizeof(T)) = on_store(x);
          }
        }
      }
      else // a segment boundary falls inside this bucket (keys tied with a splitter): destination per item
      {
        for (uint32_t p = start + threadIdx.x; p < stop; p += NT)
        {
          const T x = lds_t<T>(s_data + p * uint32_t(sizeof(T)));
          *reinterpret_cast<T*>(peer_address(s_peer, bucket_table_off, rank_table_off, b, goff + p, uint32_t(sizeof(T)))) =
            on_store(x);
        }
      }
    }
    start = end;
  }
}

// warp counters: 32-bit, or 16-bit (OPT_CTR16: two digits per word, so a warp-wide access touches at most four
// distinct words per bank instead of eight)
template <bool C16>
__device__ __forceinline__ uint32_t ctr_ld(uint32_t addr)
{
  uint32_t v;
  if (C16)
  {
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  }
  else
  {
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  }
  return v;
}
template <bool C16>
__device__ __forceinline__ void ctr_st(uint32_t addr, uint32_t v)
{
  if (C16)
  {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
  }
  else
  {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
  }
}

// match-by-ballot with the per-lane complement on the FMA pipe: `ones` is 0xffffffff passed at run time so that ptxas
// keeps x * ones + ones (== ~x) as an IMAD instead of folding it back into a LOP3.
__device__ __forceinline__ void match_digit_ballot_fma(uint32_t d, uint32_t ones, uint32_t& b, uint32_t& c)
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
    "@!p0 mad.lo.u32 v0, v0, %3, %3;\n"
    "@!p1 mad.lo.u32 v1, v1, %3, %3;\n"
    "@!p2 mad.lo.u32 v2, v2, %3, %3;\n"
    "@!p3 mad.lo.u32 v3, v3, %3, %3;\n"
    "and.b32 t, dh, 1; setp.ne.u32 p0, t, 0;\n"
    "and.b32 t, dh, 2; setp.ne.u32 p1, t, 0;\n"
    "and.b32 t, dh, 4; setp.ne.u32 p2, t, 0;\n"
    "and.b32 t, dh, 8; setp.ne.u32 p3, t, 0;\n"
    "vote.sync.ballot.b32 v4, p0, 0xffffffff;\n"
    "vote.sync.ballot.b32 v5, p1, 0xffffffff;\n"
    "vote.sync.ballot.b32 v6, p2, 0xffffffff;\n"
    "vote.sync.ballot.b32 v7, p3, 0xffffffff;\n"
    "@!p0 mad.lo.u32 v4, v4, %3, %3;\n"
    "@!p1 mad.lo.u32 v5, v5, %3, %3;\n"
    "@!p2 mad.lo.u32 v6, v6, %3, %3;\n"
    "@!p3 mad.lo.u32 v7, v7, %3, %3;\n"
    "lop3.b32 t, v0, v1, v2, 0x80;\n"
    "lop3.b32 %0, v3, v4, v5, 0x80;\n"
    "lop3.b32 %1, v6, v7, t, 0x80;\n"
    "}\n"
    : "=r"(b), "=r"(c)
    : "r"(d), "r"(ones));
}

// OPT_FOLD: shared-memory address of a 4-byte key's table entry.  rot = (shift - log2(entry bytes)) & 31,
// scaled_mask = digit mask << log2(entry bytes), base aligned to 256 * entry bytes.
template <bool FLOATK>
__device__ __forceinline__ uint32_t
digit_entry(uint32_t key, uint32_t rot, uint32_t scaled_mask, uint32_t base, uint32_t neg_zero, uint32_t pos_zero)
{
  if (FLOATK)
  {
    key = key == neg_zero ? pos_zero : key;
  }
  return (__funnelshift_r(key, key, rot) & scaled_mask) | base;
}

// match-by-ballot on the digit held in bits [1, 9) of a 16-bit-counter address (OPT_FOLD)
template <bool DH_FMA = false>
__device__ __forceinline__ void match_entry_ballot_fma(uint32_t e, uint32_t ones, uint32_t& b, uint32_t& c)
{
  uint32_t dh;
  if (DH_FMA)
  {
    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(dh) : "r"(e), "r"(ones & 0x08000000u)); // e >> 5 on the FMA pipe
  }
  else
  {
    dh = e >> 5;
  }
  asm volatile(
    "{\n"
    ".reg .pred p0, p1, p2, p3;\n"
    ".reg .b32 v0, v1, v2, v3, v4, v5, v6, v7, t;\n"
    "and.b32 t, %2, 2; setp.ne.u32 p0, t, 0;\n"
    "and.b32 t, %2, 4; setp.ne.u32 p1, t, 0;\n"
    "and.b32 t, %2, 8; setp.ne.u32 p2, t, 0;\n"
    "and.b32 t, %2, 16; setp.ne.u32 p3, t, 0;\n"
    "vote.sync.ballot.b32 v0, p0, 0xffffffff;\n"
    "vote.sync.ballot.b32 v1, p1, 0xffffffff;\n"
    "vote.sync.ballot.b32 v2, p2, 0xffffffff;\n"
    "vote.sync.ballot.b32 v3, p3, 0xffffffff;\n"
    "@!p0 mad.lo.u32 v0, v0, %3, %3;\n"
    "@!p1 mad.lo.u32 v1, v1, %3, %3;\n"
    "@!p2 mad.lo.u32 v2, v2, %3, %3;\n"
    "@!p3 mad.lo.u32 v3, v3, %3, %3;\n"
    "and.b32 t, %4, 1; setp.ne.u32 p0, t, 0;\n"
    "and.b32 t, %4, 2; setp.ne.u32 p1, t, 0;\n"
    "and.b32 t, %4, 4; setp.ne.u32 p2, t, 0;\n"
    "and.b32 t, %4, 8; setp.ne.u32 p3, t, 0;\n"
    "vote.sync.ballot.b32 v4, p0, 0xffffffff;\n"
    "vote.sync.ballot.b32 v5, p1, 0xffffffff;\n"
    "vote.sync.ballot.b32 v6, p2, 0xffffffff;\n"
    "vote.sync.ballot.b32 v7, p3, 0xffffffff;\n"
    "@!p0 mad.lo.u32 v4, v4, %3, %3;\n"
    "@!p1 mad.lo.u32 v5, v5, %3, %3;\n"
    "@!p2 mad.lo.u32 v6, v6, %3, %3;\n"
    "@!p3 mad.lo.u32 v7, v7, %3, %3;\n"
    "lop3.b32 t, v0, v1, v2, 0x80;\n"
    "lop3.b32 %0, v3, v4, v5, 0x80;\n"
    "lop3.b32 %1, v6, v7, t, 0x80;\n"
    "}\n"
    : "=r"(b), "=r"(c)
    : "r"(e), "r"(ones), "r"(dh));
}

// Decoupled look-back of one digit over the predecessor tiles (status words `base[t * RADIX]`, t < tile), W words per
// round trip.  Each word carries its own flags, so the loads need no ordering among themselves.
template <int W>
__device__ __forceinline__ void lookback_load(const uint32_t* base, int t, uint32_t (&w)[W])
{
#pragma unroll
  for (int j = 0; j < W; ++j)
  {
    w[j] = (t - j >= 0) ? ld_relaxed_u32(base + size_t(t - j) * RADIX) : LB_INCLUSIVE;
  }
}

// Folds one window into `prefix`; returns true once an inclusive prefix was reached, else moves t to the first word
// that was not published yet.
template <int W>
__device__ __forceinline__ bool lookback_consume(const uint32_t (&w)[W], uint32_t& prefix, int& t)
{
  int state = 0, used = 0; // 0 consuming, 1 met a word that is not published yet, 2 reached an inclusive prefix
#pragma unroll
  for (int j = 0; j < W; ++j)
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
        state = (w[j] & LB_INCLUSIVE) ? 2 : 0;
      }
    }
  }
  t -= used;
  return state == 2;
}

template <int W>
__device__ __forceinline__ uint32_t lookback_prefix(const uint32_t* base, uint32_t tile)
{
  uint32_t prefix = 0;
  int t           = int(tile) - 1;
  while (true)
  {
    uint32_t w[W];
    lookback_load<W>(base, t, w);
    if (lookback_consume<W>(w, prefix, t))
    {
      return prefix;
    }
  }
}

struct NoHook
{
  __device__ __forceinline__ void operator()() const {}
};

// SMEM_IN: the tile's keys were bulk-copied to shared memory at s_in (persistent kernel).  s_stage != 0: staging area
// (the persistent kernel stages in place).  after_rank() runs once every key of the tile is in a register and ranked.
template <class U, int VBYTES, int NT, int IPT, int RANK, int OPT, bool FLOATK, bool BIG, bool FULL, bool SMEM_IN = false,
          class Hook = NoHook>
__device__ __forceinline__ void onesweep_tile(
  const PassArgs& a, const uint32_t sbase, const uint32_t tile, const uint32_t tile_base, const uint32_t valid,
  const uint32_t s_in = 0, const uint32_t s_stage = 0, Hook after_rank = Hook())
{
  using L = OnesweepSmem<U, VBYTES, NT, IPT, OPT>;
  using V = typename value_of<VBYTES>::type;
  constexpr int NW       = L::NW;
  constexpr bool C16     = (OPT & OPT_CTR16) != 0;
  constexpr bool BUCKET  = (OPT & OPT_BUCKET) != 0;
  constexpr uint32_t CB  = L::CTR_BYTES;
  constexpr bool FOLD    = (OPT & OPT_FOLD) != 0 && sizeof(U) == 4 && C16 && !BUCKET && RANK == RANK_BALLOT;
  constexpr bool FOLDP   = FOLD && (OPT & OPT_FOLD_PTR) != 0 && !BIG;
  // OPT_FOLD: the warp counters, ranks and staged positions are kept in BYTES of 4-byte items (no shift when a key is
  // staged; the rank is one IMAD on the FMA pipe).  CSH converts them back to items.
  constexpr bool FOLDB   = FOLD && L::TILE * 4 < 65536;
  constexpr int CSH      = FOLDB ? 2 : 0;
  static_assert(!FOLD || (L::OFF_WARP % 512 == 0 && L::OFF_GOFF % 1024 == 0), "folded table bases must be aligned");

  const uint32_t tid   = threadIdx.x;
  const uint32_t lane  = tid & 31;
  const uint32_t warp  = tid >> 5;
  const int shift      = a.shift;
  const uint32_t dmask = a.mask;
  const U neg_zero     = U(a.xf.neg_zero);
  const U pos_zero     = U(a.xf.pos_zero);
  const uint32_t s_warp = sbase + L::OFF_WARP;
  const uint32_t s_goff = sbase + L::OFF_GOFF;
  const uint32_t s_misc = sbase + L::OFF_MISC;
  const uint32_t s_data = s_stage != 0 ? s_stage : sbase + L::OFF_DATA;
  const uint32_t s_mine = s_warp + warp * (RADIX * CB); // this warp's running offsets
  // OPT_FOLD: rotate amounts and scaled masks of the two tables addressed by digit (16-bit counters, 32-bit offsets)
  const uint32_t rot_ctr  = uint32_t(shift + 31) & 31u;
  const uint32_t rot_goff = uint32_t(shift + 30) & 31u;
  const uint32_t msk_ctr  = dmask << 1;
  const uint32_t msk_goff = dmask << 2;

  if (FOLD && (sbase & 1023u) != 0)
  {
    __trap(); // the folded table addressing relies on the 1024-byte alignment of the dynamic shared memory
  }

  // ---- load keys, warp-striped
  U key[IPT];
  const uint32_t chunk = warp * 32 * IPT + lane;
  {
    const U* kin = static_cast<const U*>(a.keys_in) + tile_base + chunk;
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      if (SMEM_IN)
      {
        key[i] = lds_t<U>(s_in + (chunk + i * 32) * uint32_t(sizeof(U)));
      }
      else
      {
        key[i] = (FULL || chunk + i * 32 < valid) ? kin[i * 32] : U(0);
      }
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

  // ---- rank: warp-private running digit counts; afterwards the warp's row holds its digit histogram.
  // rank2 keeps (position of the key among the warp's keys of its digit) + 1, two 16-bit values per register: the +1 is
  // the value the group leader stores anyway, and it is undone for free by staging at s_data - one item.  Besides
  // halving the registers, the PRMT packing stops ptxas from keeping BOTH addends of every rank alive.
  uint32_t rank2[(IPT + 1) / 2];
  uint32_t bid[BUCKET ? (IPT + 3) / 4 : 1];
  const uint32_t lt_mask = lanemask_lt();
  const uint32_t gt_mask = lanemask_gt();
  // single-digit warp: when all 32 * IPT keys of the warp share one digit their ranks are their positions.  The test
  // costs the normal path one shuffle, one compare and one vote per warp and tile (the full check only runs when the
  // first row already agrees).
  bool warp_single = false;
  if constexpr ((OPT & OPT_SHORT_WARP) != 0 && !BUCKET)
  {
    const uint32_t d0    = tile_digit<FLOATK, BUCKET>(a, key[0], shift, dmask, neg_zero, pos_zero);
    const uint32_t first = __shfl_sync(0xffffffffu, d0, 0);
    if (__all_sync(0xffffffffu, d0 == first))
    {
      uint32_t diff = 0;
#pragma unroll
      for (int i = 1; i < IPT; ++i)
      {
        diff |= tile_digit<FLOATK, BUCKET>(a, key[i], shift, dmask, neg_zero, pos_zero) ^ first;
      }
      warp_single = __all_sync(0xffffffffu, diff == 0);
      if (warp_single)
      {
#pragma unroll
        for (int i = 0; i < IPT; ++i)
        {
          put16(rank2, i, (uint32_t(i) * 32u + lane + 1u) << CSH);
        }
        if (lane == 0)
        {
          ctr_st<C16>(s_mine + first * CB, (32u * IPT) << CSH);
        }
      }
    }
  }
  if constexpr (BUCKET)
  {
    bucket_ids_of_tile<FLOATK, IPT>(sbase + L::OFF_SPLIT, key, neg_zero, pos_zero, bid);
  }
  if (!warp_single)
  {
#pragma unroll
  for (int i = 0; i < IPT; ++i)
  {
    uint32_t d = 0, ctr;
    uint32_t b, c; // peers == b & c
    if (FOLD)
    {
      ctr = digit_entry<FLOATK>(uint32_t(key[i]), rot_ctr, msk_ctr, s_mine, uint32_t(neg_zero), uint32_t(pos_zero));
      match_entry_ballot_fma<(OPT & OPT_DH_FMA) != 0>(ctr, a.all_ones, b, c);
    }
    else
    {
      if constexpr (BUCKET)
      {
        d = (bid[i / 4] >> (8 * (i & 3))) & 0xffu;
      }
      else
      {
        d = tile_digit<FLOATK, BUCKET>(a, key[i], shift, dmask, neg_zero, pos_zero);
      }
      ctr = s_mine + d * CB;
      if (RANK == RANK_MATCH)
      {
        b = c = __match_any_sync(0xffffffffu, d);
      }
      else if (OPT & OPT_FMA_NOT)
      {
        match_digit_ballot_fma(d, a.all_ones, b, c);
      }
      else
      {
        match_digit_ballot(d, b, c);
      }
    }
    // position of the key among the warp's keys of its digit, + 1 (in bytes with OPT_FOLD: peers up to and including
    // this lane, times 4, on top of the running count)
    const uint32_t next = FOLDB ? ctr_ld<C16>(ctr) + 4u * uint32_t(__popc(b & c & ~gt_mask))
                                : ctr_ld<C16>(ctr) + uint32_t(__popc(b & c & lt_mask)) + 1u;
    if ((b & c & gt_mask) == 0) // highest peer lane: its position + 1 is the new running count
    {
      ctr_st<C16>(ctr, next);
    }
    // The next row's loads of this counter must come after the leader's store.  The hardware executes a converged
    // warp's shared-memory instructions in order, but the CUDA memory model does not promise it and compute-sanitizer
    // racecheck reports the pair as a hazard without this; measured cost on C2: none (0.7023 vs 0.7047 ms per pass,
    // profiles/r2u_onesweep_ab.txt).
    __syncwarp();
    put16(rank2, i, next);
  }
  }
  after_rank();
  __syncthreads();

  // ---- per-digit tile totals (one thread per digit), publish, block-wide exclusive scan over digits
  uint32_t total = 0, excl = 0;
  uint32_t* lb_word = a.lookback + size_t(tile) * RADIX + tid;
  if (tid < RADIX)
  {
#pragma unroll
    for (int w = 0; w < NW; ++w)
    {
      total += ctr_ld<C16>(s_warp + (w * RADIX + tid) * CB);
    }
    st_relaxed_u32(lb_word, (tile == 0 ? LB_INCLUSIVE : LB_PARTIAL) | (total >> CSH));
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
    if ((OPT & OPT_SHORT_TILE) && FULL && !BUCKET && total == (uint32_t(L::TILE) << CSH))
    {
      sts32(s_misc + 44, 0x100u | tid); // single-digit tile (the word was cleared before the tile started)
    }
  }
  __syncthreads();
  if (tid < RADIX)
  {
#pragma unroll
    for (int w = 0; w < RADIX / 32; ++w)
    {
      const uint32_t ws = lds32(s_misc + w * 4);
      excl += (uint32_t(w) < warp) ? ws : 0u;
    }
    uint32_t run = excl;
#pragma unroll
    for (int w = 0; w < NW; ++w)
    {
      const uint32_t addr = s_warp + (w * RADIX + tid) * CB;
      const uint32_t c    = ctr_ld<C16>(addr);
      ctr_st<C16>(addr, run);
      run += c;
    }
  }
  __syncthreads();

  // single-digit tile: every key goes to one run in tile order -- no staging, no per-key offset look-up
  const bool short_tile = (OPT & OPT_SHORT_TILE) && FULL && !BUCKET && lds32(s_misc + 44) != 0;

  // ---- stage keys in shared memory in digit order (ranks are + 1: stage one item below s_data)
  if (short_tile)
  {
  }
  else
  {
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      uint32_t centry;
      if (FOLD)
      {
        centry = digit_entry<FLOATK>(uint32_t(key[i]), rot_ctr, msk_ctr, s_mine, uint32_t(neg_zero), uint32_t(pos_zero));
      }
      else
      {
        uint32_t d;
        if constexpr (BUCKET)
        {
          d = (bid[i / 4] >> (8 * (i & 3))) & 0xffu;
        }
        else
        {
          d = tile_digit<FLOATK, BUCKET>(a, key[i], shift, dmask, neg_zero, pos_zero);
        }
        centry = s_mine + d * CB;
      }
      const uint32_t r = get16(rank2, i) + ctr_ld<C16>(centry);
      if (VBYTES > 0)
      {
        update16(rank2, i, r);
      }
      sts_t<U>(s_data - uint32_t(sizeof(U)) + (FOLDB ? r : r * uint32_t(sizeof(U))), key[i]);
    }
  }

  // values are fetched now so their latency hides behind the look-back
  V val[VBYTES > 0 ? IPT : 1];
  if (VBYTES > 0 && !short_tile)
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

  // ---- decoupled look-back over predecessor tiles (one thread per digit)
  if (tid < RADIX)
  {
    uint32_t prefix = 0;
    if (tile > 0)
    {
      if (OPT & OPT_LB_WINDOW)
      {
        prefix = lookback_prefix<8>(a.lookback + tid, tile);
      }
      else
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
      }
      st_relaxed_u32(lb_word, LB_INCLUSIVE | (prefix + (total >> CSH)));
    }
    total >>= CSH; // items from here on
    excl >>= CSH;
    const unsigned long long gbase = a.bins[tid] + prefix;
    // element offset such that out[off + staged position] is the output slot (wraps consistently when negative)
    if (BIG)
    {
      sts64(s_goff + tid * 8, gbase - excl);
    }
    else
    {
      // OPT_FOLD biases the offsets by one tile so that they never wrap below zero (the scatter adds them to a
      // 64-bit pointer); the host selects the 64-bit-offset kernels early enough that the bias cannot overflow
      sts32(s_goff + tid * 4, uint32_t(gbase) - excl + (FOLDP ? uint32_t(L::TILE) : 0u));
    }
    if (BUCKET)
    {
      sts32(sbase + L::OFF_END + tid * 4, excl + total);
    }
    if (a.bins_next != nullptr && tile_base + valid == a.num_items)
    {
      a.bins_next[tid] = gbase + total;
    }
  }
  __syncthreads();

  U* kout = static_cast<U*>(a.keys_out);
  if (short_tile)
  {
    // the run of the tile's one digit starts at goff (its exclusive prefix inside the tile is zero); tile order is kept
    const uint32_t d = lds32(s_misc + 44) & 0xffu;
    const XformT<U> xf(a.xf);
    const unsigned long long g = (BIG ? lds64(s_goff + d * 8)
                                      : (unsigned long long) (lds32(s_goff + d * 4) - (FOLDP ? uint32_t(L::TILE) : 0u)))
                               + chunk;
    U* ko = kout + g;
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      ko[i * 32] = a.last_pass ? twiddle_out(key[i], xf) : key[i];
    }
    if (VBYTES > 0)
    {
      const V* vin = static_cast<const V*>(a.vals_in) + tile_base + chunk;
      V* vo        = static_cast<V*>(a.vals_out) + g;
#pragma unroll
      for (int i0 = 0; i0 < IPT; i0 += 8)
      {
        V v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
        {
          if (i0 + j < IPT)
          {
            v[j] = vin[(i0 + j) * 32];
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
        {
          if (i0 + j < IPT)
          {
            vo[(i0 + j) * 32] = v[j];
          }
        }
      }
    }
    return;
  }

  if constexpr (BUCKET)
  {
    if (a.peer != nullptr)
    {
      // remote destinations: line-aligned run-by-run scatter (keys, then values through the same staging buffer)
      const XformT<U> xf(a.xf);
      const uint32_t nbuckets = 2u * lds32(sbase + L::OFF_SPLIT + 128) + 1u;
      const bool last         = a.last_pass != 0;
      scatter_runs_to_peers<U, NT>(sbase + L::OFF_END, s_goff, sbase + L::OFF_PEER,
                                   uint32_t(offsetof(PeerTable, bucket_dst_keys)),
                                   uint32_t(offsetof(PeerTable, rank_dst_keys)), s_data, nbuckets, valid,
                                   [&](U k) { return last ? twiddle_out(k, xf) : k; });
      if constexpr (VBYTES > 0)
      {
        __syncthreads(); // staged keys are dead
#pragma unroll
        for (int i = 0; i < IPT; ++i)
        {
          if (FULL || chunk + i * 32 < valid)
          {
            sts_t<V>(s_data - uint32_t(sizeof(V)) + (get16(rank2, i) >> CSH) * uint32_t(sizeof(V)), val[i]);
          }
        }
        __syncthreads();
        scatter_runs_to_peers<V, NT>(sbase + L::OFF_END, s_goff, sbase + L::OFF_PEER,
                                     uint32_t(offsetof(PeerTable, bucket_dst_vals)),
                                     uint32_t(offsetof(PeerTable, rank_dst_vals)), s_data, nbuckets, valid,
                                     [](V v) { return v; });
      }
      return;
    }
  }

  // ---- coalesced scatter: consecutive threads write consecutive staged positions
  uint32_t digs[(IPT + 3) / 4];
  auto store_keys = [&](auto last_tag) {
    constexpr bool LAST = decltype(last_tag)::value;
    const XformT<U> xf(a.xf);
    // bucket mode: a thread's staged positions grow with i, so its bucket only moves forward -- follow the staged run
    // ends instead of evaluating the splitters again
    uint32_t cur = 0, cur_end = BUCKET ? lds32(sbase + L::OFF_END) : 0u;
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      const uint32_t pos = i * NT + tid;
      uint32_t d         = 0;
      if (FULL || pos < valid)
      {
        const U k = lds_t<U>(s_data + pos * uint32_t(sizeof(U)));
        if (BUCKET)
        {
          while (pos >= cur_end)
          {
            ++cur;
            cur_end = lds32(sbase + L::OFF_END + cur * 4);
          }
          d = cur;
        }
        else if (FOLD && !BIG)
        {
          // nothing: the offset entry is addressed straight from the key below
        }
        else if constexpr (!BUCKET)
        {
          d = tile_digit<FLOATK, BUCKET>(a, k, shift, dmask, neg_zero, pos_zero);
        }
        const U o = LAST ? twiddle_out(k, xf) : k;
        if (FOLD && !BIG)
        {
          const uint32_t ge =
            digit_entry<FLOATK>(uint32_t(k), rot_goff, msk_goff, s_goff, uint32_t(neg_zero), uint32_t(pos_zero));
          if (VBYTES > 0)
          {
            d = (ge >> 2) & 0xffu;
          }
          if (FOLDP)
          {
            (kout + ptrdiff_t(tid) - ptrdiff_t(L::TILE))[size_t(lds32(ge)) + size_t(i * NT)] = o;
          }
          else
          {
            kout[lds32(ge) + pos] = o;
          }
        }
        else if (BIG)
        {
          kout[lds64(s_goff + d * 8) + pos] = o;
        }
        else if (BUCKET && a.peer != nullptr)
        {
          *reinterpret_cast<U*>(peer_address(sbase + L::OFF_PEER, uint32_t(offsetof(PeerTable, bucket_dst_keys)),
                                             uint32_t(offsetof(PeerTable, rank_dst_keys)), d,
                                             lds32(s_goff + d * 4) + pos, uint32_t(sizeof(U)))) = o;
        }
        else
        {
          kout[lds32(s_goff + d * 4) + pos] = o;
        }
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
  };
  if (a.last_pass)
  {
    store_keys(std::true_type{});
  }
  else
  {
    store_keys(std::false_type{});
  }

  if (VBYTES > 0)
  {
    __syncthreads(); // staged keys are dead
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      if (FULL || chunk + i * 32 < valid)
      {
        sts_t<V>(s_data - uint32_t(sizeof(V)) + (get16(rank2, i) >> CSH) * uint32_t(sizeof(V)), val[i]);
      }
    }
    __syncthreads();
    V* vout = static_cast<V*>(a.vals_out);
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      const uint32_t pos = i * NT + tid;
      if (FULL || pos < valid)
      {
        const uint32_t d = (digs[i / 4] >> (8 * (i & 3))) & 0xffu;
        const V v        = lds_t<V>(s_data + pos * uint32_t(sizeof(V)));
        if (BIG)
        {
          vout[lds64(s_goff + d * 8) + pos] = v;
        }
        else if (BUCKET && a.peer != nullptr)
        {
          *reinterpret_cast<V*>(peer_address(sbase + L::OFF_PEER, uint32_t(offsetof(PeerTable, bucket_dst_vals)),
                                             uint32_t(offsetof(PeerTable, rank_dst_vals)), d,
                                             lds32(s_goff + d * 4) + pos, uint32_t(sizeof(V)))) = v;
        }
        else if (FOLDP)
        {
          (vout + ptrdiff_t(tid) - ptrdiff_t(L::TILE))[size_t(lds32(s_goff + d * 4)) + size_t(i * NT)] = v;
        }
        else
        {
          vout[lds32(s_goff + d * 4) + pos] = v;
        }
      }
    }
  }

}

template <class U, int VBYTES, int NT, int IPT, int RANK, int MINB, int OPT, bool FLOATK, bool BIG>
__global__ void __launch_bounds__(NT, MINB) onesweep_kernel(const PassArgs a)
{
  using L = OnesweepSmem<U, VBYTES, NT, IPT, OPT>;
  constexpr int TILE = L::TILE;
  static_assert(NT >= RADIX && NT % 32 == 0, "one thread per digit is required");
  static_assert(TILE < 65536, "staged positions (+1) are kept in 16 bits");

  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = uint32_t(__cvta_generic_to_shared(smem));
  const uint32_t tid   = threadIdx.x;

  if ((OPT & OPT_BUCKET) && a.plan != nullptr && a.plan->status != 0)
  {
    return; // the device-side plan is flagged inconsistent / over capacity (multi.cu): store nothing anywhere
  }

  // ---- dynamic tile id: a tile only starts after all its predecessors started (look-back cannot deadlock)
  if (tid == 0)
  {
    sts32(sbase + L::OFF_MISC + 32, atomicAdd(a.tile_counter, 1u));
    sts32(sbase + L::OFF_MISC + 44, 0); // single-digit-tile flag
  }
  {
    // zero the warp counters: NW * 256 counters of CTR_BYTES each, as 32-bit words
    constexpr int WORDS = L::NW * RADIX * L::CTR_BYTES / 4;
#pragma unroll
    for (int j = 0; j < WORDS / NT; ++j)
    {
      sts32(sbase + L::OFF_WARP + (j * NT + tid) * 4, 0);
    }
    static_assert(WORDS % NT == 0, "counter words must divide evenly over the threads");
  }
  if (OPT & OPT_BUCKET)
  {
    // splitters: from the kernel arguments, or from a PartitionPlan a kernel earlier in the stream wrote (multi-GPU
    // sort: the host never sees the splitters)
    if (tid < 16)
    {
      const unsigned long long v = a.plan != nullptr ? a.plan->splitters[tid] : (tid < 15 ? a.splitters[tid] : 0ull);
      sts64(sbase + L::OFF_SPLIT + tid * 8, v);
    }
    if (tid == 16)
    {
      sts32(sbase + L::OFF_SPLIT + 128, a.plan != nullptr ? a.plan->num_splitters : uint32_t(a.num_splitters));
    }
  }
  if ((OPT & OPT_BUCKET) && a.peer != nullptr)
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.peer);
    for (uint32_t w = tid; w < uint32_t(sizeof(PeerTable) / 4); w += NT)
    {
      sts32(sbase + L::OFF_PEER + w * 4, src[w]);
    }
  }
  // floating-point keys: the upsweep recorded whether ANY key of the input carries the pattern that has to be ranked as
  // the other zero; almost no input does, and then the pass runs the integer body (no compare + select per key in the
  // rank, staging and scatter loops: 0.84 -> 0.70 ms per pass on 2^28 f32 keys).  One uniform load, issued before the
  // barrier so that its latency overlaps the tile-id atomic.
  bool plain = false;
  if constexpr (FLOATK && (OPT & OPT_BUCKET) == 0)
  {
    if (a.zero_flag != nullptr)
    {
      plain = ld_relaxed_u32(a.zero_flag) == 0;
    }
  }
  __syncthreads();
  const uint32_t tile      = lds32(sbase + L::OFF_MISC + 32);
  const uint32_t tile_base = tile * uint32_t(TILE);
  const uint32_t valid     = min(uint32_t(TILE), a.num_items - tile_base);
  if constexpr (FLOATK && (OPT & OPT_BUCKET) == 0)
  {
    if (plain)
    {
      if (valid == uint32_t(TILE))
      {
        onesweep_tile<U, VBYTES, NT, IPT, RANK, OPT, false, BIG, true>(a, sbase, tile, tile_base, valid);
      }
      else
      {
        onesweep_tile<U, VBYTES, NT, IPT, RANK, OPT, false, BIG, false>(a, sbase, tile, tile_base, valid);
      }
    }
    else if (valid == uint32_t(TILE))
    {
      onesweep_tile<U, VBYTES, NT, IPT, RANK, OPT, true, BIG, true>(a, sbase, tile, tile_base, valid);
    }
    else
    {
      onesweep_tile<U, VBYTES, NT, IPT, RANK, OPT, true, BIG, false>(a, sbase, tile, tile_base, valid);
    }
  }
  else if (valid == uint32_t(TILE))
  {
    onesweep_tile<U, VBYTES, NT, IPT, RANK, OPT, FLOATK, BIG, true>(a, sbase, tile, tile_base, valid);
  }
  else
  {
    onesweep_tile<U, VBYTES, NT, IPT, RANK, OPT, FLOATK, BIG, false>(a, sbase, tile, tile_base, valid);
  }
  // the chained-scan status words of the NEXT launch are zeroed by this one, off the critical path
  if (tid < RADIX && a.lookback_next != nullptr)
  {
    for (uint32_t t = tile; t < a.lookback_next_tiles; t += gridDim.x)
    {
      a.lookback_next[size_t(t) * RADIX + tid] = 0;
    }
  }
}

} // namespace b200rs
