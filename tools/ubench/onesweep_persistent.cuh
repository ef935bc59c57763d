// EXPERIMENT (measured and rejected, kept buildable for the A/B harness tools/ubench/onesweep_ab.cu; not part of the
// library): persistent CTAs whose NEXT tile is bulk-copied (cp.async.bulk = 1-D TMA, SASS UBLKCP) into shared memory
// while the current tile is processed.  Bit-exact with the classic kernel, but 18 % slower on 2^28 uniform u32 keys
// (0.915 vs 0.776 ms per pass, B200): the pass is bound by shared-memory wavefronts, and routing the keys through
// shared memory adds two wavefronts per 32 keys (the TMA write and the LDS that replaces the LDG).  DESIGN.md section 6.
#pragma once

#include "onesweep.cuh"

namespace b200rs
{

// ---------------------------------------------------------------------------------------------------------------
// Persistent variant: one CTA per resident slot walks over tiles (dynamic tickets, so the look-back still cannot
// deadlock) and the NEXT tile's keys are bulk-copied (cp.async.bulk = 1-D TMA, completion on an mbarrier) into the
// other of two shared-memory buffers while the current tile is ranked, staged and scattered.  The load latency, the
// ticket atomic and the counter reset leave the critical path, and the prefetch costs no registers.  Keys are staged
// in place (every key is in a register before the first barrier of the tile).  Requires 16-byte aligned keys_in;
// the ragged last tile of a portion is loaded the classic way.
// ---------------------------------------------------------------------------------------------------------------
template <class U, int VBYTES, int NT, int IPT, int OPT = 0>
struct PersistSmem
{
  static constexpr int NW          = NT / 32;
  static constexpr int TILE        = NT * IPT;
  static constexpr int ITEM_BYTES  = int(sizeof(U)) > VBYTES ? int(sizeof(U)) : VBYTES;
  static constexpr int CTR_BYTES   = (OPT & OPT_CTR16) ? 2 : 4;
  // the first three offsets coincide with OnesweepSmem so that onesweep_tile addresses them the same way
  static constexpr uint32_t OFF_WARP = 0;
  static constexpr uint32_t OFF_GOFF = OFF_WARP + NW * RADIX * CTR_BYTES;
  static constexpr uint32_t OFF_END  = OFF_GOFF + RADIX * 8;
  static constexpr uint32_t OFF_PEER = OFF_END;
  static constexpr uint32_t OFF_MISC = OFF_PEER;            // u32 [16]: warp sums [8], tickets at +32 / +36
  static constexpr uint32_t OFF_MBAR = OFF_MISC + 64;       // two mbarriers
  static constexpr uint32_t OFF_BUF  = (OFF_MBAR + 16 + 16 + 127) / 128 * 128; // 16 bytes of slack below: ranks are + 1
  static constexpr uint32_t BUF_BYTES = (uint32_t(TILE) * ITEM_BYTES + 127) / 128 * 128;
  static constexpr size_t BYTES      = size_t(OFF_BUF) + 2 * size_t(BUF_BYTES);
};

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "B200RS_WAIT:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra B200RS_DONE;\n"
    "bra B200RS_WAIT;\n"
    "B200RS_DONE:\n"
    "}\n" ::"r"(bar),
    "r"(parity)
    : "memory");
}

template <class U, int VBYTES, int NT, int IPT, int RANK, int MINB, int OPT, bool FLOATK, bool BIG>
__global__ void __launch_bounds__(NT, MINB) onesweep_persistent_kernel(const PassArgs a)
{
  using L = PersistSmem<U, VBYTES, NT, IPT, OPT>;
  constexpr int TILE = L::TILE;
  static_assert(NT >= RADIX && NT % 32 == 0, "one thread per digit is required");
  static_assert(TILE < 65536, "staged positions (+1) are kept in 16 bits");
  static_assert((OPT & OPT_BUCKET) == 0, "bucket mode runs in the classic kernel");
  static_assert((TILE * sizeof(U)) % 16 == 0, "bulk copies move multiples of 16 bytes");

  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase  = uint32_t(__cvta_generic_to_shared(smem));
  const uint32_t tid    = threadIdx.x;
  const uint32_t s_misc = sbase + L::OFF_MISC;
  const uint32_t s_mbar = sbase + L::OFF_MBAR;
  const uint32_t s_buf  = sbase + L::OFF_BUF;
  constexpr uint32_t TILE_BYTES = uint32_t(TILE) * uint32_t(sizeof(U));
  const U* kin = static_cast<const U*>(a.keys_in);

  // a full tile is bulk-copied into buffer b; the ragged last tile is loaded by the tile body itself
  auto request = [&](uint32_t t, uint32_t b) {
    if (t < a.num_tiles && a.num_items - t * uint32_t(TILE) >= uint32_t(TILE))
    {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy reads of the buffer come first
      mbar_expect_tx(s_mbar + b * 8, TILE_BYTES);
      bulk_load(s_buf + b * L::BUF_BYTES, kin + size_t(t) * TILE, TILE_BYTES, s_mbar + b * 8);
    }
  };

  if (tid == 0)
  {
    mbar_init(s_mbar, 1);
    mbar_init(s_mbar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t t0 = atomicAdd(a.tile_counter, 1u);
    sts32(s_misc + 32, t0);
    request(t0, 0);
  }
  __syncthreads();
  uint32_t tile   = lds32(s_misc + 32);
  uint32_t b      = 0;
  uint32_t parity = 0; // bit b = phase of buffer b's mbarrier
  while (tile < a.num_tiles)
  {
    // ticket of the next tile (thread 0); its keys are requested once this tile's keys are in registers
    uint32_t next = 0;
    if (tid == 0)
    {
      next = atomicAdd(a.tile_counter, 1u);
      sts32(s_misc + 44, 0); // single-digit-tile flag
    }
    {
      constexpr int WORDS = L::NW * RADIX * L::CTR_BYTES / 4;
#pragma unroll
      for (int j = 0; j < WORDS / NT; ++j)
      {
        sts32(sbase + L::OFF_WARP + (j * NT + tid) * 4, 0);
      }
      static_assert(WORDS % NT == 0, "counter words must divide evenly over the threads");
    }
    const uint32_t tile_base = tile * uint32_t(TILE);
    const uint32_t valid     = min(uint32_t(TILE), a.num_items - tile_base);
    const uint32_t s_in      = s_buf + b * L::BUF_BYTES;
    if (valid == uint32_t(TILE))
    {
      mbar_wait(s_mbar + b * 8, (parity >> b) & 1u);
      parity ^= 1u << b;
    }
    __syncthreads(); // counters are zero; the previous tile's staged items (other buffer) are dead
    // after ranking (so that the ticket's round trip never stalls warp 0): publish the next ticket, request its keys
    auto prefetch = [&]() {
      if (tid == 0)
      {
        sts32(s_misc + 36, next);
        request(next, b ^ 1u);
      }
    };
    if (valid == uint32_t(TILE))
    {
      onesweep_tile<U, VBYTES, NT, IPT, RANK, OPT, FLOATK, BIG, true, true>(a, sbase, tile, tile_base, valid, s_in, s_in,
                                                                            prefetch);
    }
    else
    {
      onesweep_tile<U, VBYTES, NT, IPT, RANK, OPT, FLOATK, BIG, false, false>(a, sbase, tile, tile_base, valid, s_in,
                                                                              s_in, prefetch);
    }
    if (tid < RADIX && a.lookback_next != nullptr && tile < a.lookback_next_tiles)
    {
      a.lookback_next[size_t(tile) * RADIX + tid] = 0;
    }
    __syncthreads(); // staged items consumed; next ticket visible
    tile = lds32(s_misc + 36);
    b ^= 1u;
  }
  // rows of the next launch's status array beyond this launch's tiles
  if (tid < RADIX && a.lookback_next != nullptr)
  {
    for (uint32_t t = a.num_tiles + blockIdx.x; t < a.lookback_next_tiles; t += gridDim.x)
    {
      a.lookback_next[size_t(t) * RADIX + tid] = 0;
    }
  }
}


} // namespace b200rs
