// Whole sort inside ONE CTA for inputs of at most one tile (N <= 5120 / 2560 / 1280 items for a dominant item size of
// <= 4 / 8 / 16 bytes): one read of the input, every digit pass in shared memory, one write of the output.
//
// Replaces DeviceRadixSortSingleTileKernel -> BlockRadixSort::SortBlockedToStriped
// (/root/reference/cub/cub/device/dispatch/kernels/kernel_radix_sort.cuh:330-434, cub/cub/block/block_radix_sort.cuh:436;
//  chosen by dispatch_radix_sort.cuh:1980 for N <= 4864 / 2304 / 1024).  The capacities here are at least the
// reference's, so every N the reference sorts with its single-tile kernel is a single launch here too.  Latency-bound by
// construction (1 launch instead of 3 + passes); the float -0.0 rule of the reference's single-tile regime is carried by
// the KeyXform the host builds (common.cuh make_xform, single_tile_rule).
//
// Per pass: the same warp-private match-by-ballot ranking as the onesweep kernel, a block scan of the digit totals,
// keys (and values) staged in digit order in shared memory, then read back in the warp-striped arrangement
// (item i of lane l of warp w = element w*32*IPT + i*32 + l), which is the tile order of the next pass.
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

template <class U, int VBYTES, int IPT, bool FLOATK>
__global__ void __launch_bounds__(ST_THREADS) single_tile_kernel(const SingleTileArgs a)
{
  using L = SingleTileSmem<U, VBYTES, IPT>;
  using V = typename value_of<VBYTES>::type;
  constexpr int NW = L::NW;
  extern __shared__ __align__(16) unsigned char smem[];
  const uint32_t sbase  = uint32_t(__cvta_generic_to_shared(smem));
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
  const uint32_t chunk = warp * 32 * IPT + lane;
#pragma unroll
  for (int i = 0; i < IPT; ++i)
  {
    const uint32_t p = chunk + i * 32;
    key[i]           = p < n ? twiddle_in(static_cast<const U*>(a.keys_in)[p], xf) : U(~U(0)); // padding sorts last
    if (VBYTES > 0 && p < n)
    {
      val[i] = static_cast<const V*>(a.vals_in)[p];
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
      const uint32_t d = pass_digit<FLOATK>(key[i], bit, mask, neg_zero, pos_zero);
      const uint32_t r = rank[i] + ctr_ld<true>(s_mine + d * 2);
      sts_t<U>(s_keys + r * uint32_t(sizeof(U)), key[i]);
      if (VBYTES > 0)
      {
        sts_t<V>(s_vals + r * uint32_t(sizeof(V)), val[i]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
      const uint32_t p = chunk + i * 32;
      key[i]           = lds_t<U>(s_keys + p * uint32_t(sizeof(U)));
      if (VBYTES > 0)
      {
        val[i] = lds_t<V>(s_vals + p * uint32_t(sizeof(V)));
      }
    }
  }

#pragma unroll
  for (int i = 0; i < IPT; ++i)
  {
    const uint32_t p = chunk + i * 32;
    if (p < n)
    {
      static_cast<U*>(a.keys_out)[p] = twiddle_out(key[i], xf);
      if (VBYTES > 0)
      {
        static_cast<V*>(a.vals_out)[p] = val[i];
      }
    }
  }
}

// items per thread by dominant item size: 20 (<= 4 bytes), 10 (8 bytes), 5 (16 bytes)
template <class U, int VB>
struct SingleTileShape
{
  static constexpr int DOM = int(sizeof(U)) > VB ? int(sizeof(U)) : VB;
  static constexpr int IPT = DOM <= 4 ? 20 : (DOM <= 8 ? 10 : 5);
};

unsigned long long single_tile_capacity(int key_bytes, int value_bytes)
{
  const int dom = key_bytes > value_bytes ? key_bytes : value_bytes;
  return (unsigned long long) ST_THREADS * (dom <= 4 ? 20 : (dom <= 8 ? 10 : 5));
}

template <class U, int VB>
static cudaError_t launch_st(const SingleTileArgs& a, cudaStream_t stream)
{
  constexpr int IPT        = SingleTileShape<U, VB>::IPT;
  using L                  = SingleTileSmem<U, VB, IPT>;
  constexpr bool CAN_FLOAT = sizeof(U) >= 2; // half / bfloat16, float, double
  auto kernel              = single_tile_kernel<U, VB, IPT, false>;
  if (CAN_FLOAT && a.xf.float_mask != 0)
  {
    kernel = single_tile_kernel<U, VB, IPT, CAN_FLOAT>;
  }
  if (L::BYTES > 48 * 1024)
  {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(L::BYTES));
    if (e != cudaSuccess)
    {
      return e;
    }
  }
  kernel<<<1, ST_THREADS, L::BYTES, stream>>>(a);
  return cudaPeekAtLastError();
}

template <class U>
static cudaError_t launch_st_v(int value_bytes, const SingleTileArgs& a, cudaStream_t stream)
{
  switch (value_bytes)
  {
    case 0: return launch_st<U, 0>(a, stream);
    case 1: return launch_st<U, 1>(a, stream);
    case 2: return launch_st<U, 2>(a, stream);
    case 4: return launch_st<U, 4>(a, stream);
    case 8: return launch_st<U, 8>(a, stream);
    case 16: return launch_st<U, 16>(a, stream);
    default: return cudaErrorNotSupported;
  }
}

cudaError_t launch_single_tile(
  const void* keys_in, void* keys_out, const void* vals_in, void* vals_out, unsigned long long n, int key_bytes,
  int value_bytes, int begin_bit, int end_bit, const KeyXform& xf, cudaStream_t stream)
{
  SingleTileArgs a;
  a.keys_in   = keys_in;
  a.keys_out  = keys_out;
  a.vals_in   = vals_in;
  a.vals_out  = vals_out;
  a.num_items = uint32_t(n);
  a.begin_bit = begin_bit;
  a.end_bit   = end_bit;
  a.all_ones  = 0xffffffffu;
  a.xf        = xf;
  switch (key_bytes)
  {
    case 1: return launch_st_v<uint8_t>(value_bytes, a, stream);
    case 2: return launch_st_v<uint16_t>(value_bytes, a, stream);
    case 4: return launch_st_v<uint32_t>(value_bytes, a, stream);
    case 8: return launch_st_v<uint64_t>(value_bytes, a, stream);
    default: return cudaErrorNotSupported;
  }
}

} // namespace b200rs
