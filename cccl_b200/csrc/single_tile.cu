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
#include "single_tile.cuh"

namespace b200rs
{

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
