// Shared definitions for the B200 (sm_100a) LSD radix-sort kernels.
//
// Reference behaviour restated here (paths relative to /root/reference):
//   * key transforms            cub/cub/util_type.cuh:857-865, :906-914, :953-963 (Traits<T>::TwiddleIn/Out),
//                               cub/cub/block/radix_rank_sort_operations.cuh:533-573 (descending = extra ~)
//   * -0.0 == +0.0 for digits   cub/cub/block/radix_rank_sort_operations.cuh:44-82
//   * digit = (x >> bit) & mask cub/cub/block/radix_rank_sort_operations.cuh:108-126
// Unlike the reference (which twiddles on every pass load and un-twiddles on every pass store,
// agent_radix_sort_onesweep.cuh:355-359,535), keys live in HBM in their bit-ordered form between
// passes: the transform is applied by the upsweep histogram and the FIRST pass on load, and undone by
// the LAST pass on store.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

namespace b200rs
{

constexpr int RADIX_BITS = 8;
constexpr int RADIX      = 1 << RADIX_BITS;

// Look-back status word, one per (tile, digit): [31] inclusive prefix ready, [30] tile count ready,
// [29:0] value.  A portion therefore holds < 2^30 items (host splits larger inputs).
constexpr uint32_t LB_INCLUSIVE  = 0x80000000u;
constexpr uint32_t LB_PARTIAL    = 0x40000000u;
constexpr uint32_t LB_FLAG_MASK  = 0xC0000000u;
constexpr uint32_t LB_VALUE_MASK = 0x3FFFFFFFu;

// Runtime description of the key transform, as 64-bit fields narrowed inside the kernels.
struct KeyXform
{
  unsigned long long float_mask; // ~0 for floating-point keys, else 0
  unsigned long long sign_mask;  // high bit for signed / float keys, else 0
  unsigned long long desc_mask;  // ~0 for descending, else 0
  unsigned long long neg_zero;   // bit-ordered pattern of -0.0 (after desc), == pos_zero for non-float keys
  unsigned long long pos_zero;   // bit-ordered pattern of +0.0 (after desc)
};

template <class U>
struct XformT
{
  U float_mask, sign_mask, desc_mask, neg_zero, pos_zero;
  __host__ __device__ explicit XformT(const KeyXform& x)
      : float_mask(U(x.float_mask))
      , sign_mask(U(x.sign_mask))
      , desc_mask(U(x.desc_mask))
      , neg_zero(U(x.neg_zero))
      , pos_zero(U(x.pos_zero))
  {}
};

template <class U>
__host__ __device__ __forceinline__ U sign_fill(U x)
{
  using S = typename std::make_signed<U>::type;
  return U(S(x) >> (sizeof(U) * 8 - 1)); // all ones iff the top bit is set
}

// user bits -> bit-ordered key (ascending order of the result == requested order of the keys)
template <class U>
__host__ __device__ __forceinline__ U twiddle_in(U bits, const XformT<U>& x)
{
  U m = U((sign_fill(bits) & x.float_mask) | x.sign_mask);
  return U(bits ^ m ^ x.desc_mask);
}

// inverse of twiddle_in
template <class U>
__host__ __device__ __forceinline__ U twiddle_out(U t, const XformT<U>& x)
{
  U y = U(t ^ x.desc_mask);
  U m = U((U(~sign_fill(y)) & x.float_mask) | x.sign_mask);
  return U(y ^ m);
}

// digit-extraction view of a bit-ordered key: -0.0 is ranked as +0.0, stored bits are untouched
template <class U>
__host__ __device__ __forceinline__ U digit_view(U t, const XformT<U>& x)
{
  return t == x.neg_zero ? x.pos_zero : t;
}

template <class U>
__host__ __device__ __forceinline__ uint32_t digit_of(U view, int shift, uint32_t mask)
{
  return uint32_t(view >> shift) & mask;
}

// Largest N the reference sorts with its single-CTA kernel on sm_100 (dispatch_radix_sort.cuh:1980; policy
// tuning_radix_sort.cuh:1747,1833-1841 scaled by util_arch.cuh:128-138): 4864 / 2304 / 1024 items for a dominant
// item size of <=4 / 8 / 16 bytes.  Only used to pick the reference's float-zero rule (see make_xform).
inline unsigned long long reference_single_tile_items(int key_bytes, int value_bytes)
{
  int dom = key_bytes > value_bytes ? key_bytes : value_bytes;
  dom     = dom < 4 ? 4 : dom;
  int items = 19 * 4 / dom;
  items     = items < 1 ? 1 : items;
  int threads = (48 * 1024 / (dom * items) + 31) / 32 * 32;
  threads     = threads > 256 ? 256 : threads;
  return (unsigned long long) threads * (unsigned long long) items;
}

inline KeyXform make_xform(int key_kind, int key_bytes, int descending, bool single_tile_rule = false)
{
  const int bits                 = key_bytes * 8;
  const unsigned long long all   = bits == 64 ? ~0ull : ((1ull << bits) - 1);
  const unsigned long long high  = 1ull << (bits - 1);
  KeyXform x;
  x.float_mask = key_kind == 2 ? all : 0;
  x.sign_mask  = key_kind != 0 ? high : 0;
  x.desc_mask  = descending ? all : 0;
  if (key_kind == 2)
  {
    // Exactly the reference's rule (radix_rank_sort_operations.cuh:44-82): in the kernel's key domain (after the
    // descending inversion) the pattern TwiddleIn(-0.0) = 0x7f..f is ranked as TwiddleIn(+0.0) = 0x80..0.  For
    // ascending sorts that maps -0.0 onto +0.0; for descending sorts the inverted +0.0 IS 0x7f..f and is mapped
    // onto the inverted -0.0.  Either way both zeros share one digit view, so they tie and keep input order.
    x.neg_zero = (~high) & all;
    x.pos_zero = high;
    // The reference's single-CTA kernel (small N) does not invert keys for descending sorts; it maps -0.0 onto +0.0
    // in the un-inverted domain (radix_rank_sort_operations.cuh:44-55 "all other sorting implementations").  In the
    // inverted domain this kernel works in, that is the opposite replacement.  Only observable for descending sorts
    // on a partial bit window; callers pass single_tile_rule = (N <= reference_single_tile_items) to stay bit-exact
    // with the reference at every N.
    if (descending && single_tile_rule)
    {
      x.neg_zero = high;
      x.pos_zero = (~high) & all;
    }
  }
  else
  {
    x.neg_zero = 0;
    x.pos_zero = 0;
  }
  return x;
}

template <int BYTES>
struct uint_of;
template <>
struct uint_of<1>
{
  using type = uint8_t;
};
template <>
struct uint_of<2>
{
  using type = uint16_t;
};
template <>
struct uint_of<4>
{
  using type = uint32_t;
};
template <>
struct uint_of<8>
{
  using type = uint64_t;
};

// Value payloads are moved as opaque blobs.  16-byte values only promise 8-byte alignment.
struct alignas(8) blob16
{
  unsigned long long a, b;
};
template <int BYTES>
struct value_of
{
  using type = typename uint_of<BYTES>::type;
};
template <>
struct value_of<16>
{
  using type = blob16;
};
template <>
struct value_of<0>
{
  using type = uint8_t; // unused
};

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p)
{
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v)
{
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t lane_id()
{
  uint32_t r;
  asm("mov.u32 %0, %%laneid;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t lanemask_lt()
{
  uint32_t r;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(r));
  return r;
}

// Bucket mode with remote destinations (multi-GPU: the partition pass stores straight into the receive buffers of the
// destination GPUs over NVLink).  The partitioned order is cut into `num_dests` contiguous segments, segment r going to
// rank r; every pointer is biased so that the element with partitioned index idx is stored at ptr + idx * item size.
// bucket_dst_*[b] is the pointer of the one rank a bucket goes to, or 0 when a segment boundary falls inside it.
struct PeerTable
{
  uint32_t num_dests;
  uint32_t pad;
  unsigned long long bucket_dst_keys[32];
  unsigned long long bucket_dst_vals[32];
  unsigned long long rank_dst_keys[16];
  unsigned long long rank_dst_vals[16];
  uint32_t seg_end[16]; // seg_end[r] = first partitioned index that does NOT go to ranks <= r
};

// Bucket mode driven from the device (multi-GPU sort, multi.cu): the kernel that ends the splitter selection writes the
// plan, the partition pass that follows it in the stream reads it -- no host round trip in between.
struct PartitionPlan
{
  uint32_t num_splitters;
  uint32_t status;                      // 0 = consistent; else MULTI_ERR_* bits (multi.cu)
  unsigned long long splitters[16];     // bit-ordered, strictly increasing
  unsigned long long bins[256];         // exclusive output offset of every bucket (PassArgs::bins points here)
  PeerTable peer;                       // PassArgs::peer points here
};

// Everything one digit pass over one portion needs (type-erased; kernels cast).
struct PassArgs
{
  const void* keys_in;
  void* keys_out;
  const void* vals_in;
  void* vals_out;
  uint32_t* lookback;      // [tiles][256] status words of THIS pass (zero on entry)
  uint32_t* lookback_next; // status words of the NEXT launch, zeroed by this one (or nullptr)
  uint32_t lookback_next_tiles; // rows of lookback_next to zero
  uint32_t* tile_counter;  // dynamic tile id (zero on entry)
  const unsigned long long* bins; // [256] exclusive global offsets of this pass (+ portion)
  unsigned long long* bins_next;  // next portion's offsets, written by the last tile (or nullptr)
  uint32_t num_items;      // items in this portion, < 2^30
  uint32_t num_tiles;      // ceil(num_items / tile items of the launched configuration)
  uint32_t all_ones;       // 0xffffffff, passed at run time so that x * all_ones + all_ones stays an IMAD (== ~x)
  int shift;               // first bit of the digit
  uint32_t mask;           // (1 << digit_bits) - 1
  int first_pass;          // transform on load
  int last_pass;           // inverse transform on store
  int big;                 // whole array has >= 2^32 items: 64-bit output offsets
  KeyXform xf;
  // bucket mode (multi-GPU partition pass): the "digit" of a key is its destination bucket against these splitters
  // (bit-ordered values, strictly increasing): 2 * #{splitters below the key} + [key equals a splitter]
  int num_splitters;
  unsigned long long splitters[15];
  const PeerTable* peer; // bucket mode: remote destinations (device memory), or nullptr for keys_out / vals_out
  const PartitionPlan* plan; // bucket mode: splitters come from this device-side plan instead of the fields above
  int sm_count;          // SMs of the current device (grid size of the persistent kernel)
  // floating-point keys: word the upsweep sets when the input holds a key with the bit pattern that ranks as the other
  // zero (-0.0 for ascending sorts); while it is 0 the pass runs without the per-key zero test.  nullptr = always test.
  const uint32_t* zero_flag = nullptr;
};

} // namespace b200rs
