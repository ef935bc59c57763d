// TEST INFRASTRUCTURE ONLY -- CPU oracle for the LSD radix-sort hot path.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library, and only as the checker.  Nothing under cccl_b200/ or include/ links,
// imports or calls it; the product path fails loudly when the CUDA library is missing.
//
// Parity status: PINNED.  The restatement is checked (tests/test_oracle.py) against
//   * the reference's golden vectors (cub/test/catch2_test_device_radix_sort_env_api.cu:84-136,
//     thrust/testing/sort_by_key.cu:46-53, thrust/testing/sort.cu:40-48),
//   * the reference's own Thrust OMP/CPP sort built from /root/reference (oracle/_ref, see
//     oracle/Makefile) on seeded inputs, and the fixtures under tests/golden/ generated from it.
//
// What is restated (reference file:line, all relative to /root/reference):
//   * get_striped_keys        cub/test/catch2_radix_sort_helper.cuh:174-214
//       bit-cast -> Traits<T>::TwiddleIn -> (descending: ~) -> (float: pattern of -0 -> pattern of +0, the
//       device rule of radix_rank_sort_operations.cuh:44-82) -> mask to [begin_bit,end_bit)
//   * Traits<T>::TwiddleIn    cub/cub/util_type.cuh:857-865 (unsigned: identity),
//                             :906-914 (signed: flip sign bit),
//                             :953-963 (float: negative -> ~x, else flip sign bit)
//   * get_permutation         cub/test/catch2_radix_sort_helper.cuh:238-268
//       std::stable_sort of an index permutation with '<' (ascending) or '>' (descending)
//       on the bit-ordered keys => equal keys keep INPUT order in both directions
//   * radix_sort_reference    cub/test/catch2_radix_sort_helper.cuh:270-312 (gather keys, values)
//   * all-pass digit histogram cub/cub/agent/agent_radix_sort_histogram.cuh:205-224
//       (bins[pass][digit] of the twiddled key, used to check the upsweep kernel on its own)
//
// Plain C ABI so python (ctypes) can call it.  Build: see oracle/Makefile.

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

namespace
{

enum key_kind
{
  KIND_UINT  = 0,
  KIND_INT   = 1,
  KIND_FLOAT = 2
};

template <class U>
U twiddle_in(U bits, int kind)
{
  constexpr U high = U(1) << (sizeof(U) * 8 - 1);
  switch (kind)
  {
    case KIND_INT:
      return bits ^ high; // util_type.cuh:906-914
    case KIND_FLOAT: {
      U mask = (bits & high) ? U(~U(0)) : high; // util_type.cuh:953-963
      return bits ^ mask;
    }
    default:
      return bits; // util_type.cuh:857-865
  }
}

// Largest N the reference sorts with its single-CTA kernel on sm_100 (dispatch_radix_sort.cuh:1980):
// single_tile policy = scale_reg_bound(256 threads, 19 items, max(key, value size))
// (tuning_radix_sort.cuh:1747,1833-1841; util_arch.cuh:128-138) => 4864 / 2304 / 1024 items for a dominant item
// size of <=4 / 8 / 16 bytes.
inline uint64_t reference_single_tile_items(int key_bytes, int value_bytes)
{
  const int dom     = std::max(4, std::max(key_bytes, value_bytes));
  const int items   = std::max(1, 19 * 4 / dom);
  const int threads = std::min(256, (48 * 1024 / (dom * items) + 31) / 32 * 32);
  return uint64_t(threads) * uint64_t(items);
}

// catch2_radix_sort_helper.cuh:174-214, with the float-zero rule taken from the DEVICE code it models
// (radix_rank_sort_operations.cuh:44-82).  The reference has TWO device rules and which one runs depends on N:
//   * onesweep (N above the single-tile size): the key is twiddled, then (descending) inverted, and in THAT domain
//     the pattern TwiddleIn(-0.0) = 0x7f..f is replaced by TwiddleIn(+0.0) = 0x80..0 before the digit bits are taken;
//   * single-tile kernel (N <= reference_single_tile_items; BlockRadixSort handles descending by reversing digits,
//     not by inverting keys): the replacement happens in the NON-inverted domain, which is exactly the helper's
//     "-0 -> +0 before twiddling".
// For a full-width sort the two are indistinguishable; for a descending sort on a partial bit window they differ in
// where the zeros land.  The real cub::DeviceRadixSort outputs pin both: tests/golden/cub_*.npz (small N, single
// tile; large N, onesweep) and tests/test_vs_reference_gpu.py on the B200.
template <class U>
std::vector<U> striped_keys(
  const U* keys, uint64_t n, int kind, int begin_bit, int end_bit, bool descending, bool single_tile_rule)
{
  constexpr int total_bits = int(sizeof(U) * 8);
  constexpr U high         = U(1) << (total_bits - 1);
  std::vector<U> out(n);
  const int num_bits = end_bit - begin_bit;
  for (uint64_t i = 0; i < n; ++i)
  {
    U key = twiddle_in(keys[i], kind);
    if (kind == KIND_FLOAT && single_tile_rule && key == U(~high))
    {
      key = high; // ProcessFloatMinusZero on the un-inverted key (block_radix_sort path)
    }
    if (descending)
    {
      key = U(~key); // radix_rank_sort_operations.cuh:545-552
    }
    if (kind == KIND_FLOAT && !single_tile_rule && key == U(~high))
    {
      key = high; // ProcessFloatMinusZero, radix_rank_sort_operations.cuh:69-82 (onesweep path)
    }
    if (begin_bit > 0 || end_bit < total_bits)
    {
      // ((1 << num_bits) - 1) << begin_bit, written so num_bits == total_bits cannot overflow
      U m = (num_bits >= total_bits) ? U(~U(0)) : U((U(1) << num_bits) - 1);
      key &= U(m << begin_bit);
    }
    out[i] = key;
  }
  return out;
}

// catch2_radix_sort_helper.cuh:238-268.  The keys above are already inverted for descending sorts, so one ascending
// stable sort serves both directions (equivalent to the helper's '>' on the un-inverted keys: inversion reverses
// the order of distinct window values and keeps ties, which the stable sort leaves in input order).
template <class U>
std::vector<uint64_t> permutation(
  const U* keys, uint64_t n, int kind, int begin_bit, int end_bit, bool descending, bool single_tile_rule)
{
  std::vector<U> sk = striped_keys(keys, n, kind, begin_bit, end_bit, descending, single_tile_rule);
  std::vector<uint64_t> perm(n);
  std::iota(perm.begin(), perm.end(), uint64_t(0));
  const U* p = sk.data();
  std::stable_sort(perm.begin(), perm.end(), [p](uint64_t a, uint64_t b) {
    return p[a] < p[b];
  });
  return perm;
}

template <class U>
int sort_impl(const void* keys_in,
              void* keys_out,
              const void* vals_in,
              void* vals_out,
              uint64_t n,
              int kind,
              int value_bytes,
              int begin_bit,
              int end_bit,
              int descending,
              int zero_rule = -1) // -1: by N like cub::DeviceRadixSort; 1 / 0: force the single-tile / onesweep rule
{
  const U* kin = static_cast<const U*>(keys_in);
  U* kout      = static_cast<U*>(keys_out);
  const bool single_tile_rule =
    zero_rule < 0 ? n <= reference_single_tile_items(int(sizeof(U)), value_bytes) : zero_rule != 0;
  std::vector<uint64_t> perm  = permutation<U>(kin, n, kind, begin_bit, end_bit, descending != 0, single_tile_rule);
  // catch2_radix_sort_helper.cuh:281-284, :305-309 (gather)
  for (uint64_t i = 0; i < n; ++i)
  {
    kout[i] = kin[perm[i]];
  }
  if (value_bytes > 0 && vals_in && vals_out)
  {
    const unsigned char* vin = static_cast<const unsigned char*>(vals_in);
    unsigned char* vout      = static_cast<unsigned char*>(vals_out);
    for (uint64_t i = 0; i < n; ++i)
    {
      std::memcpy(vout + i * value_bytes, vin + perm[i] * value_bytes, value_bytes);
    }
  }
  return 0;
}

// agent_radix_sort_histogram.cuh:205-224 -- digits are taken from the twiddled (and, for descending,
// inverted: radix_rank_sort_operations.cuh:545-565) key with -0 mapped to +0 (:69-82).
template <class U>
int hist_impl(const void* keys_in, uint64_t n, int kind, int begin_bit, int end_bit, int descending, uint64_t* bins)
{
  const U* kin             = static_cast<const U*>(keys_in);
  constexpr int total_bits = int(sizeof(U) * 8);
  constexpr U high         = U(1) << (total_bits - 1);
  const int passes         = (end_bit - begin_bit + 7) / 8;
  std::fill(bins, bins + uint64_t(passes) * 256, uint64_t(0));
  for (uint64_t i = 0; i < n; ++i)
  {
    U key = twiddle_in(kin[i], kind);
    if (descending)
    {
      key = U(~key);
    }
    if (kind == KIND_FLOAT && key == U(~high))
    {
      key = high;
    }
    for (int p = 0; p < passes; ++p)
    {
      int bit   = begin_bit + 8 * p;
      int nbits = std::min(8, end_bit - bit);
      bins[p * 256 + ((key >> bit) & ((U(1) << nbits) - 1))]++;
    }
  }
  return 0;
}

} // namespace

extern "C" {

// keys are passed as raw bit patterns of width key_bytes; key_kind in {0 uint, 1 int, 2 float}.
// Returns 0 on success, -1 on unsupported key width.
int oracle_radix_sort(
  const void* keys_in,
  void* keys_out,
  const void* vals_in,
  void* vals_out,
  uint64_t n,
  int key_kind,
  int key_bytes,
  int value_bytes,
  int begin_bit,
  int end_bit,
  int descending)
{
  switch (key_bytes)
  {
    case 1:
      return sort_impl<uint8_t>(keys_in, keys_out, vals_in, vals_out, n, key_kind, value_bytes, begin_bit, end_bit, descending);
    case 2:
      return sort_impl<uint16_t>(keys_in, keys_out, vals_in, vals_out, n, key_kind, value_bytes, begin_bit, end_bit, descending);
    case 4:
      return sort_impl<uint32_t>(keys_in, keys_out, vals_in, vals_out, n, key_kind, value_bytes, begin_bit, end_bit, descending);
    case 8:
      return sort_impl<uint64_t>(keys_in, keys_out, vals_in, vals_out, n, key_kind, value_bytes, begin_bit, end_bit, descending);
    default:
      return -1;
  }
}

// One SEGMENT of cub::DeviceSegmentedRadixSort: the same order, except that the segmented kernel never inverts keys
// for descending sorts (it reverses the digit: cub/cub/device/dispatch/kernels/kernel_segmented_radix_sort.cuh:213-272,
// agent_radix_sort_downsweep.cuh), so the -0.0 -> +0.0 replacement happens in the un-inverted domain at EVERY segment
// size -- the rule the unsegmented reference only shows below its single-tile size.  Pinned by
// tests/golden/segmented/cubseg_f32_*.npz (real cub::DeviceSegmentedRadixSort outputs, short and long segments).
int oracle_radix_sort_segment(
  const void* keys_in,
  void* keys_out,
  const void* vals_in,
  void* vals_out,
  uint64_t n,
  int key_kind,
  int key_bytes,
  int value_bytes,
  int begin_bit,
  int end_bit,
  int descending)
{
  switch (key_bytes)
  {
    case 1:
      return sort_impl<uint8_t>(keys_in, keys_out, vals_in, vals_out, n, key_kind, value_bytes, begin_bit, end_bit, descending, 1);
    case 2:
      return sort_impl<uint16_t>(keys_in, keys_out, vals_in, vals_out, n, key_kind, value_bytes, begin_bit, end_bit, descending, 1);
    case 4:
      return sort_impl<uint32_t>(keys_in, keys_out, vals_in, vals_out, n, key_kind, value_bytes, begin_bit, end_bit, descending, 1);
    case 8:
      return sort_impl<uint64_t>(keys_in, keys_out, vals_in, vals_out, n, key_kind, value_bytes, begin_bit, end_bit, descending, 1);
    default:
      return -1;
  }
}

// bins: uint64[passes*256], passes = ceil((end_bit-begin_bit)/8).
int oracle_digit_histogram(
  const void* keys_in, uint64_t n, int key_kind, int key_bytes, int begin_bit, int end_bit, int descending, uint64_t* bins)
{
  switch (key_bytes)
  {
    case 1:
      return hist_impl<uint8_t>(keys_in, n, key_kind, begin_bit, end_bit, descending, bins);
    case 2:
      return hist_impl<uint16_t>(keys_in, n, key_kind, begin_bit, end_bit, descending, bins);
    case 4:
      return hist_impl<uint32_t>(keys_in, n, key_kind, begin_bit, end_bit, descending, bins);
    case 8:
      return hist_impl<uint64_t>(keys_in, n, key_kind, begin_bit, end_bit, descending, bins);
    default:
      return -1;
  }
}

// One stable 8-bit counting-sort pass (what one onesweep launch must produce;
// agent_radix_sort_onesweep.cuh:667-697 seen from outside).  Keys here are already bit-ordered.
int oracle_counting_pass_u32(const uint32_t* in, uint32_t* out, uint64_t n, int bit, int nbits)
{
  std::vector<uint64_t> off(257, 0);
  const uint32_t mask = (1u << nbits) - 1;
  for (uint64_t i = 0; i < n; ++i)
  {
    off[((in[i] >> bit) & mask) + 1]++;
  }
  for (int d = 0; d < 256; ++d)
  {
    off[d + 1] += off[d];
  }
  for (uint64_t i = 0; i < n; ++i)
  {
    out[off[(in[i] >> bit) & mask]++] = in[i];
  }
  return 0;
}

} // extern "C"
