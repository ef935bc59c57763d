// A/B harness for ONE onesweep digit pass: times a list of kernel variants on the same device-resident keys and checks
// every variant's output against variant 0 bit for bit.  Builds in seconds (only the variants listed here), unlike the
// full library -- this is where kernel experiments are measured before they go into the config tables.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I cccl_b200/csrc \
//        tools/ubench/onesweep_ab.cu -o tools/bin/onesweep_ab
//   tools/bin/onesweep_ab [log2n=28] [dist=uniform|equal|few16|sorted] [shift=0] [reps=10]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "histogram.cuh"
#include "onesweep.cuh"
#include "onesweep_persistent.cuh"

using namespace b200rs;

#define CK(x)                                                                      \
  do                                                                               \
  {                                                                                \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess)                                                         \
    {                                                                              \
      fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ uint32_t mix(uint64_t x)
{
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return uint32_t(x);
}

__global__ void gen_keys(uint32_t* k, size_t n, int dist)
{
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
  {
    uint32_t v = mix(i + 0x9e3779b97f4a7c15ull);
    if (dist == 1)
    {
      v = 0x01234567u;
    }
    else if (dist == 2)
    {
      v = mix(v & 15u);
    }
    else if (dist == 3)
    {
      v = uint32_t((i * 4294967296.0) / double(n)); // already sorted
    }
    else if (dist == 4)
    {
      v &= mix(i * 3 + 1) & mix(i * 5 + 2) & mix(i * 7 + 3) & mix(i * 11 + 4); // AND of five: bit entropy 0.201
    }
    k[i] = v;
  }
}

__global__ void count_mismatch(const uint32_t* a, const uint32_t* b, size_t n, unsigned long long* out)
{
  unsigned long long bad = 0;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
  {
    bad += a[i] != b[i];
  }
  if (bad)
  {
    atomicAdd(out, bad);
  }
}

struct Ctx
{
  uint32_t *keys, *out, *ref;
  uint32_t *lookback, *tile_counter;
  unsigned long long* bins;
  unsigned long long* mism;
  size_t n;
  int shift, reps, sms;
};

template <int NT, int IPT, int MINB, int OPT, bool PERSIST>
void run_variant(Ctx& c, const char* name, bool is_ref)
{
  using U            = uint32_t;
  constexpr int TILE = NT * IPT;
  const unsigned tiles = unsigned((c.n + TILE - 1) / TILE);
  PassArgs a;
  memset(&a, 0, sizeof(a));
  a.keys_in      = c.keys;
  a.keys_out     = is_ref ? c.ref : c.out;
  a.lookback     = c.lookback;
  a.tile_counter = c.tile_counter;
  a.bins         = c.bins + (c.shift / 8) * RADIX;
  a.num_items    = uint32_t(c.n);
  a.num_tiles    = tiles;
  a.all_ones     = 0xffffffffu;
  a.shift        = c.shift;
  a.mask         = 0xff;
  a.first_pass   = 0;
  a.last_pass    = 0;
  a.xf           = make_xform(0, 4, 0);
  a.sm_count     = c.sms;
  size_t smem;
  void (*kernel)(const PassArgs);
  unsigned grid = tiles;
  if constexpr (PERSIST)
  {
    smem   = PersistSmem<U, 0, NT, IPT, OPT>::BYTES;
    kernel = onesweep_persistent_kernel<U, 0, NT, IPT, RANK_BALLOT, MINB, OPT, false, false>;
    grid   = tiles < unsigned(c.sms * MINB) ? tiles : unsigned(c.sms * MINB);
  }
  else
  {
    smem   = OnesweepSmem<U, 0, NT, IPT, OPT>::BYTES;
    kernel = onesweep_kernel<U, 0, NT, IPT, RANK_BALLOT, MINB, OPT, false, false>;
  }
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e9f, sum = 0.f;
  for (int r = 0; r < c.reps + 2; ++r)
  {
    CK(cudaMemsetAsync(c.lookback, 0, size_t(tiles) * RADIX * 4));
    CK(cudaMemsetAsync(c.tile_counter, 0, 4));
    CK(cudaEventRecord(e0));
    kernel<<<grid, NT, smem>>>(a);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 2)
    {
      best = ms < best ? ms : best;
      sum += ms;
    }
  }
  unsigned long long bad = 0;
  if (!is_ref)
  {
    CK(cudaMemset(c.mism, 0, 8));
    count_mismatch<<<1184, 256>>>(c.out, c.ref, c.n, c.mism);
    CK(cudaMemcpy(&bad, c.mism, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemsetAsync(c.out, 0xff, c.n * 4));
  }
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, NT, smem);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kernel);
  printf("%-34s NT=%d IPT=%d occ=%d regs=%d smem=%zu  avg %.4f ms  best %.4f ms  %.0f GB/s  %s\n", name, NT, IPT, occ,
         fa.numRegs, smem, sum / c.reps, best, 2.0 * c.n * 4 / (sum / c.reps) / 1e6,
         is_ref ? "(reference output)" : (bad ? "MISMATCH" : "ok"));
  fflush(stdout);
}

int main(int argc, char** argv)
{
  Ctx c;
  const int log2n  = argc > 1 ? atoi(argv[1]) : 28;
  const char* dist = argc > 2 ? argv[2] : "uniform";
  c.shift          = argc > 3 ? atoi(argv[3]) : 0;
  c.reps           = argc > 4 ? atoi(argv[4]) : 10;
  c.n              = size_t(1) << log2n;
  const int d = !strcmp(dist, "equal") ? 1 : !strcmp(dist, "few16") ? 2 : !strcmp(dist, "sorted") ? 3
                : !strcmp(dist, "entropy5") ? 4 : 0;
  CK(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaMalloc(&c.keys, c.n * 4));
  CK(cudaMalloc(&c.out, c.n * 4));
  CK(cudaMalloc(&c.ref, c.n * 4));
  CK(cudaMalloc(&c.lookback, (c.n / 2048 + 2) * RADIX * 4));
  CK(cudaMalloc(&c.tile_counter, 256));
  CK(cudaMalloc(&c.bins, 4 * RADIX * 8));
  CK(cudaMalloc(&c.mism, 8));
  gen_keys<<<1184, 256>>>(c.keys, c.n, d);
  CK(cudaMemset(c.bins, 0, 4 * RADIX * 8));
  {
    using L     = HistLayout<4>;
    auto kernel = histogram_kernel<uint32_t, 4, 0>;
    const size_t smem = size_t(4 + 1) * RADIX * L::REPLICAS * 4; // + the alignment slack of the folded addressing
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    kernel<<<c.sms, HIST_THREADS, smem>>>(c.keys, c.n, c.bins, 0, 32, make_xform(0, 4, 0), nullptr);
    scan_bins_kernel<<<4, RADIX>>>(c.bins);
  }
  CK(cudaDeviceSynchronize());
  printf("n=2^%d dist=%s shift=%d reps=%d sms=%d\n", log2n, dist, c.shift, c.reps, c.sms);

#define V(NT, IPT, MINB, OPT, PERSIST, REF) run_variant<NT, IPT, MINB, OPT, PERSIST>(c, #NT "x" #IPT " minb" #MINB " opt" #OPT " p" #PERSIST, REF)
  V(256, 40, 3, 7, false, true);
#include "onesweep_ab_variants.inc"
  V(256, 40, 3, 7, false, false);
  return 0;
}
