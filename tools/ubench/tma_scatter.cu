// Micro-benchmark: how fast can one SM push many SMALL shared->global bulk copies (cp.async.bulk, "TMA 1D")?
// Models the onesweep scatter: each tile holds 256 digit runs of S bytes; run d of tile t goes to
// out + d*(bytes/256) + t*S (256 contiguous output streams).  Compared with the same bytes written by STG.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

template <int MODE> // 0 = bulk copies, 1 = STG from shared (coalesced inside each run)
__global__ void __launch_bounds__(256) scatter_kernel(unsigned char* out, size_t stream_bytes, int S, int tiles, unsigned* ctr)
{
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = uint32_t(__cvta_generic_to_shared(smem));
  __shared__ unsigned s_tile;
  for (int i = threadIdx.x; i < 256 * S / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = i;
  while (true)
  {
    __syncthreads();
    if (threadIdx.x == 0) s_tile = atomicAdd(ctr, 1u);
    __syncthreads();
    const unsigned t = s_tile;
    if (t >= unsigned(tiles)) break;
    if (MODE == 0)
    {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      unsigned char* dst = out + size_t(threadIdx.x) * stream_bytes + size_t(t) * S;
      bulk_s2g(dst, sbase + threadIdx.x * S, S);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    else
    {
      // thread j of every group of S/4 threads writes word j of a run; 256 threads cover 256*4/S... runs per step
      const int wpr = S / 4; // words per run
      for (int w = threadIdx.x; w < 256 * wpr; w += 256)
      {
        const int d = w / wpr, j = w % wpr;
        uint32_t v = reinterpret_cast<uint32_t*>(smem)[w];
        reinterpret_cast<uint32_t*>(out + size_t(d) * stream_bytes + size_t(t) * S)[j] = v;
      }
    }
  }
}

int main(int argc, char** argv)
{
  const size_t total = size_t(1) << 30; // 1 GiB written per launch
  unsigned char* out;
  unsigned* ctr;
  cudaMalloc(&out, total);
  cudaMalloc(&ctr, 4);
  int sizes[] = {16, 32, 64, 128, 256, 512};
  cudaFuncSetAttribute(scatter_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(scatter_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int mode = 0; mode < 2; ++mode)
    for (int S : sizes)
      for (int cps : {2, 4, 6})
      {
        const int tiles = int(total / (256 * size_t(S)));
        const size_t stream_bytes = total / 256;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float best = 1e9;
        for (int rep = 0; rep < 4; ++rep)
        {
          cudaMemset(ctr, 0, 4);
          cudaEventRecord(e0);
          if (mode == 0)
            scatter_kernel<0><<<148 * cps, 256, 256 * S>>>(out, stream_bytes, S, tiles, ctr);
          else
            scatter_kernel<1><<<148 * cps, 256, 256 * S>>>(out, stream_bytes, S, tiles, ctr);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          if (rep > 0 && ms < best) best = ms;
        }
        cudaError_t err = cudaGetLastError();
        printf("mode=%s S=%4d ctas/sm=%d  %.3f ms  %.1f GB/s  (%s)\n", mode ? "stg " : "bulk", S, cps, best,
               total / best / 1e6, cudaGetErrorString(err));
      }
  return 0;
}
