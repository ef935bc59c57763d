// Micro-benchmark: SM-issued stores into a PEER GPU's memory over NVLink -- does 128-byte line alignment of each
// warp's store (and the per-thread width) matter?  One process, two GPUs.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/ubench/peer_store.cu -o tools/bin/peer_store
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1);} } while (0)

// each warp stores runs of `run` 4-byte items starting `mis` items after a 128-byte boundary; VEC = items per thread
template <int VEC>
__global__ void store_kernel(uint32_t* dst, const uint32_t* src, size_t n, int mis)
{
  const size_t stride = size_t(gridDim.x) * blockDim.x * VEC;
  for (size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) * VEC; i + VEC <= n; i += stride)
  {
    if (VEC == 1)
    {
      dst[i + mis] = src[i];
    }
    else
    {
      const uint4 v = *reinterpret_cast<const uint4*>(src + i);
      *reinterpret_cast<uint4*>(dst + i + mis) = v; // mis must be a multiple of 4 here
    }
  }
}

int main()
{
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev < 2) { printf("needs 2 GPUs\n"); return 0; }
  const size_t n = size_t(1) << 28; // 1 GiB
  uint32_t *src, *dst_peer, *dst_local;
  CK(cudaSetDevice(1));
  CK(cudaMalloc(&dst_peer, (n + 64) * 4));
  CK(cudaSetDevice(0));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  CK(cudaMalloc(&src, n * 4));
  CK(cudaMalloc(&dst_local, (n + 64) * 4));
  CK(cudaMemset(src, 1, n * 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  auto run = [&](const char* name, auto kernel, uint32_t* dst, int mis, int grid) {
    float best = 1e9f;
    for (int r = 0; r < 5; ++r)
    {
      CK(cudaEventRecord(e0));
      kernel<<<grid, 256>>>(dst, src, n, mis);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      best = ms < best ? ms : best;
    }
    printf("%-44s mis=%2d grid=%5d  %.3f ms  %.0f GB/s\n", name, mis, grid, best, n * 4.0 / best / 1e6);
  };
  for (int grid : {148 * 4, 148 * 8, 148 * 32})
  {
    run("peer  4 B/thread", store_kernel<1>, dst_peer, 0, grid);
    run("peer  4 B/thread", store_kernel<1>, dst_peer, 1, grid);
    run("peer  4 B/thread", store_kernel<1>, dst_peer, 8, grid);
    run("peer  4 B/thread", store_kernel<1>, dst_peer, 17, grid);
    run("peer 16 B/thread", store_kernel<4>, dst_peer, 0, grid);
    run("peer 16 B/thread", store_kernel<4>, dst_peer, 4, grid);
    run("peer 16 B/thread", store_kernel<4>, dst_peer, 16, grid);
  }
  run("local 4 B/thread", store_kernel<1>, dst_local, 0, 148 * 8);
  run("local 4 B/thread", store_kernel<1>, dst_local, 1, 148 * 8);
  CK(cudaMemcpyPeerAsync(dst_peer, 1, src, 0, n * 4));
  CK(cudaEventRecord(e0));
  CK(cudaMemcpyPeerAsync(dst_peer, 1, src, 0, n * 4));
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("copy engine peer copy: %.3f ms %.0f GB/s\n", ms, n * 4.0 / ms / 1e6);
  return 0;
}
