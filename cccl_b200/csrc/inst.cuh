// Instantiation helper: each inst_k*.cu defines B200RS_KEY_BYTES and includes this file to emit the
// onesweep kernels for one key width (all value widths), so the four files compile in parallel.
#pragma once

#include "configs.h"
#include "onesweep.cuh"

namespace b200rs
{

template <class U, int VB, int NT, int IPT, int RANK, int MINB = 1>
cudaError_t launch_onesweep(const PassArgs& args, unsigned grid, cudaStream_t stream)
{
  using L     = OnesweepSmem<U, VB, NT, IPT>;
  auto kernel = onesweep_kernel<U, VB, NT, IPT, RANK, MINB>;
  if (L::BYTES > 48 * 1024)
  {
    // per-device attribute; cheap and idempotent, legal during stream capture
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(L::BYTES));
    if (e != cudaSuccess)
    {
      return e;
    }
  }
  kernel<<<grid, NT, L::BYTES, stream>>>(args);
  return cudaPeekAtLastError();
}

template <class U, int VB, int NT, int IPT, int RANK, int MINB = 1>
constexpr OnesweepConfig make_config()
{
  return OnesweepConfig{NT, IPT, RANK, MINB, NT * IPT, OnesweepSmem<U, VB, NT, IPT>::BYTES,
                        &launch_onesweep<U, VB, NT, IPT, RANK, MINB>};
}

} // namespace b200rs
