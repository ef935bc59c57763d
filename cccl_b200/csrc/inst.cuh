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
  using L = OnesweepSmem<U, VB, NT, IPT>;
  // the -0.0 == +0.0 digit rule is compiled in only where it can matter: floating-point keys (4/8 bytes);
  // 64-bit output offsets only for arrays of 2^32 items and more
  constexpr bool CAN_FLOAT = sizeof(U) >= 4;
  const bool flt           = CAN_FLOAT && args.xf.float_mask != 0;
  auto kernel              = onesweep_kernel<U, VB, NT, IPT, RANK, MINB, false, false>;
  if (args.big)
  {
    kernel = flt ? onesweep_kernel<U, VB, NT, IPT, RANK, MINB, CAN_FLOAT, true>
                 : onesweep_kernel<U, VB, NT, IPT, RANK, MINB, false, true>;
  }
  else if (flt)
  {
    kernel = onesweep_kernel<U, VB, NT, IPT, RANK, MINB, CAN_FLOAT, false>;
  }
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
