// Instantiation helper: each inst_k*.cu defines B200RS_KEY_BYTES and includes this file to emit the
// onesweep kernels for one key width (all value widths), so the four files compile in parallel.
#pragma once

#include "configs.h"
#include "onesweep.cuh"
#include "onesweep_tma.cuh"

namespace b200rs
{

template <class U, int VB, int NT, int IPT, int RANK, int MINB = 1, int OPT = 0>
cudaError_t launch_onesweep(const PassArgs& args, unsigned grid, cudaStream_t stream)
{
  using L = OnesweepSmem<U, VB, NT, IPT, OPT>;
  // the -0.0 == +0.0 digit rule is compiled in only where it can matter: floating-point keys (4/8 bytes);
  // 64-bit output offsets only for arrays of 2^32 items and more
  constexpr bool CAN_FLOAT = sizeof(U) >= 2; // half / bfloat16, float, double
  const bool flt           = CAN_FLOAT && args.xf.float_mask != 0;
  auto kernel              = onesweep_kernel<U, VB, NT, IPT, RANK, MINB, OPT, false, false>;
  if (OPT & OPT_BUCKET)
  {
    // one launch over < 2^30 items: no 64-bit-offset variant
    if (flt)
    {
      kernel = onesweep_kernel<U, VB, NT, IPT, RANK, MINB, OPT, CAN_FLOAT, false>;
    }
  }
  else if (args.big)
  {
    kernel = flt ? onesweep_kernel<U, VB, NT, IPT, RANK, MINB, OPT, CAN_FLOAT, true>
                 : onesweep_kernel<U, VB, NT, IPT, RANK, MINB, OPT, false, true>;
  }
  else if (flt)
  {
    kernel = onesweep_kernel<U, VB, NT, IPT, RANK, MINB, OPT, CAN_FLOAT, false>;
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

template <class U, int VB, int NT, int IPT, int RANK, int MINB = 1, int OPT = 0>
constexpr OnesweepConfig make_config()
{
  return OnesweepConfig{NT, IPT, RANK, MINB, NT * IPT, OnesweepSmem<U, VB, NT, IPT, OPT>::BYTES,
                        &launch_onesweep<U, VB, NT, IPT, RANK, MINB, OPT>, 0, OPT, nullptr};
}

// the same, plus its bucket-mode twin (multi-GPU partition pass)
template <class U, int VB, int NT, int IPT, int RANK, int MINB = 1, int OPT = 0>
constexpr OnesweepConfig make_config_with_bucket()
{
  return OnesweepConfig{NT, IPT, RANK, MINB, NT * IPT, OnesweepSmem<U, VB, NT, IPT, OPT>::BYTES,
                        &launch_onesweep<U, VB, NT, IPT, RANK, MINB, OPT>, 0, OPT,
                        &launch_onesweep<U, VB, NT, IPT, RANK, MINB, OPT | OPT_BUCKET>};
}

// TMA bulk-store variant (onesweep_tma.cuh); LBW = look-back window
template <class U, int VB, int NT, int IPT, int MINB, int LBW>
cudaError_t launch_onesweep_tma(const PassArgs& args, unsigned grid, cudaStream_t stream)
{
  using L = TmaSmem<U, VB, NT, IPT>;
  constexpr bool CAN_FLOAT = sizeof(U) >= 2; // half / bfloat16, float, double
  const bool flt           = CAN_FLOAT && args.xf.float_mask != 0;
  auto kernel              = onesweep_tma_kernel<U, VB, NT, IPT, MINB, LBW, false>;
  if (flt)
  {
    kernel = onesweep_tma_kernel<U, VB, NT, IPT, MINB, LBW, CAN_FLOAT>;
  }
  if (L::BYTES > 48 * 1024)
  {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(L::BYTES));
    if (e != cudaSuccess)
    {
      return e;
    }
  }
  kernel<<<grid, NT, L::BYTES, stream>>>(args);
  return cudaPeekAtLastError();
}

template <class U, int VB, int NT, int IPT, int MINB, int LBW = 4>
constexpr OnesweepConfig make_tma_config()
{
  return OnesweepConfig{NT, IPT, RANK_BALLOT, MINB, NT * IPT, TmaSmem<U, VB, NT, IPT>::BYTES,
                        &launch_onesweep_tma<U, VB, NT, IPT, MINB, LBW>, 1, 0, nullptr};
}

} // namespace b200rs
