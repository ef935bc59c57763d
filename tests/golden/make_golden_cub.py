"""Generates tests/golden/cub_*.npz from the UNMODIFIED reference's GPU path: cub::DeviceRadixSort 3.6.0 compiled from
/root/reference into oracle/_ref/ref_cub_radix_sort (oracle/Makefile).  Must run on a GPU box:

    gpurun -- python tests/golden/make_golden_cub.py gpurun_out/golden_cub     # then copy the .npz files here

These fixtures pin the corners where the reference's CPU paths say nothing or disagree with its device code:
bit windows, descending order, and +-0.0 / NaN float keys.  Small (<= 4096 items) so they live in git.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from gen import make_keys, make_values  # noqa: E402
from test_vs_reference_gpu import ref_cub_sort  # noqa: E402

CASES = [
    # name, key dtype, dist, n, value dtype, descending, (begin_bit, end_bit) or None, sprinkle zeros/nans
    ("cub_f32_desc_bits8_24_zeros", np.float32, "uniform", 4096, np.uint32, True, (8, 24), True),
    ("cub_f32_asc_bits8_24_zeros", np.float32, "uniform", 4096, np.uint32, False, (8, 24), True),
    ("cub_f32_desc_full_zeros_nans", np.float32, "uniform", 4096, np.uint32, True, None, True),
    ("cub_f64_desc_bits5_63_zeros", np.float64, "uniform", 2048, np.uint32, True, (5, 63), True),
    ("cub_f64_asc_full_zeros_nans", np.float64, "uniform", 2048, np.uint64, False, None, True),
    ("cub_i64_desc_bits16_48", np.int64, "uniform", 3000, np.uint32, True, (16, 48), False),
    ("cub_u64_u32_entropy5", np.uint64, "entropy5", 4096, np.uint32, False, None, False),
    ("cub_u32_bits3_13_desc", np.uint32, "uniform", 4096, np.uint32, True, (3, 13), False),
    ("cub_i16_bits0_9", np.int16, "uniform", 3000, np.uint32, False, (0, 9), False),
    ("cub_u8_desc", np.uint8, "uniform", 1000, np.uint32, True, None, False),
    ("cub_u32_keys_only", np.uint32, "uniform", 4096, None, False, None, False),
    # above the reference's single-tile size (4864 / 2304 items): the onesweep kernels, whose float-zero rule for
    # descending partial-window sorts differs from the single-tile kernel's (oracle/radix_sort_oracle.cpp)
    ("cub_onesweep_f32_desc_bits8_24_zeros", np.float32, "uniform", 6000, np.uint32, True, (8, 24), True),
    ("cub_onesweep_f32_asc_bits8_24_zeros", np.float32, "uniform", 6000, np.uint32, False, (8, 24), True),
    ("cub_onesweep_f32_desc_bits0_9_zeros_keys", np.float32, "uniform", 5000, None, True, (0, 9), True),
    ("cub_onesweep_f64_desc_bits5_63_zeros", np.float64, "uniform", 3000, np.uint32, True, (5, 63), True),
    ("cub_onesweep_f64_desc_bits40_64_zeros_v64", np.float64, "uniform", 2400, np.uint64, True, (40, 64), True),
    ("cub_single_f64_desc_bits40_64_zeros_v64", np.float64, "uniform", 2304, np.uint64, True, (40, 64), True),
    ("cub_single_f32_desc_bits8_24_zeros_4864", np.float32, "uniform", 4864, np.uint32, True, (8, 24), True),
    ("cub_onesweep_f32_desc_bits8_24_zeros_4865", np.float32, "uniform", 4865, np.uint32, True, (8, 24), True),
    ("cub_onesweep_i32_desc_bits5_20", np.int32, "entropy3", 7001, np.uint32, True, (5, 20), False),
]


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    for name, kdt, dist, n, vdt, desc, window, zeros in CASES:
        k = make_keys(dist, n, kdt, seed=4321)
        if zeros:
            k[::7] = -0.0
            k[::11] = 0.0
            k[3::97] = np.nan
            k[5::89] = -np.nan
        b, e = window if window else (0, np.dtype(kdt).itemsize * 8)
        v = make_values(k.size, vdt) if vdt is not None else None
        rk, rv = ref_cub_sort(k, v, desc, b, e)
        out = {"keys_in": k, "keys_out": rk, "descending": np.bool_(desc), "begin_bit": np.int32(b),
               "end_bit": np.int32(e)}
        if v is not None:
            out.update(vals_in=v, vals_out=rv)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **out)
        print("wrote", name, k.size)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else HERE)
