"""Generates tests/golden/*.npz from the UNMODIFIED reference: thrust::sort / sort_by_key compiled from
/root/reference (oracle/_ref, see oracle/Makefile).  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Each fixture stores the seeded input, the reference's output and the call parameters.  Only cases the reference's
CPU path defines bit-exactly are generated from it (full bit range; integer keys, and float keys without +-0 ties:
SURVEY.md 8c).  The fixtures are small (<= 4096 items) so they can live in git.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from gen import make_keys, make_values  # noqa: E402
from oracle_lib import ref_thrust_sort  # noqa: E402

CASES = [
    ("u32_uniform_keys", np.uint32, "uniform", 4096, None, False),
    ("u32_entropy5_pairs", np.uint32, "entropy5", 4096, np.uint32, False),
    ("u64_uniform_pairs_u32", np.uint64, "uniform", 3000, np.uint32, False),
    ("u64_entropy5_pairs_u32", np.uint64, "entropy5", 3000, np.uint32, False),
    ("i64_uniform_desc", np.int64, "uniform", 2500, None, True),
    ("i32_few16_pairs_desc", np.int32, "few16", 4096, np.uint32, True),
    ("f32_uniform_desc_pairs", np.float32, "uniform", 4096, np.uint32, True),
    ("f64_uniform_pairs_u64", np.float64, "uniform", 2048, np.uint64, False),
    ("u8_uniform_pairs_u8", np.uint8, "uniform", 1000, np.uint8, False),
    ("i16_uniform_pairs_u16", np.int16, "uniform", 3000, np.uint16, True),
    ("u32_equal_pairs", np.uint32, "equal", 1500, np.uint32, False),
]


def main():
    for name, kdt, dist, n, vdt, desc in CASES:
        k = make_keys(dist, n, kdt, seed=1234)
        if np.dtype(kdt).kind == "f":
            k = k[~np.isnan(k)]
            k = k[k != 0]
        out = {"keys_in": k, "descending": np.bool_(desc), "begin_bit": np.int32(0),
               "end_bit": np.int32(np.dtype(kdt).itemsize * 8)}
        if vdt is not None:
            v = make_values(k.size, vdt)
            rk, rv, _ = ref_thrust_sort(k, v, descending=desc, backend="cpp")
            rk2, rv2, _ = ref_thrust_sort(k, v, descending=desc, backend="omp")
            assert np.array_equal(rk.view(np.uint8), rk2.view(np.uint8)) and np.array_equal(rv, rv2)
            out.update(vals_in=v, keys_out=rk, vals_out=rv)
        else:
            rk, _ = ref_thrust_sort(k, descending=desc, backend="cpp")
            out.update(keys_out=rk)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("wrote", name, k.size)


if __name__ == "__main__":
    main()
