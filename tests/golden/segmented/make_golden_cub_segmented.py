"""Generates tests/golden/segmented/cubseg_*.npz from the UNMODIFIED reference's cub::DeviceSegmentedRadixSort (3.6.0, compiled
from /root/reference into oracle/_ref/ref_cub_radix_sort, mode `segsort`).  Must run on a GPU box:

    gpurun -- python tests/golden/segmented/make_golden_cub_segmented.py gpurun_out/golden_cubseg   # then copy the .npz files here

They pin tests/oracle_lib.py:oracle_segmented_sort -- the checker for SURVEY.md 8f-1, the next row of the scope table --
including empty segments, gaps between segments, segments of one item, a segment larger than the reference's
single-tile size, descending order, a bit window and +-0.0 keys.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/golden
sys.path.insert(0, os.path.dirname(HERE))
from gen import make_keys, make_values  # noqa: E402

BIN = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "ref_cub_radix_sort")
KT = {np.dtype(np.uint32): "u32", np.dtype(np.float32): "f32", np.dtype(np.uint64): "u64", np.dtype(np.int64): "i64"}

CASES = [
    # name, key dtype, dist, segment lengths (negative = gap of that many items), with values, descending, window
    ("cubseg_u32_contiguous", np.uint32, "uniform", [5, 0, 1, 300, 0, 0, 64, 1000, 2], True, False, None),
    ("cubseg_u32_gaps_desc", np.uint32, "few16", [10, -3, 200, -1, 0, 50, -7, 33], True, True, None),
    ("cubseg_f32_zeros_window", np.float32, "uniform", [100, 257, -5, 1, 999], True, True, (8, 24)),
    ("cubseg_u64_large_segment", np.uint64, "entropy3", [6000, 3, -2, 2500], True, False, None),
    ("cubseg_i64_keys_only", np.int64, "uniform", [0, 17, 400, -9, 1200], False, True, (16, 48)),
    # a FLOAT segment far above one tile, descending, on a bit window, full of +-0.0: pins which -0.0 rule the reference's
    # segmented kernel applies to long segments (it never inverts keys: kernel_segmented_radix_sort.cuh:213-272)
    ("cubseg_f32_large_desc_window", np.float32, "uniform", [9000, 200, -3, 6000], True, True, (8, 24)),
    ("cubseg_f32_large_asc_window", np.float32, "uniform", [7000, 5], True, False, (8, 24)),
]


def layout(lengths):
    begins, ends, pos = [], [], 0
    for ln in lengths:
        if ln < 0:
            pos += -ln
        else:
            begins.append(pos)
            ends.append(pos + ln)
            pos += ln
    return np.array(begins, dtype=np.int64), np.array(ends, dtype=np.int64), pos


def ref_cub_segsort(keys, values, begins, ends, descending, begin_bit, end_bit):
    with tempfile.TemporaryDirectory() as d:
        kf, vf, ko, vo, bf, ef = (os.path.join(d, x) for x in ("k", "v", "ko", "vo", "b", "e"))
        keys.tofile(kf)
        begins.tofile(bf)
        ends.tofile(ef)
        if values is not None:
            values.tofile(vf)
        cmd = [BIN, "segsort", KT[keys.dtype], "4" if values is not None else "0", str(keys.size), str(int(descending)),
               str(begin_bit), str(end_bit), kf, vf if values is not None else "-", ko, vo if values is not None else "-",
               str(begins.size), bf, ef]
        subprocess.run(cmd, check=True, timeout=300)
        rk = np.fromfile(ko, dtype=keys.dtype)
        rv = np.fromfile(vo, dtype=values.dtype) if values is not None else None
    return rk, rv


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    for name, kdt, dist, lengths, with_vals, desc, window in CASES:
        begins, ends, n = layout(lengths)
        k = make_keys(dist, n, kdt, seed=2468)
        if np.dtype(kdt).kind == "f":
            k[::5] = -0.0
            k[::7] = 0.0
        v = make_values(n, np.uint32) if with_vals else None
        b, e = window if window else (0, np.dtype(kdt).itemsize * 8)
        rk, rv = ref_cub_segsort(k, v, begins, ends, desc, b, e)
        out = dict(keys_in=k, keys_out=rk, begin_offsets=begins, end_offsets=ends, descending=desc, begin_bit=b,
                   end_bit=e)
        if with_vals:
            out.update(vals_in=v, vals_out=rv)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **out)
        print("wrote", name, n, "items,", begins.size, "segments")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "segmented"))
