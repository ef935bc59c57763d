"""GPU parity against the reference ITSELF: cub::DeviceRadixSort 3.6.0 compiled from the unmodified reference
headers into oracle/_ref/ref_cub_radix_sort (recipe: oracle/Makefile), run on the same B200 on the same seeded
inputs; outputs must be bit-identical (keys and values).  Skipped when the binary was never built."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from gen import make_keys, make_values
from gpu_util import assert_same_bits, gpu_sort

pytestmark = pytest.mark.gpu

BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_cub_radix_sort")
KT = {np.dtype(np.uint8): "u8", np.dtype(np.int8): "i8", np.dtype(np.uint16): "u16", np.dtype(np.int16): "i16",
      np.dtype(np.uint32): "u32", np.dtype(np.int32): "i32", np.dtype(np.float32): "f32",
      np.dtype(np.uint64): "u64", np.dtype(np.int64): "i64", np.dtype(np.float64): "f64", np.dtype(np.float16): "f16"}


def ref_cub_sort(keys, values, descending, begin_bit, end_bit, ktype=None):
    with tempfile.TemporaryDirectory() as d:
        kf, vf, ko, vo = (os.path.join(d, x) for x in ("k.bin", "v.bin", "ko.bin", "vo.bin"))
        keys.tofile(kf)
        if values is not None:
            values.tofile(vf)
        cmd = [BIN, "sort", ktype or KT[keys.dtype], str(values.dtype.itemsize if values is not None else 0), str(keys.size),
               str(int(descending)), str(begin_bit), str(end_bit), kf, vf if values is not None else "-", ko,
               vo if values is not None else "-"]
        subprocess.run(cmd, check=True, timeout=300)
        rk = np.fromfile(ko, dtype=keys.dtype)
        rv = np.fromfile(vo, dtype=values.dtype) if values is not None else None
    return rk, rv


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/ref_cub_radix_sort not built")
@pytest.mark.parametrize("kdtype,vdtype,dist,n,desc,window", [
    (np.uint32, None, "uniform", 1 << 22, False, None),
    (np.uint32, np.uint32, "entropy5", 1 << 21, False, None),
    (np.uint64, np.uint32, "uniform", 1 << 21, False, None),
    (np.uint64, np.uint32, "entropy5", 1 << 21, True, None),
    (np.float32, np.uint32, "uniform", 1 << 20, True, (8, 24)),
    (np.int64, np.uint64, "few256", 1 << 20, True, (16, 48)),
    (np.float64, np.uint32, "uniform", 300_001, False, None),
    (np.int16, np.uint32, "uniform", 200_003, True, (3, 13)),
    (np.uint8, np.uint32, "uniform", 100_000, False, None),
    (np.uint32, np.uint32, "equal", 1 << 20, False, None),
])
def test_bit_identical_to_reference_cub(kdtype, vdtype, dist, n, desc, window):
    k = make_keys(dist, n, kdtype, seed=99)
    if np.dtype(kdtype).kind == "f":
        k[::101] = -0.0
        k[::103] = 0.0
    v = make_values(n, vdtype) if vdtype is not None else None
    b, e = window if window else (0, np.dtype(kdtype).itemsize * 8)
    rk, rv = ref_cub_sort(k, v, desc, b, e)
    if v is None:
        gk, _ = gpu_sort(k, descending=desc, begin_bit=b, end_bit=e)
    else:
        gk, gv, _ = gpu_sort(k, v, descending=desc, begin_bit=b, end_bit=e)
        assert_same_bits(gv, rv, "values vs reference CUB")
    assert_same_bits(gk, rk, "keys vs reference CUB")


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/ref_cub_radix_sort not built")
@pytest.mark.parametrize("ktype", ["f16", "bf16"])
@pytest.mark.parametrize("n,desc,window", [(300_001, False, None), (4000, True, (3, 14)), (1 << 20, True, None)])
def test_16_bit_float_keys_bit_identical_to_reference_cub(ktype, n, desc, window):
    """__half / __nv_bfloat16 keys (reference: device_radix_sort.cuh:51-57, util_type.cuh:1017-1095): the same
    sign-magnitude transform on 16-bit patterns, every bit pattern included (NaNs, infinities, both zeros)."""
    bits = make_keys("uniform", n, np.uint16, seed=5)
    bits[::97] = 0x8000  # -0.0
    bits[::89] = 0x0000  # +0.0
    v = make_values(n, np.uint32)
    b, e = window if window else (0, 16)
    rk, rv = ref_cub_sort(bits, v, desc, b, e, ktype=ktype)
    gk, gv, _ = gpu_sort(bits, v, descending=desc, begin_bit=b, end_bit=e, kind=2)
    assert_same_bits(gk, rk, f"{ktype} keys vs cub n={n}")
    assert_same_bits(gv, rv, f"{ktype} values vs cub n={n}")
