"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/b200rs.h declares,
and its argument checking / size query (which touch no GPU) behave like the reference's."""
import ctypes
import os
import re

import pytest

from cccl_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    """Every entry point the headers under include/ declare: b200rs.h and the cccl.c.parallel names of b200rs_cccl_c.h."""
    text = open(os.path.join(ROOT, "include", "b200rs.h")).read()
    names = set(re.findall(r"B200RS_API\s+\w[\w\s\*]*?\b(b200rs_\w+)\s*\(", text))
    text = open(os.path.join(ROOT, "include", "b200rs_cccl_c.h")).read()
    names |= set(re.findall(r"CCCL_C_API\s+\w+\s+(cccl_\w+)\s*\(", text))
    return sorted(names)


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_native.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.b200rs_version() == 100


def _query(n, kb, vb, begin, end, overwrite, kind=0):
    return _native.sort_raw(0, 0, 0, 0, 0, 0, n, kind, kb, vb, begin, end, False, overwrite)


def test_size_query_is_pure_and_deterministic():
    # dispatch_radix_sort.cuh:1739-1743 -- no device work on the query; runs here without a GPU
    a, sel = _query(1 << 20, 4, 0, 0, 32, False)
    b, _ = _query(1 << 20, 4, 0, 0, 32, False)
    assert a == b and sel == -1
    c, _ = _query(1 << 20, 4, 0, 0, 32, True)
    assert a - c >= (1 << 20) * 4  # pointer API carries an extra key buffer (device_radix_sort.cuh:310-315)
    d, _ = _query(1 << 20, 4, 4, 0, 32, False)
    assert d - a >= (1 << 20) * 4 - (64 << 10)  # the tile (hence look-back array) size differs per type pair
    one_pass, _ = _query(1 << 20, 4, 0, 0, 8, False)
    assert one_pass < (1 << 20) * 2  # single pass goes straight from in to out: no 4 MiB key buffer, only look-back words


def test_empty_problem_needs_one_byte():
    # dispatch_radix_sort.cuh:1956-1962
    assert _query(0, 4, 0, 0, 32, False)[0] == 1
    assert _query(1000, 4, 0, 8, 8, True)[0] == 1
    assert _query(1000, 4, 0, 8, 8, False)[0] == 1


@pytest.mark.parametrize("args,code", [
    ((10, 3, 0, 0, 24, False), 801),    # unsupported key width
    ((10, 4, 3, 0, 32, False), 801),    # unsupported value width
    ((10, 4, 0, 8, 40, False), 1),      # end_bit beyond the key
    ((10, 4, 0, 9, 8, False), 1),       # begin > end
    ((10, 1, 0, 0, 8, False, 2), 801),  # 8-bit float keys (16-bit half / bfloat16 keys are supported)
])
def test_bad_arguments(args, code):
    with pytest.raises(_native.B200RSError) as e:
        _query(*args)
    assert e.value.code == code


def test_too_small_temp_is_invalid_value():
    # util_temporary_storage.cuh:75-78.  The size check comes before any device access, so fake pointers are fine.
    need, _ = _query(100000, 4, 0, 0, 32, True)
    with pytest.raises(_native.B200RSError) as e:
        _native.sort_raw(0x1000, need - 1, 0x10000, 0x20000, 0, 0, 100000, 0, 4, 0, 0, 32, False, True)
    assert e.value.code == 1


def test_config_tables_described():
    for kb in (1, 2, 4, 8):
        for vb in (0, 1, 2, 4, 8, 16):
            d = _native.describe_configs(kb, vb)
            assert len(d) >= 1 and f"k{kb}v{vb}" in d[0]


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing shipped may import, load or link it."""
    bad = []
    for base in ("cccl_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if os.sep + "build" in dirpath:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"liboracle|oracle_lib|oracle/|ref_thrust", txt):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


# ---- the adjacent callers: size queries and argument checks run without a GPU too
def _seg_query(n, segs, kb=4, vb=4, begin=0, end=None, offset_bytes=8):
    lib = _native.lib()
    need = ctypes.c_size_t(0)
    rc = lib.b200rs_segmented_sort(None, ctypes.byref(need), None, None, None, None, n, segs, None, None, offset_bytes, 0, kb, vb,
                                   begin, kb * 8 if end is None else end, 0, None)
    return rc, need.value


def test_segmented_size_query_and_argument_checks():
    rc, small = _seg_query(1000, 10)
    assert rc == 0 and small == 1  # at most one tile in total: every segment is sorted in shared memory, no scratch
    rc, big = _seg_query(1 << 24, 1000)
    assert rc == 0
    # scratch copies of keys and values (what the reference needs) + the worst-case plan of the whole-grid path
    assert (1 << 24) * 8 <= big <= int((1 << 24) * 8 * 1.4)
    assert _seg_query(1 << 24, 1000) == (0, big)  # pure
    rc1, one_pass = _seg_query(1 << 24, 1000, begin=3, end=9)
    assert rc1 == 0 and one_pass < big  # a single pass needs no scratch copies
    lib = _native.lib()
    lib.b200rs_set_segmented_long_min(0)
    try:
        rc, off = _seg_query(1 << 24, 1000)
        assert rc == 0 and off < big and off >= (1 << 24) * 8  # without the whole-grid path: the scratch copies only
    finally:
        lib.b200rs_set_segmented_long_min(1)
    assert _seg_query(10, 1, offset_bytes=2)[0] == 1  # cudaErrorInvalidValue
    assert _seg_query(10, 1, kb=3)[0] == 801  # cudaErrorNotSupported
    assert _seg_query(10, 1, begin=5, end=3)[0] == 1


def test_topk_size_query_and_argument_checks():
    lib = _native.lib()
    need = ctypes.c_size_t(0)

    def q(n, k, kb=4, vb=0, kind=0):
        rc = lib.b200rs_topk(None, ctypes.byref(need), None, None, None, None, n, k, kind, kb, vb, 1, None)
        return rc, need.value

    assert q(0, 5) == (0, 1) and q(100, 0) == (0, 1)
    rc, a = q(1 << 20, 10)
    assert rc == 0 and a > 1
    assert q(1 << 20, 1 << 19) == (0, a)  # K does not change the temporary storage
    rc, b = q(1 << 24, 10)
    assert rc == 0 and a < b <= (1 << 24) * 4  # the candidate buffer: a fraction of the keys, never a copy of them
    assert q(10, 1, kb=3)[0] == 801 and q(10, 1, vb=12)[0] == 801 and q(10, 1, kind=7)[0] == 1
    assert lib.b200rs_topk(None, None, None, None, None, None, 10, 1, 0, 4, 0, 1, None) == 1
