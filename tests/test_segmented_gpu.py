"""GPU parity tests of the segmented sort (b200rs_segmented_sort) through the C ABI.

Oracle: tests/oracle_lib.py:oracle_segmented_sort, pinned by outputs of the real cub::DeviceSegmentedRadixSort
(tests/golden/segmented/cubseg_*.npz).  Case list follows cub/test/catch2_test_device_segmented_radix_sort_keys.cu /
..._pairs.cu: empty segments, gaps, one-item segments, segments around and far above one tile, descending, bit windows,
+-0.0, every key width, offsets of both widths."""
import ctypes
import glob
import os

import numpy as np
import pytest
import torch

from cccl_b200 import _native
from gen import V16, make_keys, make_values
from gpu_util import assert_same_bits, to_dev, to_host
from oracle_lib import key_kind_of, oracle_segmented_sort

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def gpu_segmented_sort(keys, values, begins, ends, *, descending=False, begin_bit=0, end_bit=None, offset_dtype=np.int64):
    lib = _native.lib()
    n, kdt = keys.shape[0], keys.dtype
    kb = kdt.itemsize
    vb = values.dtype.itemsize if values is not None else 0
    end_bit = kb * 8 if end_bit is None else end_bit
    d_k, d_v = to_dev(keys), to_dev(values)
    # the output starts as a copy of the input: items outside every segment must come back untouched
    d_ko = d_k.clone()
    d_vo = d_v.clone() if values is not None else None
    d_b, d_e = to_dev(np.asarray(begins, dtype=offset_dtype)), to_dev(np.asarray(ends, dtype=offset_dtype))
    p = lambda t: t.data_ptr() if t is not None and t.numel() else None
    args = (n, len(begins), p(d_b), p(d_e), np.dtype(offset_dtype).itemsize, key_kind_of(kdt), kb, vb, begin_bit, end_bit,
            int(bool(descending)), torch.cuda.current_stream().cuda_stream)
    nbytes = ctypes.c_size_t(0)
    _native.check(lib.b200rs_segmented_sort(None, ctypes.byref(nbytes), None, None, None, None, *args), "size query")
    temp = torch.empty(nbytes.value + 300, dtype=torch.uint8, device="cuda")
    _native.check(lib.b200rs_segmented_sort(temp.data_ptr() + 4, ctypes.byref(nbytes), p(d_k), p(d_ko), p(d_v), p(d_vo),
                                            *args), "b200rs_segmented_sort")
    torch.cuda.synchronize()
    assert np.array_equal(to_host(d_k, kdt, n).view(np.uint8), keys.view(np.uint8)), "input keys clobbered"
    gk = to_host(d_ko, kdt, n)
    gv = to_host(d_vo, values.dtype, n) if values is not None else None
    return gk, gv


@pytest.fixture(params=["warp_per_tiny_segment", "cta_per_segment"])
def tiny_path(request):
    """Segments of at most 256 items are sorted by one warp each by default; "cta_per_segment" switches that off so the same
    cases also run the one-CTA-per-segment kernel on them."""
    lib = _native.lib()
    lib.b200rs_set_segmented_tiny_max(0 if request.param == "cta_per_segment" else 256)
    yield request.param
    lib.b200rs_set_segmented_tiny_max(256)


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


def check(keys, values, begins, ends, **kw):
    okw = {k: v for k, v in kw.items() if k != "offset_dtype"}
    gk, gv = gpu_segmented_sort(keys, values, begins, ends, **kw)
    if values is None:
        ek = oracle_segmented_sort(keys, None, begins, ends, **okw)
        assert_same_bits(gk, ek, f"segmented keys {keys.dtype} {kw}")
    else:
        ek, ev = oracle_segmented_sort(keys, values, begins, ends, **okw)
        assert_same_bits(gk, ek, f"segmented pair keys {keys.dtype} {kw}")
        assert_same_bits(gv, ev, f"segmented pair values {keys.dtype} {kw}")


def test_real_cub_segmented_fixtures():
    files = sorted(glob.glob(os.path.join(HERE, "golden", "segmented", "cubseg_*.npz")))
    assert len(files) >= 5
    for f in files:
        z = np.load(f)
        kw = dict(descending=bool(z["descending"]), begin_bit=int(z["begin_bit"]), end_bit=int(z["end_bit"]))
        vals = z["vals_in"] if "vals_in" in z.files else None
        gk, gv = gpu_segmented_sort(z["keys_in"], vals, z["begin_offsets"], z["end_offsets"], **kw)
        # the reference tool leaves items outside the segments as it found them in ITS output buffer; compare the segments
        for b, e in zip(z["begin_offsets"].tolist(), z["end_offsets"].tolist()):
            assert_same_bits(gk[b:e], z["keys_out"][b:e], f"{os.path.basename(f)} keys [{b},{e})")
            if vals is not None:
                assert_same_bits(gv[b:e], z["vals_out"][b:e], f"{os.path.basename(f)} values [{b},{e})")


@pytest.mark.parametrize("kdtype", [np.uint8, np.int16, np.float16, np.uint32, np.int32, np.float32, np.uint64, np.float64])
@pytest.mark.parametrize("descending", [False, True])
def test_segment_shapes_all_key_types(kdtype, descending, tiny_path):
    lengths = [0, 1, 2, -3, 31, 32, 33, 0, 63, 64, 65, 100, 255, 256, 257, 1279, 1280, 1281, -1, 2560, 2561, 5119, 5120, 5121, 7000,
               -11, 20_011, 3, 129, 200, 7, 96]
    begins, ends, n = layout(lengths)
    k = make_keys("uniform", n, kdtype, seed=77)
    if np.dtype(kdtype).kind == "f":
        k[::5] = -0.0
        k[::7] = 0.0
    check(k, None, begins, ends, descending=descending)
    check(k, make_values(n, np.uint32), begins, ends, descending=descending, offset_dtype=np.int32)


@pytest.mark.parametrize("vdtype", [np.uint8, np.uint16, np.uint64, V16])
def test_value_widths_and_stability(vdtype, tiny_path):
    lengths = [300, 6000, 0, 14_000, -5, 1, 250, 33, 128, 17]
    begins, ends, n = layout(lengths)
    for kdtype in (np.uint32, np.uint64):
        k = make_keys("few16", n, kdtype, seed=5)  # many ties inside every segment: stability is visible in the values
        check(k, make_values(n, vdtype), begins, ends)
        check(k, make_values(n, vdtype), begins, ends, descending=True)


@pytest.mark.parametrize("kdtype", [np.uint32, np.float32, np.int64])
def test_bit_windows_and_float_zeros(kdtype, tiny_path):
    lengths = [100, 257, -5, 1, 999, 6001, 13_000, 256, 31, 77]
    begins, ends, n = layout(lengths)
    bits = np.dtype(kdtype).itemsize * 8
    k = make_keys("entropy3", n, kdtype, seed=9)
    if np.dtype(kdtype).kind == "f":
        k[::3] = -0.0
        k[::4] = 0.0
    v = make_values(n, np.uint32)
    for b, e in ((0, bits), (8, 24), (3, bits - 5), (bits - 1, bits), (7, 7)):
        for desc in (False, True):
            check(k, v, begins, ends, descending=desc, begin_bit=b, end_bit=e)


def test_many_small_segments_and_one_large(tiny_path):
    rng = np.random.default_rng(3)
    lengths = rng.integers(0, 40, size=20_000).tolist() + [1 << 20] + rng.integers(0, 3000, size=300).tolist()
    begins, ends, n = layout(lengths)
    k = make_keys("uniform", n, np.uint32, seed=123)
    check(k, make_values(n, np.uint32), begins, ends)
    k = make_keys("equal", n, np.uint64, seed=1)
    check(k, make_values(n, np.uint32), begins, ends, descending=True)


def test_segments_in_any_order_and_unsorted_offsets():
    """Segments need not be listed in address order (the reference indexes them independently)."""
    begins, ends, n = layout([5000, 70, 9000, 0, 12])
    perm = np.array([2, 0, 4, 3, 1])
    k = make_keys("uniform", n, np.int32, seed=4)
    check(k, None, begins[perm], ends[perm])


@pytest.mark.parametrize("kdtype,vdtype", [(np.uint32, None), (np.uint32, np.uint32), (np.float32, np.uint32), (np.uint64, None),
                                           (np.int64, np.uint32), (np.float64, None), (np.uint32, np.uint64), (np.uint64, np.uint64), (np.uint32, np.uint16)])
def test_long_segments_take_whole_grid_passes(kdtype, vdtype):
    """Segments longer than 2^16 items are sorted by whole-grid onesweep passes over all long segments at once
    (csrc/segmented_long.cu: tiles never straddle segments, per-segment bins and chained-scan rows); everything shorter
    by one CTA per segment in the same call.  Long segments at both ends, next to each other, listed out of order, around
    the threshold and around multiples of the tile; +-0.0; both directions; a bit window; (u32, u16) pairs take the
    one-CTA path for every length (no whole-grid kernel for that width)."""
    lengths = [70_000, 65_536, 65_537, -3, 100, 0, 200_003, 44 * 256 * 6, 1, 5000, 131_072 + 11_264, -17, 90_001]
    begins, ends, n = layout(lengths)
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(begins))
    begins, ends = begins[perm], ends[perm]
    for dist in ("uniform", "entropy3"):
        k = make_keys(dist, n, kdtype, seed=31)
        if np.dtype(kdtype).kind == "f":
            k[::5] = -0.0
            k[::7] = 0.0
        v = make_values(n, vdtype) if vdtype is not None else None
        bits = np.dtype(kdtype).itemsize * 8
        check(k, v, begins, ends)
        check(k, v, begins, ends, descending=True, offset_dtype=np.int32)
        check(k, v, begins, ends, descending=(dist == "uniform"), begin_bit=5, end_bit=bits - 9)
        check(k, v, begins, ends, begin_bit=bits - 3, end_bit=bits)


def test_long_segments_all_equal_and_single_huge():
    begins, ends, n = layout([(1 << 21) + 77])
    k = make_keys("equal", n, np.uint32, seed=1)
    check(k, make_values(n, np.uint32), begins, ends)
    k = make_keys("few16", n, np.uint64, seed=2)
    check(k, make_values(n, np.uint32), begins, ends, descending=True)
    k = make_keys("uniform", n, np.uint32, seed=3)
    check(k, None, begins, ends)


def test_segmented_sort_under_stream_capture_with_long_segments():
    """The two halves of a call (whole-grid passes for long segments on a forked side stream, one CTA per short segment on
    the caller's stream) are joined on the caller's stream: the call is one capturable unit of work."""
    lib = _native.lib()
    begins, ends, n = layout([100_000, 300, 0, 70_000, 4000, 9])
    k = make_keys("uniform", n, np.uint32, seed=8)
    v = make_values(n, np.uint32)
    check(k, v, begins, ends)  # warm: the side stream exists before the capture below
    d_k, d_v = to_dev(k), to_dev(v)
    d_ko, d_vo = d_k.clone(), d_v.clone()
    d_b, d_e = to_dev(begins), to_dev(ends)
    nbytes = ctypes.c_size_t(0)
    tail = (n, len(begins), d_b.data_ptr(), d_e.data_ptr(), 8, 0, 4, 4, 0, 32, 0)
    _native.check(lib.b200rs_segmented_sort(None, ctypes.byref(nbytes), None, None, None, None, *tail, 0), "query")
    temp = torch.empty(nbytes.value + 256, dtype=torch.uint8, device="cuda")
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        st = torch.cuda.current_stream().cuda_stream
        _native.check(lib.b200rs_segmented_sort(temp.data_ptr(), ctypes.byref(nbytes), d_k.data_ptr(), d_ko.data_ptr(),
                                                d_v.data_ptr(), d_vo.data_ptr(), *tail, st), "segmented sort under capture")
    ek, ev = oracle_segmented_sort(k, v, begins, ends)
    for _ in range(2):
        d_ko.zero_()
        d_vo.zero_()
        g.replay()
        torch.cuda.synchronize()
        for b, e in zip(begins.tolist(), ends.tolist()):
            assert_same_bits(to_host(d_ko, np.uint32, n)[b:e], ek[b:e], "captured segmented keys")
            assert_same_bits(to_host(d_vo, np.uint32, n)[b:e], ev[b:e], "captured segmented values")


@pytest.mark.parametrize("seed", range(6))
def test_random_layouts_across_all_size_classes(seed):
    """Randomised differential test: segment lengths drawn from a mixture that hits every size class and every class
    boundary (0 / 1 / 32 / 256 / 257 / one tile +- 1 / several tiles / far above the whole-grid threshold), random gaps, random
    listing order, random key / value types, direction and bit window -- against the oracle."""
    rng = np.random.default_rng(1000 + seed)
    kdtype = [np.uint32, np.float32, np.int64, np.uint64, np.int32, np.float64][seed % 6]
    vdtype = [None, np.uint32, np.uint64, np.uint32, np.uint8, None][seed % 6]
    edges = [0, 1, 2, 31, 32, 33, 255, 256, 257, 1279, 1280, 1281, 2559, 2560, 2561, 5119, 5120, 5121, 8191, 8192, 8193, 11_263,
             11_264, 11_265]
    lengths = []
    for _ in range(160):
        r = rng.random()
        if r < 0.35:
            ln = int(rng.choice(edges))
        elif r < 0.65:
            ln = int(rng.integers(0, 300))
        elif r < 0.85:
            ln = int(rng.integers(300, 9000))
        elif r < 0.97:
            ln = int(rng.integers(9000, 60_000))
        else:
            ln = int(rng.integers(60_000, 300_000))
        lengths.append(ln)
        if rng.random() < 0.15:
            lengths.append(-int(rng.integers(1, 40)))
    begins, ends, n = layout(lengths)
    perm = rng.permutation(len(begins))
    begins, ends = begins[perm], ends[perm]
    k = make_keys(["uniform", "entropy3", "few16"][seed % 3], n, kdtype, seed=seed)
    if np.dtype(kdtype).kind == "f":
        k[::11] = -0.0
        k[::13] = 0.0
    v = make_values(n, vdtype) if vdtype is not None else None
    bits = np.dtype(kdtype).itemsize * 8
    check(k, v, begins, ends, descending=bool(seed & 1), offset_dtype=[np.int64, np.int32][seed % 2])
    b = int(rng.integers(0, bits - 1))
    e = int(rng.integers(b + 1, bits + 1))
    check(k, v, begins, ends, descending=not bool(seed & 1), begin_bit=b, end_bit=e)


def test_python_mirror_of_cuda_compute_segmented_sort():
    """cccl_b200.segmented_sort / make_segmented_sort with the reference's keyword interface and its own documentation
    examples (python/cuda_cccl/tests/compute/examples/sort/segmented_sort_basic.py, segmented_sort_buffer.py,
    segmented_sort_object.py): pointer form, DoubleBuffer form (result in .current()), reusable object with a size query."""
    import cccl_b200
    from cccl_b200 import DoubleBuffer, SortOrder

    h_keys = np.array([9, 1, 5, 4, 2, 8, 7, 3, 6], dtype=np.int32)
    h_vals = np.array([90, 10, 50, 40, 20, 80, 70, 30, 60], dtype=np.int32)
    starts, ends = np.array([0, 3, 5], dtype=np.int64), np.array([3, 5, 9], dtype=np.int64)
    d_k, d_v = torch.from_numpy(h_keys).cuda(), torch.from_numpy(h_vals).cuda()
    d_s, d_e = torch.from_numpy(starts).cuda(), torch.from_numpy(ends).cuda()

    def expected(reverse):
        ek, ev = [], []
        for s, e in zip(starts, ends):
            pairs = sorted(zip(h_keys[s:e], h_vals[s:e]), key=lambda kv: kv[0], reverse=reverse)
            ek += [k for k, _ in pairs]
            ev += [v for _, v in pairs]
        return np.array(ek, dtype=np.int32), np.array(ev, dtype=np.int32)

    d_ko, d_vo = torch.empty_like(d_k), torch.empty_like(d_v)
    cccl_b200.segmented_sort(d_in_keys=d_k, d_out_keys=d_ko, d_in_values=d_v, d_out_values=d_vo, num_items=9, num_segments=3,
                             start_offsets_in=d_s, end_offsets_in=d_e, order=SortOrder.ASCENDING)
    torch.cuda.synchronize()
    ek, ev = expected(False)
    assert np.array_equal(d_ko.cpu().numpy(), ek) and np.array_equal(d_vo.cpu().numpy(), ev)
    assert np.array_equal(d_k.cpu().numpy(), h_keys), "pointer form must not write its input"

    kdb = DoubleBuffer(d_k.clone(), torch.empty_like(d_k))
    vdb = DoubleBuffer(d_v.clone(), torch.empty_like(d_v))
    cccl_b200.segmented_sort(d_in_keys=kdb, d_out_keys=None, d_in_values=vdb, d_out_values=None, num_items=9, num_segments=3,
                             start_offsets_in=d_s, end_offsets_in=d_e, order=SortOrder.DESCENDING)
    torch.cuda.synchronize()
    ek, ev = expected(True)
    assert np.array_equal(kdb.current().cpu().numpy(), ek) and np.array_equal(vdb.current().cpu().numpy(), ev)

    # reusable object, keys only, int32 offsets, a larger input that reaches every size class
    begins, ends2, n = layout([70_000, 3, 300, 0, 9000])
    k = make_keys("uniform", n, np.float32, seed=2)
    dk, dko = torch.from_numpy(k).cuda(), torch.empty(n, dtype=torch.float32, device="cuda")
    db = torch.from_numpy(begins.astype(np.int32)).cuda()
    de = torch.from_numpy(ends2.astype(np.int32)).cuda()
    sorter = cccl_b200.make_segmented_sort(d_in_keys=dk, d_out_keys=dko, start_offsets_in=db, end_offsets_in=de,
                                           order=SortOrder.ASCENDING)
    kw = dict(d_in_keys=dk, d_out_keys=dko, d_in_values=None, d_out_values=None, num_items=n, num_segments=len(begins),
              start_offsets_in=db, end_offsets_in=de)
    need = sorter(temp_storage=None, **kw)
    assert need >= 1
    temp = torch.empty(need, dtype=torch.uint8, device="cuda")
    sorter(temp_storage=temp, **kw)
    torch.cuda.synchronize()
    want = oracle_segmented_sort(k, None, begins, ends2)
    got = dko.cpu().numpy()
    for b, e in zip(begins.tolist(), ends2.tolist()):
        assert_same_bits(got[b:e], want[b:e], "python mirror, reusable object")
