"""b200rs_topk (cub::DeviceTopK::{Max,Min}{Keys,Pairs}, /root/reference/cub/cub/device/device_topk.cuh:297,775,1238)
through the C ABI against the oracle: the reference's contract is "the K best items in any order, any subset of the ties
of the K-th key", so -- as cub/test/catch2_test_device_topk_keys.cu / _pairs.cu do -- the selected keys are sorted and
compared with the first K keys of the oracle's sort of the same input, and every returned value (an input index) must
be unique and still carry its key bit for bit."""
import ctypes

import numpy as np
import pytest
import torch

from cccl_b200 import _native
from gen import make_keys
from gpu_util import to_dev, to_host
from oracle_lib import key_kind_of, oracle_sort, oracle_topk

pytestmark = pytest.mark.gpu


def gpu_topk(keys, k, largest, with_values, vdtype=np.uint32, stream=None):
    n = keys.shape[0]
    kb = keys.dtype.itemsize
    vals = np.arange(n, dtype=vdtype) if with_values else None
    vb = vals.dtype.itemsize if with_values else 0
    kk = min(k, n)
    d_k = to_dev(keys)
    d_ko = torch.zeros(max(kk, 1) * kb, dtype=torch.uint8, device="cuda")
    d_v = to_dev(vals)
    d_vo = torch.zeros(max(kk, 1) * max(vb, 1), dtype=torch.uint8, device="cuda")
    ptr = lambda t: t.data_ptr() if t is not None and t.numel() > 0 else (0x100 if t is not None else 0)
    st = stream.cuda_stream if stream is not None else torch.cuda.current_stream().cuda_stream
    lib = _native.lib()
    need = ctypes.c_size_t(0)
    args = (ptr(d_k), ptr(d_ko), ptr(d_v) if with_values else 0, ptr(d_vo) if with_values else 0, n, k,
            key_kind_of(keys.dtype), kb, vb, int(largest), st)
    _native.check(lib.b200rs_topk(0, ctypes.byref(need), *args), "b200rs_topk (size query)")
    temp = torch.empty(need.value + 256, dtype=torch.uint8, device="cuda")
    _native.check(lib.b200rs_topk(temp.data_ptr(), ctypes.byref(need), *args), "b200rs_topk")
    torch.cuda.synchronize()
    out_k = to_host(d_ko, keys.dtype, kk)
    out_v = to_host(d_vo, vals.dtype, kk) if with_values else None
    return out_k, out_v


def check(keys, k, largest, with_values, **kw):
    n = keys.shape[0]
    kk = min(k, n)
    got_k, got_v = gpu_topk(keys, k, largest, with_values, **kw)
    want = oracle_topk(keys, None, k, largest)
    got_sorted = oracle_sort(got_k, descending=largest)
    # -0.0 and +0.0 are one key (any of the tied items may be returned): compare values, NaNs by bits
    if keys.dtype.kind == "f":
        u = {2: np.uint16, 4: np.uint32, 8: np.uint64}[keys.dtype.itemsize]
        canon = lambda a: np.where(a == 0, np.zeros_like(a), a).view(u)
        assert np.array_equal(canon(got_sorted), canon(want)), (keys.dtype, n, k, largest)
    else:
        assert np.array_equal(got_sorted, want), (keys.dtype, n, k, largest)
    if with_values:
        assert np.unique(got_v).size == kk, "an input item was returned twice"
        assert np.array_equal(keys[got_v.astype(np.int64)].view(np.uint8), got_k.view(np.uint8)), "value lost its key"


KEY_DTYPES = [np.uint8, np.int16, np.float16, np.uint32, np.int32, np.float32, np.uint64, np.int64, np.float64]


@pytest.fixture(params=["one_launch", "general"])
def topk_path(request):
    """Inputs of at most 32 MiB of keys take the one-launch cooperative kernel by default; "general" switches it off so that
    the same cases also run the multi-kernel radix select (candidate compaction included)."""
    lib = _native.lib()
    lib.b200rs_set_topk_small_max(0 if request.param == "general" else 32 << 20)
    yield request.param
    lib.b200rs_set_topk_small_max(32 << 20)


@pytest.mark.parametrize("dtype", KEY_DTYPES)
@pytest.mark.parametrize("largest", [False, True])
def test_topk_sizes_and_k(dtype, largest, topk_path):
    for n in (1, 2, 33, 1000, 70_001, (1 << 20) + 3):
        keys = make_keys("uniform", n, dtype, seed=n)
        if np.dtype(dtype).kind == "f" and n > 40:
            keys[::17] = -0.0
            keys[::19] = 0.0
        for k in sorted({1, 2, 31, 100, n // 2 + 1, n - 1, n, n + 5}):
            if k < 1:
                continue
            check(keys, k, largest, with_values=(k % 2 == 1))


@pytest.mark.parametrize("dist", ["equal", "few2", "few16", "entropy5", "sorted", "reverse"])
@pytest.mark.parametrize("dtype", [np.uint32, np.float32, np.int64])
def test_topk_ties_and_skew(dist, dtype, topk_path):
    """Many keys tied with the K-th one: exactly K items come back, every index once."""
    n = 300_007
    keys = make_keys(dist, n, dtype, seed=5)
    for k in (1, 1000, n // 3, n - 1):
        check(keys, k, True, True)
        check(keys, k, False, True)


def test_topk_float_specials(topk_path):
    n = 50_000
    keys = make_keys("uniform", n, np.float32, seed=9)
    keys[:100] = np.nan
    keys[100:200] = -np.inf
    keys[200:300] = np.inf
    keys[300:400] = -0.0
    keys[400:500] = 0.0
    for k in (50, 150, 250, 25_000):
        check(keys, k, True, True)
        check(keys, k, False, True)


@pytest.mark.parametrize("vdtype", [np.uint8, np.uint16, np.uint32, np.uint64])
def test_topk_value_widths(vdtype, topk_path):
    n = 200 if vdtype == np.uint8 else 40_000
    keys = make_keys("uniform", n, np.uint32, seed=3)
    check(keys, n // 4, True, True, vdtype=vdtype)
    check(keys, n // 4, False, True, vdtype=vdtype)


def test_topk_edge_cases():
    lib = _native.lib()
    need = ctypes.c_size_t(0)
    # empty input and k == 0: 1 byte of temp, nothing launched
    for n, k in ((0, 5), (100, 0)):
        rc = lib.b200rs_topk(0, ctypes.byref(need), 0x100, 0x100, 0, 0, n, k, 0, 4, 0, 1, 0)
        assert rc == 0 and need.value == 1
        assert lib.b200rs_topk(0x100, ctypes.byref(need), 0x100, 0x100, 0, 0, n, k, 0, 4, 0, 1, 0) == 0
    # temp too small => cudaErrorInvalidValue (1), as the sort
    keys = make_keys("uniform", 1000, np.uint32)
    d_k = to_dev(keys)
    d_o = torch.empty_like(d_k)
    small = ctypes.c_size_t(16)
    t = torch.empty(1024, dtype=torch.uint8, device="cuda")
    assert lib.b200rs_topk(t.data_ptr(), ctypes.byref(small), d_k.data_ptr(), d_o.data_ptr(), 0, 0, 1000, 10, 0, 4, 0, 1,
                           0) == 1
    # unsupported widths
    assert lib.b200rs_topk(0, ctypes.byref(need), 0, 0, 0, 0, 10, 1, 0, 3, 0, 1, 0) == 801
    assert lib.b200rs_topk(0, ctypes.byref(need), 0, 0, 0, 0, 10, 1, 0, 4, 12, 1, 0) == 801


def test_topk_on_side_stream_and_capture(topk_path):
    """Enqueued on the caller's stream without synchronising: legal inside a CUDA graph."""
    keys = make_keys("uniform", 100_000, np.uint32, seed=1)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        check(keys, 777, True, True, stream=s)
    n, k = keys.shape[0], 500
    d_k = to_dev(keys)
    d_o = torch.zeros(k * 4, dtype=torch.uint8, device="cuda")
    lib = _native.lib()
    need = ctypes.c_size_t(0)
    lib.b200rs_topk(0, ctypes.byref(need), d_k.data_ptr(), d_o.data_ptr(), 0, 0, n, k, 0, 4, 0, 0, 0)
    temp = torch.empty(need.value + 256, dtype=torch.uint8, device="cuda")
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        st = torch.cuda.current_stream().cuda_stream
        _native.check(lib.b200rs_topk(temp.data_ptr(), ctypes.byref(need), d_k.data_ptr(), d_o.data_ptr(), 0, 0, n, k, 0,
                                      4, 0, 0, st), "b200rs_topk under capture")
    for _ in range(2):
        d_o.zero_()
        g.replay()
        torch.cuda.synchronize()
        got = np.sort(to_host(d_o, np.uint32, k))
        assert np.array_equal(got, np.sort(keys)[:k])


def test_topk_full_size_property():
    """2^28 u32 keys (BASELINE size): the K-th order statistic separates output from the rest -- checked on device."""
    n, k = 1 << 28, 1 << 20
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    d_k = torch.randint(-(2**31), 2**31, (n,), dtype=torch.int32, device="cuda", generator=g)
    d_o = torch.empty(k, dtype=torch.int32, device="cuda")
    lib = _native.lib()
    need = ctypes.c_size_t(0)
    st = torch.cuda.current_stream().cuda_stream
    args = (d_k.data_ptr(), d_o.data_ptr(), 0, 0, n, k, 1, 4, 0, 1, st)
    _native.check(lib.b200rs_topk(0, ctypes.byref(need), *args), "size")
    temp = torch.empty(need.value + 256, dtype=torch.uint8, device="cuda")
    _native.check(lib.b200rs_topk(temp.data_ptr(), ctypes.byref(need), *args), "b200rs_topk")
    torch.cuda.synchronize()
    kth = int(d_o.min())
    assert int((d_k > kth).sum()) < k <= int((d_k >= kth).sum())
    assert int((d_o > kth).sum()) == int((d_k > kth).sum())


# ---- against the reference ITSELF: cub::DeviceTopK 3.6.0 compiled from the unmodified headers (oracle/ref_cub_topk.cu)
import os  # noqa: E402
import subprocess  # noqa: E402
import tempfile  # noqa: E402

TOPK_BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_cub_topk")
TOPK_KT = {np.dtype(np.uint32): "u32", np.dtype(np.int32): "i32", np.dtype(np.float32): "f32", np.dtype(np.uint64): "u64",
           np.dtype(np.int64): "i64", np.dtype(np.float64): "f64"}


@pytest.mark.skipif(not os.path.exists(TOPK_BIN), reason="oracle/_ref/ref_cub_topk not built")
@pytest.mark.parametrize("dtype,dist,n,k,largest", [
    (np.uint32, "uniform", 1 << 22, 1 << 11, True),
    (np.uint32, "entropy5", 1 << 20, 4097, False),
    (np.int32, "uniform", 300_001, 17, False),
    (np.float32, "uniform", 1 << 21, 1 << 15, True),
    (np.uint64, "uniform", 1 << 20, 1000, True),
    (np.int64, "few256", 1 << 20, 50_000, False),
    (np.float64, "uniform", 200_003, 100_000, True),
])
def test_topk_same_selection_as_reference_cub(dtype, dist, n, k, largest, topk_path):
    """Both return the K best keys in unspecified order: as sorted multisets they must be identical."""
    keys = make_keys(dist, n, dtype, seed=21)
    if np.dtype(dtype).kind == "f":
        keys[np.isnan(keys)] = 1.0  # the reference's top-k orders NaNs by its own comparison rules; not part of this case
    with tempfile.TemporaryDirectory() as d:
        kin, kout = os.path.join(d, "k.bin"), os.path.join(d, "o.bin")
        keys.tofile(kin)
        subprocess.run([TOPK_BIN, "run", TOPK_KT[np.dtype(dtype)], str(n), str(k), str(int(largest)), kin, kout], check=True,
                       timeout=300)
        ref = np.fromfile(kout, dtype=dtype)
    got, _ = gpu_topk(keys, k, largest, False)
    assert np.array_equal(np.sort(got), np.sort(ref))
