"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path through the C ABI against the oracle on the
same seeded inputs, bit-exact, plus size-independent properties at BASELINE.json's full sizes.

Case list mirrors cub/test/catch2_test_device_radix_sort_keys.cu (:75-113 sizes x order, :115-161 bit windows,
:165-219 +-0, :221-283 NaN, :287-340 entropy, :342-374 all-equal, :418-447 DoubleBuffer) and ..._pairs.cu
(:29 value widths, :81-86 stability, :89-131 DoubleBuffer)."""
import numpy as np
import pytest
import torch

from gen import V16, make_keys, make_values
from gpu_util import assert_same_bits, gpu_histogram, gpu_sort
from oracle_lib import oracle_histogram, oracle_sort

pytestmark = pytest.mark.gpu

SMALL_DEFAULT = 2**64 - 1  # b200rs_set_small_max: the library's own threshold

KEY_DTYPES = [np.uint8, np.int8, np.uint16, np.int16, np.float16, np.uint32, np.int32, np.float32, np.uint64, np.int64,
              np.float64]


@pytest.fixture(params=["fused_small", "general"])
def sort_path(request):
    """Inputs up to 2^20 items (4- / 8-byte keys with 0- / 4- / 8-byte values) take the one-launch cooperative kernel
    (csrc/small.cu) by default; "general" switches it off so that the same cases also exercise the multi-kernel path."""
    from cccl_b200 import _native

    lib = _native.lib()
    lib.b200rs_set_small_max(0 if request.param == "general" else SMALL_DEFAULT)
    yield request.param
    lib.b200rs_set_small_max(SMALL_DEFAULT)


def check_keys(k, **kw):
    got, info = gpu_sort(k, **kw)
    okw = {x: kw[x] for x in ("descending", "begin_bit", "end_bit") if x in kw}
    assert_same_bits(got, oracle_sort(k, **okw), f"keys {k.dtype} n={k.size} {kw}")
    return info


def check_pairs(k, v, **kw):
    gk, gv, info = gpu_sort(k, v, **kw)
    okw = {x: kw[x] for x in ("descending", "begin_bit", "end_bit") if x in kw}
    ok, ov = oracle_sort(k, v, **okw)
    assert_same_bits(gk, ok, f"pair keys {k.dtype}/{v.dtype} n={k.size} {kw}")
    assert_same_bits(gv, ov, f"pair values (stability) {k.dtype}/{v.dtype} n={k.size} {kw}")
    return info


def test_golden_vectors_on_gpu():
    k = np.array([8, 6, 7, 5, 3, 0, 9], dtype=np.int32)
    v = np.arange(7, dtype=np.int32)
    gk, gv, _ = gpu_sort(k, v)
    assert gk.tolist() == [0, 3, 5, 6, 7, 8, 9] and gv.tolist() == [5, 4, 3, 1, 2, 0, 6]
    gk, gv, _ = gpu_sort(k, v, descending=True)
    assert gk.tolist() == [9, 8, 7, 6, 5, 3, 0] and gv.tolist() == [6, 0, 2, 1, 3, 4, 5]
    k = np.array([1, 3, 6, 5, 2, 0, 4], dtype=np.int32)
    gk, gv, _ = gpu_sort(k, v, api="double")
    assert gk.tolist() == [0, 1, 2, 3, 4, 5, 6] and gv.tolist() == [5, 0, 4, 1, 6, 3, 2]


def test_committed_golden_fixtures_on_gpu():
    import glob
    import os

    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))
    assert files
    for f in files:
        z = np.load(f)
        kw = dict(descending=bool(z["descending"]), begin_bit=int(z["begin_bit"]), end_bit=int(z["end_bit"]))
        if "vals_in" in z:
            gk, gv, _ = gpu_sort(z["keys_in"], z["vals_in"], **kw)
            assert_same_bits(gk, z["keys_out"], f)
            assert_same_bits(gv, z["vals_out"], f)
        else:
            gk, _ = gpu_sort(z["keys_in"], **kw)
            assert_same_bits(gk, z["keys_out"], f)


@pytest.mark.parametrize("dtype", KEY_DTYPES)
@pytest.mark.parametrize("descending", [False, True])
def test_keys_all_types_sizes(dtype, descending, sort_path):
    rng = np.random.default_rng(11)
    sizes = [0, 1, 2, 31, 32, 33, 255, 4095, 4096, 4097, 8191, 8192, 8193, 12289, 100_003]
    sizes += rng.integers(32, 1 << 20, size=3).tolist()
    for n in sizes:
        check_keys(make_keys("uniform", n, dtype, seed=n + 1), descending=descending)


@pytest.mark.parametrize("dtype", [np.uint32, np.uint64, np.float32, np.int64, np.uint16, np.uint8])
def test_bit_windows(dtype, sort_path):
    bits = np.dtype(dtype).itemsize * 8
    cuts = sorted({0, bits // 3, 3 * bits // 4, bits})
    k = make_keys("uniform", 70_001, dtype, seed=9)
    v = make_values(k.size, np.uint32)
    for b in cuts:
        for e in cuts:
            if b <= e:
                for desc in (False, True):
                    check_pairs(k, v, begin_bit=b, end_bit=e, descending=desc)
    check_keys(k, begin_bit=3, end_bit=4)


@pytest.mark.parametrize("dtype", [np.float16, np.float32, np.float64])
def test_signed_zeros_and_nans(dtype, sort_path):
    udt = {np.float16: np.uint16, np.float32: np.uint32, np.float64: np.uint64}[dtype]
    n = 50_000
    rng = np.random.default_rng(2)
    k = rng.standard_normal(n).astype(dtype)
    k[rng.integers(0, n, n // 5)] = 0.0
    k[rng.integers(0, n, n // 5)] = -0.0
    nanbits = make_keys("uniform", n // 10, dtype, seed=4).view(udt)
    expmask = udt({np.float16: 0x7C00, np.float32: 0x7F800000, np.float64: 0x7FF0000000000000}[dtype])
    idx = rng.integers(0, n, n // 10)
    k.view(udt)[idx] = nanbits | expmask | udt(1)  # +-NaN with random payloads
    k[rng.integers(0, n, 100)] = np.inf
    k[rng.integers(0, n, 100)] = -np.inf
    v = make_values(n, np.uint32)
    for desc in (False, True):
        check_pairs(k, v, descending=desc)
        check_pairs(k, v, descending=desc, api="double")
        check_pairs(k, v, descending=desc, begin_bit=5, end_bit=np.dtype(dtype).itemsize * 8 - 1)


@pytest.mark.parametrize("dtype", [np.float16, np.float32, np.float64])
def test_float_zero_flag_switches_the_pass_body(dtype):
    """The general path runs the passes without the per-key zero test unless the upsweep saw a key with the aliased zero
    pattern (PassArgs::zero_flag): no such key, ONE such key at either end / in the last partial tile, many of them --
    results stay bit-exact and +-0.0 keep their input order (values), in both directions, on several portions too."""
    from cccl_b200 import _native

    lib = _native.lib()
    lib.b200rs_set_small_max(0)  # general path only: the one-launch kernel always tests
    try:
        n = 600_011
        base = make_keys("uniform", n, dtype, seed=13)
        base[base == 0] = 1.0  # start from an input with no zero of either sign
        v = make_values(n, np.uint32)
        variants = {"none": []}
        variants["one_first"] = [(0, -0.0)]
        variants["one_last"] = [(n - 1, -0.0), (5, 0.0)]
        variants["pos_only"] = [(7, 0.0), (n - 2, 0.0)]
        variants["many"] = [(i, -0.0 if i % 3 else 0.0) for i in range(0, n, 1001)]
        for name, edits in variants.items():
            k = base.copy()
            for i, z in edits:
                k[i] = z
            for desc in (False, True):
                check_pairs(k, v, descending=desc)
                check_keys(k, descending=desc, api="double")
                check_pairs(k, v, descending=desc, begin_bit=3, end_bit=np.dtype(dtype).itemsize * 8 - 2)
        lib.b200rs_set_portion_items(50_000)
        k = base.copy()
        k[n - 3] = -0.0
        k[n // 2] = 0.0
        check_pairs(k, v)
        check_pairs(k, v, descending=True)
    finally:
        lib.b200rs_set_portion_items(0)
        lib.b200rs_set_small_max(SMALL_DEFAULT)


@pytest.mark.parametrize("dist", ["entropy2", "entropy3", "entropy5", "equal", "few2", "few16", "few256", "sorted",
                                  "reverse"])
@pytest.mark.parametrize("dtype", [np.uint32, np.uint64])
def test_skewed_distributions_pairs_stable(dist, dtype, sort_path):
    n = 300_007
    k = make_keys(dist, n, dtype, seed=21)
    v = make_values(n, np.uint32)
    check_pairs(k, v)
    check_pairs(k, v, descending=True, api="double")


@pytest.mark.parametrize("vdtype", [np.uint8, np.uint16, np.uint32, np.uint64, V16])
@pytest.mark.parametrize("kdtype", [np.uint8, np.uint16, np.uint32, np.uint64])
def test_value_widths(kdtype, vdtype, sort_path):
    for n in (1, 777, 40_000):
        k = make_keys("few256" if n > 1000 else "uniform", n, kdtype, seed=n)
        v = make_values(n, vdtype)
        check_pairs(k, v)
        check_pairs(k, v, api="double", descending=True)


@pytest.mark.parametrize("kdtype", [np.uint8, np.int16, np.uint32, np.float32, np.int64, np.float64])
def test_single_tile_kernel_and_general_path_agree(kdtype):
    """Inputs of at most one tile take the single-CTA kernel (reference: DeviceRadixSortSingleTileKernel,
    kernel_radix_sort.cuh:330-434, chosen at dispatch_radix_sort.cuh:1980).  Sizes straddle its capacities
    (5120 / 2560 / 1280 items for <= 4 / 8 / 16-byte items); the same inputs are also pushed through the general
    multi-kernel path (diagnostic switch) -- both must equal the oracle, one launch vs several."""
    from cccl_b200 import _native

    lib = _native.lib()
    bits = np.dtype(kdtype).itemsize * 8
    sizes = [1, 2, 33, 1000, 1279, 1280, 1281, 2559, 2560, 2561, 4864, 5119, 5120, 5121]
    for vdtype in (None, np.uint8, np.uint32, np.uint64, V16):
        for n in sizes:
            k = make_keys("few256" if n % 2 else "uniform", n, kdtype, seed=n)
            if np.dtype(kdtype).kind == "f":
                k[::7] = -0.0
                k[::11] = 0.0
            v = make_values(n, vdtype) if vdtype is not None else None
            dom = max(np.dtype(kdtype).itemsize, np.dtype(vdtype).itemsize if vdtype is not None else 0)
            cap = 256 * (20 if dom <= 4 else 10 if dom <= 8 else 5)
            for kw in (dict(), dict(descending=True, api="double"), dict(begin_bit=bits // 4, end_bit=bits - 3,
                                                                          descending=True)):
                info = check_pairs(k, v, **kw) if v is not None else check_keys(k, **kw)
                # the single-CTA kernel needs no temp storage (1 byte, like the reference's empty layout)
                assert (info["temp_bytes"] == 1) == (n <= cap), (n, cap, info)
                if n <= cap:
                    assert info["launches"] == 1
                    try:
                        lib.b200rs_set_single_tile(0)
                        # without the single-CTA kernel: the one-launch cooperative kernel where it is compiled ...
                        info2 = check_pairs(k, v, **kw) if v is not None else check_keys(k, **kw)
                        fused = np.dtype(kdtype).itemsize in (4, 8) and (vdtype is None or np.dtype(vdtype).itemsize in (4, 8))
                        assert (info2["launches"] == 1) == fused, (n, info2)
                        # ... and the general multi-kernel path
                        lib.b200rs_set_small_max(0)
                        info3 = check_pairs(k, v, **kw) if v is not None else check_keys(k, **kw)
                        assert info3["launches"] > 1
                    finally:
                        lib.b200rs_set_single_tile(1)
                        lib.b200rs_set_small_max(SMALL_DEFAULT)


def test_double_buffer_selector_and_pass_parity(sort_path):
    # result is wherever selector says (catch2_test_device_radix_sort_keys.cu:441-446)
    k = make_keys("uniform", 50_000, np.uint32, seed=8)
    for e, want_sel in ((8, 1), (16, 0), (24, 1), (32, 0)):
        info = check_keys(k, begin_bit=0, end_bit=e, api="double")
        assert info["selector"] == want_sel
    info = check_keys(k, begin_bit=8, end_bit=8, api="double")
    assert info["selector"] == 0 and info["temp_bytes"] == 1
    info = check_keys(k, begin_bit=8, end_bit=8, api="pointer")  # pointer API: plain copy
    assert info["selector"] == 1


def test_unaligned_temp_storage_and_streams(sort_path):
    k = make_keys("uniform", 123_457, np.uint32, seed=1)
    v = make_values(k.size, np.uint64)
    check_pairs(k, v, temp_misalign=3)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        check_pairs(k, v, stream=s)


def test_multi_portion_path():
    """Forces tiny portions so several launches per pass chain through bins_next
    (reference: portion loop, dispatch_radix_sort.cuh:1859-1933; large-N tests ..._keys.cu:488-519)."""
    from cccl_b200 import _native

    lib = _native.lib()
    try:
        lib.b200rs_set_portion_items(20_000)
        for dtype in (np.uint32, np.uint64):
            k = make_keys("entropy3", 131_072 + 77, dtype, seed=5)
            v = make_values(k.size, np.uint32)
            info = check_pairs(k, v)
            assert info["launches"] > 3 + np.dtype(dtype).itemsize
            check_pairs(k, v, api="double", descending=True)
    finally:
        lib.b200rs_set_portion_items(0)


def test_every_compiled_config_is_correct():
    from cccl_b200 import _native

    lib = _native.lib()
    try:
        for kdt, vdt in ((np.uint32, None), (np.uint32, np.uint32), (np.uint64, None), (np.uint64, np.uint32)):
            n_cfg = len(_native.describe_configs(np.dtype(kdt).itemsize, np.dtype(vdt).itemsize if vdt else 0))
            # all-equal / two-value inputs drive the single-digit short circuits (reference:
            # agent_radix_sort_onesweep.cuh:386-467) of the configurations that have them
            for dist in ("entropy2", "equal", "few2"):
                k = make_keys(dist, 200_003, kdt, seed=77)
                for c in range(n_cfg):
                    lib.b200rs_set_config(c)
                    if vdt is None:
                        check_keys(k)
                        check_keys(k, descending=True)
                    else:
                        check_pairs(k, make_values(k.size, vdt))
    finally:
        lib.b200rs_set_config(-1)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int16, np.float16, np.uint32, np.int32, np.uint64, np.float32])
def test_upsweep_histogram_alone(dtype):
    for n, off in ((0, 0), (5, 1), (100_003, 0), (100_003, 3), (1 << 20, 1)):
        k = make_keys("entropy2", n, dtype, seed=3)
        for desc in (False, True):
            got = gpu_histogram(k, descending=desc, offset_items=off)
            want = oracle_histogram(k, descending=desc)
            assert np.array_equal(got, want), (dtype, n, off, desc)
    k = make_keys("uniform", 50_000, dtype, seed=3)
    bits = np.dtype(dtype).itemsize * 8
    got = gpu_histogram(k, begin_bit=bits // 3, end_bit=bits - 1)
    assert np.array_equal(got, oracle_histogram(k, begin_bit=bits // 3, end_bit=bits - 1))
    # every window start modulo 8 and narrow last digits (the folded addressing rotates by begin_bit - 7 modulo 32)
    for b in range(0, min(bits, 12)):
        for e in sorted({b + 1, min(bits, b + 9), bits}):
            for desc in (False, True):
                got = gpu_histogram(k, descending=desc, begin_bit=b, end_bit=e)
                assert np.array_equal(got, oracle_histogram(k, descending=desc, begin_bit=b, end_bit=e)), (dtype, b, e, desc)


def test_cuda_graph_capture(sort_path):
    """The execute call is allocation-free and stream-ordered: legal under capture
    (cub/test/catch2_test_launch_helper.h:91-121)."""
    from cccl_b200 import _native
    from gpu_util import to_dev, to_host

    k = make_keys("uniform", 300_000, np.uint32, seed=6)
    d_in, d_out = to_dev(k), to_dev(np.zeros_like(k))
    need, _ = _native.sort_raw(0, 0, d_in.data_ptr(), d_out.data_ptr(), 0, 0, k.size, 0, 4, 0, 0, 32, False, False)
    temp = torch.empty(need, dtype=torch.uint8, device="cuda")
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            _native.sort_raw(temp.data_ptr(), need, d_in.data_ptr(), d_out.data_ptr(), 0, 0, k.size, 0, 4, 0, 0, 32,
                             False, False, s.cuda_stream)
    for seed in (6, 7):
        k = make_keys("uniform", 300_000, np.uint32, seed=seed)
        d_in.copy_(to_dev(k))
        g.replay()
        torch.cuda.synchronize()
        assert_same_bits(to_host(d_out, np.uint32, k.size), np.sort(k), "graph replay")


def test_python_mirror_api():
    from cccl_b200 import DoubleBuffer, SortOrder, make_radix_sort, radix_sort

    h_k = np.array([-5, 4, 2, -3, 2, 4, 0, -1, 2, 8], dtype=np.int32)
    h_v = np.arange(10, dtype=np.float32)
    d_k, d_v = torch.from_numpy(h_k).cuda(), torch.from_numpy(h_v).cuda()
    o_k, o_v = torch.empty_like(d_k), torch.empty_like(d_v)
    radix_sort(d_in_keys=d_k, d_out_keys=o_k, d_in_values=d_v, d_out_values=o_v, num_items=10,
               order=SortOrder.ASCENDING)
    torch.cuda.synchronize()
    order = np.argsort(h_k, kind="stable")
    assert o_k.cpu().numpy().tolist() == h_k[order].tolist() and o_v.cpu().numpy().tolist() == h_v[order].tolist()
    kb, vb = DoubleBuffer(d_k, o_k), DoubleBuffer(d_v, o_v)
    sorter = make_radix_sort(d_in_keys=kb, d_out_keys=None, d_in_values=vb, d_out_values=None,
                             order=SortOrder.DESCENDING)
    kw = dict(d_in_keys=kb, d_out_keys=None, d_in_values=vb, d_out_values=None, num_items=10)
    nbytes = sorter(temp_storage=None, **kw)
    temp = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    sorter(temp_storage=temp, **kw)
    torch.cuda.synchronize()
    order = np.argsort(-h_k.astype(np.int64), kind="stable")
    assert kb.current().cpu().numpy().tolist() == h_k[order].tolist()
    assert vb.current().cpu().numpy().tolist() == h_v[order].tolist()
    assert kb.selector == vb.selector


def _checksum(t: torch.Tensor):
    return int(t.view(torch.int64).sum().item()) if t.numel() % 1 == 0 else 0


@pytest.mark.parametrize("case", ["u32_keys", "u64_u32_pairs_uniform", "u64_u32_pairs_entropy5", "f32_desc_window",
                                  "i64_desc_window"])
def test_full_size_properties(case):
    """BASELINE.json sizes (2^28): sortedness under the sort's own order, multiset preservation (order-independent
    checksums), value stability (equal-key runs carry increasing original indices), key/value pairing."""
    n = 1 << 28
    g = torch.Generator(device="cuda").manual_seed(5)

    def rnd64(count):
        return torch.randint(-(2**63), 2**63 - 1, (count,), dtype=torch.int64, device="cuda", generator=g)

    from cccl_b200 import _native

    if case == "u32_keys":
        keys = rnd64(n // 2).view(torch.int32)  # raw bits of n u32 keys
        kind, kb, vb, desc, b, e = 0, 4, 0, False, 0, 32
    elif case.startswith("u64_u32_pairs"):
        keys = rnd64(n)
        if case.endswith("entropy5"):
            for _ in range(4):
                keys &= rnd64(n)
        kind, kb, vb, desc, b, e = 0, 8, 4, False, 0, 64
    elif case == "f32_desc_window":
        keys = rnd64(n // 2).view(torch.int32)
        kind, kb, vb, desc, b, e = 2, 4, 0, True, 8, 24
    else:
        keys = rnd64(n)
        kind, kb, vb, desc, b, e = 1, 8, 0, True, 16, 48
    out = torch.empty_like(keys)
    vals = torch.arange(n, dtype=torch.int32, device="cuda") if vb else None
    vout = torch.empty_like(vals) if vb else None
    p = lambda t: t.data_ptr() if t is not None else 0
    need, _ = _native.sort_raw(0, 0, p(keys), p(out), p(vals), p(vout), n, kind, kb, vb, b, e, desc, False)
    temp = torch.empty(need, dtype=torch.uint8, device="cuda")
    _, sel = _native.sort_raw(temp.data_ptr(), need, p(keys), p(out), p(vals), p(vout), n, kind, kb, vb, b, e, desc,
                              False)
    torch.cuda.synchronize()
    assert sel == 1
    del temp

    # order key: the transformed, window-masked key as a signed-comparable int64 (computed with torch on device;
    # this is the CHECK, not the product)
    def order_key(t):
        if kb == 4:
            u = t.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
            if kind == 2:
                neg = (u >> 31) == 1
                u = torch.where(u == 0x80000000, torch.zeros_like(u), u)  # -0 -> +0
                neg = (u >> 31) == 1
                u = torch.where(neg, (~u) & 0xFFFFFFFF, u ^ 0x80000000)
            w = (u >> b) & ((1 << (e - b)) - 1)
            return w
        u = t.view(torch.int64)
        if kind == 1:
            u = u ^ (-(2**63))
        # unsigned 64-bit window -> keep as int64 after shifting so it compares correctly
        if e - b < 64:
            w = (u >> b) & ((1 << (e - b)) - 1)
            return w
        return u ^ (-(2**63))  # unsigned order as signed order

    chunk = 1 << 26
    prev_last = None
    for s in range(0, n, chunk):
        ok = order_key(out[s:s + chunk])
        d = ok[1:] - ok[:-1] if (e - b) < 63 else None
        if desc:
            good = ok[1:] <= ok[:-1]
        else:
            good = ok[1:] >= ok[:-1]
        assert bool(good.all()), f"{case}: not sorted in chunk at {s}"
        if vb:
            same = ok[1:] == ok[:-1]
            vv = vout[s:s + chunk].to(torch.int64)
            assert bool((~same | (vv[1:] > vv[:-1])).all()), f"{case}: stability broken in chunk at {s}"
        if prev_last is not None:
            assert (ok[0] <= prev_last) if desc else (ok[0] >= prev_last)
        prev_last = ok[-1]
        del ok, good
    # multiset preserved: sum and xor-like checksums of raw bits are order independent
    a64, b64 = keys.view(torch.int64), out.view(torch.int64)
    if kb == 8:
        assert int(a64.sum()) == int(b64.sum())
        assert int((a64 * a64).sum()) == int((b64 * b64).sum())
    else:
        a32, b32 = keys.view(torch.int32).to(torch.int64), out.view(torch.int32).to(torch.int64)
        assert int(a32.sum()) == int(b32.sum())
        assert int((a32 * a32).sum()) == int((b32 * b32).sum())
        del a32, b32
    if vb:
        # pairing: out[i] must equal keys[vout[i]]
        for s in range(0, n, chunk):
            idx = vout[s:s + chunk].to(torch.int64)
            assert bool((keys.view(torch.int64)[idx] == out.view(torch.int64)[s:s + chunk]).all())


@pytest.mark.parametrize("kb,extra", [(1, 12_345), (2, 4_099)])
def test_more_than_2_to_the_32_items(kb, extra):
    """NumItemsT = 64-bit (reference: choose_offset.cuh:35-52, large-N tests catch2_test_device_radix_sort_keys.cu:488-519):
    N > 2^32 items takes the 64-bit-offset kernels and several chained portions per pass for real (not through the
    diagnostic switches).  1- and 2-byte keys keep it to one or two passes; the check is exact and runs on the device:
    the output of a key sort is the digit histogram written out in order."""
    from cccl_b200 import _native

    n = (1 << 32) + extra
    free, _ = torch.cuda.mem_get_info()
    if free < 3 * n * kb + (4 << 30):
        pytest.skip("not enough free device memory")
    g = torch.Generator(device="cuda").manual_seed(17)
    words = (n * kb + 7) // 8
    keys = torch.randint(-(2**63), 2**63 - 1, (words,), dtype=torch.int64, device="cuda", generator=g).view(torch.uint8)[: n * kb]
    out = torch.empty_like(keys)
    need, _ = _native.sort_raw(0, 0, keys.data_ptr(), out.data_ptr(), 0, 0, n, 0, kb, 0, 0, 8 * kb, True, False)
    temp = torch.empty(need, dtype=torch.uint8, device="cuda")
    _, sel = _native.sort_raw(temp.data_ptr(), need, keys.data_ptr(), out.data_ptr(), 0, 0, n, 0, kb, 0, 0, 8 * kb, True,
                              False)
    torch.cuda.synchronize()
    assert sel == 1
    launches = _native.lib().b200rs_last_launch_count()
    assert launches >= 3 + kb * 5, launches  # 5 portions of < 2^30 items per pass
    del temp
    dt = torch.uint8 if kb == 1 else torch.int16
    kin, kout = keys.view(dt), out.view(dt)
    nbins = 1 << (8 * kb)
    counts = torch.zeros(nbins, dtype=torch.int64, device="cuda")
    chunk = 1 << 28
    for lo in range(0, n, chunk):  # histogram of the input in pieces (bincount wants int64 copies)
        piece = kin[lo:lo + chunk].to(torch.int64) & (nbins - 1)
        counts += torch.bincount(piece, minlength=nbins)
    assert int(counts.sum()) == n
    # descending: bin nbins-1 first.  Compare run boundaries instead of materialising the expected array.
    order = torch.arange(nbins - 1, -1, -1, device="cuda")
    ends = torch.cumsum(counts[order], 0)
    starts = ends - counts[order]
    nz = counts[order] > 0
    first = (kout[starts[nz].clamp(max=n - 1)].to(torch.int64) & (nbins - 1))
    last = (kout[(ends[nz] - 1)].to(torch.int64) & (nbins - 1))
    assert torch.equal(first, order[nz]) and torch.equal(last, order[nz])
    # and every adjacent pair is non-increasing
    for lo in range(0, n - 1, chunk):
        a = kout[lo:lo + chunk + 1].to(torch.int32) & (nbins - 1)
        assert bool((a[1:] <= a[:-1]).all())


def test_debug_sync_logging():
    """B200RS_DEBUG_SYNC=1 (the counterpart of CUB_DEBUG_SYNC, cub/cub/util_debug.cuh:37-72): every stream operation of
    the sort is synchronised and logged with its status; results are unchanged."""
    import os
    import subprocess
    import sys

    code = (
        "import numpy as np, sys; sys.path.insert(0, 'tests');"
        "from gen import make_keys; from gpu_util import gpu_sort; from oracle_lib import oracle_sort;"
        "k = make_keys('uniform', 3_000_000, np.uint32, seed=1); g, _ = gpu_sort(k);"
        "assert np.array_equal(g, oracle_sort(k)); print('ok')"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, B200RS_DEBUG_SYNC="1"),
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr[-2000:]
    for op in ("memset", "histogram", "scan", "onesweep"):
        assert f"({op})" in res.stderr, res.stderr[-2000:]
    assert res.stderr.count("(onesweep)") == 4 and "no error" in res.stderr
