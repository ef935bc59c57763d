"""CPU tests: pin the oracle against the reference's golden vectors, the compiled reference (oracle/_ref,
built from /root/reference when present) and the committed fixtures under tests/golden/."""
import glob
import os

import numpy as np
import pytest

from gen import make_keys, make_values
from oracle_lib import oracle_histogram, oracle_sort, ref_thrust, ref_thrust_sort

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_golden_cub_api_example():
    # cub/test/catch2_test_device_radix_sort_env_api.cu:84-136, device_radix_sort.cuh:345-366
    k = np.array([8, 6, 7, 5, 3, 0, 9], dtype=np.int32)
    v = np.arange(7, dtype=np.int32)
    ok, ov = oracle_sort(k, v)
    assert ok.tolist() == [0, 3, 5, 6, 7, 8, 9] and ov.tolist() == [5, 4, 3, 1, 2, 0, 6]
    ok, ov = oracle_sort(k, v, descending=True)
    assert ok.tolist() == [9, 8, 7, 6, 5, 3, 0] and ov.tolist() == [6, 0, 2, 1, 3, 4, 5]


def test_golden_thrust_sort_by_key():
    # thrust/testing/sort_by_key.cu:46-53, thrust/testing/sort.cu:40-48
    k = np.array([1, 3, 6, 5, 2, 0, 4], dtype=np.int32)
    v = np.array([0, 1, 2, 3, 4, 5, 6], dtype=np.int32)
    ok, ov = oracle_sort(k, v)
    assert ok.tolist() == [0, 1, 2, 3, 4, 5, 6] and ov.tolist() == [5, 0, 4, 1, 6, 3, 2]


def test_stability_and_descending_stability():
    # catch2_radix_sort_helper.cuh:227-235: equal keys keep input order in BOTH directions
    k = np.array([2, 1, 2, 1, 2], dtype=np.uint8)
    v = np.arange(5, dtype=np.uint32)
    assert oracle_sort(k, v)[1].tolist() == [1, 3, 0, 2, 4]
    assert oracle_sort(k, v, descending=True)[1].tolist() == [0, 2, 4, 1, 3]


def test_signed_zero_and_nan_order():
    # catch2_test_device_radix_sort_keys.cu:165-219 (+-0 equal, stable), :221-283 (NaNs by bit pattern)
    bits = np.array([0x80000000, 0x00000000, 0x80000000, 0x7FC00000, 0xFFC00000, 0x3F800000, 0xBF800000],
                    dtype=np.uint32)
    k = bits.view(np.float32)
    v = np.arange(k.size, dtype=np.uint32)
    ok, ov = oracle_sort(k, v)
    # -NaN, -1, (-0,+0,-0 in input order), 1, +NaN
    assert ov.tolist() == [4, 6, 0, 1, 2, 5, 3]
    assert ok.view(np.uint32).tolist() == bits[[4, 6, 0, 1, 2, 5, 3]].tolist()


def test_bit_window_only_compares_window():
    # device_radix_sort.cuh:108-112; catch2_test_device_radix_sort_keys.cu:115-161
    k = np.array([0x0100, 0x00FF, 0x0201, 0x0001], dtype=np.uint16)
    v = np.arange(4, dtype=np.uint32)
    ok, ov = oracle_sort(k, v, begin_bit=8, end_bit=16)
    assert ov.tolist() == [1, 3, 0, 2]
    ok, ov = oracle_sort(k, v, begin_bit=0, end_bit=0)  # nothing compared: identity
    assert ov.tolist() == [0, 1, 2, 3]


@pytest.mark.parametrize("dtype", [np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32, np.uint64, np.int64,
                                   np.float32, np.float64])
@pytest.mark.parametrize("descending", [False, True])
def test_oracle_matches_compiled_reference(dtype, descending):
    """thrust::sort_by_key of the unmodified reference (OMP and CPP backends) == oracle, integer keys and
    floats without +-0 ties (SURVEY.md 8c caveat)."""
    if ref_thrust("omp") is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    n = 50_000
    for dist in ("uniform", "entropy3", "few16"):
        k = make_keys(dist, n, dtype, seed=3)
        if np.dtype(dtype).kind == "f":
            k = k[~np.isnan(k)]
            k = k[k != 0]  # +-0 ties order differently in thrust's CPU encoder
        v = make_values(k.size, np.uint32)
        ok, ov = oracle_sort(k, v, descending=descending)
        for backend in ("omp", "cpp"):
            rk, rv, _ = ref_thrust_sort(k, v, descending=descending, backend=backend)
            assert np.array_equal(ok.view(np.uint8), rk.view(np.uint8)), (dist, backend)
            assert np.array_equal(ov, rv), (dist, backend)


def test_histogram_oracle_consistent_with_sort():
    k = make_keys("uniform", 10_000, np.int32, seed=5)
    h = oracle_histogram(k, begin_bit=4, end_bit=29)
    assert h.shape == (4, 256) and (h.sum(axis=1) == k.size).all()
    assert h[3, 2:].sum() == 0  # last pass has 1 bit


def test_committed_golden_fixtures():
    files = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))
    assert files, "tests/golden/*.npz missing (run tests/golden/make_golden.py where /root/reference exists)"
    for f in files:
        z = np.load(f)
        meta = {k: z[k].item() for k in ("descending", "begin_bit", "end_bit")}
        got = oracle_sort(z["keys_in"], z["vals_in"] if "vals_in" in z else None, **meta)
        if "vals_in" in z:
            assert np.array_equal(got[0].view(np.uint8), z["keys_out"].view(np.uint8)), f
            assert np.array_equal(got[1], z["vals_out"]), f
        else:
            assert np.array_equal(got.view(np.uint8), z["keys_out"].view(np.uint8)), f


def test_16_bit_float_keys_in_the_oracle():
    """half keys (reference: util_type.cuh:1017-1095, device_radix_sort.cuh:51-57): the generic transform on a 16-bit
    pattern.  Checked against numpy's own float16 ordering on NaN-free data, plus +-0 stability and NaN placement."""
    rng = np.random.default_rng(5)
    k = rng.standard_normal(20_000).astype(np.float16)
    k[::13] = np.float16(-0.0)
    k[::17] = np.float16(0.0)
    v = np.arange(k.size, dtype=np.uint32)
    for desc in (False, True):
        ok, ov = oracle_sort(k, v, descending=desc)
        want = np.sort(k, kind="stable")
        assert np.array_equal(ok.astype(np.float32), want[::-1].astype(np.float32) if desc else want.astype(np.float32))
        z = ok == 0  # both zeros tie: they keep input order (values increasing inside the run of zeros)
        assert (np.diff(ov[z].astype(np.int64)) > 0).all()
    bits = np.array([0x7E00, 0xFE00, 0x3C00, 0xBC00, 0x7C00, 0xFC00], dtype=np.uint16).view(np.float16)
    ok = oracle_sort(bits)
    assert ok.view(np.uint16).tolist() == [0xFE00, 0xFC00, 0xBC00, 0x3C00, 0x7C00, 0x7E00]  # -NaN -inf -1 1 +inf +NaN


def test_segmented_oracle_against_reference_cub_fixtures():
    """oracle_segmented_sort (the checker for SURVEY.md 8f-1) against outputs of the UNMODIFIED reference's
    cub::DeviceSegmentedRadixSort run on a B200 (tests/golden/segmented/make_golden_cub_segmented.py): empty segments, gaps,
    one-item segments, a segment above the single-tile size, descending, bit window, +-0.0."""
    import glob

    from oracle_lib import oracle_segmented_sort

    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "segmented", "cubseg_*.npz")))
    assert len(files) >= 5
    for f in files:
        z = np.load(f)
        kw = dict(descending=bool(z["descending"]), begin_bit=int(z["begin_bit"]), end_bit=int(z["end_bit"]))
        if "vals_in" in z.files:
            ok, ov = oracle_segmented_sort(z["keys_in"], z["vals_in"], z["begin_offsets"], z["end_offsets"], **kw)
            assert np.array_equal(ov, z["vals_out"]), f
        else:
            ok = oracle_segmented_sort(z["keys_in"], None, z["begin_offsets"], z["end_offsets"], **kw)
        assert np.array_equal(ok.view(np.uint8), z["keys_out"].view(np.uint8)), f


def test_topk_oracle_against_the_reference_documentation_examples():
    """cub/test/catch2_test_device_topk_api.cu:115-158 (MinPairs, k = 4) and :165-210 (MaxPairs, k = 4): keys
    {5, -3, 1, 7, 8, 2, 4, 6} with their indices as values."""
    from oracle_lib import oracle_topk

    keys = np.array([5, -3, 1, 7, 8, 2, 4, 6], dtype=np.int32)
    vals = np.arange(8, dtype=np.int32)
    k, v = oracle_topk(keys, vals, 4, largest=False)
    assert k.tolist() == [-3, 1, 2, 4] and v.tolist() == [1, 2, 5, 6]
    k, v = oracle_topk(keys, vals, 4, largest=True)
    assert k.tolist() == [8, 7, 6, 5] and v.tolist() == [4, 3, 7, 0]
    assert oracle_topk(keys, None, 100, largest=True).tolist() == sorted(keys.tolist(), reverse=True)  # k capped to n
    assert oracle_topk(keys, None, 0, largest=True).size == 0
