"""The two multi-GPU support kernels (b200rs_select_histogram, b200rs_bucket_ids) against their numpy restatement
(tests/test_multi_gpu_host.py OracleOps), through CudaOps, for every key type, both orders, skewed inputs."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from gen import make_keys  # noqa: E402
from test_multi_gpu_host import OracleOps, to_torch  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.uint32, np.int32, np.float32, np.uint64, np.int64, np.float64])
@pytest.mark.parametrize("dist", ["uniform", "entropy5", "equal", "few16"])
def test_select_histogram_and_bucket_ids(dtype, dist):
    from cccl_b200.multi_gpu import CudaOps

    cuda, ref = CudaOps(), OracleOps()
    for n in (0, 1, 31, 33, 100_003):
        k = make_keys(dist, n, dtype, seed=11)
        if np.dtype(dtype).kind == "f" and n > 40:
            k[::17] = -0.0
            k[::19] = 0.0
        hk = to_torch(k)
        dk = hk.cuda()
        kb = np.dtype(dtype).itemsize
        for desc in (False, True):
            v, _ = ref._kv(hk, desc)
            top = cuda.top_digit_histogram(dk, desc).cpu()
            assert torch.equal(top, ref.top_digit_histogram(hk, desc)), (n, desc, "top digit")
            # prefixes that exist in the data plus one that (probably) does not
            for rnd in range(1, kb):
                shift = np.uint64(8 * kb - 8 * rnd)
                have = np.unique(v >> shift)[:6] if n else np.zeros(0, dtype=np.uint64)
                prefixes = np.unique(np.concatenate([have, np.array([1], dtype=np.uint64)]))
                got = cuda.select_histogram(dk, prefixes, rnd, desc).cpu()
                assert torch.equal(got, ref.select_histogram(hk, prefixes, rnd, desc)), (n, desc, rnd)
            vs = np.sort(v)
            picks = vs[[n // 10, n // 2, n // 2, (9 * n) // 10]] if n else np.zeros(0, dtype=np.uint64)
            qs = np.unique(np.concatenate([picks, np.array([0, 5], dtype=np.uint64)]))
            got = cuda.bucket_ids(dk, qs, desc).cpu()
            assert torch.equal(got, ref.bucket_ids(hk, qs, desc)), (n, desc, "ids")
            if n:
                pk, pv = cuda.partition(cuda.bucket_ids(dk, qs, desc), int(2 * len(qs)).bit_length(), dk,
                                        torch.arange(n, dtype=torch.int32, device="cuda"))
                rk, rv = ref.partition(ref.bucket_ids(hk, qs, desc), int(2 * len(qs)).bit_length(), hk,
                                       torch.arange(n, dtype=torch.int32))
                assert torch.equal(pk.cpu().view(torch.uint8), rk.view(torch.uint8)), (n, desc, "partition keys")
                assert torch.equal(pv.cpu(), rv), (n, desc, "partition values")
