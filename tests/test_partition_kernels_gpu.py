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
                tp = torch.from_numpy(prefixes.view(np.int64).copy())
                got = cuda.select_histogram(dk, tp.cuda(), rnd, desc).cpu()
                assert torch.equal(got, ref.select_histogram(hk, tp, rnd, desc)), (n, desc, rnd)
            if kb >= 4 and n:
                # candidate compaction: round 1 emits the keys that carry a 1-byte prefix, round 2 scans only those
                # (all-equal keys overflow the candidate buffer, which must fall back to scanning every key)
                p2 = np.unique(v >> np.uint64(8 * kb - 16))[:5]
                p2 = np.unique(np.concatenate([p2, np.array([3], dtype=np.uint64)]))
                p1 = np.unique(p2 >> np.uint64(8))
                t1 = torch.from_numpy(p1.view(np.int64).copy())
                t2 = torch.from_numpy(p2.view(np.int64).copy())
                got1 = cuda.select_histogram(dk, t1.cuda(), 1, desc, candidates="emit").cpu()
                assert torch.equal(got1, ref.select_histogram(hk, t1, 1, desc)), (n, desc, "emit round")
                got2 = cuda.select_histogram(dk, t2.cuda(), 2, desc, candidates="use").cpu()
                assert torch.equal(got2, ref.select_histogram(hk, t2, 2, desc)), (n, desc, "candidate round")
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
                # the fused pass (one onesweep launch whose digit is the bucket) must give the same arrangement
                sizes = np.bincount(ref.bucket_ids(hk, qs, desc).numpy(), minlength=2 * len(qs) + 1)
                for vals in (torch.arange(n, dtype=torch.int32, device="cuda"), None):
                    fused = cuda.partition_by_splitters(dk, vals, qs, sizes, desc)
                    if fused is None:
                        assert kb < 4, "the fused partition pass must exist for 4- and 8-byte keys"
                        continue
                    assert torch.equal(fused[0].cpu().view(torch.uint8), rk.view(torch.uint8)), (n, desc, "fused keys")
                    if vals is not None:
                        assert torch.equal(fused[1].cpu(), rv), (n, desc, "fused values")
                    assert torch.equal(dk.cpu().view(torch.uint8), hk.view(torch.uint8)), "input modified"
