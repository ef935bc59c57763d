"""One b200rs_segmented_sort call for profiling: 2^24 (u32,u32) pairs, mean segment length argv[1] (default 20000).
ncu --metrics gpu__time_duration.sum --clock-control none python tools/seg_one.py 20000"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cccl_b200 import _native  # noqa: E402

if __name__ == "__main__":
    mean = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    n = 1 << 24
    rng = np.random.default_rng(7)
    keys = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    lens = rng.integers(0, 2 * mean + 1, size=max(2, 2 * n // mean)).astype(np.int64)
    ends = np.cumsum(lens)
    ends = ends[ends <= n]
    begins = np.concatenate([[0], ends[:-1]]).astype(np.int64)
    dk = torch.from_numpy(keys.view(np.int32)).cuda()
    dv = torch.arange(n, dtype=torch.int32, device="cuda")
    ko, vo = torch.empty_like(dk), torch.empty_like(dv)
    db, de = torch.from_numpy(begins).cuda(), torch.from_numpy(ends.astype(np.int64)).cuda()
    lib = _native.lib()
    a = (n, len(begins), db.data_ptr(), de.data_ptr(), 8, 0, 4, 4, 0, 32, 0, torch.cuda.current_stream().cuda_stream)
    nb = ctypes.c_size_t(0)
    _native.check(lib.b200rs_segmented_sort(None, ctypes.byref(nb), None, None, None, None, *a), "query")
    temp = torch.empty(nb.value, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        _native.check(lib.b200rs_segmented_sort(temp.data_ptr(), ctypes.byref(nb), dk.data_ptr(), ko.data_ptr(), dv.data_ptr(),
                                                vo.data_ptr(), *a), "sort")
    torch.cuda.synchronize()
    print("segments", len(begins), "temp", nb.value)
