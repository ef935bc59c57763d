#!/usr/bin/env python
"""Runs a few sorts of one workload/config so ncu can capture the kernels: python tools/prof_one.py <workload> <config> [log2n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from cccl_b200 import _native
from cccl_b200.radix_sort import key_kind_of

name = sys.argv[1] if len(sys.argv) > 1 else "sortkeys_u32_2^28_uniform"
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
kdt, vdt, log2n, dist, desc, b, e = bench.WORKLOADS[name]
n = 1 << log2n
kb = np.dtype(kdt).itemsize
vb = np.dtype(vdt).itemsize if vdt else 0
kind = key_kind_of(np.dtype(kdt))
lib = _native.lib()
lib.b200rs_set_config(cfg)
keys, vals = bench.make_device_input(torch, np, name)
keys_out = torch.empty_like(keys)
vals_out = torch.empty_like(vals) if vals is not None else None
p = lambda t: t.data_ptr() if t is not None else 0
st = torch.cuda.current_stream().cuda_stream
need, _ = _native.sort_raw(0, 0, p(keys), p(keys_out), p(vals), p(vals_out), n, kind, kb, vb, b, e, desc, False, st)
temp = torch.empty(need, dtype=torch.uint8, device="cuda")
for _ in range(iters):
    _native.sort_raw(temp.data_ptr(), need, p(keys), p(keys_out), p(vals), p(vals_out), n, kind, kb, vb, b, e, desc,
                     False, st)
torch.cuda.synchronize()
print("done", name, _native.describe_configs(kb, vb)[cfg])
