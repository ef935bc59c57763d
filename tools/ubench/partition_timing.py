"""Device time of the multi-GPU support kernels on one GPU: 2^28 uniform u32 keys (+ u32 values)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cccl_b200.multi_gpu import CudaOps

ops = CudaOps()
n = 1 << 28
g = torch.Generator(device="cuda").manual_seed(1)
k = torch.randint(-(2**31), 2**31 - 1, (n,), dtype=torch.int32, device="cuda", generator=g).view(torch.uint32)
v = torch.arange(n, dtype=torch.int32, device="cuda")

def timeit(label, fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{label:50s} {e0.elapsed_time(e1)/reps:8.3f} ms", flush=True)

timeit("top_digit_histogram", lambda: ops.top_digit_histogram(k, False))
for m in (1, 3, 7):
    for rnd in (1, 2, 3):
        pref = torch.arange(1, m + 1, dtype=torch.int64, device="cuda") * (37 if rnd == 1 else 37 * 256**(rnd-1) + 5)
        timeit(f"select_histogram prefixes={m} round={rnd}", lambda: ops.select_histogram(k, pref, rnd, False))
for m in (1, 3, 7):
    sp = (np.arange(1, m + 1, dtype=np.uint64) * np.uint64(2**32 // (m + 1)))
    ids = ops.bucket_ids(k, sp, False)
    sizes = torch.bincount(ids.to(torch.int64), minlength=2 * m + 1).cpu().numpy()
    timeit(f"bucket_ids splitters={m}", lambda: ops.bucket_ids(k, sp, False))
    timeit(f"partition_by_splitters pairs splitters={m}", lambda: ops.partition_by_splitters(k, v, sp, sizes, False))
    timeit(f"partition_by_splitters keys  splitters={m}", lambda: ops.partition_by_splitters(k, None, sp, sizes, False))
eq = torch.full((n,), 7, dtype=torch.int32, device="cuda").view(torch.uint32)
timeit("select_histogram all-equal keys, round 1 hit", lambda: ops.select_histogram(eq, torch.zeros(1, dtype=torch.int64, device="cuda"), 1, False))
