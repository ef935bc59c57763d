"""Where does the one-launch cooperative kernel (csrc/small.cu) stop winning against the general multi-kernel path?
python tools/ubench/small_crossover.py  -> one line per size: ms of both paths (device time, CUDA events)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cccl_b200 import _native  # noqa: E402

lib = _native.lib()


def run(n, kb, vb, iters):
    g = torch.Generator(device="cuda").manual_seed(1)
    keys = torch.randint(-(2**63), 2**63 - 1, (max(1, n * kb // 8),), dtype=torch.int64, device="cuda", generator=g)
    vals = torch.arange(max(1, n * vb // 8), dtype=torch.int64, device="cuda") if vb else None
    ko, vo = torch.empty_like(keys), (torch.empty_like(vals) if vb else None)
    p = lambda t: t.data_ptr() if t is not None else 0
    st = torch.cuda.current_stream().cuda_stream
    need, _ = _native.sort_raw(0, 0, p(keys), p(ko), p(vals), p(vo), n, 0, kb, vb, 0, kb * 8, False, False, st)
    temp = torch.empty(need, dtype=torch.uint8, device="cuda")
    step = lambda: _native.sort_raw(temp.data_ptr(), need, p(keys), p(ko), p(vals), p(vo), n, 0, kb, vb, 0, kb * 8,
                                    False, False, st)
    for _ in range(5):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, lib.b200rs_last_launch_count()


for kb, vb in ((4, 0), (4, 4), (8, 0), (8, 4), (8, 8)):
    for log2n in range(13, 24):
        n = 1 << log2n
        iters = 200 if log2n < 20 else 50
        lib.b200rs_set_small_max(1 << 30)
        a, la = run(n, kb, vb, iters)
        lib.b200rs_set_small_max(0)
        b, lb = run(n, kb, vb, iters)
        print(f"k{kb}v{vb} 2^{log2n}: fused {a * 1000:8.1f} us ({la} launch)   general {b * 1000:8.1f} us ({lb} ops)   "
              f"ratio {b / a:5.2f}", flush=True)
lib.b200rs_set_small_max(2**64 - 1)
