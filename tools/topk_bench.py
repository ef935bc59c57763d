"""Top-k timing on one B200: b200rs_topk against the unmodified cub::DeviceTopK (oracle/_ref/ref_cub_topk) on the same GPU.
Axes follow the reference's bench (cub/benchmarks/bench/topk/keys.cu:109-111): Elements 2^16..2^28, SelectedElements
2^3..2^23, entropy.  Prints one JSON line per point; algorithmic bytes = one read of the keys + K outputs."""
import ctypes
import json
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cccl_b200 import _native  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_cub_topk")


def ours(log2n, log2k, and_rounds, iters=20):
    n, k = 1 << log2n, 1 << log2k
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    keys = torch.randint(-(2**31), 2**31, (n,), dtype=torch.int32, device="cuda", generator=g)
    for _ in range(and_rounds - 1):
        keys &= torch.randint(-(2**31), 2**31, (n,), dtype=torch.int32, device="cuda", generator=g)
    if and_rounds == 0:
        keys.fill_(4)
    out = torch.empty(k, dtype=torch.int32, device="cuda")
    lib = _native.lib()
    need = ctypes.c_size_t(0)
    st = torch.cuda.current_stream().cuda_stream
    args = (keys.data_ptr(), out.data_ptr(), 0, 0, n, k, 0, 4, 0, 1, st)
    _native.check(lib.b200rs_topk(0, ctypes.byref(need), *args), "size")
    temp = torch.empty(need.value + 256, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        _native.check(lib.b200rs_topk(temp.data_ptr(), ctypes.byref(need), *args), "topk")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        lib.b200rs_topk(temp.data_ptr(), ctypes.byref(need), *args)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    return {"ms": ms, "gkeys_per_s": n / ms / 1e6, "read_once_gbs": n * 4 / ms / 1e6, "temp_bytes": need.value}


def cub(log2n, log2k, and_rounds, iters=20):
    if not os.path.exists(BIN):
        return None
    r = subprocess.run([BIN, "bench", "u32", str(log2n), str(log2k), "1", str(and_rounds), str(iters)], capture_output=True,
                       text=True, timeout=600)
    return json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stderr[-300:]}


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--small-sweep":
        # one-launch kernel vs the multi-kernel select around the crossover (ours only)
        for log2n in (20, 21, 22, 23, 24, 25, 26):
            for log2k in (11, 19):
                row = {"workload": f"topk_u32_2^{log2n}_k2^{log2k}"}
                for name, mx in (("general", 0), ("one_launch", 1 << 40)):
                    _native.lib().b200rs_set_topk_small_max(mx)
                    row[name + "_ms"] = ours(log2n, log2k, 1)["ms"]
                print(json.dumps(row), flush=True)
        sys.exit(0)
    for log2n in (16, 20, 22, 24, 28):
        for log2k in (3, 11, 19, 23):
            if log2k >= log2n:
                continue
            for rounds in (1, 5):
                o, c = ours(log2n, log2k, rounds), cub(log2n, log2k, rounds)
                print(json.dumps({"workload": f"topk_u32_2^{log2n}_k2^{log2k}_and{rounds}", "b200rs": o, "cub": c,
                                  "speedup": (c["ms"] / o["ms"]) if c and "ms" in c else None}), flush=True)
