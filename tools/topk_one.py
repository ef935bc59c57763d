"""One b200rs_topk call (2^log2n u32 keys, K = 2^log2k) for profiling: ncu --metrics gpu__time_duration.sum python tools/topk_one.py 28 11"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import topk_bench  # noqa: E402

if __name__ == "__main__":
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    log2k = int(sys.argv[2]) if len(sys.argv) > 2 else 11
    print(topk_bench.ours(log2n, log2k, 1, iters=1))
