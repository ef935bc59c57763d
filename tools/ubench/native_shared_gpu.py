"""Experiment: can b200rs_sort_multi run with TWO processes sharing ONE GPU (time-sliced contexts)?  The kernels of one
process spin on flags the other process's kernels set, so this only works if the driver time-slices between the two
contexts.  Run under `timeout`; prints the wall time of a few sorts.  python tools/ubench/native_shared_gpu.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def worker(rank, world, port):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cccl_b200.multi_gpu import distributed_sort

    g = torch.Generator(device="cuda").manual_seed(rank)
    for n in (1000, 100_000, 1_000_000):
        k = torch.randint(-(2**31), 2**31 - 1, (n,), dtype=torch.int32, device="cuda", generator=g)
        v = torch.arange(n, dtype=torch.int32, device="cuda") + rank * n
        t0 = time.time()
        st = {}
        ok, ov = distributed_sort(k, v, protocol="native", stats=st)
        torch.cuda.synchronize()
        print(f"rank {rank} n {n}: {time.time() - t0:.3f} s status {st.get('status')} sorted "
              f"{bool((ok[1:] >= ok[:-1]).all())}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    mp.spawn(worker, args=(2, 29533), nprocs=2, join=True)
