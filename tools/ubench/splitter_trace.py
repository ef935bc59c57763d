"""Diagnostic: where the splitter-selection phase spends its time (host wall clock with device syncs between steps)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from cccl_b200 import multi_gpu as mg

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 28
g = torch.Generator(device="cuda").manual_seed(42 + rank)
keys = torch.randint(-(2**31), 2**31 - 1, (n,), dtype=torch.int32, device="cuda", generator=g).view(torch.uint32)
ops = mg._default_ops()
targets = np.cumsum(np.full(world, n, dtype=np.int64))[:-1]
nt = len(targets)

def sync():
    torch.cuda.synchronize()
    return time.perf_counter()

for it in range(4):
    dist.barrier(); t = [sync()]
    marks = []
    def mark(name):
        t.append(sync()); marks.append((name, (t[-1] - t[-2]) * 1e3))
    tgt = torch.from_numpy(np.maximum(targets, 1)).cuda()
    prefix = torch.zeros(nt, dtype=torch.int64, device="cuda")
    mark("setup")
    for rnd in range(4):
        if rnd == 0:
            h = ops.top_digit_histogram(keys, False).expand(nt, 256)
        else:
            h = ops.select_histogram(keys, prefix, rnd, False, candidates="emit" if rnd == 1 else "use")
        mark(f"r{rnd} kernel")
        hg = h.contiguous().clone()
        dist.all_reduce(hg)
        mark(f"r{rnd} allreduce")
        cum = torch.cumsum(hg, 1)
        b = torch.searchsorted(cum, (tgt).unsqueeze(1)).squeeze(1).clamp_(max=255)
        prefix = prefix * 256 + b
        mark(f"r{rnd} pick ops")
    if rank == 0 and it == 3:
        print(" | ".join(f"{k} {v:.3f}" for k, v in marks), flush=True)
    # whole function, unsynchronised inside
    dist.barrier(); t0 = sync()
    mg.select_splitters_unsorted(keys, targets, kind=0, key_bytes=4, descending=False, ops=ops, group=None, dist=dist)
    t1 = sync()
    if rank == 0 and it == 3:
        print(f"select_splitters_unsorted total {1e3*(t1-t0):.3f} ms", flush=True)
dist.destroy_process_group()
