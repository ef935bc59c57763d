"""Diagnostic: SortPairs (u32,u32) 2^28 reading its input from ordinary vs symmetric (peer-mapped) device memory."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
from cccl_b200 import _native
from cccl_b200.radix_sort import SortOrder, radix_sort, DoubleBuffer

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = _native.lib()
n = 1 << 28
g = torch.Generator(device="cuda").manual_seed(1 + rank)
k = torch.randint(-(2**31), 2**31 - 1, (n,), dtype=torch.int32, device="cuda", generator=g)
v = torch.arange(n, dtype=torch.int32, device="cuda")
sk = symm.empty(n, dtype=torch.int32, device=f"cuda:{local}"); symm.rendezvous(sk, dist.group.WORLD)
sv = symm.empty(n, dtype=torch.int32, device=f"cuda:{local}"); symm.rendezvous(sv, dist.group.WORLD)
sk.copy_(k); sv.copy_(v)
ok, ov = torch.empty_like(k), torch.empty_like(v)

def run(kin, vin, label, db=False):
    for it in range(3):
        torch.cuda.synchronize()
        lib.b200rs_timing_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if db:
            kb, vb = DoubleBuffer(kin, ok), DoubleBuffer(vin, ov)
            t = radix_sort(d_in_keys=kb, d_out_keys=None, d_in_values=vb, d_out_values=None, num_items=n, order=SortOrder.ASCENDING)
        else:
            t = radix_sort(d_in_keys=kin.view(torch.uint32), d_out_keys=ok.view(torch.uint32), d_in_values=vin, d_out_values=ov, num_items=n, order=SortOrder.ASCENDING)
        e1.record(); torch.cuda.synchronize()
        ops = _native.timing_read(); lib.b200rs_timing_enable(0)
    if rank == 0:
        print(label, f"{e0.elapsed_time(e1):.3f} ms", " ".join(f"{o}={t:.3f}" for o, t in ops), flush=True)

run(k, v, "plain  pointer")
run(sk, sv, "symm   pointer")
sk.copy_(k); sv.copy_(v)
run(sk, sv, "symm   doublebuffer", db=True)
k2 = k.clone(); v2 = v.clone()
run(k2, v2, "plain  doublebuffer", db=True)
dist.barrier(); dist.destroy_process_group()
