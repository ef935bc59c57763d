"""Probe: torch symmetric memory on this box -- peer-mapped buffers and the NVLink copy bandwidth they give."""
import os, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 28  # 1 GiB of int32
buf = symm.empty(n, dtype=torch.int32, device=f"cuda:{local}")
hdl = symm.rendezvous(buf, dist.group.WORLD)
print(rank, "rendezvous ok", [hex(p) for p in hdl.buffer_ptrs][:4], flush=True)
src = torch.arange(n, dtype=torch.int32, device="cuda")
peer = (rank + 1) % world
pbuf = hdl.get_buffer(peer, (n,), torch.int32, 0)
hdl.barrier()
for sz in (1 << 22, 1 << 25, 1 << 28):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        pbuf[:sz].copy_(src[:sz])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(rank, f"push {sz*4/1e6:.0f} MB to peer {peer}: {ms:.3f} ms = {sz*4/ms/1e6:.1f} GB/s", flush=True)
hdl.barrier()
ok = bool((buf[:1000].cpu() == torch.arange(1000, dtype=torch.int32)).all())
print(rank, "data from peer correct:", ok, flush=True)
# NCCL all_to_all for comparison
out = torch.empty(n, dtype=torch.int32, device="cuda")
splits = [n // world] * world
for _ in range(2):
    dist.all_to_all_single(out, src, splits, splits)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    dist.all_to_all_single(out, src, splits, splits)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(rank, f"nccl all_to_all 1 GiB/rank: {ms:.3f} ms, off-GPU {n*4*(world-1)/world/ms/1e6:.1f} GB/s per direction", flush=True)
dist.destroy_process_group()
