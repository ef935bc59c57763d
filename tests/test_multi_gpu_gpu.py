"""Multi-GPU sort on the GPU box: the CUDA primitives (b200rs_sort / b200rs_splitter_ranks through the C ABI) under the
distributed protocol.  Two or three ranks share cuda:0 with a gloo group (collectives staged through the host), so the
whole path is exercised on a one-GPU box; when the box has >= 2 GPUs the same cases also run over NCCL, one GPU per
rank.  Oracle = one stable CPU sort of the rank-order concatenation."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from gen import make_keys  # noqa: E402
from oracle_lib import oracle_sort  # noqa: E402

pytestmark = pytest.mark.gpu

CASES = [
    ("uniform_u32", np.uint32, [300_000, 200_001, 150_000], "uniform", False, True),
    ("entropy5_u64_desc", np.uint64, [100_000, 140_000, 60_000], "entropy5", True, True),
    ("equal_u32", np.uint32, [50_000, 70_001, 10], "equal", False, True),
    ("few16_i32", np.int32, [80_000, 0, 90_000], "few16", False, True),
    ("f32_zeros_desc", np.float32, [60_000, 50_000, 40_000], "uniform", True, True),
    ("i64_keys_only", np.int64, [70_000, 30_000, 1], "uniform", False, False),
    ("u8", np.uint8, [40_000, 40_000, 40_000], "uniform", False, True),
]


def _dev_tensor(a, device):
    from test_multi_gpu_host import TDT

    if a.size == 0:
        return torch.empty(0, dtype=TDT[a.dtype], device=device)
    return torch.from_numpy(a.view(np.uint8).copy()).to(device).view(TDT[a.dtype])


def _host(t):
    from cccl_b200.radix_sort import _torch_np_dtype

    dt = _torch_np_dtype(t)
    return t.view(torch.uint8).cpu().numpy().view(dt).copy() if t.numel() else np.empty(0, dtype=dt)


def _worker(rank, world, port, backend, results, protocol="partition", exchange="auto"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    device = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(device)
    dist.init_process_group(backend, rank=rank, world_size=world)
    from cccl_b200.multi_gpu import distributed_sort

    failures = []
    for name, dtype, sizes, dist_name, desc, with_vals in CASES:
        if protocol == "native" and np.dtype(dtype).itemsize < 4:
            continue  # b200rs_sort_multi: 4- and 8-byte keys (the Python protocols cover the narrow ones)
        ns = sizes[:world]
        shards, vshards, base = [], [], 0
        for r, n in enumerate(ns):
            k = make_keys(dist_name, n, dtype, seed=13 * r + 5)
            if np.dtype(dtype).kind == "f" and n:
                k[::17] = -0.0
                k[::19] = 0.0
            shards.append(k)
            vshards.append(np.arange(base, base + n, dtype=np.uint32))
            base += n
        d_k = _dev_tensor(shards[rank], device)
        d_v = _dev_tensor(vshards[rank], device) if with_vals else None
        stats = {}
        ok, ov = distributed_sort(d_k, d_v, descending=desc, stats=stats, protocol=protocol, exchange=exchange)
        torch.cuda.synchronize()
        if protocol == "native" and stats.get("status") != 0:
            failures.append((name, f"device-side status {stats.get('status')}"))
        if backend == "nccl" and sum(ns) > 0 and protocol == "partition":
            want = {"auto": ("fused" if np.dtype(dtype).itemsize >= 4 else "peer"), "peer": "peer",
                    "collective": "collective"}[exchange]
            if stats.get("exchange") != want:
                failures.append((name, f"exchange {stats.get('exchange')} != {want}: "
                                       + str(stats.get("peer_exchange_unavailable"))))
        if not np.array_equal(_host(d_k).view(np.uint8), shards[rank].view(np.uint8)):
            failures.append((name, "input shard modified"))
        allk, allv = np.concatenate(shards), np.concatenate(vshards)
        ek, ev = oracle_sort(allk, allv, descending=desc) if with_vals else (oracle_sort(allk, descending=desc), None)
        lo = int(np.sum(ns[:rank]))
        hi = lo + ns[rank]
        gk = _host(ok)
        if gk.shape[0] != ns[rank] or not np.array_equal(gk.view(np.uint8), ek[lo:hi].view(np.uint8)):
            failures.append((name, "keys"))
        if with_vals and not np.array_equal(_host(ov), ev[lo:hi]):
            failures.append((name, "values"))
    results[rank] = failures
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, backend, protocol="partition", exchange="auto"):
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), backend, results, protocol, exchange), nprocs=world, join=True)
    for r in range(world):
        assert results[r] == [], f"rank {r}: {results[r]}"


@pytest.mark.parametrize("protocol", ["partition", "sort"])
@pytest.mark.parametrize("world", [2, 3])
def test_distributed_sort_cuda_ranks_sharing_one_gpu(world, protocol):
    _run(world, "gloo", protocol)


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_sort_native_cxx_host_ranks_sharing_one_gpu(world):
    """b200rs_sort_multi with the ranks as processes on ONE GPU (CUDA IPC between contexts of the same device; the
    kernels of one rank wait for flags the other ranks' kernels set, which the driver's time slicing lets happen), so
    the C++ host, the device-side all-reduce and the fused exchange are exercised on a one-GPU box too."""
    _run(world, "gloo", "native")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_distributed_sort_native_cxx_host():
    """b200rs_sort_multi (C++ host, kernels only: device-side all-reduce over peer memory, fused exchange)."""
    _run(min(torch.cuda.device_count(), 3), "nccl", "native")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("protocol,exchange", [("partition", "auto"), ("partition", "peer"), ("partition", "collective"),
                                               ("sort", "collective")])
def test_distributed_sort_nccl(protocol, exchange):
    _run(min(torch.cuda.device_count(), 3), "nccl", protocol, exchange)
