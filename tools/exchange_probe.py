"""torchrun probe (N >= 2): the in-fabric gradient exchange (hg_nvls_allreduce_f32) against NCCL — values and time.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/exchange_probe.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hidegs_b200 import parallel  # noqa: E402


def timed(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world, "nvls": parallel.nvls_available(dev)}
    n = 1_000_000 * 59 + 3  # odd tail on purpose
    arena = parallel.SymmetricArena(n, dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for trial in range(3):
        x = torch.randn(n, device=dev, generator=g)
        want = x.clone()
        dist.all_reduce(want)
        arena.tensor.copy_(x)
        arena.all_reduce_()
        torch.cuda.synchronize()
        err = float((arena.tensor - want).abs().max())
        out["max_abs_err_%d" % trial] = err
        # every rank must hold the SAME bits (the sum is formed once, in the switch, and broadcast)
        mine = arena.tensor.view(torch.int32).to(torch.int64).sum()
        lo, hi = mine.clone(), mine.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out["replicas_identical_%d" % trial] = bool(lo.item() == hi.item())
    plain = torch.zeros(n, device=dev)
    out["nccl_ms"] = timed(lambda: dist.all_reduce(plain))
    for blocks in (16, 32, 48, 74, 148, 296):
        arena.blocks = blocks
        out["nvls_ms_b%d" % blocks] = timed(lambda: arena.all_reduce_())
    big = parallel.SymmetricArena(2 * n, dev)
    plain2 = torch.zeros(2 * n, device=dev)
    out["nccl_ms_2M"] = timed(lambda: dist.all_reduce(plain2))
    out["nvls_ms_2M"] = timed(lambda: big.all_reduce_())
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
