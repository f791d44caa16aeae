"""Probe (torchrun, N>=2): what the gradient exchange can use on this box.

Prints, per rank 0: NCCL all-reduce time at the two arena sizes of the bench (236 MB / 472 MB),
whether torch symmetric memory rendezvous works, the multicast pointer (NVLS), and the time of
torch's library multimem all-reduce on the same buffer.  Plumbing probe only, nothing is shipped from it.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def timed(fn, iters=10, warm=3):
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
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {"world": world}
    for n_gauss in (1_000_000, 2_000_000):
        n = n_gauss * 59
        buf = torch.zeros(n, device="cuda")
        ms = timed(lambda: dist.all_reduce(buf))
        out["nccl_ms_%dM" % (n_gauss // 1_000_000)] = ms
        out["nccl_busbw_%dM" % (n_gauss // 1_000_000)] = n * 4 / ms / 1e6 * 2 * (world - 1) / world
        del buf
    try:
        import torch.distributed._symmetric_memory as symm
        n = 1_000_000 * 59
        n = (n + 1023) // 1024 * 1024
        t = symm.empty(n, dtype=torch.float32, device=torch.device("cuda", local))
        h = symm.rendezvous(t, dist.group.WORLD.group_name)
        out["symm_ok"] = True
        out["multicast_ptr"] = int(h.multicast_ptr)
        try:
            out["has_multicast"] = bool(symm._SymmetricMemory.has_multicast_support(torch._C._autograd.DeviceType.CUDA, local))
        except Exception as e:  # noqa: BLE001
            out["has_multicast_err"] = repr(e)[:200]
        out["buffer_ptrs"] = [int(p) for p in h.buffer_ptrs]
        out["signal_pad_size"] = int(h.signal_pad_size)
        t.fill_(1.0)
        h.barrier()
        for name in ("multimem_all_reduce_", "one_shot_all_reduce", "two_shot_all_reduce_"):
            try:
                op = getattr(torch.ops.symm_mem, name)
                ms = timed(lambda: op(t, "sum", dist.group.WORLD.group_name))
                out[name + "_ms"] = ms
            except Exception as e:  # noqa: BLE001
                out[name + "_err"] = repr(e)[:200]
    except Exception as e:  # noqa: BLE001
        out["symm_ok"] = False
        out["symm_err"] = repr(e)[:400]
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
