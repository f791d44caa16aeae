"""torchrun probe (N >= 2): the chunk-overlapped gradient exchange of the rasterizer backward against the plain path —
same sums (bit for bit: every element is reduced once by the switch in both), and the step time of each.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/overlap_probe.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hidegs_b200 import parallel, synthetic as syn  # noqa: E402
from hidegs_b200.diff_gaussian_rasterization import _C as C  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N = int(os.environ.get("HG_PROBE_N", 1_000_000))
    bench.N_GAUSS = N
    scene = {k: v.to(dev) for k, v in syn.make_scene(N, seed=0).items()}
    cam = bench.camera_for(rank, 1).to(dev)
    am = syn.geometry_all_map(scene["means3D"], scene["scales"], scene["rotations"], cam)
    g = {k: v.to(dev) for k, v in syn.upstream_grads(bench.WIDTH, bench.HEIGHT, seed=1).items()}
    bg = torch.zeros(3, device=dev)
    fa = bench.op_tuple(C, scene, cam, am, dev, bg)
    out = {"world": world, "N": N}

    arena = parallel.SymmetricArena(N * 80, dev)

    def plain():
        fwd = C.rasterize_gaussians(*fa)
        C.rasterize_gaussians_backward(*bench.bwd_tuple(fa, fwd, g), grad_arena=arena.tensor)
        arena.all_reduce_(59 * N)

    results = {}
    for chunks in (0, 2, 4, 8):
        ov = parallel.OverlappedBackwardExchange(arena, N, 16, n_chunks=chunks) if chunks else None

        def step():
            if ov is None:
                plain()
            else:
                fwd = C.rasterize_gaussians(*fa)
                C.rasterize_gaussians_backward(*bench.bwd_tuple(fa, fwd, g), **ov.backward_kwargs())
                ov.finish()
        for _ in range(4):
            step()
        torch.cuda.synchronize()
        results[chunks] = arena.tensor[:59 * N].clone()
        dist.barrier()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(20):
            step()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["ms_chunks_%d" % chunks] = round(float(t), 4)
    ref = results[0]
    # the blend backward accumulates with float atomics, so two runs differ in the last bits; compare relative L2
    for c in (2, 4, 8):
        out["rel_l2_vs_plain_%d" % c] = float((results[c] - ref).norm() / ref.norm())
    # exactness of the exchange itself: reduce a fixed arena both ways
    fixed = torch.randn(59 * N, device=dev, generator=torch.Generator(device=dev).manual_seed(5 + rank))
    arena.tensor[:59 * N].copy_(fixed)
    arena.all_reduce_(59 * N)
    whole = arena.tensor[:59 * N].clone()
    arena.tensor[:59 * N].copy_(fixed)
    torch.cuda.synchronize()
    dist.barrier()
    ov = parallel.OverlappedBackwardExchange(arena, N, 16, n_chunks=4)
    per = ((N + 3) // 4 + 127) // 128 * 128
    for c, p0 in enumerate(range(0, N, per)):
        ov._on_chunk(c, p0, min(N, p0 + per))
    ov.finish()
    torch.cuda.synchronize()
    out["ranges_equal_whole"] = bool(torch.equal(arena.tensor[:59 * N], whole))
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
