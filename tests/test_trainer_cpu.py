"""CPU / gloo, world size 2: the view-sharded step's host logic — views[rank::world], one SUM all-reduce of the
gradient arena, update scaled by 1 / total views — gives every rank the parameters a single process gets from all
views.  The CUDA pieces (render, losses, fused Adam) are replaced by a synthetic differentiable per-view loss and a
plain-torch update; their own parity is covered by the GPU tests."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _scene(n=64):
    g = torch.Generator().manual_seed(0)
    return dict(means3D=torch.randn(n, 3, generator=g), shs=torch.randn(n, 16, 3, generator=g),
                opacity=torch.rand(n, 1, generator=g) * 0.8 + 0.1, scales=torch.rand(n, 3, generator=g) * 0.1 + 0.01,
                rotations=torch.nn.functional.normalize(torch.randn(n, 4, generator=g)))


def _make_trainer():
    from hidegs_b200 import trainer as tr
    params = tr.GaussianParams.from_scene(_scene(), torch.device("cpu"))
    t = tr.ViewShardedTrainer.__new__(tr.ViewShardedTrainer)
    t.params, t.bg, t.opt, t.pipe, t.group = params, None, tr.OptimizationParams, tr.PipelineParams, None
    t.iteration = 0
    t.sparse_adam = t.densification_stats = False  # CUDA-only paths
    t.world = dist.get_world_size() if dist.is_initialized() else 1
    t.rank = dist.get_rank() if dist.is_initialized() else 0

    def view_loss(cam, gt, iteration, gt_ready=None):  # cam: a seed-like float, gt: a target vector
        p = params
        val = ((p.get_xyz * cam).sum(1) + p.get_opacity[:, 0] * p.get_scaling.sum(1) + p.get_features.mean((1, 2))
               + (p.get_rotation * gt).sum(1))
        return (val ** 2).mean(), None

    class _Sgd:
        def step(self, grad_scale=1.0, visible_mask=None):
            with torch.no_grad():
                params.param_arena.add_(params.grad_arena, alpha=-0.1 * grad_scale)

    t.view_loss, t.adam = view_loss, _Sgd()
    return t


def _views():
    g = torch.Generator().manual_seed(1)
    return [(float(i + 1) * 0.3, torch.randn(4, generator=g)) for i in range(4)]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hidegs_b200 import parallel
    t = _make_trainer()
    mine = parallel.shard_views(_views())
    for _ in range(3):
        t.step(mine, total_views=4)
    out[rank] = t.params.param_arena.clone()
    dist.destroy_process_group()


def test_two_ranks_equal_one_process():
    single = _make_trainer()
    for _ in range(3):
        single.step(_views())
    want = single.params.param_arena.clone()
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    for r in range(2):
        assert torch.allclose(out[r], want, rtol=1e-5, atol=1e-7), r
    assert torch.equal(out[0], out[1])


def test_params_and_grads_alias_their_arenas():
    from hidegs_b200 import trainer as tr
    p = tr.GaussianParams.from_scene(_scene(16), torch.device("cpu"))
    assert p.param_arena.numel() == 16 * 59 and p.grad_arena.numel() == 16 * 59
    off = 0
    for name, w in tr.GROUPS:
        leaf = p.leaves[name]
        assert leaf.data_ptr() == p.param_arena[off:].data_ptr() and leaf.grad.data_ptr() == p.grad_arena[off:].data_ptr()
        off += 16 * w
    (p.get_xyz.sum() + p.get_features.sum() * 2).backward()
    assert torch.all(p.grad_arena[:48] == 1) and torch.all(p.grad_arena[48:48 + 16 * 48] == 2)
    p.zero_grad()
    assert not p.grad_arena.any() and p.leaves["xyz"].grad.data_ptr() == p.grad_arena.data_ptr()


def test_morton_order_is_a_locality_preserving_permutation():
    from hidegs_b200.trainer import morton_order
    g = torch.Generator().manual_seed(0)
    xyz = torch.rand(4000, 3, generator=g) * torch.tensor([400.0, 224.0, 30.0])
    order = morton_order(xyz)
    assert sorted(order.tolist()) == list(range(4000))
    step_sorted = (xyz[order][1:] - xyz[order][:-1]).norm(dim=1).mean()
    step_raw = (xyz[1:] - xyz[:-1]).norm(dim=1).mean()
    assert float(step_sorted) < 0.25 * float(step_raw)
    assert torch.equal(morton_order(xyz), order)  # deterministic (stable sort)


def test_balance_views_equal_counts_and_near_equal_cost():
    from hidegs_b200.trainer import balance_views
    import random
    rnd = random.Random(0)
    for world, n in ((2, 8), (8, 64), (4, 10), (3, 9)):
        costs = [rnd.randint(1, 1000) for _ in range(n)]
        shards = balance_views(costs, world)
        assert sorted(i for s in shards for i in s) == list(range(n))
        sizes = [len(s) for s in shards]
        assert max(sizes) - min(sizes) <= 1
        loads = [sum(costs[i] for i in s) for s in shards]
        naive = [sum(costs[i] for i in range(n)[r::world]) for r in range(world)]
        assert max(loads) <= max(naive)
        assert balance_views(costs, world) == shards  # deterministic: every rank computes the same table
        # ranks that run at different paces: the slowest rank's TIME (speed x cost) is what gets levelled
        speeds = [1.0 + 0.15 * (r % 2) for r in range(world)]
        sh2 = balance_views(costs, world, speeds)
        assert sorted(i for s in sh2 for i in s) == list(range(n)) and max(map(len, sh2)) - min(map(len, sh2)) <= 1
        t_aware = max(speeds[r] * sum(costs[i] for i in s) for r, s in enumerate(sh2))
        t_blind = max(speeds[r] * sum(costs[i] for i in s) for r, s in enumerate(shards))
        assert t_aware <= t_blind
