"""CPU (gloo, world_size 2) tests of the view-sharded data-parallel plumbing."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hidegs_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_arena(n, seed, shared=True):
    """Gradient tensors laid out like the backward's arena (xyz 3 | sh 48 | opacity 1 | scale 3 | rot 4)."""
    g = torch.Generator().manual_seed(seed)
    widths = (3, 48, 1, 3, 4)
    if shared:
        arena = torch.randn(n * 76, generator=g)
        out, off = [], 0
        for w in widths:
            out.append(arena[off:off + n * w].view(n, w))
            off += n * w
        return arena, out
    return None, [torch.randn(n, w, generator=g) for w in widths]


def _worker(rank, world, port, shared, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 257
        arena, grads = _make_arena(n, seed=100 + rank, shared=shared)
        tail_before = arena[59 * n:].clone() if shared else None
        flat = parallel.flat_view(grads)
        assert (flat is not None) == shared
        if shared:
            assert flat.numel() == 59 * n and flat.data_ptr() == arena.data_ptr()
        _, nbytes = parallel.allreduce_gradients(grads, average=False)
        assert nbytes == 59 * n * 4
        expect = [sum(_make_arena(n, seed=100 + r, shared=shared)[1][i] for r in range(world)) for i in range(5)]
        for a, b in zip(grads, expect):
            assert torch.allclose(a, b, rtol=0, atol=1e-6)
        if shared:  # nothing outside the 59-float parameter block is touched
            assert torch.equal(arena[59 * n:], tail_before)
        views = list(range(10))
        mine = parallel.shard_views(views)
        assert mine == views[rank::world]
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _run(shared):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, shared, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_allreduce_in_place_on_shared_arena():
    _run(shared=True)


def test_allreduce_packs_when_storage_is_not_shared():
    _run(shared=False)


def test_flat_view_detects_gaps_and_single_process_path():
    arena, grads = _make_arena(16, 1)
    assert parallel.flat_view(grads).numel() == 59 * 16
    assert parallel.flat_view([grads[0], grads[2]]) is None  # not back to back
    assert parallel.flat_view([grads[1].t()]) is None  # not contiguous
    before = [g.clone() for g in grads]
    work, nbytes = parallel.allreduce_gradients(grads)  # no process group: a no-op
    assert work is None and nbytes == 59 * 16 * 4 and all(torch.equal(a, b) for a, b in zip(grads, before))


def _factored_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N = 64
        fx = parallel.FactoredExchange(N, "cpu")
        assert fx.symmetric is None and fx.tensor.numel() == 59 * N and fx.block == 3 * N + 4
        assert fx.factor_offset % 4 == 0 and fx.factors.numel() == world * fx.block
        kw = fx.backward_kwargs()
        assert kw["grad_arena"].data_ptr() == fx.buffer.data_ptr() and kw["sh_factor"].data_ptr() == fx.my_factors.data_ptr()

        def fill(r):
            g = torch.Generator().manual_seed(300 + r)
            return torch.randn(59 * N, generator=g), torch.randn(fx.block, generator=g)
        a, f = fill(rank)
        fx.tensor.copy_(a)
        fx.my_factors.copy_(f)
        sh_before = fx.tensor[3 * N:51 * N].clone()
        fx.exchange()
        every = [fill(r) for r in range(world)]
        total = sum(e[0] for e in every)
        assert torch.allclose(fx.tensor[:3 * N], total[:3 * N], atol=1e-6)            # xyz summed
        assert torch.allclose(fx.tensor[51 * N:], total[51 * N:], atol=1e-6)          # opacity | scale | rotation summed
        assert torch.equal(fx.tensor[3 * N:51 * N], sh_before)                        # the SH block did not travel
        for r in range(world):                                                        # every rank's factors gathered
            assert torch.equal(fx.factors[r * fx.block:(r + 1) * fx.block], every[r][1])
        assert fx.bytes == 4 * (11 * N + fx.block)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_factored_exchange_moves_only_the_non_sh_blocks_and_the_factors():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_factored_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_pin_host_to_gpu_node_never_raises_without_a_gpu():
    """The placement helper is an optimisation: without NVML / a GPU it reports why and leaves the affinity alone."""
    import os
    from hidegs_b200.parallel import pin_host_to_gpu_node
    before = os.sched_getaffinity(0)
    info = pin_host_to_gpu_node("cuda:0")
    assert isinstance(info, dict) and "pinned" in info
    if not info["pinned"]:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
