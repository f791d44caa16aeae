"""The per-step gradient exchange (include/hidegs_exchange.h): argument checks on the CPU, and — where the box has at
least two GPUs behind an NVSwitch — the in-fabric all-reduce kernel against NCCL on identical inputs."""
import json
import os
import subprocess
import sys

import pytest
import torch

from hidegs_b200 import _lib, parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_argument_checks_without_a_gpu():
    L = parallel._exchange_lib()
    assert L.hg_nvls_flag_words(8, 1024) == 8 * 1024 and L.hg_nvls_flag_words(0, 4) == 0
    assert L.hg_nvls_allreduce_f32(None, None, None, 0, 1, 1 << 20, 0, None) == 0      # world 1: nothing to exchange
    assert L.hg_nvls_allreduce_f32(256, 256, 256, 0, 2, 0, 0, None) == 0               # empty arena
    assert L.hg_nvls_allreduce_f32(256, 256, 256, 2, 2, 16, 0, None) == 1              # rank out of range
    assert b"bad argument" in L.hg_last_error()
    assert L.hg_nvls_allreduce_f32(None, 256, 256, 0, 2, 16, 0, None) == 1
    assert b"mandatory" in L.hg_last_error()
    assert L.hg_nvls_allreduce_f32(264, 256, 256, 0, 2, 16, 0, None) == 1              # multicast pointer not 16-byte aligned
    assert b"aligned" in L.hg_last_error()


def test_nvls_is_not_claimed_without_a_process_group():
    assert parallel.nvls_available() is False
    t, arena = parallel.make_exchange_arena(100, "cpu")
    assert arena is None and t.shape == (100,) and float(t.abs().sum()) == 0.0


@pytest.mark.gpu
def test_in_fabric_allreduce_matches_nccl(cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs behind an NVSwitch")
    world = torch.cuda.device_count()  # every GPU of the box (2, 4 or 8 ranks)
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "exchange_probe.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    if not out["nvls"]:
        pytest.skip("no multicast (NVLS) support on this box")
    assert out["world"] == world
    for i in range(3):
        # identical replicas at any world size; same bits as NCCL's sum at world 2 (two addends: the order cannot
        # matter), within fp32 summation-order noise of it (59M standard-normal values, |sum| <~ 6 sqrt(world)) beyond
        assert out["replicas_identical_%d" % i]
        assert out["max_abs_err_%d" % i] <= (0.0 if world == 2 else 1e-5), out


def test_overlap_helper_ranges_cover_the_soa_arena_once():
    """OverlappedBackwardExchange: a slot range maps to five ranges of the SoA arena (xyz | sh | opacity | scale |
    rotation); over all chunks every float of the 59 N arena is exchanged exactly once, in multiples of 4 floats."""
    import numpy as np

    class FakeArena:
        class _B:
            device = torch.device("cpu")
        _buf = _B()

        def __init__(self):
            self.calls = []

        def all_reduce_ranges_(self, offs, cnts):
            self.calls.append((list(offs), list(cnts)))
            return 4 * sum(cnts)

    N = 1000
    ov = parallel.OverlappedBackwardExchange.__new__(parallel.OverlappedBackwardExchange)
    ov.arena, ov.N, ov.n_chunks, ov.widths, ov.bytes = FakeArena(), N, 4, (3, 48, 1, 3, 4), 0
    seen = np.zeros(59 * N, np.int32)
    for p0 in range(0, N, 256):  # chunk boundaries as hg_raster_backward_chunked makes them (multiples of 128 slots)
        p1 = min(N, p0 + 256)
        offs, cnts, base = [], [], 0
        for w in ov.widths:
            offs.append(base + w * p0)
            cnts.append(w * (p1 - p0))
            base += w * N
        ov.bytes += ov.arena.all_reduce_ranges_(offs, cnts)
        for o, c in zip(offs, cnts):
            assert o % 4 == 0 and c % 4 == 0
            seen[o:o + c] += 1
    assert (seen == 1).all() and ov.bytes == 4 * 59 * N
