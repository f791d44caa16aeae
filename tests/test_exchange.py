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
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "exchange_probe.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    if not out["nvls"]:
        pytest.skip("no multicast (NVLS) support on this box")
    for i in range(3):  # same bits as NCCL's sum at world 2 (two addends: the order cannot matter), identical replicas
        assert out["max_abs_err_%d" % i] == 0.0 and out["replicas_identical_%d" % i]
