"""
GPU tests of the single large NTT sharded over ranks (galois_b200._dist.ShardedNtt, ``-m gpu``).

world = 1 runs the whole pipeline (column pass, twiddle, row pass) on one GPU and must equal the direct transform of
the same vector (which round 1 pinned to the oracle / live reference); with at least two GPUs the test relaunches itself
under torchrun with 2 ranks (NCCL all-to-all between the passes) and every rank checks its output slab against the direct
transform computed locally.
"""

import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

P = 998244353
ROOT = Path(__file__).resolve().parent.parent


def _check(log_n: int, inverse: bool):
    import torch

    import galois_b200 as gb
    from galois_b200._dist import ShardedNtt

    GF = gb.GF(P)
    n = 1 << log_n
    sh = ShardedNtt(GF, n)
    g = torch.Generator(device="cuda").manual_seed(11 + log_n)
    x = torch.randint(0, P, (n,), dtype=torch.int32, device="cuda", generator=g)
    full = gb.DeviceArray(GF, x.view(torch.uint8).reshape(-1), (n,), np.uint32)
    direct = gb.ntt_device(full, inverse=inverse).tensor.view(torch.int32)
    slab = sh.input_slab(x)
    local = gb.DeviceArray(GF, slab.view(torch.uint8).reshape(-1), tuple(slab.shape), np.uint32)
    out = sh(local, inverse=inverse).tensor.view(torch.int32).view(sh.rows, sh.n2)
    want = direct[sh.output_indices().reshape(-1)].view(sh.rows, sh.n2)
    assert torch.equal(out, want), (log_n, inverse, sh.rank)
    return sh


@pytest.mark.parametrize("log_n", [12, 16, 20, 22, 23])
def test_single_rank_equals_direct_transform(log_n):
    for inverse in (False, True):
        _check(log_n, inverse)


def test_oracle_spot_check():
    import torch

    import galois_b200 as gb
    import oracle
    from galois_b200._dist import ShardedNtt

    GF = gb.GF(P)
    n = 1 << 18
    sh = ShardedNtt(GF, n)
    x = np.random.default_rng(3).integers(0, P, n, dtype=np.int64)
    t = torch.from_numpy(x.astype(np.int32)).cuda()
    slab = sh.input_slab(t)
    out = sh(gb.DeviceArray(GF, slab.view(torch.uint8).reshape(-1), tuple(slab.shape), np.uint32))
    got = np.empty(n, dtype=np.int64)
    got[sh.output_indices().reshape(-1).cpu().numpy()] = out.tensor.view(torch.int32).cpu().numpy().astype(np.int64)
    want = oracle.Field.prime(P).fft(x, int(GF.primitive_root_of_unity(n)))
    assert np.array_equal(got, want)


def test_two_ranks_nccl():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", str(Path(__file__)), "worker"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "worker":
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist

    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for log_n in (16, 22, 23):  # 998244353 - 1 = 2^23 * 7 * 17: 2^23 is the largest power-of-two size
        for inverse in (False, True):
            sh = _check(log_n, inverse)
    if dist.get_rank() == 0:
        print("sharded NTT ok", sh.timings)
    dist.barrier()
    dist.destroy_process_group()
