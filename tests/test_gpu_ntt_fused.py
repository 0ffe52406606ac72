"""
GPU tests of the persistent fused 2^20-point NTT kernel (gfb_ntt_fused.cuh, ``-m gpu``).

The fused kernel serves uint32 batches; 8-byte element types still run the two-launch kernels, which were pinned in
round 1 against the oracle and the live reference (SHA-256 of its 2^20-point outputs).  So the same batch transformed
in both element types must agree bit for bit, at batch sizes that exercise the task queue's edges: fewer transforms
than the pass-2 lag, ring wrap-around (more than 8), uneven tails.  Oracle rows and round trips on top.
"""

import hashlib

import numpy as np
import pytest

import galois_b200 as gb
import oracle

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

N = 2**20
P = 998244353
P_NONLAZY = 2013265921  # 15 * 2^27 + 1 >= 2^30: the non-lazy butterflies


def _rand(GF, batch, seed):
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.randint(0, GF.order, (batch, N), dtype=torch.int64, device="cuda", generator=g)
    return t


def _dev(GF, t, dtype):
    import torch

    tt = t.to(torch.int32 if dtype == np.uint32 else torch.int64).contiguous()
    return gb.DeviceArray(GF, tt.view(torch.uint8).reshape(-1), tuple(t.shape), dtype)


def _as_i64(d):
    import torch

    if d.dtype == np.uint32:
        return d.tensor.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    return d.tensor.view(torch.int64)


@pytest.mark.parametrize("batch", [1, 2, 3, 4, 5, 8, 9, 17, 41])
@pytest.mark.parametrize("p", [P, P_NONLAZY])
def test_fused_equals_two_launch_path(batch, p):
    import torch

    GF = gb.GF(p)
    t = _rand(GF, batch, 100 + batch)
    x32, x64 = _dev(GF, t, np.uint32), _dev(GF, t, np.int64)
    for inverse in (False, True):
        a = gb.ntt_device(x32, inverse=inverse)
        b = gb.ntt_device(x64, inverse=inverse)
        assert torch.equal(_as_i64(a), _as_i64(b)), (batch, p, inverse)
    # round trip, out of place and in place (y aliases x)
    X = gb.ntt_device(x32)
    back = gb.ntt_device(X, inverse=True)
    assert torch.equal(back.tensor, x32.tensor)
    gb.ntt_device(X, inverse=True, out=X)
    assert torch.equal(X.tensor, x32.tensor)


def test_fused_against_oracle_and_reference_digest(live_ntt):
    """Rows of a 12-transform batch against the oracle, and the reference's own 2^20-point output (SHA-256 recorded from
    the live reference by tests/golden/make_golden.py) through the fused kernel inside a batch."""
    import torch

    data, meta = live_ntt
    GF = gb.GF(P)
    F = oracle.Field.prime(P)
    omega = int(GF.primitive_root_of_unity(N))
    t = _rand(GF, 12, 7)
    X = gb.ntt_device(_dev(GF, t, np.uint32))
    got = _as_i64(X).view(12, N)
    for row in (0, 5, 11):
        want = F.fft(t[row].cpu().numpy(), omega)
        assert np.array_equal(got[row].cpu().numpy(), want), row
    # the reference-generated input of the golden digest (GF.Random(2^20, seed=3), live-reference output hashed by
    # tests/golden/make_golden.py), placed in the middle of a batch
    m = meta["p998244353/n1048576/seed3"]
    x = np.asarray(GF.Random(N, seed=3, dtype=np.uint32))
    assert hashlib.sha256(x.tobytes()).hexdigest() == m["x_sha256"]
    t[6] = torch.from_numpy(x.astype(np.int64)).cuda()
    X = gb.ntt_device(_dev(GF, t, np.uint32))
    row = _as_i64(X).view(12, N)[6].cpu().numpy().astype(np.uint32)
    assert hashlib.sha256(row.tobytes()).hexdigest() == m["ntt_sha256"]
    Xi = gb.ntt_device(_dev(GF, t, np.uint32), inverse=True)
    row = _as_i64(Xi).view(12, N)[6].cpu().numpy().astype(np.uint32)
    assert hashlib.sha256(row.tobytes()).hexdigest() == m["intt_sha256"]


def test_config3_batch_4096_on_one_gpu():
    """BASELINE configs[2] at its full size on ONE GPU: 4096 x 2^20 uint32 (17.2 GB in, 17.2 GB out).
    intt(ntt(x)) == x on every row; rows {0, 1, 511, 512, 4095} against the oracle."""
    import torch

    GF = gb.GF(P)
    batch = 4096
    g = torch.Generator(device="cuda").manual_seed(3)
    t = torch.randint(0, P, (batch, N), dtype=torch.int32, device="cuda", generator=g)
    x = gb.DeviceArray(GF, t.view(torch.uint8).reshape(-1), (batch, N), np.uint32)
    X = gb.ntt_device(x)
    F = oracle.Field.prime(P)
    omega = int(GF.primitive_root_of_unity(N))
    Xv = X.tensor.view(torch.int32).view(batch, N)
    for row in (0, 1, 511, 512, 4095):
        want = F.fft(t[row].cpu().numpy().astype(np.int64), omega)
        assert np.array_equal(Xv[row].cpu().numpy().astype(np.int64), want), row
    gb.ntt_device(X, inverse=True, out=X)
    assert torch.equal(X.tensor, x.tensor)
