"""
Multi-GPU plumbing: one process per GPU (``torchrun``), batch-dimension sharding, optional final gather.

The hot path never communicates: element-wise arrays, NTT batches and Reed-Solomon codewords are
independent along their leading dimension, so each rank processes a contiguous shard
(SURVEY.md §8e).  The only collective is the optional gather of the result shards
(``all_gather_into_tensor`` over NCCL on GPUs, gloo in the CPU tests).
"""

from __future__ import annotations

import numpy as np


def shard_bounds(total: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the contiguous, balanced shard owned by ``rank``; the first ``total % world`` ranks get one more."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(total: int, world: int) -> list[int]:
    return [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]


def gather_rows(local, total_rows: int, group=None):
    """
    All-gather row shards ([rows_r, ...] tensors laid out by ``shard_bounds``) into the full [total_rows, ...]
    tensor on every rank.  Uneven shards are padded to the largest shard for the collective and trimmed after.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if world == 1:
        return local
    sizes = shard_sizes(total_rows, world)
    biggest = max(sizes)
    tail = tuple(local.shape[1:])
    if local.shape[0] != biggest:
        padded = torch.zeros((biggest,) + tail, dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    else:
        padded = local.contiguous()
    out = torch.empty((world * biggest,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    if all(s == biggest for s in sizes):
        return out
    parts = [out[r * biggest: r * biggest + sizes[r]] for r in range(world)]
    return torch.cat(parts, dim=0)


# -------------------------------------------------------------------------------------------------
# One large NTT sharded over the GPUs of a node (second half)
# -------------------------------------------------------------------------------------------------
class ShardedNtt:
    """
    A single N = N1 * N2 point NTT over GF(p) spread over the ranks of a process group: the four-step factorisation
    the batched kernels already use inside one GPU (X[k1 + N1 k2] = sum_i2 w2^(i2 k2) w^(i2 k1) sum_i1 x[N2 i1 + i2] w1^(i1 k1),
    w1 = w^N2, w2 = w^N1, w = the reference's primitive N-th root of unity — fft_jit, _domains/_function.py:177-384,
    FieldArray.primitive_root_of_unity _fields/_array.py:1127-1187), with the ONE data exchange the algorithm needs done as
    an NCCL all-to-all between the two local passes:

        rank r owns the column slab  x[:, r*N2/G : (r+1)*N2/G]  of the [N1][N2] row-major view of x   (input layout)
        1. local:  N2/G column transforms of length N1 (batched power-of-two kernels), times w^(i2 k1)
        2. all-to-all: rank r receives the rows k1 in its row slab from every rank
        3. local:  N1/G row transforms of length N2
        rank r then owns X[k1 + N1 k2] for k1 in [r*N1/G, (r+1)*N1/G), all k2, as an [N1/G][N2] array     (output layout)

    Everything on a rank is the existing device API (ntt_device, element-wise multiply); torch supplies the local
    transposes (data movement) and torch.distributed the collective.  ``timings`` holds the device time of the last
    call's three phases in milliseconds.
    """

    def __init__(self, field, n: int, group=None):
        import torch
        import torch.distributed as dist

        from . import _ntt

        self.field, self.n = field, int(n)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        log_n = self.n.bit_length() - 1
        if self.n != 1 << log_n or (field.order - 1) % self.n != 0:
            raise ValueError(f"Argument 'n' must be a power of two dividing {field.order - 1}, not {n}.")
        self.n1 = 1 << (log_n // 2)
        self.n2 = self.n // self.n1
        if self.n1 % self.world or self.n2 % self.world:
            raise ValueError(f"The world size {self.world} must divide both factors {self.n1} x {self.n2}.")
        self.cols = self.n2 // self.world  # columns of the input slab
        self.rows = self.n1 // self.world  # rows of the output slab
        p = field.order
        self.omega = int(field.primitive_root_of_unity(self.n))
        self._ntt = _ntt
        self._torch = torch
        self._tw = {}  # inverse flag -> DeviceArray [cols][n1] of w^(+-(c0 + c) * k1) (times 1/N for the inverse)
        self.timings = {}
        self._p = p

    def _twiddle(self, inverse: bool):
        from ._field import DeviceArray

        tw = self._tw.get(inverse)
        if tw is None:
            torch = self._torch
            f, p = self.field, self._p
            w = pow(self.omega, -1, p) if inverse else self.omega
            c = torch.arange(self.rank * self.cols, (self.rank + 1) * self.cols, dtype=torch.int64, device="cuda")
            k1 = torch.arange(self.n1, dtype=torch.int64, device="cuda")
            e = (c[:, None] * k1[None, :]) % self.n  # exponent of w, [cols][n1]
            base = torch.full((self.cols * self.n1,), w, dtype=torch.int32, device="cuda")
            base_d = DeviceArray(f, base.view(torch.uint8).reshape(-1), (self.cols * self.n1,), np.uint32)
            tw = base_d ** e.reshape(-1).contiguous()
            if inverse:  # the 1/N of the inverse transform rides on the twiddle
                tw = tw * f(pow(self.n % p, -1, p))
            self._tw[inverse] = tw
        return tw

    # The sub-transforms need w1 = w^N2 and w2 = w^N1.  ntt_device takes its root from primitive_root_of_unity(N1) /
    # (N2), and alpha^((p-1)/N * N2) = alpha^((p-1)/N1): they are the same elements, nothing to pass.

    def __call__(self, x_local, inverse: bool = False):
        """x_local: DeviceArray [N1][N2/G] (uint32).  Returns DeviceArray [N1/G][N2]."""
        import torch.distributed as dist

        from ._field import DeviceArray

        torch = self._torch
        f = self.field
        if tuple(x_local.shape) != (self.n1, self.cols) or x_local.dtype != np.uint32:
            raise ValueError(f"x_local must be a uint32 DeviceArray of shape {(self.n1, self.cols)}, not {x_local.shape} {x_local.dtype}.")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        # 1. columns -> contiguous rows [cols][n1], batched length-N1 transforms, twiddle
        xt = x_local.tensor.view(torch.int32).view(self.n1, self.cols).t().contiguous()
        a = DeviceArray(f, xt.view(torch.uint8).reshape(-1), (self.cols, self.n1), np.uint32)
        y = self._ntt.ntt_device(a, inverse=inverse, scaled=False)
        y = self._mul(y, self._twiddle(inverse))
        ev[1].record()
        # 2. all-to-all: block for rank h = rows k1 of its slab -> make k1 the leading axis
        z = y.tensor.view(torch.int32).view(self.cols, self.n1).t().contiguous()  # [n1][cols]
        if self.world > 1:
            recv = torch.empty_like(z)
            dist.all_to_all_single(recv, z, group=self.group)  # recv[g] = block [rows][cols] from rank g
            zz = recv.view(self.world, self.rows, self.cols).permute(1, 0, 2).contiguous()  # [rows][G * cols] = [rows][n2]
        else:
            zz = z
        ev[2].record()
        # 3. rows: length-N2 transforms
        b = DeviceArray(f, zz.view(torch.uint8).reshape(-1), (self.rows, self.n2), np.uint32)
        out = self._ntt.ntt_device(b, inverse=inverse, scaled=False)
        ev[3].record()
        torch.cuda.synchronize()
        self.timings = {"columns_ms": ev[0].elapsed_time(ev[1]), "all_to_all_ms": ev[1].elapsed_time(ev[2]),
                        "rows_ms": ev[2].elapsed_time(ev[3])}
        return out

    def _mul(self, a, b):
        from . import _lib
        from ._field import DeviceArray

        out = DeviceArray.empty(self.field, a.shape, a.dtype)
        st = _lib.lib().gfb_binary(self.field._handle(), _lib.GFB_MUL, 4, a.data_ptr, b.data_ptr, out.data_ptr, a.size,
                                   DeviceArray._stream())
        _lib.check(st, "gfb_binary")
        return out

    # helpers for tests / benches: cut a full vector into this rank's input slab, and say which outputs a rank holds
    def input_slab(self, x_full):
        """x_full: torch int32 tensor [N] on the GPU (the whole vector) -> this rank's [N1][N2/G] slab."""
        v = x_full.view(self.n1, self.n2)[:, self.rank * self.cols:(self.rank + 1) * self.cols].contiguous()
        return v

    def output_indices(self):
        """k (natural output index) of every element of this rank's [N1/G][N2] result: k = k1 + N1 * k2."""
        torch = self._torch
        k1 = torch.arange(self.rank * self.rows, (self.rank + 1) * self.rows, dtype=torch.int64, device="cuda")
        k2 = torch.arange(self.n2, dtype=torch.int64, device="cuda")
        return k1[:, None] + self.n1 * k2[None, :]
