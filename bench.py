#!/usr/bin/env python
"""
bench.py — headline benchmark of the galois_b200 hot path (driver contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload gfp_mul|ntt|rs_decode|gf2_128_mul]
    python bench.py --impl reference ...      # the UNMODIFIED reference (baseline/_ref: galois + Numba) on the host
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Default workload = BASELINE.json configs[1]: GF(2^31-1) element-wise multiply of two 1e8-element int64
arrays ("uint64" in BASELINE.json; galois has no uint64 dtype).  One "step" = one pass of
the multiply kernel over the whole 1e8-element batch resident in HBM.  `value` is field multiplications
per second over all ranks; `e2e` is the same metric through the NumPy-facing FieldArray API
(`x * y` on pinned host ndarrays: H2D + kernel + D2H inside the timed region).

One JSON line on stdout (rank 0).  Multi-GPU: the element-wise arrays and NTT batches shard by rank with
no data-path collective; timing is the max over ranks.
"""

from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

P_M31 = 2**31 - 1
P_NTT = 998244353
REF_DIR = ROOT / "baseline" / "_ref"
PORT_THREADS_CAP = 64  # OpenMP threads of the secondary CPU figure (oracle port): fixed, the same at every N


def load_reference():
    """The unmodified reference (mhostetter/galois installed at baseline/_ref), or None."""
    if not (REF_DIR / "galois").is_dir():
        return None
    if str(REF_DIR) not in sys.path:
        sys.path.insert(0, str(REF_DIR))
    os.environ.setdefault("NUMBA_CACHE_DIR", str(ROOT / "baseline" / "_numba_cache"))
    # the reference is single-threaded on this path except two prange loops (matmul_jit, evaluate_elementwise_jit); on a
    # 128-core host Numba's default of one worker per core makes every tiny parallel region cost milliseconds
    os.environ.setdefault("NUMBA_NUM_THREADS", "8")
    try:
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import galois
        return galois
    except Exception as exc:  # numba / llvmlite missing
        print(f"[bench] reference not importable: {type(exc).__name__}: {exc}", file=sys.stderr)
        return None
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


# ----------------------------------------------------------------------------------------------- helpers
def measured_hbm_peak():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.load(open(f))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(kernel_key: str, units_per_step: int):
    """DRAM bytes per step (dram__bytes_read.sum + dram__bytes_write.sum per unit from the committed
    `ncu --set full` capture, profiles/*traffic*.json, times the units one step processes), or None."""
    for f in sorted((ROOT / "profiles").glob("*traffic*.json"), reverse=True):
        try:
            d = json.load(open(f))
            if kernel_key in d:
                return d[kernel_key]["dram_bytes_per_unit"] * units_per_step
        except Exception:
            continue
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus: int):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    return rank, world, local


def barrier_sync(world):
    import torch

    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms: float, world: int) -> float:
    import torch

    if world == 1:
        return ms
    import torch.distributed as dist

    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def time_steps(step_fn, steps: int, warmup: int, world: int, per_step=None):
    """W untimed warm-ups, then EXACTLY `steps` steps between barrier+synchronize pairs; CUDA events on the
    current torch stream (the stream every kernel here is launched on).  Returns max-over-ranks milliseconds of the
    whole region; `per_step` (a list) receives this rank's per-step durations (events between the steps)."""
    import torch

    for _ in range(warmup):
        step_fn()
    barrier_sync(world)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)] if per_step is not None and steps <= 4096 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if marks:
        marks[0].record()
    for i in range(steps):
        step_fn()
        if marks:
            marks[i + 1].record()
    e1.record()
    barrier_sync(world)
    if marks:
        per_step.extend(marks[i].elapsed_time(marks[i + 1]) for i in range(steps))
    return max_over_ranks(e0.elapsed_time(e1), world)


def time_steps_graph(step_fn, steps: int, warmup: int, world: int):
    """Launch-bound workloads (a 1e6-element call is a few microseconds of GPU work, less than the Python + driver launch
    path): the K steps are captured ONCE into a CUDA graph — K back-to-back launches of the same kernels on the same
    buffers — and the timed region is the replay of that graph.  Same work, no per-step host overhead."""
    import torch

    for _ in range(warmup):
        step_fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        step_fn()
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            for _ in range(steps):
                step_fn()
    torch.cuda.synchronize()
    g.replay()  # untimed: instantiation / upload
    barrier_sync(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    barrier_sync(world)
    return max_over_ranks(e0.elapsed_time(e1), world)


# ----------------------------------------------------------------------------------------------- workloads
class GfpMul:
    """BASELINE configs[1]: GF(2^31-1) x*y, 1e8 int64 elements per GPU (weak scaling)."""

    name = "GF(2^31-1) multiply, 1e8-element int64 arrays per GPU (BASELINE configs[1])"
    metric, unit, dtype, scaling = "gfp_mul_per_sec", "field-mul/s", "int64", "weak"
    kernel = "k_binary<u64, PrimeMersenne31, MUL>"
    n = 10**8
    bytes_per_unit = 24  # 2 x 8 B read + 8 B write

    def __init__(self, rank, world):
        import torch

        import galois_b200 as gb

        self.gb, self.torch = gb, torch
        self.GF = gb.GF(P_M31)
        g = torch.Generator(device="cuda").manual_seed(1234 + rank)
        x = torch.randint(0, P_M31, (self.n,), dtype=torch.int64, device="cuda", generator=g)
        y = torch.randint(1, P_M31, (self.n,), dtype=torch.int64, device="cuda", generator=g)
        self.x = gb.DeviceArray(self.GF, x.view(torch.uint8), (self.n,), np.int64)
        self.y = gb.DeviceArray(self.GF, y.view(torch.uint8), (self.n,), np.int64)
        self.out = gb.DeviceArray.empty(self.GF, (self.n,), np.int64)
        self.units_per_step = self.n
        self.launches_per_step = 1
        self.h2d = self.d2h = 0

    def step(self):
        from galois_b200 import _lib

        st = _lib.lib().gfb_binary(self.GF._handle(), _lib.GFB_MUL, 8, self.x.data_ptr, self.y.data_ptr,
                                   self.out.data_ptr, self.n, self.gb.DeviceArray._stream())
        _lib.check(st, "gfb_binary")

    def check(self):
        import oracle

        idx = slice(0, 200_000)
        xs = self.x.tensor.view(self.torch.int64)[idx].cpu().numpy()
        ys = self.y.tensor.view(self.torch.int64)[idx].cpu().numpy()
        got = self.out.tensor.view(self.torch.int64)[idx].cpu().numpy()
        want = oracle.Field.prime(P_M31).mul(xs, ys)
        assert np.array_equal(got, want), "GPU result differs from the oracle"

    # end-to-end through the reference-facing API with HOST ndarrays: x * y on the reference's own galois.GF(2^31-1)
    # arrays (ordinary pageable NumPy memory) with galois_b200.enable(GF); the mirror class on pinned arrays is the
    # second figure
    def e2e_setup(self, galois=None):
        self.h2d, self.d2h = 2 * 8 * self.n, 8 * self.n
        self.px = self.GF.Random(self.n, seed=1, dtype=np.int64)
        self.py = self.GF.Random(self.n, seed=2, low=1, dtype=np.int64)
        if galois is None:
            self.hx, self.hy = self.px, self.py
            self.e2e_api = "galois_b200.GF(2^31-1): x * y on pinned host ndarrays (gfb_binary_host)"
            return
        self.RGF = self.gb.enable(galois.GF(P_M31))
        self.hx = self.RGF.Random(self.n, seed=1, dtype=np.int64)
        self.hy = self.RGF.Random(self.n, seed=2, low=1, dtype=np.int64)
        self.e2e_api = ("galois.GF(2^31-1) (the unmodified reference, baseline/_ref) after galois_b200.enable(GF): x * y on "
                        "pageable host ndarrays -> gfb_binary_host")

    def e2e_step(self):
        self.hz = self.hx * self.hy

    def e2e_pinned_step(self):
        self.pz = self.px * self.py

    def e2e_teardown(self):
        if hasattr(self, "RGF"):
            assert type(self.hz) is self.RGF and self.hz.dtype == np.int64
            import oracle

            want = oracle.Field.prime(P_M31).mul(np.asarray(self.hx[:100000]), np.asarray(self.hy[:100000]))
            assert np.array_equal(np.asarray(self.hz[:100000]), want), "e2e result differs from the oracle"
            self.gb.disable(self.RGF)

    # the reference itself: galois.GF(2^31-1), x * y on the full arrays (Numba jit-calculate ufunc, single thread)
    def ref_setup(self, galois):
        GF = galois.GF(P_M31)
        self.rx = GF.Random(self.n, seed=1, dtype=np.int64)
        self.ry = GF.Random(self.n, seed=2, low=1, dtype=np.int64)

    def ref_step(self):
        self.rz = self.rx * self.ry
        return self.n

    ref_sample = "full 1e8-element int64 arrays: x * y on galois.GF(2^31-1) FieldArrays (Numba jit-calculate ufunc, _calculate.py:337-340)"

    def cpu_step(self, threads):
        import oracle

        if not hasattr(self, "cx"):
            rng = np.random.default_rng(1)
            self.cx = rng.integers(0, P_M31, self.n, dtype=np.int64)
            self.cy = rng.integers(1, P_M31, self.n, dtype=np.int64)
            self.cz = np.empty_like(self.cx)
        oracle.prime_mul_i64(P_M31, self.cx, self.cy, self.cz, threads=threads)
        return self.n

    cpu_sample = "full 1e8-element int64 arrays, (a*b) % p loop (_calculate.py:337-340) split over OpenMP threads"


class Ntt:
    """BASELINE configs[2]: 2^20-point NTT over GF(998244353), batch 4096 sharded over the ranks (strong)."""

    metric, unit, dtype, scaling = "ntt_2p20_per_sec", "NTT/s", "uint32", "strong"
    kernel = "k_ntt_fused (persistent, TMA-staged; GFB_NTT_FUSED=0: k_ntt_pass1 + k_ntt_pass2)"
    N = 2**20
    bytes_per_unit = 2 * 4 * 2**20  # 8 MiB per transform

    def __init__(self, rank, world, batch=4096):
        import torch

        import galois_b200 as gb

        self.gb, self.torch = gb, torch
        self.GF = gb.GF(P_NTT)
        from galois_b200._dist import shard_bounds

        self.batch_total = batch
        lo, hi = shard_bounds(batch, rank, world)
        self.rows = hi - lo
        self.name = f"NTT 2^20 over GF(998244353), batch {batch} sharded over {world} GPU(s) (BASELINE configs[2])"
        g = torch.Generator(device="cuda").manual_seed(3 + rank)
        x = torch.randint(0, P_NTT, (self.rows, self.N), dtype=torch.int32, device="cuda", generator=g)
        self.x = gb.DeviceArray(self.GF, x.view(torch.uint8).reshape(-1), (self.rows, self.N), np.uint32)
        self.out = gb.DeviceArray.empty(self.GF, (self.rows, self.N), np.uint32)
        self.units_per_step = self.rows
        self.launches_per_step = 2 * ((self.rows + 73) // 74)  # pass 1 + pass 2 per chunk of 74 transforms
        self.h2d = self.d2h = 0

    def step(self):
        self.gb.ntt_device(self.x, out=self.out)

    def gather_tensor(self):
        return self.out.tensor.view(self.rows, -1)

    def check(self):
        import oracle

        F = oracle.Field.prime(P_NTT)
        omega = int(self.GF.primitive_root_of_unity(self.N))
        row = self.x.tensor[: 4 * self.N].view(self.torch.int32).cpu().numpy().astype(np.int64)
        got = self.out.tensor[: 4 * self.N].view(self.torch.int32).cpu().numpy().astype(np.int64)
        assert np.array_equal(got, F.fft(row, omega)), "GPU NTT differs from the oracle"

    def e2e_setup(self, galois=None):
        self.e2e_rows = min(self.rows, 256)
        self.hx = self.GF.Random((self.e2e_rows, self.N), seed=3)
        self.h2d = self.d2h = 4 * self.N * self.e2e_rows
        self.e2e_units = self.e2e_rows
        self.e2e_api = "galois_b200.ntt on a pinned [rows, 2^20] host batch (gfb_ntt_exec_host)"

    def e2e_step(self):
        self.hz = self.gb.ntt(self.hx)

    # the reference is 1-D only (_function.py:180-181): one galois.ntt call per row, single thread
    ref_rows = 4

    def ref_setup(self, galois):
        self.rgalois = galois
        GF = galois.GF(P_NTT)
        self.rx = GF.Random((self.ref_rows, self.N), seed=3)

    def ref_step(self):
        for i in range(self.ref_rows):
            self.rz = self.rgalois.ntt(self.rx[i])
        return self.ref_rows

    ref_sample = "4 rows of the 4096 per step: galois.ntt(x[i]) on GF(998244353) arrays (Numba fft_jit, _function.py:337-352)"

    def cpu_step(self, threads):
        import oracle

        F = oracle.Field.prime(P_NTT)
        rows = max(threads, 8)
        if not hasattr(self, "cx"):
            self.cx = np.random.default_rng(3).integers(0, P_NTT, (rows, self.N), dtype=np.int64)
            self.omega = int(self.GF.primitive_root_of_unity(self.N))
        F.fft(self.cx, self.omega, threads=threads)
        return rows

    cpu_sample = "max(threads, 8) rows of the 4096 (one 1-D fft_jit-style transform per row, rows over OpenMP threads)"


class RsDecode:
    """BASELINE configs[4]: RS(255,223) decode of codewords with 16 symbol errors each (t = 16)."""

    n_err, n_era = 16, 0

    metric, unit, dtype, scaling = "rs255_223_decode_cw_per_sec", "codeword/s", "u8", "strong"
    kernel = "k_rs_decode_bs"
    d1 = 32  # parity symbols (the bit-sliced decoder takes codes with at most 32; above that k_rs_decode runs)
    bytes_per_unit = 518

    def __init__(self, rank, world, total=10**6):
        import torch

        import galois_b200 as gb

        self.gb, self.torch = gb, torch
        from galois_b200._dist import shard_bounds

        self.rs = gb.ReedSolomon(255, 223)
        lo, hi = shard_bounds(total, rank, world)
        self.N = hi - lo
        self.batch_total = total
        self.name = f"RS(255,223) decode, {total} codewords x 16 errors sharded over {world} GPU(s) (BASELINE configs[4])"
        GF = self.rs.field
        M = GF.Random((self.N, 223), seed=6 + rank)
        dC = self.rs.encode_device(M.to_device())
        C = dC.tensor.view(self.N, 255)
        g = torch.Generator(device="cuda").manual_seed(7 + rank)
        # distinct positions per row: top-k of random keys; the first n_err get a non-zero error, the rest are erased
        tot = self.n_err + self.n_era
        keys = torch.rand((self.N, 255), device="cuda", generator=g)
        pos = keys.topk(tot, dim=1).indices
        err = torch.randint(1, 256, (self.N, tot), device="cuda", generator=g, dtype=torch.int32).to(torch.uint8)
        R = C.clone()
        R.scatter_(1, pos, R.gather(1, pos) ^ err)
        self.E = None
        if self.n_era:
            E = torch.zeros((self.N, 255), dtype=torch.uint8, device="cuda")
            E.scatter_(1, pos[:, self.n_err:], 1)
            self.E = E
        self.C = C
        self.R = gb.DeviceArray(GF, R.reshape(-1), (self.N, 255), np.uint8)
        self.name = self.name.replace("x 16 errors", f"x {self.n_err} errors + {self.n_era} erasures")
        self.units_per_step = self.N
        self.launches_per_step = 1
        self.h2d = self.d2h = 0

    def step(self):
        self.dec, self.nerr = self.rs.decode_device(self.R, erasures=self.E)

    def gather_tensor(self):
        return self.dec.tensor.view(self.N, 255)

    def secondary_roofline(self, units_per_sec):
        """the decoder is not HBM-bound.  The bit-sliced kernel (d1 <= 32) is bound by the SM's logic (ALU) pipe:
        its useful work is the constant-multiply Horner steps of the syndromes and the Chien search, 64 LOP3 warp-instructions
        per step for 32 codewords, against 4 sub-partitions x one ALU warp-instruction every 2 clocks.  The log-domain kernel
        (d1 > 32) is bound by shared-memory table lookups: field multiply-adds against one 32-lane wavefront per clock."""
        n, d1, v, u = 255, self.d1, self.n_err, self.n_era
        sm_mhz = 1965.0
        try:
            sm_mhz = float(json.load(open(ROOT / "MEASURED_PEAKS.json"))["sm_max_mhz"])
        except Exception:
            pass
        n_sm = self.torch.cuda.get_device_properties(0).multi_processor_count
        if d1 <= 32:
            lop3 = (n + 8 * (u + v + 1)) * 64 / 32.0  # warp-level LOP3 per codeword: syndromes + 8 Chien rounds of 32 positions
            peak = n_sm * sm_mhz * 1e6 * 4 * 0.5
            ach = units_per_sec * lop3
            return {"bound": "alu", "achieved": ach, "peak": peak, "unit": "warp-LOP3/s", "frac": ach / peak,
                    "ops_per_unit": lop3, "kernel": "k_rs_decode_bs",
                    "note": "useful logic work only: (n + 8 (u+v+1)) Horner steps x 64 LOP3 per 32 codewords; peak = SMs x 4 "
                            "sub-partitions x 0.5 ALU-pipe warp-instructions per clock x SM clock.  ncu (profiles/r02_summary.md): "
                            "ALU pipe 63.5 % active, the busiest unit of the kernel, once the bit transposes, byte packing and "
                            "the thread-per-codeword key equation are counted"}
        ops = d1 * n + (d1 - u) * 2 * (v + 1) + n * (u + v + 1) + (v + u) * (d1 + u + v) + d1 * (v + 1) + d1 * (u + 1) * (u > 0)
        peak = n_sm * sm_mhz * 1e6 * 32
        ach = units_per_sec * ops
        return {"bound": "smem", "achieved": ach, "peak": peak, "unit": "GF multiply-add/s", "frac": ach / peak,
                "ops_per_unit": ops, "kernel": "k_rs_decode",
                "note": "syndromes d1*n + Berlekamp-Massey + Chien n*(u+v+1) + Forney + the Omega / Gamma products; peak = SMs x 32 lanes x SM clock (one conflict-free shared-memory wavefront per clock)"}

    def check(self):
        assert bool((self.nerr == self.n_err).all()), "not every codeword reported the injected number of errors"
        assert bool((self.dec.tensor.view(self.N, 255) == self.C).all()), "decoded codewords differ from the originals"

    def e2e_setup(self, galois=None):
        self.hR = self.R.cpu()
        self.hE = self.E.cpu().numpy().astype(bool) if self.E is not None else None
        self.h2d, self.d2h = 255 * self.N * (2 if self.E is not None else 1), 255 * self.N + 8 * self.N
        self.e2e_units = self.N
        self.e2e_api = "galois_b200.ReedSolomon(255,223).decode on pinned host ndarrays (gfb_rs_decode_host)"
        if galois is not None and type(self) in (RsDecode, RsDecodeErasures):
            self.rrs = galois.ReedSolomon(255, 223)
            self.RGF = self.gb.enable(self.rrs.field)
            self.hR = np.array(np.asarray(self.hR)).view(self.RGF)  # pageable copy owned by the reference class
            self.e2e_api = ("galois.ReedSolomon(255,223).decode (the unmodified reference) after galois_b200.enable(rs.field), "
                            "pageable host ndarrays -> bch_decode_jit hook -> gfb_rs_decode_host")

    def e2e_step(self):
        rs = getattr(self, "rrs", self.rs)
        self.hz = rs.decode(self.hR, erasures=self.hE, output="codeword", errors=True)

    def e2e_teardown(self):
        if hasattr(self, "RGF"):
            dec, nerr = self.hz
            assert type(dec) is self.RGF and bool((nerr == self.n_err).all())
            assert np.array_equal(np.asarray(dec), self.C.cpu().numpy()), "e2e decode differs from the original codewords"
            self.gb.disable(self.RGF)

    ref_cw = 200

    def ref_setup(self, galois):
        self.rrs = galois.ReedSolomon(255, 223)
        GF = self.rrs.field
        rng = np.random.default_rng(7)
        C = self.rrs.encode(GF.Random((self.ref_cw, 223), seed=6))
        E = np.zeros((self.ref_cw, 255), dtype=bool)
        for i in range(self.ref_cw):
            pos = rng.choice(255, self.n_err + self.n_era, replace=False)
            C[i, pos] ^= GF(rng.integers(1, 256, pos.size, dtype=np.uint8))
            E[i, pos[self.n_err:]] = True
        self.rR, self.rE = C, (E if self.n_era else None)

    def ref_step(self):
        self.rz = self.rrs.decode(self.rR, erasures=self.rE, errors=True)
        return self.ref_cw

    ref_sample = "200 codewords per step: galois.ReedSolomon(255,223).decode(..., errors=True) (Numba bch_decode_jit, _bch.py:1337-1578)"

    def cpu_step(self, threads):
        import oracle

        n = 250 * threads
        if not hasattr(self, "ors"):
            self.ors = oracle.ReedSolomon(oracle.Field.binary(8, 0x11D, alpha=2), 255, 223)
            self.cR = np.asarray(self.R.cpu())[:n].astype(np.int64)
            self.cE = self.E[:n].cpu().numpy() if self.E is not None else None
        self.ors.decode(self.cR, self.cE, threads=threads)
        return len(self.cR)

    cpu_sample = "250 codewords per thread with 16 errors each (bch_decode_jit restatement, codewords over OpenMP threads)"


class Gf128Mul:
    """BASELINE configs[3]: GF(2^128) (GCM polynomial) multiply, 1e7 elements."""

    name = "GF(2^128) x^128+x^7+x^2+x+1 multiply, 1e7 elements (BASELINE configs[3])"
    metric, unit, dtype, scaling = "gf2_128_mul_per_sec", "field-mul/s", "u128", "weak"
    kernel = "k_binary<u128, Binary128, MUL>"
    n = 10**7
    bytes_per_unit = 48

    def __init__(self, rank, world):
        import torch

        import galois_b200 as gb

        self.gb, self.torch = gb, torch
        self.GF = gb.GF(2**128, irreducible_poly="x^128 + x^7 + x^2 + x + 1", primitive_element=2, verify=False)
        g = torch.Generator(device="cuda").manual_seed(5 + rank)
        a = torch.randint(-(2**63), 2**63 - 1, (self.n, 2), dtype=torch.int64, device="cuda", generator=g)
        b = torch.randint(-(2**63), 2**63 - 1, (self.n, 2), dtype=torch.int64, device="cuda", generator=g)
        self.a = gb.DeviceArray(self.GF, a.view(torch.uint8).reshape(-1), (self.n,), np.object_)
        self.b = gb.DeviceArray(self.GF, b.view(torch.uint8).reshape(-1), (self.n,), np.object_)
        self.out = gb.DeviceArray.empty(self.GF, (self.n,), np.object_)
        self.units_per_step, self.launches_per_step, self.h2d, self.d2h = self.n, 1, 0, 0

    def step(self):
        from galois_b200 import _lib

        st = _lib.lib().gfb_binary(self.GF._handle(), _lib.GFB_MUL, 16, self.a.data_ptr, self.b.data_ptr,
                                   self.out.data_ptr, self.n, self.gb.DeviceArray._stream())
        _lib.check(st, "gfb_binary")

    def check(self):
        import oracle

        O = oracle.Field128(128, (1 << 128) | 0x87)
        k = 20000
        la = self.a.tensor[: 16 * k].cpu().numpy().view(np.uint64).reshape(-1, 2)
        lb = self.b.tensor[: 16 * k].cpu().numpy().view(np.uint64).reshape(-1, 2)
        got = self.out.tensor[: 16 * k].cpu().numpy().view(np.uint64).reshape(-1, 2)
        assert np.array_equal(got, O.mul(la, lb, threads=8)), "GPU GF(2^128) product differs from the oracle"

    def e2e_setup(self, galois=None):
        self.e2e_units = None

    ref_n = 20000

    def ref_setup(self, galois):
        GF = galois.GF(2**128, irreducible_poly="x^128 + x^7 + x^2 + x + 1", primitive_element=2, verify=False)
        rng = np.random.default_rng(5)
        limbs = rng.integers(0, 2**64, (2, self.ref_n, 2), dtype=np.uint64)
        mk = lambda l: GF([int(lo) | (int(hi) << 64) for lo, hi in l])  # noqa: E731
        self.rx, self.ry = mk(limbs[0]), mk(limbs[1])

    def ref_step(self):
        self.rz = self.rx * self.ry
        return self.ref_n

    ref_sample = "2e4 elements per step: x * y on galois.GF(2^128) object arrays (python-calculate, _calculate.py:309-324)"

    def cpu_step(self, threads):
        import oracle

        O = oracle.Field128(128, (1 << 128) | 0x87)
        n = 200_000 * threads
        if not hasattr(self, "ca"):
            rng = np.random.default_rng(5)
            self.ca = rng.integers(0, 2**64, (n, 2), dtype=np.uint64)
            self.cb = rng.integers(0, 2**64, (n, 2), dtype=np.uint64)
        O.mul(self.ca, self.cb, threads=threads)
        return n

    cpu_sample = "2e5 elements per thread, shift-and-xor multiply on 128-bit integers (the reference itself runs this field on Python ints)"


class Ewise:
    """Generic element-wise workload: one ufunc over n-element arrays of one field / dtype (device-resident)."""

    scaling = "weak"
    field_order, field_kw, op, np_dtype, n, dtype = P_M31, {}, "mul", np.int64, 10**8, "int64"
    metric, unit = "gfp_op_per_sec", "field-op/s"

    def __init__(self, rank, world):
        import torch

        import galois_b200 as gb
        from galois_b200 import _lib

        self.gb, self.torch, self._lib = gb, torch, _lib
        self.GF = gb.GF(self.field_order, **self.field_kw)
        self.eb = self.GF._elem_bytes(self.np_dtype)
        self.bytes_per_unit = (2 if self.op in ("inv", "neg") else 3) * self.eb
        type(self).bytes_per_unit = self.bytes_per_unit
        g = torch.Generator(device="cuda").manual_seed(99 + rank)
        order = self.GF.order

        def rand(low):
            if self.eb == 16:
                t = torch.randint(-(2**63), 2**63 - 1, (self.n, 2), dtype=torch.int64, device="cuda", generator=g)
                t[:, 0] |= 1  # never zero
                return t
            tdt = {1: torch.uint8, 2: torch.int16, 4: torch.int32, 8: torch.int64}[self.eb]
            return torch.randint(low, order, (self.n,), dtype=torch.int64, device="cuda", generator=g).to(tdt)

        a, b = rand(0), rand(1)
        self.a = gb.DeviceArray(self.GF, a.view(torch.uint8).reshape(-1), (self.n,), self.np_dtype)
        self.b = gb.DeviceArray(self.GF, b.view(torch.uint8).reshape(-1), (self.n,), self.np_dtype)
        self.out = gb.DeviceArray.empty(self.GF, (self.n,), self.np_dtype)
        self.units_per_step, self.launches_per_step, self.h2d, self.d2h = self.n, 1, 0, 0
        self.name = f"{self.GF.name} {self.op}, {self.n:.0e}-element {self.dtype} arrays per GPU"

    def step(self):
        L, lib, h, st = self._lib, self._lib.lib(), self.GF._handle(), self.gb.DeviceArray._stream()
        if self.op in ("inv", "neg"):
            code = L.GFB_INV if self.op == "inv" else L.GFB_NEG
            rc = lib.gfb_unary(h, code, self.eb, self.b.data_ptr, self.out.data_ptr, self.n, st)
        else:
            code = {"add": L.GFB_ADD, "sub": L.GFB_SUB, "mul": L.GFB_MUL, "div": L.GFB_DIV}[self.op]
            rc = lib.gfb_binary(h, code, self.eb, self.a.data_ptr, self.b.data_ptr, self.out.data_ptr, self.n, st)
        L.check(rc, "gfb ewise")

    def _oracle(self):
        import oracle

        if self.GF._is_prime_field:
            return oracle.Field.prime(self.GF.order)
        if self.GF.degree >= 64:
            return oracle.Field128(self.GF.degree, int(self.GF.irreducible_poly))
        return oracle.Field.binary(self.GF.degree, int(self.GF.irreducible_poly),
                                   alpha=int(self.GF.primitive_element) if self.GF.degree <= 16 else None)

    def _host(self, arr, k):
        raw = arr.tensor[: k * self.eb].cpu().numpy()
        if self.eb == 16:
            return raw.view(np.uint64).reshape(-1, 2)
        return raw.view(self.np_dtype).astype(np.int64)

    def check(self):
        import oracle

        k = min(self.n, 20000 if self.eb == 16 and self.op == "inv" else 100_000)
        if self.eb == 16 and self.op == "inv":
            k = 300
        a, b, got = self._host(self.a, k), self._host(self.b, k), self._host(self.out, k)
        F = self._oracle()
        opc = {"add": oracle.ADD, "sub": oracle.SUB, "mul": oracle.MUL, "div": oracle.DIV}
        want = F.unary_op(oracle.INV if self.op == "inv" else oracle.NEG, b, threads=8) if self.op in ("inv", "neg") \
            else F.binary_op(opc[self.op], a, b, threads=8)
        assert np.array_equal(got, want), f"GPU {self.name} differs from the oracle"

    def ref_setup(self, galois):
        GF = galois.GF(self.field_order, **self.field_kw)
        n = self.ref_n if self.ref_n else min(self.n, 10**7)
        if GF.dtypes == [np.object_]:
            rng = np.random.default_rng(1)
            mk = lambda: GF([int(lo) | (int(hi) << 64) | 1 for lo, hi in rng.integers(0, 2**63, (n, 2))])  # noqa: E731
            self.rx, self.ry = mk(), mk()
        else:
            self.rx = GF.Random(n, seed=1, dtype=self.np_dtype)
            self.ry = GF.Random(n, seed=2, low=1, dtype=self.np_dtype)
        self.ref_count = n

    def ref_step(self):
        if self.op == "inv":
            self.rz = np.reciprocal(self.ry)
        elif self.op == "neg":
            self.rz = -self.ry
        else:
            self.rz = {"add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.true_divide}[self.op](self.rx, self.ry)
        return self.ref_count

    ref_n = 0
    ref_sample = "min(n, 1e7)-element arrays through the reference's own ufunc on galois.GF arrays (single thread)"

    def e2e_setup(self, galois=None):
        self.e2e_api = "galois_b200 FieldArray ufunc on pinned host ndarrays (gfb_*_host)"
        if self.eb == 16:
            self.e2e_units = None  # object-dtype host arrays: conversion cost dominates, not a transfer benchmark
            return
        self.hx = self.GF.Random(self.n, seed=1, dtype=self.np_dtype)
        self.hy = self.GF.Random(self.n, seed=2, low=1, dtype=self.np_dtype)
        nin = 1 if self.op in ("inv", "neg") else 2
        self.h2d, self.d2h = nin * self.eb * self.n, self.eb * self.n

    def e2e_step(self):
        if self.op == "inv":
            self.hz = np.reciprocal(self.hy)
        elif self.op == "neg":
            self.hz = -self.hy
        else:
            self.hz = {"add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.true_divide}[self.op](self.hx, self.hy)

    def cpu_step(self, threads):
        import oracle

        if not hasattr(self, "cF"):
            self.cF = self._oracle() if hasattr(self, "GF") else None
        n = self.cpu_n * threads if self.cpu_per_thread else self.cpu_n
        if not hasattr(self, "ca"):
            rng = np.random.default_rng(1)
            if self.eb_cpu == 16:
                self.ca = rng.integers(0, 2**64, (n, 2), dtype=np.uint64) | np.uint64(1)
                self.cb = rng.integers(0, 2**64, (n, 2), dtype=np.uint64) | np.uint64(1)
            else:
                self.ca = rng.integers(0, self.field_order, n, dtype=np.int64)
                self.cb = rng.integers(1, self.field_order, n, dtype=np.int64)
        opc = {"add": oracle.ADD, "sub": oracle.SUB, "mul": oracle.MUL, "div": oracle.DIV}
        if self.op in ("inv", "neg"):
            self.cF.unary_op(oracle.INV if self.op == "inv" else oracle.NEG, self.cb, threads=threads)
        else:
            self.cF.binary_op(opc[self.op], self.ca, self.cb, threads=threads)
        return n

    cpu_n, cpu_per_thread, eb_cpu = 10**7, False, 8
    cpu_sample = "1e7-element int64 arrays through the reference's scalar kernel (int64 DUFunc semantics), elements over OpenMP threads"
    kernel = "k_binary / k_unary"


def _ewise(name, **kw):
    kw.setdefault("metric", name + "_per_sec")
    return type("Ewise_" + name, (Ewise,), kw)


GCM = {"irreducible_poly": "x^128 + x^7 + x^2 + x + 1", "primitive_element": 2, "verify": False}
EWISE = {
    "gfp_add": _ewise("gfp_add", op="add", kernel="k_binary<u64, PrimeMersenne31, ADD>"),
    "gfp_inv": _ewise("gfp_inv", op="inv", kernel="k_unary<u64, PrimeMersenne31, INV> (batch inversion)"),
    "gfp_div": _ewise("gfp_div", op="div", kernel="k_binary<u64, PrimeMersenne31, DIV> (batch inversion)"),
    "gfp_mul_u32": _ewise("gfp_mul_u32", op="mul", np_dtype=np.uint32, dtype="uint32",
                          kernel="k_binary<u32, PrimeMersenne31, MUL>"),
    "gfp_mul_barrett": _ewise("gfp_mul_barrett", op="mul", field_order=P_NTT, np_dtype=np.uint32, dtype="uint32",
                              kernel="k_binary<u32, PrimeBarrett, MUL>"),
    "gf256_mul": _ewise("gf256_mul", op="mul", field_order=2**8, np_dtype=np.uint8, dtype="u8",
                        kernel="k_tab8_aa (64 KiB product table)"),
    "gf256_mul_1e6": _ewise("gf256_mul_1e6", op="mul", field_order=2**8, np_dtype=np.uint8, dtype="u8", n=10**6,
                            kernel="k_tab8_aa (64 KiB product table)", cpu_n=10**6, prefer_graph=True),
    "gf2_32_mul": _ewise("gf2_32_mul", op="mul", field_order=2**32, np_dtype=np.uint32, dtype="uint32", n=10**7,
                         kernel="k_binary<u32, BinaryCalc, MUL>", cpu_n=10**6),
    "gf2_128_inv": _ewise("gf2_128_inv", op="inv", field_order=2**128, field_kw=GCM, np_dtype=np.object_, dtype="u128",
                          n=10**6, kernel="k_unary<u128, Binary128, INV>", cpu_n=200, cpu_per_thread=True, eb_cpu=16,
                          ref_n=100),
}

WORKLOADS = {"gfp_mul": GfpMul, "ntt": Ntt, "rs_decode": RsDecode, "gf2_128_mul": Gf128Mul}
WORKLOADS.update(EWISE)


class RsDecodeErasures(RsDecode):
    """RS(255,223) decode with 32 erasures per codeword (the correctable reading of BASELINE's "32 errors")."""

    n_err, n_era = 0, 32
    metric = "rs255_223_decode_erasures_cw_per_sec"


class RsEncode(RsDecode):
    """RS(255,223) systematic encode of 1e6 messages."""

    metric, kernel, bytes_per_unit = "rs255_223_encode_cw_per_sec", "k_rs_encode", 478

    def __init__(self, rank, world, total=10**6):
        import galois_b200 as gb
        from galois_b200._dist import shard_bounds

        self.gb = gb
        self.rs = gb.ReedSolomon(255, 223)
        lo, hi = shard_bounds(total, rank, world)
        self.N = hi - lo
        self.name = f"RS(255,223) encode, {total} messages sharded over {world} GPU(s)"
        self.M = self.rs.field.Random((self.N, 223), seed=6 + rank)
        self.dM = self.M.to_device()
        self.units_per_step, self.launches_per_step, self.h2d, self.d2h = self.N, 1, 0, 0

    def step(self):
        self.dC = self.rs.encode_device(self.dM)

    def check(self):
        import oracle

        ors = oracle.ReedSolomon(oracle.Field.binary(8, 0x11D, alpha=2), 255, 223)
        got = self.dC.tensor[: 255 * 2000].cpu().numpy().reshape(2000, 255)
        assert np.array_equal(got, ors.encode(np.asarray(self.M)[:2000], threads=8)), "GPU RS encode differs from the oracle"

    def e2e_setup(self, galois=None):
        self.h2d, self.d2h, self.e2e_units = 223 * self.N, 255 * self.N, self.N

    def e2e_step(self):
        self.hz = self.rs.encode(self.M)

    def cpu_step(self, threads):
        import oracle

        n = 2000 * threads
        if not hasattr(self, "ors"):
            self.ors = oracle.ReedSolomon(oracle.Field.binary(8, 0x11D, alpha=2), 255, 223)
            self.cM = np.random.default_rng(6).integers(0, 256, (n, 223), dtype=np.int64)
        self.ors.encode(self.cM, threads=threads)
        return n

    cpu_sample = "2000 messages per thread (message @ P over GF(2^8) lookups, _linear.py:270-284), messages over OpenMP threads"


class Intt(Ntt):
    """Inverse transform of BASELINE configs[2] (scaled by 1/N)."""

    metric = "intt_2p20_per_sec"

    def step(self):
        self.gb.ntt_device(self.x, inverse=True, out=self.out)

    def check(self):
        import oracle

        F = oracle.Field.prime(P_NTT)
        p = P_NTT
        omega = pow(int(self.GF.primitive_root_of_unity(self.N)), -1, p)
        row = self.x.tensor[: 4 * self.N].view(self.torch.int32).cpu().numpy().astype(np.int64)
        got = self.out.tensor[: 4 * self.N].view(self.torch.int32).cpu().numpy().astype(np.int64)
        assert np.array_equal(got, F.fft(row, omega, scale=pow(self.N % p, -1, p))), "GPU INTT differs from the oracle"

    def e2e_step(self):
        self.hz = self.gb.intt(self.hx)


class BchDecode(RsDecode):
    """Binary BCH(255,131), t = 18: decode 1e6 codewords with 18 bit errors each."""

    metric, n_err, n_era, bytes_per_unit = "bch255_131_decode_cw_per_sec", 18, 0, 518
    kernel, d1 = "k_rs_decode", 36

    def __init__(self, rank, world, total=10**6):
        import torch

        import galois_b200 as gb
        from galois_b200._dist import shard_bounds

        self.gb, self.torch = gb, torch
        self.rs = gb.BCH(255, 131)
        lo, hi = shard_bounds(total, rank, world)
        self.N = hi - lo
        self.name = f"BCH(255,131) decode, {total} codewords x 18 bit errors sharded over {world} GPU(s)"
        M = self.rs.field.Random((self.N, 131), seed=6 + rank)
        C = self.rs.encode_device(M.to_device()).tensor.view(self.N, 255)
        g = torch.Generator(device="cuda").manual_seed(7 + rank)
        pos = torch.rand((self.N, 255), device="cuda", generator=g).topk(18, dim=1).indices
        R = C.clone()
        R.scatter_(1, pos, R.gather(1, pos) ^ 1)
        self.E, self.C = None, C
        self.R = gb.DeviceArray(self.rs.field, R.reshape(-1), (self.N, 255), np.uint8)
        self.units_per_step, self.launches_per_step, self.h2d, self.d2h = self.N, 1, 0, 0

    def cpu_step(self, threads):
        import oracle

        n = 250 * threads
        if not hasattr(self, "ob"):
            self.ob = oracle.BCH(oracle.Field.binary(8, 0x11D, alpha=2), 255, 37)
            if not hasattr(self, "cR"):
                rng = np.random.default_rng(7)
                C = self.ob.encode(rng.integers(0, 2, (n, 131)), threads=threads)
                for i in range(n):
                    C[i, rng.choice(255, 18, replace=False)] ^= 1
                self.cR = C
        self.ob.decode(self.cR, threads=threads)
        return len(self.cR)

    cpu_sample = "250 codewords per thread with 18 bit errors each (bch_decode_jit restatement, field GF(2), extension GF(2^8))"


RsEncode.secondary_roofline = None
RsEncode.ref_setup = None  # no galois-driven leg for these two (oracle port only)
BchDecode.ref_setup = None
WORKLOADS.update({"rs_decode_erasures": RsDecodeErasures, "rs_encode": RsEncode, "intt": Intt, "bch_decode": BchDecode})


# ----------------------------------------------------------------------------------------------- CPU arms
class MatInv:
    """inverse of a 1024 x 1024 matrix over GF(998244353) by row-reducing [A | I] (inv_jit)."""

    metric, unit, dtype, scaling = "gfp_matinv_1024_per_sec", "inv/s", "uint32", "weak"
    kernel = "k_rr_panel_cluster (+ k_rr_panel_w2, k_rr_panel_update per 32-column panel)"
    n = 1024
    bytes_per_unit = 2 * 4 * 1024 * 1024  # compulsory: read A, write A^-1 (the n^3 update traffic is the algorithm's own)

    def __init__(self, rank, world):
        import torch

        import galois_b200 as gb
        from galois_b200 import _linalg

        self.gb, self.torch, self._linalg = gb, torch, _linalg
        self.GF = gb.GF(P_NTT)
        n = self.n
        self.name = f"inverse of a {n}x{n} matrix over GF(998244353) via row_reduce([A | I], ncols={n}), one matrix per GPU per step"
        g = torch.Generator(device="cuda").manual_seed(11 + rank)
        a = torch.randint(0, P_NTT, (n, n), dtype=torch.int32, device="cuda", generator=g)
        ai = torch.cat([a, torch.eye(n, dtype=torch.int32, device="cuda")], dim=1).contiguous()
        self.a = a
        self.src = gb.DeviceArray(self.GF, ai.view(torch.uint8).reshape(-1), (n, 2 * n), np.uint32)
        self.out = gb.DeviceArray.empty(self.GF, (n, 2 * n), np.uint32)
        self.units_per_step = 1
        self.launches_per_step = 5 * (n // 32) + 2  # panel + (w2, update) on the next panel's columns and on the far columns
        self.h2d = self.d2h = 0

    def step(self):
        _, self.info, _ = self._linalg.row_reduce_device(self.src, ncols=self.n, out=self.out)

    def check(self):
        # A . A^-1 == I through the matmul kernel, and the left block reduced to the identity with n pivots
        n, torch = self.n, self.torch
        assert int(self.info[0].item()) == n, "matrix was singular (seeded input should not be)"
        R = self.out.tensor.view(torch.int32).view(n, 2 * n)
        assert torch.equal(R[:, :n], torch.eye(n, dtype=torch.int32, device="cuda"))
        A = self.gb.DeviceArray(self.GF, self.a.contiguous().view(torch.uint8).reshape(-1), (n, n), np.uint32).cpu()
        Ainv = R[:, n:].contiguous().cpu().numpy().astype(np.uint32).view(self.GF)
        assert np.array_equal(np.asarray(A @ Ainv), np.identity(n, dtype=np.uint32)), "A @ A^-1 != I"

    def e2e_setup(self, galois=None):
        self.hA = self.GF.Random((self.n, self.n), seed=11)
        self.h2d = self.n * 2 * self.n * 4
        self.d2h = self.n * 2 * self.n * 4
        self.e2e_units = 1

    def e2e_step(self):
        self.hz = np.linalg.inv(self.hA)

    def cpu_step(self, threads):
        import oracle

        F = oracle.Field.prime(P_NTT)
        if not hasattr(self, "cA"):
            self.cA = np.random.default_rng(11).integers(0, P_NTT, (self.n, self.n), dtype=np.int64)
        F.matrix_inv(self.cA)
        return 1

    cpu_sample = "the same 1024x1024 inversion, one thread (row_reduce_jit is a serial NumPy/ufunc loop in the reference)"
    cpu_cores = 1


class PolyRoots:
    """all roots of a degree-64 polynomial over GF(2^31 - 1) (roots_jit: exhaustive search)."""

    metric, unit, dtype, scaling = "gfp_poly_roots_points_per_sec", "points/s", "uint32", "weak"
    kernel = "k_roots (count pass + write pass)"
    degree = 64
    bytes_per_unit = 0  # no array traffic: 2^31 field elements are generated from the thread index (ALU-bound)

    def __init__(self, rank, world):
        import torch

        import galois_b200 as gb
        from galois_b200 import _linalg

        self.gb, self.torch, self._linalg = gb, torch, _linalg
        self.GF = gb.GF(2**31 - 1)
        p = 2**31 - 1
        self.name = f"roots of a degree-{self.degree} polynomial with 24 planted roots over GF(2^31-1): every field element tested"
        import oracle

        F = oracle.Field.prime(p)  # test infrastructure builds the planted polynomial (not timed)
        rng = np.random.default_rng(21 + rank)
        self.planted = np.unique(rng.integers(0, p, 24))
        poly = rng.integers(1, p, self.degree + 1 - self.planted.size)
        for z in self.planted:
            poly = F.convolve(poly, np.array([1, (p - int(z)) % p]))
        self.coeffs_host = self.GF(poly.astype(np.int64))
        nz = np.nonzero(poly)[0]
        self.degrees = (poly.size - 1 - nz).astype(np.int64)
        self.coeffs = self.GF(poly[nz]).to_device()
        self.units_per_step = p
        self.launches_per_step = 3
        self.h2d = self.d2h = 0

    def step(self):
        self.roots, self.count = self._linalg.poly_roots_device(self.degrees, self.coeffs)

    def check(self):
        n = int(self.count.item())
        got = self.roots.cpu()[:n]
        assert set(self.planted.tolist()) <= set(np.asarray(got).tolist()), "planted roots missing"
        assert np.all(np.diff(np.asarray(got).astype(np.int64)) > 0), "roots are not in increasing order"
        assert not np.asarray(self.gb.poly_evaluate(self.coeffs_host, got)).any(), "a reported root is not a root"

    def e2e_setup(self, galois=None):
        self.poly = self.gb.Poly(self.coeffs_host)
        self.h2d = 4 * (self.degree + 1)
        self.d2h = 4 * self.degree
        self.e2e_units = 2**31 - 1

    def e2e_step(self):
        self.hz = self.poly.roots()

    def cpu_step(self, threads):
        import oracle

        F = oracle.Field.prime(2**31 - 1)
        n = 1 << 20
        if not hasattr(self, "cx"):
            self.cc = np.random.default_rng(21).integers(1, 2**31 - 1, self.degree + 1, dtype=np.int64)
            self.cx = np.arange(1, n + 1, dtype=np.int64)
        F.poly_evaluate(self.cc, self.cx)
        return n

    cpu_sample = ("2^20 of the 2^31 - 1 points, one thread, Horner evaluation (degree multiplies + adds per point, the "
                  "same count as the reference's Chien step lambda_j *= alpha^j)")
    cpu_cores = 1


class FftMixed:
    """mixed-radix DFT, N = 7 * 17 * 2^13 = 974848 over GF(998244353) (fft_jit general branch)."""

    metric, unit, dtype, scaling = "fft_mixed_974848_per_sec", "FFT/s", "uint32", "weak"
    kernel = "k_mixed_stage_v (radix 17, radix 7 + twiddle) -> 8192-point four-step kernels on rows -> k_transpose_rows"
    N = 7 * 17 * 2**13
    bytes_per_unit = 2 * 4 * 7 * 17 * 2**13

    def __init__(self, rank, world, batch=64):
        import torch

        import galois_b200 as gb

        self.gb, self.torch = gb, torch
        self.GF = gb.GF(P_NTT)
        self.rows = batch
        self.name = f"mixed-radix DFT N=7*17*2^13 over GF(998244353), batch {batch} per GPU"
        g = torch.Generator(device="cuda").manual_seed(31 + rank)
        x = torch.randint(0, P_NTT, (self.rows, self.N), dtype=torch.int32, device="cuda", generator=g)
        self.x = gb.DeviceArray(self.GF, x.view(torch.uint8).reshape(-1), (self.rows, self.N), np.uint32)
        self.out = gb.DeviceArray.empty(self.GF, (self.rows, self.N), np.uint32)
        self.units_per_step = self.rows
        chunk_rows = max(1, (64 << 20) // (self.N * 4))  # rows per workspace chunk (gfb_ntt.cu: kGenericWsBytes)
        self.launches_per_step = 5 * ((self.rows + chunk_rows - 1) // chunk_rows)  # 2 odd stages, 2 row passes, permutation
        self.h2d = self.d2h = 0

    def step(self):
        self.gb.ntt_device(self.x, out=self.out)

    def check(self):
        import oracle

        F = oracle.Field.prime(P_NTT)
        omega = int(self.GF.primitive_root_of_unity(self.N))
        row = self.x.tensor[: 4 * self.N].view(self.torch.int32).cpu().numpy().astype(np.int64)
        got = self.out.tensor[: 4 * self.N].view(self.torch.int32).cpu().numpy().astype(np.int64)
        assert np.array_equal(got, F.fft(row, omega)), "GPU mixed-radix DFT differs from the oracle"

    def e2e_setup(self, galois=None):
        self.e2e_rows = min(self.rows, 32)
        self.hx = self.GF.Random((self.e2e_rows, self.N), seed=31)
        self.h2d = self.d2h = 4 * self.N * self.e2e_rows
        self.e2e_units = self.e2e_rows

    def e2e_step(self):
        self.hz = np.fft.fft(self.hx)

    def cpu_step(self, threads):
        import oracle

        F = oracle.Field.prime(P_NTT)
        rows = max(threads, 8)
        if not hasattr(self, "cx"):
            self.cx = np.random.default_rng(31).integers(0, P_NTT, (rows, self.N), dtype=np.int64)
            self.omega = int(self.GF.primitive_root_of_unity(self.N))
        F.fft(self.cx, self.omega, threads=threads)
        return rows

    cpu_sample = "max(threads, 8) transforms (one fft_jit-style mixed-radix transform per row, rows over OpenMP threads)"


class Convolve:
    """np.convolve of two 4096-coefficient polynomials over GF(998244353) (reference: 41 ms,
    schoolbook because np.convolve would overflow, _function.py:147-149).  Here: three 8192-point NTTs."""

    metric, unit, dtype, scaling = "gfp_convolve_4096x4096_per_sec", "convolution/s", "uint32", "weak"
    kernel = "k_conv_pad -> k_ntt_small/four-step (x2 batched, x1) -> k_binary MUL -> k_conv_unpad"
    na = nb = 4096
    field_order = P_NTT
    bytes_per_unit = 4 * (4096 + 4096 + 8191)  # compulsory: read a and b, write the 8191 coefficients

    def __init__(self, rank, world):
        import torch

        import galois_b200 as gb
        from galois_b200 import _lib

        self.gb, self.torch, self._lib = gb, torch, _lib
        self.GF = gb.GF(self.field_order)
        self.np_dtype = self.GF.dtypes[0]
        self.eb = np.dtype(self.np_dtype).itemsize
        self.name = f"np.convolve, {self.na} x {self.nb} coefficients over {self.GF.name}, one pair per GPU per step"
        self.ha = self.GF.Random(self.na, seed=41 + rank)
        self.hb = self.GF.Random(self.nb, seed=42 + rank)
        self.a, self.b = self.ha.to_device(), self.hb.to_device()
        self.out = gb.DeviceArray.empty(self.GF, (self.na + self.nb - 1,), self.np_dtype)
        self.units_per_step, self.launches_per_step, self.h2d, self.d2h = 1, 6, 0, 0

    def step(self):
        st = self._lib.lib().gfb_convolve(self.GF._handle(), self.eb, self.a.data_ptr, self.na, self.b.data_ptr, self.nb,
                                          self.out.data_ptr, self.gb.DeviceArray._stream())
        self._lib.check(st, "gfb_convolve")

    def _oracle(self):
        import oracle

        if self.GF._is_prime_field:
            return oracle.Field.prime(self.GF.order)
        return oracle.Field.binary(self.GF.degree, int(self.GF.irreducible_poly), alpha=int(self.GF.primitive_element))

    def check(self):
        want = self._oracle().convolve(np.asarray(self.ha).astype(np.int64), np.asarray(self.hb).astype(np.int64))
        assert np.array_equal(np.asarray(self.out.cpu()).astype(np.int64), want), "GPU convolution differs from the oracle"

    def e2e_setup(self, galois=None):
        self.h2d, self.d2h = self.eb * (self.na + self.nb), self.eb * (self.na + self.nb - 1)
        self.e2e_units = 1
        self.e2e_api = "np.convolve on galois_b200 FieldArrays (gfb_convolve_host)"
        if galois is not None:
            self.RGF = self.gb.enable(galois.GF(self.field_order))
            self.ha = self.RGF(np.asarray(self.ha))
            self.hb = self.RGF(np.asarray(self.hb))
            self.e2e_api = "np.convolve on the reference's galois.GF arrays after galois_b200.enable(GF) (convolve_jit hook -> gfb_convolve_host)"

    def e2e_step(self):
        self.hz = np.convolve(self.ha, self.hb)

    def e2e_teardown(self):
        if hasattr(self, "RGF"):
            self.gb.disable(self.RGF)

    def ref_setup(self, galois):
        GF = galois.GF(self.field_order)
        self.ra, self.rb = GF.Random(self.na, seed=41), GF.Random(self.nb, seed=42)

    def ref_step(self):
        self.rz = np.convolve(self.ra, self.rb)
        return 1

    ref_sample = "the same 4096 x 4096 np.convolve on galois.GF arrays (Numba convolve_jit, _function.py:111-167)"

    def cpu_step(self, threads):
        if not hasattr(self, "ca"):
            rng = np.random.default_rng(41)
            self.ca = rng.integers(0, self.field_order, self.na, dtype=np.int64)
            self.cb = rng.integers(0, self.field_order, self.nb, dtype=np.int64)
            import oracle

            self.cF = oracle.Field.prime(self.field_order) if self.field_order == P_NTT else oracle.Field.binary(8, 0x11D, alpha=2)
        self.cF.convolve(self.ca, self.cb)
        return 1

    cpu_sample = "the same 4096 x 4096 convolution, schoolbook, one thread (convolve_jit's fall-back loop)"
    cpu_cores = 1


class ConvolveGf256(Convolve):
    """np.convolve 4096 x 4096 over GF(2^8) (reference: 36 ms): schoolbook kernel, EXP/LOG tables in shared memory."""

    metric = "gf256_convolve_4096x4096_per_sec"
    kernel = "k_convolve<u8, BinaryLut>"
    field_order = 2**8
    dtype = "u8"
    bytes_per_unit = 4096 + 4096 + 8191


class MatmulRs:
    """the RS(255,223) encode-shaped product (1e6 x 223) @ (223 x 32) over GF(2^8) through np.matmul's
    kernel (the reference: single-threaded triple loop, 1.74e4 rows/s)."""

    metric, unit, dtype, scaling = "gf256_matmul_rows_per_sec", "row/s", "u8", "weak"
    kernel = "k_matmul<u8, BinaryLut>"
    M, K, N = 10**6, 223, 32
    bytes_per_unit = 223 + 32  # per row of A: read the row, write the 32 products (B is 7 KB, read once)

    def __init__(self, rank, world):
        import torch

        import galois_b200 as gb
        from galois_b200 import _lib

        self.gb, self.torch, self._lib = gb, torch, _lib
        self.GF = gb.GF(2**8)
        self.name = f"({self.M} x {self.K}) @ ({self.K} x {self.N}) over GF(2^8), one product per GPU per step"
        g = torch.Generator(device="cuda").manual_seed(51 + rank)
        A = torch.randint(0, 256, (self.M, self.K), dtype=torch.uint8, device="cuda", generator=g)
        B = torch.randint(0, 256, (self.K, self.N), dtype=torch.uint8, device="cuda", generator=g)
        self.A = gb.DeviceArray(self.GF, A.reshape(-1), (self.M, self.K), np.uint8)
        self.B = gb.DeviceArray(self.GF, B.reshape(-1), (self.K, self.N), np.uint8)
        self.C = gb.DeviceArray.empty(self.GF, (self.M, self.N), np.uint8)
        self.units_per_step, self.launches_per_step, self.h2d, self.d2h = self.M, 1, 0, 0

    def step(self):
        st = self._lib.lib().gfb_matmul(self.GF._handle(), 1, self.A.data_ptr, self.B.data_ptr, self.C.data_ptr, 1, self.M,
                                        self.K, self.N, 0, self.gb.DeviceArray._stream())
        self._lib.check(st, "gfb_matmul")

    def check(self):
        import oracle

        F = oracle.Field.binary(8, int(self.GF.irreducible_poly), alpha=int(self.GF.primitive_element))
        k = 2000
        A = self.A.tensor.view(self.M, self.K)[:k].cpu().numpy().astype(np.int64)
        B = self.B.tensor.view(self.K, self.N).cpu().numpy().astype(np.int64)
        got = self.C.tensor.view(self.M, self.N)[:k].cpu().numpy().astype(np.int64)
        assert np.array_equal(got, F.matmul(A, B, threads=8)), "GPU matmul differs from the oracle"

    def e2e_setup(self, galois=None):
        self.hA = self.A.cpu()
        self.hB = self.B.cpu()
        self.h2d, self.d2h, self.e2e_units = self.M * self.K + self.K * self.N, self.M * self.N, self.M
        self.e2e_api = "A @ B on galois_b200 FieldArrays (gfb_matmul_host)"

    def e2e_step(self):
        self.hz = self.hA @ self.hB

    ref_rows = 2000

    def ref_setup(self, galois):
        GF = galois.GF(2**8)
        self.rA = GF.Random((self.ref_rows, self.K), seed=51)
        self.rB = GF.Random((self.K, self.N), seed=52)

    def ref_step(self):
        self.rz = self.rA @ self.rB
        return self.ref_rows

    ref_sample = "2000 rows per step: A @ B on galois.GF(2^8) arrays (Numba matmul_jit, _linalg.py:286-308)"

    def cpu_step(self, threads):
        import oracle

        if not hasattr(self, "cA"):
            self.cF = oracle.Field.binary(8, 0x11D, alpha=2)
            rng = np.random.default_rng(51)
            self.cA = rng.integers(0, 256, (2000 * threads, self.K), dtype=np.int64)
            self.cB = rng.integers(0, 256, (self.K, self.N), dtype=np.int64)
        self.cF.matmul(self.cA, self.cB, threads=threads)
        return len(self.cA)

    cpu_sample = "2000 rows per thread (matmul_jit's triple loop with EXP/LOG lookups), rows over OpenMP threads"


class NttSharded:
    """SURVEY.md §8f rank 3 (second half): ONE 2^26-point NTT spread over the ranks — local column pass, NCCL all-to-all,
    local row pass (galois_b200._dist.ShardedNtt).  Over GF(2013265921) = GF(15 * 2^27 + 1): 998244353 - 1 = 2^23 * 7 * 17
    has no 2^26-th roots of unity.  The only workload whose curve a collective can limit; `phases_ms` names the all-to-all's
    share."""

    metric, unit, dtype, scaling = "ntt_2p26_sharded_per_sec", "NTT/s", "uint32", "strong"
    kernel = "k_ntt_* (column pass) -> all_to_all_single (NCCL) -> k_ntt_* (row pass)"
    LOG_N = 26
    P = 2013265921  # 15 * 2^27 + 1, smallest primitive root 31
    bytes_per_unit = 2 * 4 * 2**26

    def __init__(self, rank, world, batch=None):
        import torch

        import galois_b200 as gb
        from galois_b200._dist import ShardedNtt

        self.gb, self.torch = gb, torch
        self.GF = gb.GF(self.P)
        self.N = 1 << self.LOG_N
        self.sh = ShardedNtt(self.GF, self.N)
        self.name = f"one 2^{self.LOG_N}-point NTT over GF({self.P}) sharded over {world} GPU(s) (column slabs in, row slabs out)"
        # two impulses: X[k] = w^(j0 k) + 5 w^(j1 k) checks every index map and twiddle at full size
        self.j0, self.j1 = 12345677 % self.N, (3 * self.N // 7) | 1
        g = torch.Generator(device="cuda").manual_seed(17)
        slab = torch.randint(0, self.P, (self.sh.n1, self.sh.cols), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
        self.x = gb.DeviceArray(self.GF, slab.view(torch.uint8).reshape(-1), (self.sh.n1, self.sh.cols), np.uint32)
        self.units_per_step = 1 if rank == 0 else 0  # one transform for the whole job
        type(self).bytes_per_unit = 2 * 4 * self.N // world  # per-GPU share of the compulsory traffic (roofline is per GPU)
        self.launches_per_step = 8
        self.h2d = self.d2h = 0
        self.phases = []

    def step(self):
        self.out = self.sh(self.x)
        self.phases.append(dict(self.sh.timings))

    def check(self):
        torch, sh, p = self.torch, self.sh, self.P
        imp = torch.zeros((sh.n1, sh.n2), dtype=torch.int32, device="cuda")
        imp.view(-1)[self.j0] = 1
        imp.view(-1)[self.j1] = 5
        slab = imp[:, sh.rank * sh.cols:(sh.rank + 1) * sh.cols].contiguous()
        X = sh(self.gb.DeviceArray(self.GF, slab.view(torch.uint8).reshape(-1), (sh.n1, sh.cols), np.uint32))
        got = X.tensor.view(torch.int32).view(sh.rows, sh.n2)
        idx = sh.output_indices()
        sel = torch.randint(0, idx.numel(), (4096,), device="cuda")
        ks = idx.reshape(-1)[sel].cpu().numpy().astype(object)
        w = sh.omega
        want = [(pow(w, int(self.j0 * k % self.N), p) + 5 * pow(w, int(self.j1 * k % self.N), p)) % p for k in ks]
        have = (got.reshape(-1)[sel].cpu().numpy().astype(np.int64) & 0xFFFFFFFF).tolist()
        assert have == want, "sharded NTT differs from w^(jk)"

    def e2e_setup(self, galois=None):
        self.e2e_units = None

    ref_setup = None

    def cpu_step(self, threads):
        import oracle

        F = oracle.Field.prime(self.P)
        if not hasattr(self, "cx"):
            self.cN = 1 << 22
            self.comega = pow(31, (self.P - 1) // self.cN, self.P)  # 31 is the smallest primitive root of 2013265921
            self.cx = np.random.default_rng(3).integers(0, self.P, self.cN, dtype=np.int64)
        F.fft(self.cx, self.comega)
        return self.cN / (1 << self.LOG_N)  # fraction of a 2^26-point transform by element count (log factor ignored: 22 vs 26 stages)

    cpu_sample = "one 2^22-point transform (1/16 of the elements), one thread, counted as 1/16 transform"
    cpu_cores = 1


WORKLOADS.update({"matinv": MatInv, "poly_roots": PolyRoots, "fft_mixed": FftMixed, "convolve": Convolve, "ntt_sharded": NttSharded,
                  "convolve_gf256": ConvolveGf256, "matmul_rs": MatmulRs})


def cpu_threads() -> int:
    """Threads of the oracle-port CPU figure: fixed by the machine, NOT by OMP_NUM_THREADS (torchrun sets that to 1)."""
    return max(1, min(PORT_THREADS_CAP, os.cpu_count() or 1))


def time_cpu(work, target_seconds=6.0, max_reps=50):
    """Time `work.cpu_step` (oracle port, OpenMP) for roughly target_seconds; returns (units/s, reps, seconds, threads)."""
    threads = cpu_threads()
    work.cpu_step(threads)  # warm-up (allocations, page faults, OpenMP pool)
    units, reps, t0 = 0, 0, time.perf_counter()
    while reps < max_reps:
        units += work.cpu_step(threads)
        reps += 1
        if time.perf_counter() - t0 > target_seconds:
            break
    dt = time.perf_counter() - t0
    return units / dt, reps, dt, threads


def has_ref(cls) -> bool:
    return callable(getattr(cls, "ref_setup", None))


def time_reference_live(work, galois, target_seconds=12.0, max_reps=30):
    """Time the unmodified reference (`work.ref_step`) for roughly target_seconds -> (units/s, reps, seconds)."""
    work.ref_setup(galois)
    work.ref_step()  # JIT compilation + first-touch page faults
    units, reps, t0 = 0, 0, time.perf_counter()
    while reps < max_reps:
        units += work.ref_step()
        reps += 1
        if time.perf_counter() - t0 > target_seconds:
            break
    dt = time.perf_counter() - t0
    return units / dt, reps, dt


def port_setup(cls, work, threads):
    """Inputs of the oracle-port leg when no CUDA object exists (the --impl reference process)."""
    if issubclass(cls, Ewise):
        import galois_b200 as gb

        work.GF = gb.GF(cls.field_order, **cls.field_kw)  # host-side description of the field only
        work.name = f"{work.GF.name} {cls.op}"
    if issubclass(cls, (Ntt, FftMixed)):
        import galois_b200 as gb

        work.GF = gb.GF(P_NTT)  # host-side constants only (omega)
    if cls in (RsDecode, RsDecodeErasures):
        import oracle

        F = oracle.Field.binary(8, 0x11D, alpha=2)
        work.ors = oracle.ReedSolomon(F, 255, 223)
        rng = np.random.default_rng(7)
        n = 250 * threads
        C = work.ors.encode(rng.integers(0, 256, (n, 223)), threads=threads)
        cE = np.zeros((n, 255), dtype=np.uint8)
        for i in range(n):
            pos = rng.choice(255, cls.n_err + cls.n_era, replace=False)
            C[i, pos] ^= rng.integers(1, 256, len(pos))
            cE[i, pos[cls.n_err:]] = 1
        work.cR = C
        work.cE = cE if cls.n_era else None


def workload_label(cls, work):
    return getattr(work, "name", None) or getattr(cls, "name", None) or (cls.__doc__ or cls.__name__).strip().split("\n")[0].rstrip(".")


def run_reference(args):
    """--impl reference: the UNMODIFIED reference (baseline/_ref = `pip install --no-deps --target` of /root/reference)
    through its own public API and stock code path — galois.GF arrays, Numba ufuncs / JIT functions, one thread (the
    reference is single-threaded on this path) — on rank 0 only.  Each step is a bounded sample of the workload
    (`cpu_baseline.sample`).  The OpenMP oracle port is printed next to it as `cpu_port` (fixed thread count).  If the
    reference cannot be imported on this machine the port is the line's value and `cpu_baseline.kind` says "port"."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cls = WORKLOADS[args.workload]
    work = cls.__new__(cls)  # CPU legs only: no CUDA objects are created
    threads = cpu_threads()
    galois = load_reference() if has_ref(cls) else None
    port = None
    try:
        port_setup(cls, work, threads)
        pv, preps, psecs, pth = time_cpu(work, target_seconds=4.0 if galois is not None else 0.0)
        port = {"value": pv, "unit": cls.unit, "cores": getattr(cls, "cpu_cores", pth), "kind": "port",
                "sample": f"{cls.cpu_sample}; {preps} repetitions in {psecs:.1f} s"}
    except Exception as exc:  # the secondary figure must not take the line down
        port = {"error": f"{type(exc).__name__}: {exc}"}

    if galois is not None:
        work.ref_setup(galois)
        step, kind, cores, sample = work.ref_step, "reference", 1, cls.ref_sample
        note = "the unmodified reference (baseline/_ref: galois + Numba) through its public API, single thread"
    else:
        step = lambda: work.cpu_step(threads)  # noqa: E731
        kind, cores, sample = "port", getattr(cls, "cpu_cores", threads), cls.cpu_sample
        note = "reference not importable here: CPU oracle port of the reference path (oracle/gf_oracle.c)"
    for _ in range(max(args.warmup, 1)):
        step()
    # bound the run: stop early (and say so) rather than exceed ~4 minutes
    t0 = time.perf_counter()
    units = done = 0
    for _ in range(args.steps):
        units += step()
        done += 1
        if time.perf_counter() - t0 > 240.0:
            break
    dt = time.perf_counter() - t0
    value = units / dt
    line = {
        "impl": "reference", "metric": cls.metric, "value": value, "unit": cls.unit, "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / done, "higher_is_better": True,
        "scaling": cls.scaling, "vs_baseline": None, "dtype": cls.dtype, "data": "synthetic",
        "config": {"workload": workload_label(cls, work), "note": note},
        "cpu_baseline": {"value": value, "unit": cls.unit, "cores": cores, "kind": kind, "sample": sample},
        "cpu_port": port,
        "e2e": {"value": value, "unit": cls.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if done != args.steps:
        line["config"]["note"] += f"; stopped after {done} of {args.steps} steps (240 s bound)"
    print(json.dumps(line), flush=True)
    return 0


def all_units(units: int, world: int) -> int:
    if world == 1:
        return units
    import torch
    import torch.distributed as dist

    t = torch.tensor([units], dtype=torch.int64, device="cuda")
    dist.all_reduce(t)
    return int(t.item())


def time_gather(work, world):
    """The path's only collective: the optional final all-gather of the row-sharded result."""
    import torch

    rows_total = getattr(work, "batch_total", None)
    if world == 1 or rows_total is None or not hasattr(work, "gather_tensor"):
        return None
    from galois_b200._dist import gather_rows

    flat = work.gather_tensor()
    for _ in range(2):
        gather_rows(flat, rows_total)
    barrier_sync(world)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    full = gather_rows(flat, rows_total)
    g1.record()
    barrier_sync(world)
    del full
    return max_over_ranks(g0.elapsed_time(g1), world)


# ----------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="gfp_mul", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="ntt: total transforms; rs_decode: total codewords")
    ap.add_argument("--gather", action="store_true", help="N>1: also time the final all-gather of the result shards")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads in the default run")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--graph", action="store_true", help="time the K steps as one CUDA-graph replay (launch-bound workloads)")
    ap.add_argument("--no-ref-e2e", action="store_true", help="e2e through the galois_b200 mirror classes even if baseline/_ref exists")
    ap.add_argument("--extras-ntt-batch", type=int, default=4096, help="total transforms of the secondary NTT workload")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference(args)

    import torch

    rank, world, local = dist_setup(args.gpus)
    cls = WORKLOADS[args.workload]
    kw = {}
    if args.batch is not None and issubclass(cls, (Ntt, FftMixed)):
        kw["batch"] = args.batch
    if args.batch is not None and issubclass(cls, RsDecode):
        kw["total"] = args.batch
    work = cls(rank, world, **kw)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    per_step = []
    use_graph = args.graph or getattr(cls, "prefer_graph", False)
    if use_graph:
        ms = time_steps_graph(work.step, args.steps, args.warmup, world)
    else:
        ms = time_steps(work.step, args.steps, args.warmup, world, per_step=per_step)
    clocks = sampler.stop() if rank == 0 else None
    work.check()
    workload_name = work.name
    launches = args.steps * work.launches_per_step
    per_gpu_units = work.units_per_step

    # whole-job units per step: per-rank shards may differ by one row when the batch does not divide evenly
    units_all = all_units(work.units_per_step, world)
    total_units = units_all * args.steps
    value = total_units / (ms / 1e3)

    # optional: the final gather of the result shards (the only collective of the path), timed on its own
    gather_ms = time_gather(work, world) if args.gather else None
    peak, peak_src = measured_hbm_peak()
    # dominant kernel: algorithmic bytes of one step on this rank / mean device time of one step
    achieved = work.units_per_step * cls.bytes_per_unit / (ms / 1e3 / args.steps) / 1e9

    sec_roofline = (work.secondary_roofline(work.units_per_step * args.steps / (ms / 1e3))
                    if callable(getattr(work, "secondary_roofline", None)) else None)

    # ---- end to end through the reference-facing API with HOST buffers (same metric): the reference's own classes
    #      with galois_b200.enable() when baseline/_ref is importable, else the mirror classes
    galois = load_reference() if not args.no_ref_e2e else None
    e2e = None
    work.e2e_setup(galois)
    if getattr(work, "e2e_units", work.units_per_step) is not None:
        e2e_units = getattr(work, "e2e_units", work.units_per_step)
        e2e_steps = max(3, min(args.steps, 10))

        def timed(fn):
            for _ in range(2):
                fn()
            barrier_sync(world)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                fn()
            barrier_sync(world)
            return max_over_ranks((time.perf_counter() - t0) * 1e3, world) / 1e3

        dt = timed(work.e2e_step)
        e2e = {"value": e2e_units * world * e2e_steps / dt, "unit": cls.unit, "h2d_bytes_per_step": work.h2d,
               "d2h_bytes_per_step": work.d2h, "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
               "api": getattr(work, "e2e_api", "galois_b200 host API (gfb_*_host pipeline)")}
        if hasattr(work, "e2e_pinned_step"):
            dtp = timed(work.e2e_pinned_step)
            e2e["pinned_value"] = e2e_units * world * e2e_steps / dtp
            e2e["pinned_api"] = "galois_b200.GF mirror class, operands and result in pooled pinned memory"
        if hasattr(work, "e2e_teardown"):
            work.e2e_teardown()

    # ---- secondary workloads (default run, EVERY N): the other BASELINE configs, device-resident, strong-scaled
    #      over the ranks where the config shards (NTT batch 4096, RS 1e6 codewords), each with its own roofline
    #      fraction and the time of the optional final gather
    extras = {}
    if args.workload == "gfp_mul" and not args.no_extras:
        for name in ("x", "y", "out", "hx", "hy", "hz", "px", "py", "pz"):
            if hasattr(work, name):
                delattr(work, name)
        torch.cuda.empty_cache()
        for key, ecls, kw in (("ntt", Ntt, {"batch": args.extras_ntt_batch}), ("rs_decode", RsDecode, {"total": 10**6}),
                              ("gf2_128_mul", Gf128Mul, {})):
            try:
                w = ecls(rank, world, **kw)
                esteps = 5
                ems = time_steps(w.step, esteps, 3, world)
                w.check()
                eunits = all_units(w.units_per_step, world)
                ach = w.units_per_step * ecls.bytes_per_unit / (ems / esteps / 1e3) / 1e9
                extras[key] = {"metric": ecls.metric, "value": eunits * esteps / (ems / 1e3), "unit": ecls.unit,
                               "workload": w.name, "scaling": ecls.scaling, "ms_per_step": ems / esteps,
                               "units_per_gpu_per_step": w.units_per_step,
                               "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                            "kernel": ecls.kernel, "algorithmic_bytes_per_unit": ecls.bytes_per_unit}}
                if callable(getattr(w, "secondary_roofline", None)):
                    extras[key]["roofline_secondary"] = w.secondary_roofline(w.units_per_step * esteps / (ems / 1e3))
                if world > 1 and ecls.scaling == "strong":
                    extras[key]["final_gather_ms"] = time_gather(w, world)
                del w
                torch.cuda.empty_cache()
            except Exception as exc:  # a secondary workload must not take the headline down
                extras[key] = {"error": f"{type(exc).__name__}: {exc}"}

    if rank != 0:
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- CPU baseline on the host cores (rank 0, N = 1 only), bounded samples: the unmodified reference when it is
    #      importable (single thread, as the reference runs), and the OpenMP oracle port as a second figure
    cpu = port = None
    if world == 1 and not args.no_cpu:
        try:
            v, reps, secs, threads = time_cpu(work)
            port = {"value": v, "unit": cls.unit, "cores": getattr(cls, "cpu_cores", threads), "kind": "port",
                    "sample": f"{cls.cpu_sample}; {reps} repetitions in {secs:.1f} s"}
        except Exception as exc:
            port = {"error": f"{type(exc).__name__}: {exc}"}
        if galois is not None and has_ref(cls):
            v, reps, secs = time_reference_live(work, galois)
            cpu = {"value": v, "unit": cls.unit, "cores": 1, "kind": "reference",
                   "sample": f"{cls.ref_sample}; {reps} repetitions in {secs:.1f} s"}
        else:
            cpu = port

    med = statistics.median(per_step) if per_step else None
    line = {
        "metric": cls.metric, "value": value, "unit": cls.unit, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "ms_per_step_median": med, "higher_is_better": True,
        "scaling": cls.scaling, "vs_baseline": None, "dtype": cls.dtype, "data": "synthetic",
        "config": {"workload": workload_name, "parallelism": f"shard{world}",
                   "l2": "inputs larger than L2 (no flush needed)" if not use_graph else "working set inside L2 (launch-bound size), K steps replayed from one CUDA graph",
                   "units_per_gpu_per_step": per_gpu_units},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": recorded_traffic(args.workload, per_gpu_units), "kernel": cls.kernel, "peak_source": peak_src,
                     "frac_of_nominal": achieved / 8000.0,  # HBM3e nominal 8.0 TB/s, alongside the measured peak
                     "algorithmic_bytes_per_unit": cls.bytes_per_unit,
                     "algorithmic_bytes_per_step": cls.bytes_per_unit * per_gpu_units},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    if sec_roofline is not None:
        line["roofline_secondary"] = sec_roofline
    if port is not None and port is not cpu:
        line["cpu_port"] = port
    if med is not None:
        line["value_at_median"] = units_all / (med / 1e3)
    if getattr(work, "phases", None):
        ph = work.phases[-args.steps:]
        line["phases_ms"] = {k: statistics.median(d[k] for d in ph) for k in ph[0]}
    if gather_ms is not None:
        line["final_gather_ms"] = gather_ms
    if extras:
        line["extras"] = extras
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
