"""GPU-box debugging aid: the binding's convolve / matmul / Reed-Solomon calls step by step, with a watchdog that dumps
every thread's stack if a step does not return (a hang inside a C call cannot be interrupted by pytest-timeout)."""
import faulthandler
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
faulthandler.enable()
from refimport import import_reference  # noqa: E402

galois = import_reference()
import galois_b200 as gb  # noqa: E402


def step(name, fn, limit=120):
    faulthandler.dump_traceback_later(limit, exit=True)
    t0 = time.time()
    try:
        r = fn()
        print(f"[ok] {name}: {time.time() - t0:.2f}s", flush=True)
        return r
    except Exception as e:  # noqa: BLE001
        print(f"[EXC] {name}: {type(e).__name__}: {e}", flush=True)
        return None
    finally:
        faulthandler.cancel_dump_traceback_later()


def cmp(name, GF, fn):
    gb.disable(GF)
    want = step(name + " [reference]", fn)
    gb.enable(GF)
    got = step(name + " [cuda]", fn)
    gb.disable(GF)
    if want is None or got is None:
        return
    ok = type(got) is type(want) and np.array_equal(np.asarray(got).astype(object), np.asarray(want).astype(object)) and \
        getattr(got, "dtype", None) == getattr(want, "dtype", None)
    print(f"     {name}: {'EQUAL' if ok else 'DIFFERENT'} type={type(got).__name__}/{type(want).__name__} "
          f"dtype={getattr(got, 'dtype', None)}/{getattr(want, 'dtype', None)} shape={np.shape(got)}/{np.shape(want)}", flush=True)
    if not ok:
        print("     got ", np.asarray(got).ravel()[:8], "\n     want", np.asarray(want).ravel()[:8], flush=True)


for order in (998244353, 31):
    GF = galois.GF(order)
    a, b = GF.Random(300, seed=1), GF.Random(77, seed=2)
    A, B = GF.Random((40, 33), seed=3), GF.Random((33, 21), seed=4)
    cmp(f"convolve GF({order})", GF, lambda: np.convolve(a, b))
    cmp(f"convolve mode=same GF({order})", GF, lambda: np.convolve(a, b, mode="same"))
    cmp(f"matmul GF({order})", GF, lambda: A @ B)
    cmp(f"matmul mat-vec GF({order})", GF, lambda: A @ B[:, 0])
    cmp(f"matmul vec-vec GF({order})", GF, lambda: A[0, :] @ B[:, 0])

rs = galois.ReedSolomon(255, 223)
GF = rs.field
M = GF.Random((300, 223), seed=6)
cmp("rs.encode", GF, lambda: rs.encode(M))
C = rs.encode(M)
R = C.copy()
rng = np.random.default_rng(7)
for i in range(300):
    pos = rng.choice(255, int(rng.integers(0, 18)), replace=False)
    R[i, pos] ^= GF(rng.integers(1, 256, pos.size, dtype=np.uint8))
cmp("rs.decode", GF, lambda: rs.decode(R))
cmp("rs.decode errors", GF, lambda: rs.decode(R, errors=True)[1])
cmp("rs.decode 1-D", GF, lambda: rs.decode(R[0]))
cmp("rs.detect", GF, lambda: rs.detect(R))
era = np.zeros((300, 255), dtype=bool)
for i in range(300):
    era[i, rng.choice(255, 20, replace=False)] = True
R2 = C.copy()
R2[era] = 0
cmp("rs.decode erasures", GF, lambda: rs.decode(R2, erasures=era, errors=True)[0])
cmp("rs shortened", GF, lambda: rs.decode(rs.encode(M[:10, :100]), errors=True)[0])
print("done", flush=True)
