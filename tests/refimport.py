"""Import the UNMODIFIED reference from baseline/_ref (installed by `pip install --no-deps --target baseline/_ref
/root/reference`).  baseline/_ref is git-ignored but travels to the GPU box with the snapshot."""

from __future__ import annotations

import os
import sys
import warnings
from pathlib import Path

REF_DIR = Path(__file__).resolve().parent.parent / "baseline" / "_ref"


def import_reference():
    """The reference package `galois`, or None when it is not installed / not importable (numba missing)."""
    if not (REF_DIR / "galois").is_dir():
        return None
    if str(REF_DIR) not in sys.path:
        sys.path.insert(0, str(REF_DIR))
    os.environ.setdefault("NUMBA_CACHE_DIR", str(REF_DIR.parent / "_numba_cache"))
    # two of the reference's kernels use prange (matmul_jit, evaluate_elementwise_jit, called once per Chien position
    # by the decoder): with Numba's default of one worker per core a 128-core host spends milliseconds per tiny region
    os.environ.setdefault("NUMBA_NUM_THREADS", "8")
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import galois
    except Exception:  # noqa: BLE001  (numba / llvmlite absent on this machine)
        return None
    return galois
