"""Shared helpers for the parity tests: field-name parsing and oracle construction."""

from __future__ import annotations

import json
import re
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"

# irreducible polynomial / primitive element of every binary field the fixtures use
# (values recorded from the live reference by tests/golden/make_golden.py)
_BINARY_META: dict[int, dict] = {}


def _load_binary_meta():
    if _BINARY_META:
        return
    m1 = json.load(open(GOLDEN / "ref_fields_meta.json"))
    m2 = json.load(open(GOLDEN / "live_elementwise_meta.json"))
    for name, d in m1.items():
        if d["characteristic"] == 2:
            _BINARY_META[d["degree"]] = {"irreducible_poly": d["irreducible_poly"], "primitive_element": d["primitive_element"]}
    for name, d in m2.items():
        m = int(re.match(r"GF\(2\^(\d+)\)", name).group(1))
        _BINARY_META[m] = d
    _BINARY_META[1] = {"irreducible_poly": 3, "primitive_element": 1}


def parse_field(name: str) -> dict:
    """'GF(2^8)' → {p: 2, m: 8, poly, alpha};  'GF(31)' → {p: 31, m: 1}."""
    _load_binary_meta()
    mt = re.match(r"GF\((\d+)(?:\^(\d+))?\)$", name)
    if not mt:
        raise ValueError(name)
    p = int(mt.group(1))
    m = int(mt.group(2) or 1)
    if p == 2:
        meta = _BINARY_META[m]
        return {"p": 2, "m": m, "poly": int(meta["irreducible_poly"]), "alpha": int(meta["primitive_element"])}
    if m != 1:
        raise ValueError("odd-characteristic extension fields are out of scope")
    return {"p": p, "m": 1, "poly": 0, "alpha": None}


def oracle_field(name: str, lookup: bool | None = None):
    """The oracle's field object.  lookup=None mimics the reference default: tables when order <= 2^20."""
    import oracle

    spec = parse_field(name)
    if spec["p"] != 2:
        return oracle.Field.prime(spec["p"])
    if spec["m"] > 64:
        return oracle.Field128(spec["m"], spec["poly"])
    if lookup is None:
        lookup = spec["m"] <= 16
    return oracle.Field.binary(spec["m"], spec["poly"], alpha=spec["alpha"] if lookup else None)


def limbs_to_obj(limbs: np.ndarray) -> np.ndarray:
    """[..., 2] uint64 → object array of Python ints."""
    flat = limbs.reshape(-1, 2)
    out = np.empty(flat.shape[0], dtype=object)
    for i, (lo, hi) in enumerate(flat):
        out[i] = int(lo) | (int(hi) << 64)
    return out.reshape(limbs.shape[:-1])
