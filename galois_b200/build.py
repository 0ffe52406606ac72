"""
Build the in-tree native libraries with nvcc / gcc (no torch extension machinery: the product
is a plain C-ABI shared library, see include/galois_b200.h).

    python -m galois_b200.build            # build what is stale
    python -m galois_b200.build --force

Outputs (git-ignored, but they travel to the GPU box with the gpurun snapshot):
    galois_b200/lib/libgalois_b200.so      CUDA kernels + C ABI, sm_100a only
    oracle/_build/libgforacle.so           CPU restatement of the reference (test infrastructure)
"""

from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "galois_b200" / "csrc"
LIBDIR = ROOT / "galois_b200" / "lib"
OBJDIR = ROOT / "build" / "obj"
INCLUDE = ROOT / "include"
LIB = LIBDIR / "libgalois_b200.so"

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-O3",
    "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler",
    "-fPIC",
    f"-I{INCLUDE}",
    f"-I{CSRC}",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA toolkit is required to build galois_b200")


def _digest(paths: list[Path], extra: str = "") -> str:
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def _run(cmd: list[str]) -> None:
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        raise RuntimeError(f"command failed: {' '.join(cmd[:3])} ...")
    if proc.stderr.strip() and os.environ.get("GFB_BUILD_VERBOSE"):
        sys.stderr.write(proc.stderr)


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + sorted(INCLUDE.glob("*.h"))
    OBJDIR.mkdir(parents=True, exist_ok=True)
    LIBDIR.mkdir(parents=True, exist_ok=True)
    hdr_digest = _digest(headers, " ".join(NVCC_FLAGS))

    todo = []
    objs = []
    for src in sources:
        obj = OBJDIR / (src.stem + ".o")
        stamp = OBJDIR / (src.stem + ".sha")
        want = _digest([src], hdr_digest)
        objs.append(obj)
        if force or not obj.exists() or not stamp.exists() or stamp.read_text() != want:
            todo.append((src, obj, stamp, want))

    def compile_one(item):
        src, obj, stamp, want = item
        extra = ["-Xptxas", "-v"] if verbose else []
        _run([nvcc, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)])
        stamp.write_text(want)
        return src.name

    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(os.cpu_count() or 8, 16, len(todo))) as pool:
            for name in pool.map(compile_one, todo):
                print(f"[galois_b200.build] compiled {name}", flush=True)
    if todo or not LIB.exists():
        _run(
            [
                nvcc,
                "-shared",
                "-gencode",
                "arch=compute_100a,code=sm_100a",
                "-Xcompiler",
                "-fPIC",
                "-cudart",
                "static",
                "-o",
                str(LIB),
                *[str(o) for o in objs],
            ]
        )
        print(f"[galois_b200.build] linked {LIB.relative_to(ROOT)}", flush=True)
    return LIB


def build_oracle(force: bool = False) -> Path | None:
    """Compile oracle/'s C restatement (test infrastructure, never loaded by the product path)."""
    src = ROOT / "oracle" / "gf_oracle.c"
    if not src.exists():
        return None
    outdir = ROOT / "oracle" / "_build"
    outdir.mkdir(parents=True, exist_ok=True)
    out = outdir / "libgforacle.so"
    stamp = outdir / "libgforacle.sha"
    want = _digest([src])
    if force or not out.exists() or not stamp.exists() or stamp.read_text() != want:
        gcc = shutil.which("gcc") or "gcc"
        _run([gcc, "-O2", "-fPIC", "-shared", "-fopenmp", "-std=c11", "-o", str(out), str(src), "-lm"])
        stamp.write_text(want)
        print(f"[galois_b200.build] built {out.relative_to(ROOT)}", flush=True)
    return out


def main(argv: list[str]) -> int:
    force = "--force" in argv
    verbose = "--verbose" in argv
    build_cuda(force=force, verbose=verbose)
    build_oracle(force=force)
    return 0


if __name__ == "__main__":
    raise SystemExit(main(sys.argv[1:]))
