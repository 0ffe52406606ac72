#!/usr/bin/env python
"""Merge gpurun_out/<tag>_bench_changed.jsonl into profiles/<tag>_bench_all.jsonl (key: metric + impl; new lines win),
write gpurun_out/<tag>_bench_all.jsonl for scripts/bench_table.py."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"


def load(path):
    return [json.loads(l) for l in open(path) if l.strip().startswith("{")] if path.exists() else []


def key(r):
    return (r["metric"], r.get("impl", "b200"))


merged = {key(r): r for r in load(ROOT / "profiles" / f"{tag}_bench_all.jsonl")}
new = load(ROOT / "gpurun_out" / f"{tag}_bench_changed.jsonl")
for r in new:
    merged[key(r)] = r
with open(ROOT / "gpurun_out" / f"{tag}_bench_all.jsonl", "w") as f:
    for r in merged.values():
        f.write(json.dumps(r) + "\n")
print(f"merged {len(new)} new line(s); {len(merged)} total")
