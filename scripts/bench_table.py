#!/usr/bin/env python
"""Render gpurun_out/<tag>_bench_all.jsonl as profiles/<tag>_bench_all.md (and keep the jsonl beside it)."""
import json
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src = ROOT / "gpurun_out" / f"{tag}_bench_all.jsonl"
rows = [json.loads(l) for l in open(src) if l.strip().startswith("{")]
ours = {r["metric"]: r for r in rows if r.get("impl") != "reference"}
ref = {r["metric"]: r for r in rows if r.get("impl") == "reference"}
out = [f"# bench_all {tag}: every workload, device-resident value / end-to-end (host buffers) / CPU reference arm", "",
       "One B200, `python bench.py --workload W --steps 20 --warmup 3`; reference arm (`--impl reference`) = the unmodified",
       "reference from baseline/_ref on one thread (kind `reference`) where the workload has a galois-driven leg, the OpenMP oracle",
       "port otherwise (kind `port`).  `frac` = achieved algorithmic GB/s over the measured HBM copy peak.", "",
       "| workload | value | unit | ms/step | alg. GB/s | frac | e2e (host API) | CPU ref arm | kind | cores | value/CPU | e2e/CPU |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|"]
for m, r in ours.items():
    c = ref.get(m)
    e2e = (r.get("e2e") or {}).get("value")
    cv = c["value"] if c else None
    out.append(f"| {r['config']['workload']} | {r['value']:.4g} | {r['unit']} | {r['ms_per_step']:.4g} | "
               f"{r['roofline']['achieved']:.0f} | {r['roofline']['frac']:.3f} | {e2e:.4g} | " if e2e else
               f"| {r['config']['workload']} | {r['value']:.4g} | {r['unit']} | {r['ms_per_step']:.4g} | "
               f"{r['roofline']['achieved']:.0f} | {r['roofline']['frac']:.3f} | n/a | ")
    out[-1] += (f"{cv:.4g} | {c['cpu_baseline']['kind']} | {c['cpu_baseline']['cores']} | {r['value']/cv:.1f}x | " + (f"{e2e/cv:.1f}x |" if e2e else "n/a |")) if cv else "n/a | | | | |"
(ROOT / "profiles" / f"{tag}_bench_all.md").write_text("\n".join(out) + "\n")
shutil.copy(src, ROOT / "profiles" / src.name)
print("\n".join(out))
