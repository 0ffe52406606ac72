#!/usr/bin/env python
"""
Summarise the ncu exports that scripts/profile_gpu.sh left in gpurun_out/ into profiles/:

    profiles/<tag>_launches.csv        every launch of `bench.py --steps 2 --warmup 3` with its device time
    profiles/<tag>_<kernel>_raw.csv    ncu --set full, raw metric page (one row per captured launch)
    profiles/<tag>_summary.md          the numbers DESIGN.md / bench.py cite
    profiles/<tag>_traffic.json        dram bytes per launch, read by bench.py for `roofline.traffic`

    python scripts/summarize_ncu.py r01
"""
import collections
import csv
import json
import shutil
import sys
from pathlib import Path

csv.field_size_limit(10**9)
ROOT = Path(__file__).resolve().parent.parent
SRC, DST = ROOT / "gpurun_out", ROOT / "profiles"
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__grid_size", "launch__block_size"]


def raw_rows(fn):
    rows = list(csv.reader(open(fn)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[hi], rows[hi + 1]
    for r in rows[hi + 2:]:
        if len(r) >= len(hdr):
            yield dict(zip(hdr, r)), dict(zip(hdr, units))


def main(tag):
    DST.mkdir(exist_ok=True)
    lines = [f"# ncu summary {tag}", "",
             "Source: `scripts/profile_gpu.sh` under gpurun on one B200 (`--clock-control none`).  Durations under ncu are",
             "cold-cache and serialised: use them for shares and per-launch traffic, not as benchmark values.", ""]
    traffic = {}
    workload_of = {"mul": "gfp_mul", "ntt": "ntt", "rs": "rs_decode", "gf128": "gf2_128_mul", "gf256": "gf256_mul",
                   "ntt_fused": "ntt", "elim": "matinv_panel"}
    for name in ("mul", "ntt", "rs", "gf256", "gf128", "elim"):
        fn = SRC / f"{tag}_{name}_raw.csv"
        if not fn.exists():
            continue
        shutil.copy(fn, DST / fn.name)
        det = SRC / f"{tag}_{name}_details.csv"
        if det.exists():
            shutil.copy(det, DST / det.name)
        lines += [f"## {name}", "", "| kernel | time | dram read | dram write | regs | warps active % | issue active % | SM thr % | ALU pipe % | FMA-heavy pipe % | inst (warp) | smem bank conflicts / wavefronts |", "|---|---|---|---|---|---|---|---|---|---|---|---|"]
        tot = 0.0
        units = {"mul": 10**8, "rs": 200000, "gf128": 10**7, "gf256": 10**8, "elim": 1}.get(name)  # elements / codewords per captured launch
        for d, u in raw_rows(fn):
            if name in ("ntt", "ntt_fused"):
                if "fused" in d["Kernel Name"]:
                    units = 592  # the persistent kernel: one launch for the whole --batch 592 of the capture command
                else:
                    units = int(float(d["launch__grid_size"])) // 64  # one CTA per (transform, 16-wide tile)
            t = float(d["gpu__time_duration.sum"]) * UNIT[u["gpu__time_duration.sum"]]
            rd = float(d["dram__bytes_read.sum"]) * UNIT[u["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"]) * UNIT[u["dram__bytes_write.sum"]]
            tot += rd + wr
            kn = d["Kernel Name"].split("(")[0].replace("void ", "")[:60]
            lines.append(f"| `{kn}` | {t*1e6:.1f} us | {rd/1e6:.1f} MB | {wr/1e6:.1f} MB | {d['launch__registers_per_thread']} | "
                         f"{float(d['sm__warps_active.avg.pct_of_peak_sustained_active']):.1f} | "
                         f"{float(d['smsp__issue_active.avg.pct_of_peak_sustained_active']):.1f} | "
                         f"{float(d['sm__throughput.avg.pct_of_peak_sustained_elapsed']):.1f} | "
                         f"{float(d.get('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active') or 0):.1f} | "
                         f"{float(d.get('sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed') or 0):.1f} | "
                         f"{float(d['smsp__inst_executed.sum']):.3g} | "
                         f"{d['l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']} / {d['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']} |")
        # dram bytes of the captured launch(es) (ntt: pass 1 + pass 2 of one chunk) and the units they processed
        traffic[workload_of[name]] = {"dram_bytes": tot, "units": units, "dram_bytes_per_unit": tot / units}
        lines.append("")
    # launch list: share of each kernel in the default bench command
    fn = SRC / f"{tag}_launches_default.csv"
    if fn.exists():
        shutil.copy(fn, DST / f"{tag}_launches.csv")
        rows = list(csv.reader(open(fn)))
        hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        hdr = rows[hi]
        agg = collections.defaultdict(lambda: [0, 0.0])
        for r in rows[hi + 1:]:
            d = dict(zip(hdr, r))
            if d.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(d["Metric Value"].replace(",", "")) * UNIT.get(d["Metric Unit"], 1e-9)
            kn = d["Kernel Name"].split("(")[0].replace("void ", "")[:70]
            agg[kn][0] += 1
            agg[kn][1] += v
        total = sum(v for _, v in agg.values())
        lines += ["## launch list of `python bench.py --steps 2 --warmup 3 --no-cpu` (all kernels, incl. torch data generation)", "",
                  "| kernel | launches | total time | share |", "|---|---|---|---|"]
        for kn, (cnt, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
            lines.append(f"| `{kn}` | {cnt} | {v*1e3:.3f} ms | {100*v/total:.1f}% |")
        lines.append("")
    (DST / f"{tag}_summary.md").write_text("\n".join(lines))
    json.dump(traffic, open(DST / f"{tag}_traffic.json", "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
