#!/usr/bin/env python
"""Turn the raw ncu outputs in gpurun_out/ into the small tracked summaries under profiles/.

    python scripts/summarize_ncu.py <tag>        # e.g. r1
Reads  gpurun_out/launches.csv           (ncu --metrics gpu__time_duration.sum launch list of bench.py)
       gpurun_out/prof_hotpath.ncu-rep   (ncu --set full capture of the hot kernels)
Writes profiles/<tag>_launches_summary.md, profiles/<tag>_hotpath_metrics.csv,
       profiles/roofline_traffic.json (per-launch DRAM bytes of k_eval, read by bench.py)
"""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)


def short(name):
    m = re.match(r"(?:void )?(?:bh::)?(\w+)", name)
    n = m.group(1) if m else name
    t = re.search(r"<(\d+), ?(\d+)", name)
    return f"{n}<{t.group(1)},{t.group(2)}>" if t else n


# ---- launch list -----------------------------------------------------------
lp = os.path.join(ROOT, "gpurun_out", "launches.csv")
if os.path.exists(lp):
    lines = [l for l in open(lp) if not l.startswith("==")]
    rows = list(csv.reader(io.StringIO("".join(lines))))
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    agg = defaultdict(list)
    order = []
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v = v / 1000.0 if r[iu] in ("ns", "nsecond") else v      # -> us
        agg[short(r[ik])].append(v)
        order.append(short(r[ik]))
    total = sum(sum(v) for v in agg.values())
    with open(os.path.join(out_dir, f"{tag}_launches_summary.md"), "w") as f:
        f.write(f"# ncu launch list of `python bench.py --steps 2 --warmup 1 --rollout 16 --no-cpu-baseline` ({tag})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 900`; per-launch times are "
                "cold-cache and serialised: compare shares, not absolutes.\n\n")
        f.write("| kernel | launches | total us | share | mean us | min us | max us |\n|---|---|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| {k} | {len(v)} | {sum(v):.1f} | {100 * sum(v) / total:.1f}% | {sum(v) / len(v):.2f} | {min(v):.2f} | {max(v):.2f} |\n")
        # share inside one env step: consecutive k_eval, k_commit pairs
        ev = [x for k, v in agg.items() if k.startswith("k_eval_t") for x in v]
        cm = [x for k, v in agg.items() if k.startswith("k_commit") for x in v]
        if ev and cm:
            f.write(f"\nInside an env step (k_eval + k_commit): k_eval mean {sum(ev) / len(ev):.2f} us, "
                    f"k_commit mean {sum(cm) / len(cm):.2f} us -> k_eval share "
                    f"{100 * (sum(ev) / len(ev)) / (sum(ev) / len(ev) + sum(cm) / len(cm)):.1f}%\n")
    print("wrote launches summary:", len(order), "launches")

# ---- full capture ------------------------------------------------------------
def raw_page(stem):
    """Raw metric page of a capture: the CSV written on the GPU box, else the .ncu-rep itself."""
    csv_path = os.path.join(ROOT, "gpurun_out", stem + "_raw.csv")
    rep_path = os.path.join(ROOT, "gpurun_out", stem + ".ncu-rep")
    if os.path.exists(csv_path):
        return open(csv_path).read()
    if os.path.exists(rep_path):
        return subprocess.run(["ncu", "-i", rep_path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return None


EXTRA = [("prof_prop", "propagation"), ("prof_eval_warm", "eval_warm_cache"), ("prof_bundle", "eval_bundle")]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
raw = raw_page("prof_hotpath")
if raw:
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {k: hdr.index(k) for k in KEYS if k in hdr}
    ik = hdr.index("Kernel Name")
    with open(os.path.join(out_dir, f"{tag}_hotpath_metrics.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{k} [{units[i]}]" for k, i in idx.items()])
        for r in data:
            w.writerow([short(r[ik])] + [r[i] for i in idx.values()])
    # per-launch DRAM traffic of k_eval (steady-state launches)
    ev = [r for r in data if short(r[ik]).startswith("k_eval_t")]
    if ev:
        def mb(r, key):
            i = hdr.index(key)
            v = float(r[i].replace(",", ""))
            u = units[i].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        tr = [mb(r, "dram__bytes_read.sum") + mb(r, "dram__bytes_write.sum") for r in ev]
        dur = [float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")) for r in ev]
        json.dump({"k_eval_dram_bytes_per_launch": sum(tr) / len(tr), "launches": len(tr),
                   "k_eval_duration_under_ncu": dur, "unit_duration": units[hdr.index("gpu__time_duration.sum")],
                   "source": f"profiles/{tag}_hotpath_metrics.csv (ncu --set full, 8 candidates per launch)"},
                  open(os.path.join(out_dir, "roofline_traffic.json"), "w"), indent=1)
    print("wrote hotpath metrics:", len(data), "launches")

for stem, label in EXTRA:
    raw = raw_page(stem)
    if not raw:
        continue
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {k: hdr.index(k) for k in KEYS if k in hdr}
    ik = hdr.index("Kernel Name")
    with open(os.path.join(out_dir, f"{tag}_{label}_metrics.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{k} [{units[i]}]" for k, i in idx.items()])
        for r in data:
            w.writerow([short(r[ik])] + [r[i] for i in idx.values()])
    if label == "eval_warm_cache":
        def mb(r, key):
            i = hdr.index(key)
            v = float(r[i].replace(",", ""))
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(units[i].lower(), 1)
        tr = [mb(r, "dram__bytes_read.sum") + mb(r, "dram__bytes_write.sum") for r in data]
        tp = os.path.join(out_dir, "roofline_traffic.json")
        cur = json.load(open(tp)) if os.path.exists(tp) else {}
        cur["k_eval_dram_bytes_per_launch_warm_cache"] = sum(tr) / len(tr)
        cur["note"] = ("cold = ncu default cache control (L2 flushed before every replay: the 25 MB impulse "
                       "table is re-read from DRAM); warm = --cache-control none (h stays L2 resident as in the live run)")
        json.dump(cur, open(tp, "w"), indent=1)
    print("wrote", label, len(data), "launches")
