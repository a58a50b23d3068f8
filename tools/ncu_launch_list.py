"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X ...`) of the
bench command: per kernel and per kernel family, launches / total time / share.  Times under ncu are serialised and
cold-cache, so only the SHARES are comparable with the live CUDA-event profile that bench.py prints (`roofline.families`).

usage: python tools/ncu_launch_list.py gpurun_out/launches.csv [bench_line.json] > profiles/rN_ncu_launch_list_summary.md
"""
import collections
import csv
import json
import re
import sys

FAMILIES = [  # first match wins
    ("wgrad", r"wgrad"),
    ("lang", r"sgemm|lang_|infonce_lang|col_sum|factor"),
    ("conv_igemm", r"conv_igemm_kernel|halo3x3"),
    ("norm", r"bn_|stem_bwd|preprocess_stem"),
    ("pool", r"stem_pool|avgpool"),
    ("loss", r"loss_|tcn|lp_"),
    ("optim", r"adam|pack_dgrad|pack_"),
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"r3m::(<unnamed>::)?", "", name)
    return re.sub(r"\(.*$", "", name)


def family(name):
    for fam, pat in FAMILIES:
        if re.search(pat, name):
            return fam
    return "other"


def main():
    rows = []
    with open(sys.argv[1], newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
            rows.append((short(r["Kernel Name"]), ns))
    total = sum(ns for _, ns in rows)
    per_k = collections.defaultdict(lambda: [0, 0.0])
    per_f = collections.defaultdict(lambda: [0, 0.0])
    for k, ns in rows:
        per_k[k][0] += 1
        per_k[k][1] += ns
        fam = family(k)
        per_f[fam][0] += 1
        per_f[fam][1] += ns
    live = None
    if len(sys.argv) > 2:
        line = json.load(open(sys.argv[2]))
        fams = line["roofline"]["families"]
        live_total = sum(v["ms"] for v in fams.values())
        live = {k: v["ms"] / live_total for k, v in fams.items()}
    print(f"captured launches: {len(rows)}, total {total / 1e6:.2f} ms (serialised, cold cache: compare shares)\n")
    print("| family | launches | total us | share under ncu | share in the live step (bench.py CUDA events) |")
    print("|---|---|---|---|---|")
    for fam, (n, ns) in sorted(per_f.items(), key=lambda kv: -kv[1][1]):
        lv = f"{100 * live[fam]:.1f} %" if live and fam in live else ""
        print(f"| {fam} | {n} | {ns / 1e3:.1f} | {100 * ns / total:.1f} % | {lv} |")
    print("\n| kernel | family | launches | total us | share |")
    print("|---|---|---|---|---|")
    for k, (n, ns) in sorted(per_k.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {family(k)} | {n} | {ns / 1e3:.1f} | {100 * ns / total:.1f} % |")


if __name__ == "__main__":
    main()
