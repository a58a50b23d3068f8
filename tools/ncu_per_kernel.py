"""Turn an `ncu --csv` log of ONE training step into (a) a per-kernel markdown table and (b) the conv_igemm family's
DRAM traffic per launch (bench.py's roofline.traffic).  Usage (on the GPU box, then here):

  ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,\
launch__registers_per_thread -s <launches of the warm-up steps> -c <launches of one step> --csv \
--log-file gpurun_out/per_kernel.csv python tools/profile_step.py
  python tools/ncu_per_kernel.py gpurun_out/per_kernel.csv profiles/<round>_ncu_per_kernel.md profiles/<round>_conv_traffic.json
"""
import collections
import csv
import json
import re
import sys


def to_float(v):
    return float(v.replace(",", "")) if v not in ("", "n/a") else 0.0


def main(src, md, js):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(r[col["ID"]], {"kernel": r[col["Kernel Name"]]})
        unit, val = r[col["Metric Unit"]], to_float(r[col["Metric Value"]])
        name = r[col["Metric Name"]]
        if name.startswith("dram__bytes"):
            val *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        if name == "gpu__time_duration.sum":
            val *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(unit, 1)
        d[name] = val
    agg = collections.OrderedDict()
    for d in per.values():
        k = re.sub(r"^void ", "", d["kernel"])
        k = re.sub(r"\(.*$", "", k).replace("<unnamed>::", "").replace("unnamed>::", "").replace("r3m::", "")
        a = agg.setdefault(k, collections.defaultdict(float))
        a["n"] += 1
        t = d.get("gpu__time_duration.sum", 0.0)
        a["us"] += t
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["tensor_t"] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        a["dram_t"] += t * d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.0)
        a["regs"] = max(a["regs"], d.get("launch__registers_per_thread", 0.0))
        a["lts"] += d.get("lts__t_bytes.sum", 0.0)
    total = sum(a["us"] for a in agg.values())
    with open(md, "w") as f:
        f.write(f"captured {int(sum(a['n'] for a in agg.values()))} launches, {total / 1e3:.2f} ms (serialised, cold cache: "
                "compare shares)\n\n| kernel | launches | total us | share | dram read MB | dram write MB | HBM GB/s | "
                "of 6532 | dram % (ncu) | tensor % | L2 GB/s | regs |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            gbs = (a["rd"] + a["wr"]) / a["us"] / 1e3 if a["us"] else 0
            f.write(f"| `{k}` | {int(a['n'])} | {a['us']:.0f} | {100 * a['us'] / total:.1f} % | {a['rd'] / 1e6:.0f} | "
                    f"{a['wr'] / 1e6:.0f} | {gbs:.0f} | {gbs / 6532:.2f} | {a['dram_t'] / a['us']:.0f} | "
                    f"{a['tensor_t'] / a['us']:.1f} | {a['lts'] / a['us'] / 1e3:.0f} | {int(a['regs'])} |\n")
    # DRAM traffic per kernel family, in the shape bench.py's roofline.traffic reads (per engine op: a wgrad op is the
    # tcgen05 kernel plus its split-K reduce)
    fam_of = [("conv_igemm_kernel", "conv_igemm"), ("halo3x3_kernel", "conv_igemm"), ("wgrad_kernel", "wgrad"),
              ("wgrad_pair_kernel", "wgrad"), ("wgrad_reduce_kernel", "wgrad"),
              ("bn_", "norm"), ("stem_bwd", "norm"), ("preprocess_stem", "norm"), ("stem_pool", "pool"),
              ("avgpool", "pool"), ("adam", "optim"), ("pack_dgrad", "optim"), ("stem_pack", "optim"),
              ("stem_unpack", "optim"), ("cast_bf16", "optim"), ("sgemm", "lang"), ("lang_", "lang"),
              ("splitk_reduce", "lang"), ("col_sum", "lang"), ("vec_sum", "lang"), ("loss_", "loss"),
              ("lp_finalize", "loss"), ("tcn_finalize", "loss"), ("publish_flag", "loss")]
    ops_kernel = {"wgrad": ("wgrad_kernel", "wgrad_pair_kernel")}  # families whose ops are counted on their main kernels
    fams = collections.OrderedDict()
    for k, a in agg.items():
        fam = next((f for pre, f in fam_of if k.startswith(pre)), "other")
        if k.startswith("conv_igemm_kernel") and re.search(r", 1, [01]>$", k):
            fam = "lang"  # the tf32 tier inside a training step: the language head's split-tf32 GEMMs
        r = fams.setdefault(fam, {"launches": 0, "ops": 0, "dram_bytes_per_step": 0.0, "l2_bytes_per_step": 0.0,
                                  "us_under_ncu": 0.0})
        r["launches"] += int(a["n"])
        if fam not in ops_kernel or k.split("<")[0] in ops_kernel[fam]:
            r["ops"] += int(a["n"])
        r["dram_bytes_per_step"] += a["rd"] + a["wr"]
        r["l2_bytes_per_step"] += a["lts"]
        r["us_under_ncu"] += a["us"]
    for r in fams.values():
        r["dram_bytes_per_launch"] = r["dram_bytes_per_step"] / max(1, r["ops"])
    out = {"source": src, "note": "ncu --clock-control none, one ResNet-50 c3 step (tools/profile_step.py); cold cache and "
           "serialised: DRAM bytes are an upper bound of the in-step traffic", "families": fams}
    json.dump(out, open(js, "w"), indent=1)
    print({f: (r["ops"], round(r["dram_bytes_per_step"] / 1e9, 2)) for f, r in fams.items()})


if __name__ == "__main__":
    main(*sys.argv[1:4])
