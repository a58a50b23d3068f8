"""The profile-summarising tools parse the committed ncu artefacts (profiles/) — the numbers DESIGN.md / bench.py quote
come out of these scripts, so a parsing regression would silently change them."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_launch_list_family_classifier():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import ncu_launch_list as T
    finally:
        sys.path.pop(0)
    cases = {
        "void r3m::<unnamed>::conv_igemm_kernel<256, 3, 2, 8, 1, 1, 0, 0>(CUtensorMap_st, int)": "conv_igemm",
        "void r3m::<unnamed>::halo3x3_kernel(CUtensorMap_st)": "conv_igemm",
        "void r3m::<unnamed>::wgrad_pair_kernel(CUtensorMap_st)": "wgrad",
        "void r3m::<unnamed>::wgrad_reduce_kernel<8>(float*)": "wgrad",
        "void r3m::<unnamed>::bn_bwd_apply_kernel<0, 2, 1>(r3m::BnBwdArgs)": "norm",
        "void r3m::<unnamed>::stem_bwd_kernel<1>(r3m::StemBwdArgs)": "norm",
        "void r3m::<unnamed>::stem_pool_kernel(r3m::StemPoolArgs)": "pool",
        "void r3m::<unnamed>::adam_kernel(float*)": "optim",
        "void r3m::<unnamed>::loss_tcn_kernel<0>(float*)": "loss",
        "void r3m::<unnamed>::sgemm_kernel<1, 0>(float*)": "lang",
    }
    for name, fam in cases.items():
        assert T.family(T.short(name)) == fam, (name, T.family(T.short(name)))
    assert T.short("void r3m::<unnamed>::adam_kernel(float*)") == "adam_kernel"


def test_traffic_json_matches_what_bench_reads():
    """bench.py's roofline.traffic = profiles/r2_traffic.json -> families[f].dram_bytes_per_launch; the dominant family's
    DRAM bytes per step must stay in the neighbourhood of its algorithmic bytes (51.7 GB for the c3 step)."""
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    fams = d["families"]
    for f in ("norm", "conv_igemm", "wgrad"):
        assert fams[f]["ops"] > 0 and fams[f]["dram_bytes_per_launch"] > 0
        assert abs(fams[f]["dram_bytes_per_launch"] * fams[f]["ops"] - fams[f]["dram_bytes_per_step"]) < 1e-3 * fams[f]["dram_bytes_per_step"]
    assert 0.8 * 51.7e9 < fams["norm"]["dram_bytes_per_step"] < 1.1 * 51.7e9


def test_launch_list_summary_runs_on_the_committed_capture(tmp_path):
    src = os.path.join(ROOT, "profiles", "r2_ncu_launch_list.csv")
    # the committed list is the reduced form (id, kernel, grid, block, duration_ns): rebuild an ncu-shaped CSV from it
    import csv

    rows = list(csv.DictReader(open(src)))
    assert len(rows) >= 700
    ncu = tmp_path / "launches.csv"
    with open(ncu, "w", newline="") as f:
        w = csv.writer(f, quoting=csv.QUOTE_ALL)
        w.writerow(["ID", "Process ID", "Process Name", "Host Name", "Kernel Name", "Context", "Stream", "Block Size",
                    "Grid Size", "Device", "CC", "Section Name", "Metric Name", "Metric Unit", "Metric Value"])
        for r in rows:
            w.writerow([r["id"], "1", "python", "h", "void r3m::<unnamed>::" + r["kernel"] + "(int)", "1", "7", r["block"],
                        r["grid"], "0", "10.0", "s", "gpu__time_duration.sum", "ns", r["duration_ns"]])
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_launch_list.py"), str(ncu),
                          os.path.join(ROOT, "profiles", "r2_bench_n1.json")], capture_output=True, text=True, check=True)
    text = out.stdout
    assert "| norm |" in text and "| conv_igemm |" in text and "| wgrad |" in text
    shares = {}
    for line in text.splitlines():
        cells = [c.strip() for c in line.strip("|").split("|")]
        if len(cells) == 5 and cells[0] in ("norm", "conv_igemm", "wgrad"):
            shares[cells[0]] = (float(cells[3].rstrip(" %")), float(cells[4].rstrip(" %")))
    # contract: the kernel families' SHARE of the step under ncu agrees with the live CUDA-event split
    for fam, (under_ncu, live) in shares.items():
        assert abs(under_ncu - live) < 5.0, (fam, under_ncu, live)
