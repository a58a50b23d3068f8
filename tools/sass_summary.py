"""`cuobjdump -sass` of the shipped library -> per-kernel counts of the Blackwell-native mnemonics (tcgen05 MMA: UTCHMMA /
UTCQMMA..., TMEM load: LDTM, TMA loads / stores / reductions: UTMALDG / UTMASTG / UTMAREDG) and of the legacy tensor
ops that must NOT appear (HMMA).  Usage: python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "r3m_b200", "libr3m_b200.so")
PATTERNS = ["UTCHMMA", "UTCMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "HMMA", "ATOMG", "REDG", "RED."]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        for p in PATTERNS:
            if re.search(r"(?<![A-Z])" + re.escape(p), line):  # whole mnemonics: "HMMA" must not match "UTCHMMA"
                cur[p] += 1
        if re.search(r"\bUTMALDG\S*IM2COL", line):
            cur["UTMALDG.IM2COL"] += 1
        if "UTCHMMA" in line or "UTCMMA" in line:
            cur["tcgen05.mma"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} ({len(kernels)} kernels, sm_100a)\n")
    print("| kernel | tcgen05.mma (UTC*MMA) | LDTM | UTMALDG (im2col) | UTMASTG | UTMAREDG | HMMA | ATOMG/RED |")
    print("|---|---|---|---|---|---|---|---|")
    for (name, c), dn in zip(kernels.items(), demangle):
        total.update(c)
        short = dn.replace("r3m::(anonymous namespace)::", "").replace("void ", "")
        short = re.sub(r"\((CUtensorMap_st|r3m::|float|__nv|unsigned|int|void).*$", "", short)
        if not any(c[k] for k in ("tcgen05.mma", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA")):
            continue
        print(f"| `{short}` | {c['tcgen05.mma']} | {c['LDTM']} | {c['UTMALDG']} ({c['UTMALDG.IM2COL']}) | {c['UTMASTG']} | "
              f"{c['UTMAREDG']} | {c['HMMA']} | {c['ATOMG'] + c['REDG'] + c['RED.']} |")
    print(f"\ntotals: tcgen05.mma {total['tcgen05.mma']}, LDTM {total['LDTM']}, UTMALDG {total['UTMALDG']} "
          f"(im2col {total['UTMALDG.IM2COL']}), UTMASTG {total['UTMASTG']}, UTMAREDG {total['UTMAREDG']}, HMMA {total['HMMA']}")
    assert total["HMMA"] == 0, "legacy mma.sync tensor instructions found"


if __name__ == "__main__":
    sys.exit(main())
