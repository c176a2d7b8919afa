"""Where a kernel's warps stall, from the source page of a committed-elsewhere ncu report (no GPU needed):
    ncu -i REPORT.ncu-rep --page source --csv --print-source sass > src.csv
    python tools/ncu_hotspots.py src.csv k_admm k_ppass > profiles/<name>.md
For the LAST launch of every kernel whose name contains one of the given substrings: share of the warp-stall samples
per stall reason, per instruction class, and the instructions that collect the most samples."""
import csv
import re
import sys
from collections import Counter, defaultdict

CLASSES = [("DMMA", r"\bDMMA"), ("FP64 (DFMA/DADD/DMUL/DSETP/DMNMX)", r"\b(DFMA|DADD|DMUL|DSETP|DMNMX|MUFU)"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"),
           ("mbarrier (SYNCS)", r"\bSYNCS"), ("TMA (UTMA*/UBLKCP)", r"\b(UTMA|UBLKCP)"), ("bar.sync (BAR)", r"\bBAR\b"),
           ("global/local (LDG/STG/LDL/STL/LD/ST/ATOM/RED)", r"\b(LDG|STG|LDL|STL|LD|ST|ATOM|ATOMG|RED)\b"), ("shuffle / vote", r"\b(SHFL|VOTE|VOTEU)")]


def main():
    path, wanted = sys.argv[1], sys.argv[2:]
    launches, cur = [], None
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = dict(name=row[1], hdr=None, rows=[])
            launches.append(cur)
        elif row[0] == "Address" and cur is not None:
            cur["hdr"] = row
        elif cur is not None and cur["hdr"] is not None and row[0].startswith("0x"):
            cur["rows"].append(row)
    last = {}
    for l in launches:
        if any(w in l["name"] for w in wanted):
            last[l["name"]] = l
    for name, l in last.items():
        h = {k: i for i, k in enumerate(l["hdr"])}
        reasons = [k for k in l["hdr"] if k.startswith("stall_") and "Not Issued" not in k]
        tot = sum(int(r[h["# Samples"]]) for r in l["rows"])
        short = re.sub(r"\(int\)|\(bool\)|tritd::", "", name.split("(CUtensorMap")[0].split("(AdmmMaps")[0].replace("void ", ""))
        print(f"## `{short}` — {tot} warp-stall samples, {len(l['rows'])} SASS instructions\n")
        by_reason = Counter()
        for r in l["rows"]:
            for k in reasons:
                by_reason[k] += int(r[h[k]])
        rs = sum(by_reason.values())
        print("| stall reason | share of samples |\n|---|---|")
        for k, v in by_reason.most_common(8):
            print(f"| {k[6:]} | {100.0 * v / rs:.1f} % |")
        by_class, inst_class = Counter(), Counter()
        for r in l["rows"]:
            src = r[h["Source"]]
            cls = next((c for c, pat in CLASSES if re.search(pat, src)), "other (integer, address, control)")
            by_class[cls] += int(r[h["# Samples"]])
            inst_class[cls] += int(r[h["Instructions Executed"]])
        ti = sum(inst_class.values())
        print("\n| instruction class | samples | warp instructions executed |\n|---|---|---|")
        for c, v in by_class.most_common():
            print(f"| {c} | {100.0 * v / tot:.1f} % | {100.0 * inst_class[c] / ti:.1f} % |")
        print("\n| # | samples | top stall | instruction |\n|---|---|---|---|")
        top = sorted(l["rows"], key=lambda r: -int(r[h["# Samples"]]))[:12]
        for i, r in enumerate(top, 1):
            why = max(reasons, key=lambda k: int(r[h[k]]))
            print(f"| {i} | {100.0 * int(r[h['# Samples']]) / tot:.1f} % | {why[6:]} | `{' '.join(r[h['Source']].split())}` |")
        print()


if __name__ == "__main__":
    main()
