"""Summarise `ncu -i X.ncu-rep --page source --csv` output: opcode histogram weighted by executed
instructions, stall-reason totals, and the hottest SASS lines. Usage: ncu_src_summary.py src.csv [top]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
# several kernels may be concatenated; take sections starting with a "Kernel Name" row
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for sec in sections[:1]:
    hdr, data = sec["hdr"], sec["data"]
    print(sec["name"][:120])
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    tot_ex = sum(int(r[iex]) for r in data)
    tot_s = sum(int(r[isamp]) for r in data)
    print("instructions executed", tot_ex, "samples", tot_s, "sass lines", len(data))
    c, cs = Counter(), Counter()
    for r in data:
        toks = r[isrc].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        c[op] += int(r[iex])
        cs[op] += int(r[isamp])
    for op, n in c.most_common(top):
        print(f"  {op:10s} exec {100*n/tot_ex:5.1f}%   samples {100*cs[op]/max(tot_s,1):5.1f}%")
    stalls = [(h, sum(int(r[i]) for r in data)) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    print("stall reasons (all samples):")
    for h, n in sorted(stalls, key=lambda x: -x[1])[:10]:
        print(f"  {h:24s} {100*n/max(tot_s,1):5.1f}%")
    print("hottest lines by samples:")
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:top]:
        st = sorted(((h, int(r[i])) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h), key=lambda x: -x[1])[:2]
        print(f"  {int(r[isamp]):6d} {100*int(r[isamp])/max(tot_s,1):4.1f}%  {r[isrc].strip()[:70]:70s} {st}")
