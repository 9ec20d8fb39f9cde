#!/usr/bin/env python
"""Per-region instruction / stall summary of an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv
--print-source sass; usage: ncu_source.py file.csv [section-index]).  Splits the chosen kernel's SASS at BAR.SYNC instructions and prints, per region,
executed warp instructions, sampled stalls and the top stall reasons and opcodes."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # kernel section index in the CSV
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"][which]
print("kernel:", rows[hdr_i - 1][1][:100] if hdr_i > 0 and rows[hdr_i - 1][0] == "Kernel Name" else "?")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
regions = [[]]
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) or r[0] == "Address" or r[0] == "Kernel Name":
        if r and r[0] == "Kernel Name":
            break
        continue
    regions[-1].append(r)
    if "BAR.SYNC" in r[col["Source"]]:
        regions.append([])
tot_inst = sum(int(r[col["Instructions Executed"]]) for reg in regions for r in reg)
tot_samp = sum(int(r[col["# Samples"]]) for reg in regions for r in reg)
print("total warp instructions %d, samples %d" % (tot_inst, tot_samp))
for i, reg in enumerate(regions):
    inst = sum(int(r[col["Instructions Executed"]]) for r in reg)
    samp = sum(int(r[col["# Samples"]]) for r in reg)
    st = Counter()
    ops = Counter()
    for r in reg:
        for s in stall_cols:
            st[s] += int(r[col[s]])
        op = r[col["Source"]].split()
        op = [o for o in op if not o.startswith("@")]
        ops[op[0].split(".")[0]] += int(r[col["Instructions Executed"]])
    print("region %d: %d SASS lines, inst %.1f%%, samples %.1f%%" % (i, len(reg), 100.0 * inst / tot_inst, 100.0 * samp / max(tot_samp, 1)))
    print("   stalls:", ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / max(samp, 1)) for k, v in st.most_common(6)))
    print("   ops:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(inst, 1)) for k, v in ops.most_common(10)))
