#!/usr/bin/env python
"""SASS evidence of the built library: per kernel the instruction count and the Blackwell / tensor / async mnemonics it
contains.  usage: python scripts/sass_summary.py [library] > profiles/<prefix>_sass_summary.txt   (CPU only: cuobjdump)"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "cryo_ralib_b200/libcryo_ralib.so"
KEYS = ["HMMA", "UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "LDG.E.128", "LDG.E.ENL2.256", "ST.E.ENL2.256",
        "ATOMS", "RED", "FFMA2", "FADD2", "FMUL2", "FMNMX", "DFMA", "BAR.SYNC"]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
print("SASS evidence of the built library (cuobjdump -sass %s; sm_100a, CUDA 12.9)." % so)
print("Per kernel: instruction count and the Blackwell / tensor / async mnemonics it contains")
print("(HMMA = mma.sync, UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / .st, UTCBAR = tcgen05.commit,")
print(" UBLKCP = cp.async.bulk (TMA unit), SYNCS = mbarrier, LDGSTS = cp.async, FFMA2 / FADD2 / FMUL2 = packed f32x2).\n")
name, cnt, n = None, collections.Counter(), 0


def flush():
    if name:
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
        d = re.sub(r"\(anonymous namespace\)::", "", d)
        d = re.sub(r"\(.*", "", d).replace("void ", "")
        print("%-44s %6d instr  %s" % (d[:44], n, "  ".join("%s %d" % (k, cnt[k]) for k in KEYS if cnt[k])))


for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        flush()
        name, cnt, n = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and name:
        op = m.group(1)
        n += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k.count(".") and op.startswith(k)):
                cnt[k] += 1
flush()
