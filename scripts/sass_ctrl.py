#!/usr/bin/env python
"""Decode the scheduling control bits of sm_100 SASS (cuobjdump -sass prints the 128-bit encodings).

    python scripts/sass_ctrl.py <object or .so> <function-name substring> [opcode regex]

High 64-bit word, bits 41..63 (Volta and later): stall[4] yield[1] wbar[3] rbar[3] wait-mask[6] reuse[4].
Prints  addr  stall  wr-barrier  rd-barrier  wait-mask  instruction  -- which shows on which hardware scoreboard
each variable-latency instruction (LDG, LDS, tcgen05.ld ...) signals and which instruction drains which scoreboards.
"""
import re
import subprocess
import sys


def decode(path, func, opre=None):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout.splitlines()
    on = False
    rows = []
    i = 0
    while i < len(out):
        ln = out[i]
        if "Function :" in ln:
            on = func in ln
        elif on:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", ln)
            if m and i + 1 < len(out):
                m2 = re.search(r"/\* (0x[0-9a-f]+) \*/", out[i + 1])
                if m2:
                    hi = int(m2.group(1), 16)
                    ctrl = hi >> 41
                    stall = ctrl & 15
                    yld = (ctrl >> 4) & 1
                    wbar = (ctrl >> 5) & 7
                    rbar = (ctrl >> 8) & 7
                    wait = (ctrl >> 11) & 63
                    rows.append((m.group(1), stall, yld, wbar, rbar, wait, m.group(2).strip()))
                    i += 1
        i += 1
    for a, stall, yld, wbar, rbar, wait, ins in rows:
        if opre and not re.search(opre, ins) and wait == 0:
            continue
        w = "".join(str(b) if (wait >> b) & 1 else "-" for b in range(6))
        print("%s st%-2d %s W%s R%s wait[%s]  %s" % (a, stall, "Y" if yld else " ", wbar if wbar != 7 else "-",
                                                    rbar if rbar != 7 else "-", w, ins))
    return rows


if __name__ == "__main__":
    decode(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
