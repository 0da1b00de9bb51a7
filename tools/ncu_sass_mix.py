"""Execution-weighted SASS opcode mix of one kernel of an .ncu-rep (source page), plus the hottest stall samples.
   python tools/ncu_sass_mix.py file.ncu-rep [n_top_opcodes]"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ops, samples = Counter(), Counter()
    total = tot_s = 0
    body = []
    for r in rows:
        if not r or not r[0].startswith("0x"):
            continue
        src = r[iS].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
        op = m.group(2) if m else "?"
        full = op
        n, s = int(r[iE] or 0), int(r[iSm] or 0)
        ops[op] += n
        samples[op] += s
        total += n
        tot_s += s
        body.append((r[0], src, n, s, full))
    print(f"total warp instructions {total:,}; samples {tot_s:,}")
    for op, n in ops.most_common(top):
        print(f"  {op:14s} {n:14,d} {100.0 * n / total:6.2f} %   samples {100.0 * samples[op] / max(tot_s, 1):6.2f} %")
    return body


if __name__ == "__main__":
    main()
