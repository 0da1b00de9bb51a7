"""Instruction share per source line / opcode for one kernel of an .ncu-rep (needs --import-source on, -lineinfo).
   python tools/ncu_lines.py file.ncu-rep [kernel_index] [top_n]"""
import csv
import io
import subprocess
import sys
from collections import Counter


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rep, kidx, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    secs, cur = [], None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = {"file": r[1], "rows": []}
            secs.append(cur)
        elif r[0] == "Function Name":
            cur["fn"] = r[1]
        elif r[0] == "Line No":
            cur["hdr"] = r
        else:
            cur["rows"].append(r)
    files = []
    for s in secs:
        if s["file"] not in files:
            files.append(s["file"])
    nf = len(files)
    mine = secs[kidx * nf:(kidx + 1) * nf]
    iI = mine[0]["hdr"].index("Instructions Executed")
    iS = mine[0]["hdr"].index("# Samples")
    lines, ops = [], Counter()
    seen_addr = set()
    for s in mine:
        cur_line = None
        for r in s["rows"]:
            if r[0] != "":
                cur_line = (s["file"].split("/")[-1], r[0], r[1].strip()[:100])
                lines.append([0, 0, cur_line])
            else:
                lines[-1][0] += I(r[iI])
                lines[-1][1] += I(r[iS])
                if r[2] not in seen_addr:  # SASS rows repeat under every inlined frame: count each address once
                    seen_addr.add(r[2])
                    op = r[3].split()
                    o = op[1] if op[0].startswith("@") else op[0]
                    ops[o.split(".")[0]] += I(r[iI])
    tot = sum(ops.values())
    print(f"kernel {kidx}: {mine[0].get('fn', '?')[:70]}  instructions executed {tot}")
    print("by opcode:", ", ".join(f"{o} {n / tot * 100:.1f}%" for o, n in ops.most_common(24)))
    ts = max(1, sum(l[1] for l in lines))
    for n, smp, (f, ln, src) in sorted(lines, key=lambda l: -l[0])[:top]:
        print(f"{n / tot * 100:5.1f}% smp {smp / ts * 100:5.1f}%  {f}:{ln}  {src}")


main()
