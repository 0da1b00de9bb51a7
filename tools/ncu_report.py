"""profiles/<name>.md from one GPU pass (tools/gpu_pass.sh <tag> launches ncu): launch list of one config-2 step, `ncu --set full` metrics of the
captured adjoint / forward sweeps, the execution-weighted SASS mix and shared-memory wavefronts of the heaviest captured launch of each.
   python tools/ncu_report.py <tag> > profiles/r2_ncu_summary.md"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter, defaultdict

import ncu_summary  # noqa: F401  (same directory; its KEYS)


def launch_list(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit() and r[12] == "gpu__time_duration.sum"]
    agg, order = defaultdict(lambda: [0, 0.0]), []
    sweeps = []
    for r in rows:
        name, val, unit = r[4], float(r[14].replace(",", "")), r[13]
        ms = val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        if name not in agg:
            order.append(name)
        agg[name][0] += 1
        agg[name][1] += ms
        if "sweep_flat_kernel" in name:
            sweeps.append((name, ms))
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | ms | share |\n|---|---|---|---|")
    for name in sorted(order, key=lambda n: -agg[n][1]):
        print(f"| `{name[:110]}` | {agg[name][0]} | {agg[name][1]:.3f} | {100 * agg[name][1] / tot:.1f} % |")
    fwd = [round(ms, 2) for n, ms in sweeps if "<0," in n]
    bwd = [round(ms, 2) for n, ms in sweeps if "<1," in n]
    print(f"\nforward sweeps in launch order (ms): {fwd}\n\nadjoint sweeps in launch order (ms): {bwd}\n")


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"### {d['Kernel Name'][:70]} (id {d['ID']})\n\n| metric | value | unit |\n|---|---|---|")
        for k in ncu_summary.KEYS:
            if k in d:
                print(f"| {k} | {d[k]} | {units[hdr.index(k)]} |")
        stalls = {k.split("issue_stalled_")[1].split("_per")[0]: float(d[k]) for k in hdr
                  if k.startswith("smsp__average_warp") and "issue_stalled" in k and k.endswith("_per_issue_active.ratio") and "not_issued" not in k and d[k]}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:10]
        print("| warp stall cycles per issued instruction | " + ", ".join(f"{k} {v:.2f}" for k, v in top) + " | |\n")
        res.append(float(d["gpu__time_duration.sum"].replace(",", "")))
    return res


def sass_mix(rep, which):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    secs, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = []
            secs.append(cur)
        elif r and r[0] == "Address":
            hdr = r
        elif cur is not None and r and r[0].startswith("0x"):
            cur.append(r)
    iS, iE, iSm, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
    body = secs[2 * which]  # every kernel appears twice in the source page
    ops, samp, wav = Counter(), Counter(), Counter()
    for r in body:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS].strip())
        op = m.group(2) if m else "?"
        ops[op] += int(r[iE] or 0)
        samp[op] += int(r[iSm] or 0)
        w = int(r[iW] or 0)
        if w:
            full = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS].strip()).group(2)
            wav[full + (" (uniform address)" if re.search(r"\[UR\d+(\+0x[0-9a-f]+)?\]", r[iS]) else "")] += w
    tot, tots, totw = sum(ops.values()), max(sum(samp.values()), 1), max(sum(wav.values()), 1)
    print(f"total warp instructions {tot:,}; stall samples {tots:,}\n\n| opcode | executed | share | share of stall samples |\n|---|---|---|---|")
    for op, n in ops.most_common(16):
        print(f"| {op} | {n:,} | {100.0 * n / tot:.2f} % | {100.0 * samp[op] / tots:.2f} % |")
    fp = sum(n for o, n in ops.items() if o in ("FFMA2", "FMUL2", "FFMA", "FMUL", "FADD", "FADD2"))
    print(f"\nFP instructions {100.0 * fp / tot:.1f} % (packed FFMA2 + FMUL2 {100.0 * (ops['FFMA2'] + ops['FMUL2']) / tot:.1f} %); HBM<->shared tile movement "
          f"(LDGSTS + STG + LDG) {100.0 * (ops['LDGSTS'] + ops['STG'] + ops['LDG']) / tot:.2f} % of issued instructions.\n")
    print(f"shared-memory wavefronts by instruction ({totw:,} from LSU instructions):\n\n| instruction | wavefronts | share |\n|---|---|---|")
    for k, v in wav.most_common(6):
        print(f"| {k} | {v:,} | {100.0 * v / totw:.1f} % |")
    print()


def static_sass():
    import os

    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qandle_b200", "libqandle_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            funcs[cur][m.group(2)] += 1
    want = {"streaming adjoint sweep `<1,1,1,1,0>`": "sweep_flat_kernelILb1ELb1ELb1ELb1ELb0", "forward sweep `<0,1,0,1,0>`": "sweep_flat_kernelILb0ELb1ELb0ELb1ELb0",
            "push exchange (TMA)": "exchange_push_tma", "pull exchange (p2p)": "exchange_p2p_kernel"}
    ops = ["FFMA2", "FMUL2", "FFMA", "LDS", "STS", "MOV", "LOP3", "BRA", "SHFL", "LDGSTS", "STG", "LDG", "BAR", "UBLKCP", "SYNCS", "LDL", "STL"]
    print("## Static SASS counts of the shipped `libqandle_b200.so` (`cuobjdump -sass`)\n")
    print("| kernel | instructions | " + " | ".join(ops) + " |\n|---|---|" + "---|" * len(ops))
    for name, key in want.items():
        c = Counter()
        for f, cc in funcs.items():
            if key in f:
                c += cc
        print(f"| {name} | {sum(c.values())} | " + " | ".join(str(c[o]) for o in ops) + " |")
    print()


def main():
    tag = sys.argv[1]
    out = "gpurun_out"
    print(f"# Round 2: ncu evidence for the default build (1x B200, gpurun pass {tag}; commands: tools/gpu_pass.sh; this file: tools/ncu_report.py {tag})\n")
    print("All captures: `ncu --set full --clock-control none --import-source on -k regex:sweep_flat` on `tools/profile_step.py c2 4096 2` (BASELINE config 2, "
          "second step), i.e. the kernels `bench.py` times: `fl::sweep_flat_kernel<BWD, FULL, STREAM, DYN, FUSE>` = `<1,1,1,1,0>` streaming adjoint sweep with "
          "persistent CTAs, `<0,1,0,1,0>` forward sweep with persistent CTAs.  Numbers under ncu are cold-cache and serialised: shares, counters and traffic are "
          "what they are for, not bench values.  `profiles/ncu_traffic.json` (read by `bench.py` for `roofline.traffic`) is the DRAM traffic of the adjoint capture.\n")
    print(f"## Launch list of one config-2 step (`{tag}_launches_c2.csv`, `--metrics gpu__time_duration.sum`)\n")
    launch_list(f"{out}/{tag}_launches_c2.csv")
    print("## ncu --set full: three consecutive adjoint sweeps of config 2's second step (ids 0, 1, 2)\n")
    t = raw_metrics(f"{out}/{tag}_prof_bwd.ncu-rep")
    heavy = max(range(len(t)), key=lambda i: t[i])
    print(f"### Execution-weighted SASS mix of the heaviest captured adjoint sweep (id {heavy})\n")
    sass_mix(f"{out}/{tag}_prof_bwd.ncu-rep", heavy)
    print("## ncu --set full: two consecutive forward sweeps of config 2's second step\n")
    t = raw_metrics(f"{out}/{tag}_prof_fwd.ncu-rep")
    print("### Execution-weighted SASS mix of the second captured forward sweep\n")
    sass_mix(f"{out}/{tag}_prof_fwd.ncu-rep", 1)
    static_sass()


if __name__ == "__main__":
    main()
