"""Static shared-memory bank-conflict estimate of the flat complex64 stages of a workload's plan (host-only; no GPU).

For every stage and side (load through the inverse prefix CNOTs / store through the suffix CNOTs) the 16-byte unit slot of
thread g's pack j is rebuilt exactly as flat64.cuh builds its tables (ins0 of the register bits, pk::absorb_maps,
pk::slot_off); an LDS.128 / STS.128 of a warp is served per quarter-warp (8 threads x 16 B), conflict-free when the 8 slots
fall into 8 different 16-byte bank groups.  Prints wavefronts per warp instruction (1.0 = conflict-free ... 8.0).
   python tools/bank_conflicts.py c2 [--bwd] [-v]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from qandle_b200 import engine

K_CX, K_CX_EXT = 4, 5


def ins0(k, p):
    return ((k >> p) << (p + 1)) | (k & ((1 << p) - 1))


def slot_unit(x, identity="--identity" in sys.argv):
    """pk::slot_off (packed64.cuh) in units; --identity: the old fold (-DQB_SWIZZLE_IDENTITY)"""
    q = x >> 1
    h = (q >> 3) & 7
    return q ^ (h if identity else (0x8D53B8 >> (3 * h)) & 7)


def absorb(x, ops, reverse):
    for op in (reversed(ops) if reverse else ops):
        if op["kind"] == K_CX:
            x ^= ((x >> op["c"]) & 1) << op["a"]
    return x


def stage_wavefronts(st, ops):
    rb = st["regbits"]
    pre = ops[st["op_begin"]:st["pre_end"]]
    suf = ops[st["suf_begin"]:st["op_end"]]
    res = []
    for side, lst, rev in (("load", pre, True), ("store", suf, False)):
        tot = n = 0
        for j in range(8):
            xj = 0
            for k in range(3):
                if (j >> k) & 1:
                    xj |= 1 << rb[k + 1]
            for warp in range(8):
                for quarter in range(4):
                    groups = {}
                    for t in range(8):
                        g = warp * 32 + quarter * 8 + t
                        x = g << 1
                        for r in rb[1:]:
                            x = ins0(x, r)
                        u = slot_unit(absorb(x | xj, lst, rev))
                        groups.setdefault(u & 7, set()).add(u)
                    tot += max(len(v) for v in groups.values())
                    n += 1
        res.append((side, tot / n))
    return res


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    wl = bench.WORKLOADS[name]
    rows, _ = bench.oracle_rows(wl)
    plan = engine.Plan(torch.tensor(rows, dtype=torch.int32).reshape(-1, 4), wl["n"], engine.C64, (0, 0, 0, 0, 1))
    d = engine.parse_plan_dump(plan.dump().tolist())
    for key, okey in (("stages", "ops"), ("stages_bwd", "ops_bwd")):
        tot = cnt = 0
        worst = []
        for si, sw in enumerate(d["sweeps"]):
            for ti, st in enumerate(sw[key]):
                if len(st["regbits"]) < 4:
                    continue
                for side, w in stage_wavefronts(st, sw[okey]):
                    tot += w
                    cnt += 1
                    if w > 1.01:
                        worst.append((w, si, ti, side, st["regbits"], st["xthread"]))
        print(f"{name} {key}: mean wavefronts per quarter-warp access {tot / max(cnt, 1):.3f} over {cnt} stage sides; {len(worst)} with conflicts")
        if "-v" in sys.argv:
            for w in sorted(worst, reverse=True)[:40]:
                print("   ", w)


if __name__ == "__main__":
    main()
