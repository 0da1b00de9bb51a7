"""Print the stage structure of a workload's plan (host-only; no GPU needed).
   python tools/plan_stats.py c2 [-v]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from oracle import statevec as O
from qandle_b200 import engine

KN = {1: "U1", 2: "D1", 3: "D1x", 4: "CX", 5: "CXx", 6: "CZ", 7: "CZx1", 8: "CZx2", 9: "SWAP"}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    verbose = "-v" in sys.argv
    wl = bench.WORKLOADS[name]
    rows, _ = bench.oracle_rows(wl)
    prog = torch.tensor(rows, dtype=torch.int32).reshape(-1, 4)
    dtype = engine.C128 if wl.get("c128") else engine.C64
    plan = engine.Plan(prog, wl["n"], dtype, (0, 0, 0, 0, 1))
    d = engine.parse_plan_dump(plan.dump().tolist())
    tot_st = tot_ops = tot_u1 = tot_body_cx = tot_abs = 0
    bwd = "--bwd" in sys.argv
    from collections import Counter
    shapes = Counter()
    for si, sw in enumerate(d["sweeps"]):
        ops = sw["ops_bwd"] if bwd else sw["ops"]
        stages = sw["stages_bwd"] if bwd else sw["stages"]
        print(f"sweep {si}: tile_bits={sw['tile_bits']} ops={len(ops)} stages={len(stages)} kslots={len(sw['kslots'])}")
        for st in stages:
            shapes[bin(st.get("shape", 0)).count("1")] += 1
            pre = ops[st["op_begin"]:st["pre_end"]]
            body = ops[st["pre_end"]:st["suf_begin"]]
            suf = ops[st["suf_begin"]:st["op_end"]]
            tot_st += 1
            tot_ops += len(pre) + len(body) + len(suf)
            tot_u1 += sum(1 for o in body if o["kind"] in (1, 2, 3))
            tot_body_cx += sum(1 for o in body if o["kind"] in (4, 5))
            tot_abs += len(pre) + len(suf)
            if verbose:
                fmt = lambda os_: " ".join(f"{KN[o['kind']]}({o['a']}{',' + str(o['c']) if o['c'] >= 0 else ''}|r{o['r']}{',' + str(o['rc']) if o['c'] >= 0 else ''})" for o in os_)
                print(f"   regbits={st['regbits']} pre[{fmt(pre)}] body[{fmt(body)}] suf[{fmt(suf)}]")
    print("2x2s per stage:", dict(sorted(shapes.items())))
    print(f"total: sweeps={len(d['sweeps'])} stages={tot_st} ops={tot_ops} 1q-ops={tot_u1} body-CX={tot_body_cx} absorbed-CX={tot_abs}")


main()
