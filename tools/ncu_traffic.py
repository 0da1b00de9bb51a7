"""DRAM traffic per launch of the dominant kernel, from an `ncu --set full` capture -> profiles/ncu_traffic.json (read by bench.py).

   python tools/ncu_traffic.py gpurun_out/<tag>_prof_bwd.ncu-rep c2 4096 [out.json]

The capture is `tools/gpu_pass.sh ncu` (adjoint sweeps of one config-2 step).  The file records the kernel name, the workload, the
per-launch dram__bytes_read.sum + dram__bytes_write.sum and a hash of the kernel sources: bench.py reports `roofline.traffic` from it only
while that hash equals the hash of the sources it runs (a changed kernel needs a new capture, otherwise `traffic` is null)."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL_SOURCES = ["flat64.cuh", "packed64.cuh", "kernels.cuh", "capi.cu"]


def sources_sha16():
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        h.update(open(os.path.join(ROOT, "qandle_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def to_bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(value.replace(",", "")) * scale


def main():
    rep, workload, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
    out_path = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", "ncu_traffic.json")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rd = to_bytes(d["dram__bytes_read.sum"], units[hdr.index("dram__bytes_read.sum")])
        wr = to_bytes(d["dram__bytes_write.sum"], units[hdr.index("dram__bytes_write.sum")])
        launches.append({"id": int(d["ID"]), "kernel": d["Kernel Name"], "dram_bytes_read": rd, "dram_bytes_write": wr,
                         "gpu_time_ms_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")) *
                         {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[hdr.index("gpu__time_duration.sum")]]})
    kernels = sorted({l["kernel"] for l in launches})
    assert len(kernels) == 1, f"one kernel per capture expected, got {kernels}"
    mean = sum(l["dram_bytes_read"] + l["dram_bytes_write"] for l in launches) / len(launches)
    res = {"kernel": kernels[0], "workload": workload, "batch": batch, "dram_bytes_per_launch": mean, "launches": launches,
           "capture": "ncu --set full --clock-control none (tools/gpu_pass.sh ncu) of " + os.path.basename(rep),
           "kernel_sources": KERNEL_SOURCES, "kernel_sources_sha16": sources_sha16()}
    json.dump(res, open(out_path, "w"), indent=1)
    print(out_path, kernels[0], f"{mean / 1e9:.3f} GB per launch over {len(launches)} launches")


if __name__ == "__main__":
    main()
