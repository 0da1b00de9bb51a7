"""Recipe for oracle/_ref: the UNMODIFIED reference package, made importable on the GPU box.  TEST / BENCH INFRASTRUCTURE ONLY.

The reference (gstenzel/qandle v0.1.8) is pure Python: "building" it is copying /root/reference/src/qandle (read-only, present in
the build container only) into oracle/_ref/qandle and putting the qw_map stand-in (oracle/_shim/qw_map.py: the un-vendored
dependency qW-Map 0.1.2 restated, parity unpinned -- SURVEY.md 8c) next to it.  oracle/_ref/ is git-ignored -- no reference source
enters the history -- but not gpurun-ignored, so it travels with the snapshot like the built .so files.  Nothing under qandle_b200/
imports it; bench.py's reference arm and cpu_baseline leg time it (config 1 is the one BASELINE config the reference can
instantiate: every gate is a dense 2^n x 2^n matrix, reference operators.py:277-287).

    python -m oracle.make_ref        # idempotent
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src/qandle"
OUT = os.path.join(HERE, "_ref")


def make(force: bool = False) -> str:
    """Returns the directory to put on sys.path ('' when neither the reference nor an earlier copy exists)."""
    dst = os.path.join(OUT, "qandle")
    if os.path.isdir(REF_SRC) and (force or not os.path.isdir(dst)):
        os.makedirs(OUT, exist_ok=True)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(REF_SRC, dst, ignore=shutil.ignore_patterns("__pycache__", "test", "*.pyc"))
        shutil.copy(os.path.join(HERE, "_shim", "qw_map.py"), os.path.join(OUT, "qw_map.py"))
    return OUT if os.path.isdir(dst) else ""


def import_reference():
    """The reference package as a module object (None if oracle/_ref was never made), without leaving it on sys.path."""
    path = make()
    if not path:
        return None
    sys.path.insert(0, path)
    try:
        import qandle  # noqa: PLC0415

        return qandle
    finally:
        sys.path.remove(path)


if __name__ == "__main__":
    print(make(force="--force" in sys.argv) or "reference not available here")
