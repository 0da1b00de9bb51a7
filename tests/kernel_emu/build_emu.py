"""Build the kernel emulator: the UNMODIFIED CUDA sources of qandle_b200/csrc, textually adapted for g++ (kernel launches,
dynamic shared memory declarations and the eight inline-PTX sites) and linked with the fiber scheduler.
TEST INFRASTRUCTURE (see shim/cuda_runtime.h).  Usage: python build_emu.py [--force]  ->  _build/libqandle_b200_emu.so"""
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.environ.get("QB_REPO_ROOT") or os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "qandle_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libqandle_b200_emu.so")
SOURCES = ["capi.cu", "plan.cpp", "plan.h", "kernels.cuh", "packed64.cuh", "flat64.cuh", "flat128.cuh"]


def _split_top(s):
    """split at top-level commas (parentheses / brackets / braces / template angle brackets of casts are balanced)"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def _rewrite_launches(src):
    """NAME<<<grid, block, smem, stream>>>(args);  ->  kemu::launch(grid, block, smem, [&] { NAME(args); });"""
    out, pos, n = [], 0, 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            out.append(src[pos:])
            break
        # kernel expression: back to the start of the statement
        j = k
        while j > 0 and src[j - 1] not in ";{}\n":
            j -= 1
        head = src[j:k]
        indent = head[: len(head) - len(head.lstrip())]
        prefix = ""
        name = head.strip()
        m = re.match(r"^(else\s+|if\s*\(.*\)\s+)?(.*)$", name, re.S)  # `else kernel<<<...` / `if (c) kernel<<<...`
        if m and m.group(1):
            prefix, name = m.group(1), m.group(2)
        e = src.index(">>>", k)
        cfg = _split_top(src[k + 3:e])
        assert len(cfg) in (2, 3, 4), cfg
        # argument list
        a0 = src.index("(", e)
        depth, a1 = 0, a0
        while True:
            if src[a1] == "(":
                depth += 1
            elif src[a1] == ")":
                depth -= 1
                if depth == 0:
                    break
            a1 += 1
        args = src[a0 + 1:a1]
        semi = src.index(";", a1)
        smem = cfg[2] if len(cfg) > 2 else "0"
        out.append(src[pos:j])
        out.append(f"{indent}{prefix}kemu::launch(dim3({cfg[0]}), dim3({cfg[1]}), (size_t)({smem}), [&] {{ {name}({args}); }});")
        pos = semi + 1
        n += 1
    return "".join(out), n


PTX = [
    (re.compile(r'asm volatile\("cp\.async\.cg\.shared\.global \[%0\], \[%1\], 16;" ::"r"\((.*?)\), "l"\((.*?)\)\);'), r"kemu::cp_async16(\1, \2);"),
    (re.compile(r'asm volatile\("cp\.async\.commit_group;"\);'), r"(void)0;"),
    (re.compile(r'asm volatile\("cp\.async\.wait_group %0;" ::"n"\(N\)\);'), r"(void)0;"),
    (re.compile(r'asm volatile\("bar\.sync %0, (\d+);" ::"r"\((.*?)\) : "memory"\);'), r"kemu::barrier(\2, \1);"),
    (re.compile(r'asm volatile\("xor\.b32 %0, %0, %1;\\n\\txor\.b32 %1, %1, %0;\\n\\txor\.b32 %0, %0, %1;" : "\+r"\(x\), "\+r"\(y\)\);'),
     r"{ const unsigned t_ = x; x = y; y = t_; }"),
]


def transform(name, src):
    stats = {}
    src, stats["launches"] = _rewrite_launches(src)
    src, stats["dyn_smem"] = re.subn(r"extern __shared__ __align__\(16\) unsigned char smem_raw\[\];", "unsigned char* const smem_raw = kemu::dyn_smem();", src)
    n_ptx = 0
    for rx, rep in PTX:
        src, k = rx.subn(rep, src)
        n_ptx += k
    stats["ptx"] = n_ptx
    if "asm volatile" in src or "asm(" in src or "<<<" in src:
        raise RuntimeError(f"{name}: an inline-PTX site or kernel launch was not recognised by the emulator build")
    return src, stats


def build(force=False, verbose=False, defines=(), tag=""):
    """defines: extra -D macros (a kernel variant under test); tag: suffix of the library built with them"""
    global LIB
    lib = LIB if not tag else LIB.replace(".so", f"_{tag}.so")
    return _build(lib, force, verbose, tuple(defines))


def _build(LIB, force, verbose, defines):
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(HERE, "kernel_emu.cpp"), os.path.join(HERE, "shim", "cuda_runtime.h"),
                                                        os.path.join(ROOT, "include", "qandle_b200.h"), os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    gen = os.path.join(OUT, "gen", "qandle_b200", "csrc")
    os.makedirs(gen, exist_ok=True)
    os.makedirs(os.path.join(OUT, "gen", "include"), exist_ok=True)
    shutil.copy(os.path.join(ROOT, "include", "qandle_b200.h"), os.path.join(OUT, "gen", "include", "qandle_b200.h"))
    total = {"launches": 0, "dyn_smem": 0, "ptx": 0}
    for s in SOURCES:
        text, st = transform(s, open(os.path.join(CSRC, s)).read())
        for k in total:
            total[k] += st[k]
        dst = os.path.join(gen, s.replace(".cu", ".cpp") if s.endswith(".cu") else s)
        open(dst, "w").write(text)
    if verbose:
        print("kernel_emu: rewrote", total)
    cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-w"] + [f"-D{d}" for d in defines] + ["-I", os.path.join(HERE, "shim"),
           os.path.join(gen, "capi.cpp"), os.path.join(gen, "plan.cpp"), os.path.join(HERE, "kernel_emu.cpp"), "-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("kernel_emu build failed:\n" + " ".join(cmd) + "\n" + r.stderr[-6000:])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
