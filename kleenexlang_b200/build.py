"""Builds libkexcuda.so (sm_100a) in-tree with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "kexcuda.cu")
LIB = os.path.join(HERE, "libkexcuda.so")


def build(force=False, verbose=False):
    deps = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))]
    deps.append(os.path.join(HERE, "..", "include", "kexcuda.h"))
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
        return LIB
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    import ctypes
    ctypes.CDLL(LIB)            # unresolved symbols show here, not on the GPU box
    return LIB


KEXRUN = os.path.join(HERE, "kexrun")


def build_kexrun(force=False):
    """The native launcher (tools/kexrun.c): binds the C ABI from C with dlopen; `kexc compile --out`
    installs a copy of it next to the program blob."""
    src = os.path.join(HERE, "..", "tools", "kexrun.c")
    if not force and os.path.exists(KEXRUN) and os.path.getmtime(KEXRUN) >= os.path.getmtime(src):
        return KEXRUN
    r = subprocess.run(["cc", "-O2", "-I", os.path.join(HERE, "..", "include"), src, "-ldl", "-o", KEXRUN],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("cc failed:\n" + r.stdout + r.stderr)
    return KEXRUN


if __name__ == "__main__":
    print(build_kexrun(force=True))
    print(build(force=True, verbose="-v" in sys.argv))
