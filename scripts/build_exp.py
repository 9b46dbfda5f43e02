"""Kernel experiments: builds kleenexlang_b200/exp/libkexcuda_<tag>.so with -DKEX_EXP_<NAME>... so that several
variants of a kernel can be timed in one GPU call (KEX_LIB=<path> selects the library).
Usage: build_exp.py tag [DEFINE ...]"""
import os, subprocess, sys
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "kleenexlang_b200")
tag, defs = sys.argv[1], sys.argv[2:]
os.makedirs(os.path.join(HERE, "exp"), exist_ok=True)
out = os.path.join(HERE, "exp", "libkexcuda_%s.so" % tag)
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
       "-shared", "-o", out, os.path.join(HERE, "csrc", "kexcuda.cu")] + ["-D" + d for d in defs]
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode:
    sys.exit("nvcc failed:\n" + r.stdout + r.stderr)
print(out)
