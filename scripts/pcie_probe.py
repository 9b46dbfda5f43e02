"""PCIe probe: H2D / D2H alone and concurrently (pinned), then kex_run_host at a few sub-wave sizes."""
import sys, time, os, ctypes
sys.path.insert(0, "/root/repo")
import numpy as np, torch
GIB = 1 << 30
h_a = torch.empty(4 * GIB, dtype=torch.uint8).pin_memory()
h_b = torch.empty(8 * GIB, dtype=torch.uint8).pin_memory()
d_a = torch.empty(4 * GIB, dtype=torch.uint8, device="cuda")
d_b = torch.empty(8 * GIB, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); return time.perf_counter() - t0
for _ in range(2):
    a = t(lambda: d_a.copy_(h_a, non_blocking=True))
    b = t(lambda: h_b.copy_(d_b, non_blocking=True))
    def both():
        with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
        with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
    c = t(both)
    print("H2D 4GiB %.1f ms (%.1f GB/s)  D2H 8GiB %.1f ms (%.1f GB/s)  both %.1f ms" % (a*1e3, 4*GIB/a/1e9, b*1e3, 8*GIB/b/1e9, c*1e3), flush=True)
del d_a, d_b, h_b
from kleenexlang_b200 import workloads
from kleenexlang_b200.kexprog import compile_kex
from kleenexlang_b200.runtime import CompiledProgram
src = open("/root/repo/programs/csv2json.kex").read()
prog = CompiledProgram(compile_kex(src))
block = workloads.gen_csv(64 << 20, seed=100)
rows = int((block == 10).sum())
reps = 63
h_in = h_a[:reps * len(block)]
h_in.copy_(torch.from_numpy(np.tile(block, reps)))
n = h_in.numel(); expect = n + 127 * rows * reps
h_out = torch.empty(expect + 4096, dtype=torch.uint8).pin_memory()
L = prog._L
ol, stt, fc = ctypes.c_size_t(), ctypes.c_int(), ctypes.c_size_t()
for wave in ("16", "32", "64", "128", "256", "off"):
    if wave == "off": os.environ["KEX_NO_HOST_PIPELINE"] = "1"
    else: os.environ["KEX_HOST_WAVE_MIB"] = wave
    for i in range(3):
        t0 = time.perf_counter()
        rc = L.kex_run_host(prog._h, ctypes.c_char_p(h_in.data_ptr()), n, h_out.data_ptr(), h_out.numel(), ctypes.byref(ol), ctypes.byref(stt), ctypes.byref(fc))
        dt = time.perf_counter() - t0
        assert rc == 0 and ol.value == expect
    print("wave", wave, "%.1f ms  %.2f GiB/s" % (dt * 1e3, n / dt / GIB), flush=True)
