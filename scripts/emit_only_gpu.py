"""csv2json device-resident run (for ncu captures of k3_emit): argv[1] = GiB."""
import sys, time
sys.path.insert(0, "/root/repo")
import torch
from kleenexlang_b200.runtime import CompiledProgram
from kleenexlang_b200.kexprog import compile_kex
from kleenexlang_b200 import workloads
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
name = sys.argv[2] if len(sys.argv) > 2 else "csv2json"
prog = CompiledProgram(compile_kex(open("/root/repo/programs/%s.kex" % name).read()))
prog.set_timing(True)
block = workloads.GENERATORS[name](64 << 20, seed=100)
reps = max(1, int(gib * (1 << 30)) // len(block))
d_in = torch.from_numpy(block).cuda().repeat(reps)
n = d_in.numel()
d_out = torch.empty(int(n * 4.4) + (1 << 20), dtype=torch.uint8, device="cuda")
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st, olen, _ = prog.run_device(d_in.data_ptr(), n, d_out.data_ptr(), d_out.numel())
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
k = prog.kernel_ms()
inf = prog.info()
print("%s %.2f GiB: %.1f GiB/s in (fwd %.2f seams %.2f emit %.2f all %.2f ms) emit_kernel %d exact_tiles %d" % (
    name, n / 2**30, n / dt / 2**30, k[0], k[1], k[2], k[3], inf["emit_kernel"], inf["exact_tiles"]))
