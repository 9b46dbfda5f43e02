"""Small driver for ncu captures: csv2json over <MiB> of synthetic CSV, <runs> runs."""
import sys
sys.path.insert(0, "/root/repo")
import torch
from kleenexlang_b200.runtime import CompiledProgram
from kleenexlang_b200.kexprog import compile_kex
from kleenexlang_b200 import workloads
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 512
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
name = sys.argv[3] if len(sys.argv) > 3 else "csv2json"
src = open("/root/repo/programs/%s.kex" % name).read()
prog = CompiledProgram(compile_kex(src))
block = workloads.GENERATORS[name](64 << 20, seed=100)
reps = max(1, (mib << 20) // len(block))
d_in = torch.from_numpy(block).cuda().repeat(reps)
n = d_in.numel()
d_out = torch.empty(int(n * 4.5) + (1 << 20), dtype=torch.uint8, device="cuda")
for i in range(runs):
    st, olen, _ = prog.run_device(d_in.data_ptr(), n, d_out.data_ptr(), d_out.numel())
    torch.cuda.synchronize()
    print("run", i, st, olen, flush=True)
