"""Device-resident throughput of programs with register actions (transducer phase + action interpreter)."""
import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import torch
from kleenexlang_b200.runtime import CompiledProgram
from kleenexlang_b200.kexprog import compile_kex
from action_cases import source, gen
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["swap_fields", "partition", "nested", "reverse_items"]
for name in names:
    prog = CompiledProgram(compile_kex(source(name)))
    block = np.frombuffer(gen(name, 4 << 20, 3), dtype=np.uint8)
    reps = max(1, (mib << 20) // len(block))
    d_in = torch.from_numpy(block.copy()).cuda().repeat(reps)
    n = d_in.numel()
    d_out = torch.empty(int(n * 2.2) + (1 << 20), dtype=torch.uint8, device="cuda")
    for i in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st, olen, _ = prog.run_device(d_in.data_ptr(), n, d_out.data_ptr(), d_out.numel())
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    # size-independent check: every program here works record by record (or reverses a list of
    # records), so the output of `reps` copies of the block is `reps` copies of the block's output
    bo = prog.run(block.tobytes())
    ok = bo[0] == 0 and olen == len(bo[1]) * reps and bool(
        (d_out[:olen].view(reps, len(bo[1])) == torch.frombuffer(bytearray(bo[1]), dtype=torch.uint8).cuda()).all())
    print("   verified against %d x the block's output: %s" % (reps, "OK" if ok else "MISMATCH"), flush=True)
    prog.select_phase(1)
    d_mid = torch.empty(int(n * 14) + (1 << 20), dtype=torch.uint8, device="cuda")
    for i in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st1, olen1, _ = prog.run_device(d_in.data_ptr(), n, d_mid.data_ptr(), d_mid.numel())
        torch.cuda.synchronize(); dt1 = time.perf_counter() - t0
    del d_mid
    print("%-16s %.2f GiB in, status %d, out/in %.3f: %6.2f GiB/s in (%.1f ms; transducer phase alone %.1f ms, stream/in %.2f), %d launches" % (
        name, n / 2**30, st, olen / n, n / dt / 2**30, dt * 1e3, dt1 * 1e3, olen1 / n, prog.launch_count()), flush=True)
    del d_in, d_out
