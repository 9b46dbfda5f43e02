"""Quick GPU check: parity on a few sizes per program, then a timed csv2json run."""
import sys, time
sys.path.insert(0, "/root/repo")
import torch
from kleenexlang_b200.runtime import CompiledProgram
from kleenexlang_b200.kexprog import compile_kex
from kleenexlang_b200.frontend.driver import build_ssts
from kleenexlang_b200 import workloads
from oracle.sstbin import oracle_run
bad = 0
for name in ["csv2json", "iso_datetime_to_json", "thousand_sep", "add-commas", "fastq2fasta"]:
    src = open("/root/repo/programs/%s.kex" % name).read()
    prog = CompiledProgram(compile_kex(src))
    ssts = build_ssts(src)
    print(name, prog.info(), flush=True)
    for nb in (1000, 70000, 3 << 20):
        d = workloads.GENERATORS[name](nb, seed=3).tobytes()
        for variant in ("ok", "bad"):
            dd = d if variant == "ok" else d[:len(d) // 2] + b"\x01" + d[len(d) // 2 + 1:]
            got = prog.run(dd)
            exp = oracle_run(ssts, dd)
            ok = got[:2] == exp[:2] and (got[0] == 0 or got[2] == exp[2])
            print(name, nb, variant, "OK" if ok else "MISMATCH", len(got[1]), len(exp[1]), got[0], exp[0], got[2], exp[2], flush=True)
            if not ok:
                bad += 1
                a, b = got[1], exp[1]
                k = next((i for i in range(min(len(a), len(b))) if a[i] != b[i]), min(len(a), len(b)))
                print("first diff at", k, a[max(0, k - 40):k + 40], b[max(0, k - 40):k + 40])
print("mismatches:", bad)
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
src = open("/root/repo/programs/csv2json.kex").read()
prog = CompiledProgram(compile_kex(src))
prog.set_timing(True)
block = workloads.gen_csv(64 << 20, seed=100)
rows = int((block == 10).sum())
reps = max(1, int(gib * (1 << 30)) // len(block))
d_in = torch.from_numpy(block).cuda().repeat(reps)
n = d_in.numel()
expect = n + 127 * rows * reps
d_out = torch.empty(expect + (1 << 20), dtype=torch.uint8, device="cuda")
for i in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st, olen, _ = prog.run_device(d_in.data_ptr(), n, d_out.data_ptr(), d_out.numel())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("run", i, "status", st, "olen ok", olen == expect, "%.2f ms" % (dt * 1e3), "%.1f GiB/s" % (n / dt / 2**30),
          "kernel ms", [round(x, 3) for x in prog.kernel_ms()], "launches", prog.launch_count(), flush=True)
