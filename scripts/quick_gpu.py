import sys
sys.path.insert(0, "/root/repo")
from kleenexlang_b200.runtime import CompiledProgram
from kleenexlang_b200.kexprog import compile_kex
from kleenexlang_b200.frontend.driver import build_ssts
from kleenexlang_b200 import workloads
from oracle.sstbin import oracle_run
for name in ["fastq2fasta", "csv2json", "thousand_sep"]:
    src = open("/root/repo/programs/%s.kex" % name).read()
    prog = CompiledProgram(compile_kex(src))
    ssts = build_ssts(src)
    for nb in (1000, 70000):
        d = workloads.GENERATORS[name](nb, seed=3).tobytes()
        got = prog.run(d)
        exp = oracle_run(ssts, d)
        print(name, nb, "OK" if got[:2] == exp[:2] else "MISMATCH", len(got[1]), len(exp[1]), got[0], exp[0], flush=True)
        if got[:2] != exp[:2]:
            a, b = got[1], exp[1]
            k = next((i for i in range(min(len(a), len(b))) if a[i] != b[i]), min(len(a), len(b)))
            print("first diff at", k, a[max(0,k-40):k+40], b[max(0,k-40):k+40])
