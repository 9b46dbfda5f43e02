"""Small parity run for compute-sanitizer (memcheck / racecheck)."""
import sys
sys.path.insert(0, "/root/repo")
from kleenexlang_b200.runtime import CompiledProgram
from kleenexlang_b200.kexprog import compile_kex
from kleenexlang_b200.frontend.driver import build_ssts
from kleenexlang_b200 import workloads
from oracle.sstbin import oracle_run
names = sys.argv[1:] or ["csv2json", "iso_datetime_to_json", "thousand_sep", "fastq2fasta"]
for name in names:
    src = open("/root/repo/programs/%s.kex" % name).read()
    prog = CompiledProgram(compile_kex(src))
    ssts = build_ssts(src)
    for nb in (777, 300000):
        d = workloads.GENERATORS[name](nb, seed=5).tobytes()
        for dd in (d, d[:len(d) // 2] + b"\x01" + d[len(d) // 2 + 1:], d[:len(d) - 3]):
            got, exp = prog.run(dd), oracle_run(ssts, dd)
            assert got[:2] == exp[:2], (name, nb)
    print(name, "ok", flush=True)
