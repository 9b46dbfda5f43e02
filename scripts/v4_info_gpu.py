"""Prints which emit kernel ran and how many tiles were evaluated exactly, per program and run."""
import sys
sys.path.insert(0, "/root/repo")
from kleenexlang_b200.runtime import CompiledProgram
from kleenexlang_b200.kexprog import compile_kex
from kleenexlang_b200 import workloads
for name in sys.argv[1:] or ["csv2json", "iso_datetime_to_json", "fastq2fasta", "thousand_sep"]:
    prog = CompiledProgram(compile_kex(open("/root/repo/programs/%s.kex" % name).read()))
    d = workloads.GENERATORS[name](4 << 20, seed=63).tobytes()
    for i in range(3):
        st, out, _ = prog.run(d)
        print(name, "run", i, "status", st, "out", len(out), prog.info(), flush=True)
