import os, sys
sys.path.insert(0, "/root/repo")
from kleenexlang_b200.runtime import CompiledProgram
from kleenexlang_b200.kexprog import compile_kex
from kleenexlang_b200 import workloads
name = "csv2json"
src = open("/root/repo/programs/%s.kex" % name).read()
size = int(sys.argv[1]); runs = int(sys.argv[2])
d = workloads.GENERATORS[name](size, seed=5).tobytes()
prog = CompiledProgram(compile_kex(src))
for i in range(runs):
    st, out, _ = prog.run(d)
    print("run", i, st, len(out), prog.info()["emit_kernel"], flush=True)
