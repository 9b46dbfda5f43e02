"""Parity run of the G-mode emit kernel (kex_v4.cuh) for compute-sanitizer: first run (G learnt, exact
live sets everywhere), second and third run (tail evaluation), truncated, rejecting, and the forced
exact / tiny-window paths."""
import os, sys
sys.path.insert(0, "/root/repo")
from kleenexlang_b200.runtime import CompiledProgram
from kleenexlang_b200.kexprog import compile_kex
from kleenexlang_b200.frontend.driver import build_ssts
from kleenexlang_b200 import workloads
from oracle.sstbin import oracle_run
names = sys.argv[1:] or ["csv2json", "fastq2fasta"]
for name in names:
    src = open("/root/repo/programs/%s.kex" % name).read()
    ssts = build_ssts(src)
    d = workloads.GENERATORS[name](5 << 19, seed=5).tobytes()
    exp = oracle_run(ssts, d)
    bad = d[:len(d) // 2] + b"\x01" + d[len(d) // 2 + 1:]
    ebad = oracle_run(ssts, bad)
    for knob in ({}, {"KEX_V4_STAGE": "256", "KEX_V4_RECCAP": "8"}, {"KEX_V4_EXACT": "1"}, {"KEX_V4_TAIL_TEST": "1"}):
        for k, v in knob.items():
            os.environ[k] = v
        prog = CompiledProgram(compile_kex(src))
        for i in range(3):
            assert prog.run(d)[:2] == exp[:2], (name, knob, i)
        got = prog.run(bad)
        assert got[0] == ebad[0] and (got[0] == 0 or got[2] == ebad[2]) and got[1] == ebad[1], (name, knob, "reject")
        assert prog.run(d[:700001])[:2] == oracle_run(ssts, d[:700001])[:2]
        print(name, knob, prog.info()["emit_kernel"], "ok", flush=True)
        prog.close()
        for k in knob:
            del os.environ[k]
