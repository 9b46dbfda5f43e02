"""Heavy mixing of the deferred and the immediate (window exceeded) paths of k4_emit / k3_emit within
the same groups: a staging window that about half of the tiles exceed.  Output against the oracle, repeatedly."""
import os, sys
sys.path.insert(0, "/root/repo")
from kleenexlang_b200.runtime import CompiledProgram
from kleenexlang_b200.kexprog import compile_kex
from kleenexlang_b200.frontend.driver import build_ssts
from kleenexlang_b200 import workloads
from oracle.sstbin import oracle_run
name = "csv2json"
src = open("/root/repo/programs/%s.kex" % name).read()
ssts = build_ssts(src)
d = workloads.GENERATORS[name](6 << 20, seed=9).tobytes()
exp = oracle_run(ssts, d)
for stage in ("1920", "2048", "2176"):
    for knob in ("KEX_V4_STAGE", "KEX_V3_STAGE"):
        os.environ[knob] = stage
        if knob == "KEX_V3_STAGE":
            os.environ["KEX_NO_V4"] = "1"
        prog = CompiledProgram(compile_kex(src))
        ok = all(prog.run(d)[:2] == exp[:2] for _ in range(20))
        print(knob, stage, "kernel", prog.info()["emit_kernel"], "20 runs", "ok" if ok else "MISMATCH", flush=True)
        prog.close()
        os.environ.pop(knob); os.environ.pop("KEX_NO_V4", None)
