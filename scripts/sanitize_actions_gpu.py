"""Small parity run of the action-interpreter kernels for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from kleenexlang_b200.runtime import CompiledProgram
from kleenexlang_b200.kexprog import compile_kex
from kleenexlang_b200.frontend.driver import build_ssts
from action_cases import NAMES, source, gen, rejecting
from oracle.sstbin import oracle_run
for tile in ("16", "1024"):
    os.environ["KEX_ACT_TILE"] = tile
    for name in NAMES:
        src = source(name)
        prog = CompiledProgram(compile_kex(src))
        ssts = build_ssts(src, 3, actions=True)
        for nb in (0, 333, 70000):
            d = gen(name, nb, seed=5)
            cases = [d]
            if rejecting(name, d) is not None:
                cases.append(rejecting(name, d))
            for dd in cases:
                got, exp = prog.run(dd), oracle_run(ssts, dd)
                assert got[:2] == exp[:2], (name, nb)
        print(name, tile, "ok", flush=True)
