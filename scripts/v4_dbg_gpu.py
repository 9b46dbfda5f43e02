"""Debug run for compute-sanitizer: one action program through the C ABI vs the oracle."""
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from kleenexlang_b200.runtime import CompiledProgram
from kleenexlang_b200.kexprog import compile_kex
from kleenexlang_b200.frontend.driver import build_ssts
from action_cases import source, gen
from oracle.sstbin import oracle_run
name = sys.argv[1] if len(sys.argv) > 1 else "partition"
os.environ["KEX_ACT_TILE"] = "16"
src = source(name)
prog = CompiledProgram(compile_kex(src))
print(prog.info(), flush=True)
ssts = build_ssts(src, 3, actions=True)
for seed, size in [(0, 0), (1, 40), (2, 700), (3, 20000), (4, 300000)]:
    d = gen(name, size, seed)
    exp = oracle_run(ssts, d)
    got = prog.run(d)
    print(name, size, "ok" if got[:2] == exp[:2] else "MISMATCH", flush=True)
