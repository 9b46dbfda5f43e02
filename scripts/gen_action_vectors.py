"""Writes tests/golden/action_vectors.json: input/output pairs for the programs under programs/actions/,
produced by the restated lockstep FST simulation followed by the action interpretation
(`kexc simulate --sim=lockstep`, src/KMC/Frontend/Commands.hs:277-289; src/KMC/Kleenex/Actions.hs:14-58) --
independent of the SST construction, the action-stream split, the C oracle and the CUDA kernels."""
import base64
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from action_cases import NAMES, source, gen, rejecting          # noqa: E402
from kleenexlang_b200.frontend.driver import simulate_lockstep  # noqa: E402

vecs = []
for name in NAMES:
    for seed, size in [(11, 0), (12, 60), (13, 900), (14, 6000)]:
        d = gen(name, size, seed)
        out = simulate_lockstep(source(name), d)
        assert out is not None
        vecs.append({"program": name, "input": base64.b64encode(d).decode(), "output": base64.b64encode(out).decode(),
                     "accept": True})
    bad = rejecting(name, gen(name, 200, 15))
    if bad is not None:
        assert simulate_lockstep(source(name), bad) is None
        vecs.append({"program": name, "input": base64.b64encode(bad).decode(), "output": "", "accept": False})
with open(os.path.join(ROOT, "tests", "golden", "action_vectors.json"), "w") as f:
    json.dump(vecs, f, indent=0)
print(len(vecs), "vectors")
