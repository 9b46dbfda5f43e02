"""Parity of the action-interpreter kernels (csrc/kex_act.cuh) and of whole
programs with register actions, through the C ABI, with the CPU oracle
(oracle/kex_oracle.c: kex_oracle_run + kex_oracle_act).  Bit-exact."""
import os
import random
import struct

import pytest

from action_cases import NAMES, source, gen, rejecting
from test_actions import random_stream
from kleenexlang_b200.frontend.actions import ESC, run_act_stream
from kleenexlang_b200.frontend.driver import build_ssts
from kleenexlang_b200.kexprog import compile_kex, MAGIC_ACT, MAGIC_PIPE, VERSION
from oracle.sstbin import oracle_run

pytestmark = pytest.mark.gpu
_cache = {}


def gpu_prog(name):
    from kleenexlang_b200.runtime import CompiledProgram
    if name not in _cache:
        src = source(name)
        _cache[name] = (CompiledProgram(compile_kex(src)), build_ssts(src, 3, actions=True))
    return _cache[name]


@pytest.fixture(params=[16, 48, 1024])
def act_tile(request):
    # small tiles put a seam at (nearly) every token boundary
    os.environ["KEX_ACT_TILE"] = str(request.param)
    yield request.param
    del os.environ["KEX_ACT_TILE"]


@pytest.mark.parametrize("name", NAMES)
def test_action_programs(name, act_tile):
    prog, ssts = gpu_prog(name)
    for seed, size in [(0, 0), (1, 40), (2, 700), (3, 20000), (4, 300000)]:
        d = gen(name, size, seed)
        est, eout, _ = oracle_run(ssts, d)
        assert est == 0
        st, out, _ = prog.run(d)
        assert (st, out) == (est, eout), (name, size)


@pytest.mark.parametrize("name", [n for n in NAMES if rejecting(n, b"") is not None])
def test_action_programs_reject(name):
    prog, ssts = gpu_prog(name)
    for seed, size in [(1, 40), (2, 40000)]:
        bad = rejecting(name, gen(name, size, seed))
        est, eout, ecnt = oracle_run(ssts, bad)
        assert est == 1
        assert prog.run(bad) == (est, eout, ecnt)


def test_action_program_large():
    # 24 MiB: ~24 K tiles, ~100 groups
    prog, ssts = gpu_prog("swap_fields")
    d = gen("swap_fields", 1 << 20, 9) * 24
    st, out, _ = prog.run(d)
    est, eout, _ = oracle_run(ssts, d)
    assert (st, out) == (est, eout)
    assert prog.run(out)[:2] == oracle_run(ssts, out)[:2]
    prog, ssts = gpu_prog("reverse_items")
    d = gen("reverse_items", 1 << 18, 10) * 6        # (the CPU oracle copies on every `!acc`: quadratic)
    st, out, _ = prog.run(d)
    assert (st, out) == oracle_run(ssts, d)[:2]
    assert prog.run(out)[1] == d                      # reversing twice


def act_only_blob(nregs):
    ph = struct.pack("<8I", MAGIC_ACT, VERSION, nregs, ESC, 32, 0, 0, 0)
    return struct.pack("<8I", MAGIC_PIPE, VERSION, 1, 32 + len(ph), 32, len(ph), 0, 0) + ph


def test_interpreter_kernels_on_random_streams(act_tile):
    from kleenexlang_b200.runtime import CompiledProgram
    rng = random.Random(act_tile)
    progs = {k: CompiledProgram(act_only_blob(k)) for k in (1, 2, 4)}
    for it in range(60):
        nregs = rng.choice([1, 2, 4])
        s = random_stream(rng, rng.choice([0, 5, 50, 400, 4000, 30000]), nregs, rng.randrange(1, 8))
        if rng.random() < 0.3 and s:
            s = s[:rng.randrange(len(s))]
        st, out, _ = progs[nregs].run(s)
        assert st == 0 and out == run_act_stream(s), (it, len(s))


def test_interpreter_refuses_what_it_cannot_hold():
    from kleenexlang_b200.runtime import CompiledProgram, KexError
    prog = CompiledProgram(act_only_blob(2))
    with pytest.raises(KexError):
        prog.run(bytes([ESC, 2]))                      # pop on the bottom builder: not an action stream
    with pytest.raises(KexError):
        prog.run(bytes([ESC, 1]) * 40)                 # 40 nested builders > 32 slots
    with pytest.raises(KexError):
        prog.run(bytes([ESC, 3 + 2 * 7]))              # register 7 of 2


def test_committed_action_vectors_on_gpu():
    from test_actions import load_action_vectors
    for v in load_action_vectors():
        prog, _ = gpu_prog(v["program"])
        st, out, _ = prog.run(v["input"])
        assert (st == 0) == v["accept"], v["program"]
        if v["accept"]:
            assert out == v["output"], v["program"]


from test_actions import REF_ACTION_VECS


@pytest.mark.parametrize("v", REF_ACTION_VECS, ids=[v["name"] for v in REF_ACTION_VECS])
def test_reference_register_programs_gpu(v):
    """The reference's own benchmark programs with register actions, executed
    by the CUDA path (transducer phase + action-interpreter kernels) on the
    committed vectors (tests/golden/reference_action_vectors.json)."""
    from kleenexlang_b200.runtime import CompiledProgram
    prog = CompiledProgram(compile_kex(v["program"]))
    assert prog.run(v["input"])[:2] == (0, v["output"])
    # and tiled: most of these grammars are record* -- where tiling keeps the input in the
    # language the output tiles as well (checked against the oracle, not assumed)
    big = v["input"] * 50
    est, eout, _ = oracle_run(build_ssts(v["program"], 3, actions=True), big)
    got = prog.run(big)
    assert got[0] == est and (est != 0 or got[1] == eout)
    prog.close()


def test_reject_inside_an_action_stage_gpu():
    """Reject of a stage with register actions after several 16 KiB of stream:
    one truncation, on the stage's output (tests/test_actions.py has the oracle
    side of this rule)."""
    prog, ssts = gpu_prog("swap_fields")
    good = gen("swap_fields", 200000, 7)
    for at in (50000, 150000):
        cut = good.index(b"\n", at) + 1
        bad = good[:cut] + b"a line without the separator\n" + good[cut:]
        exp = oracle_run(ssts, bad)
        assert exp[0] == 1 and len(exp[1]) >= 2 * 16384
        assert prog.run(bad) == exp


@pytest.mark.parametrize("v", REF_ACTION_VECS, ids=[v["name"] for v in REF_ACTION_VECS])
def test_reference_phases_of_register_programs_gpu(v):
    """`--phases=reference` on the reference's programs with register actions: oracle phase + action-SST
    phase where the device can evaluate the action SST as it is, transducer + action-interpreter phase
    otherwise (tests/test_actions.py has the CPU-model side)."""
    from kleenexlang_b200.runtime import CompiledProgram
    from kleenexlang_b200.kexprog import compile_reference_phases
    prog = CompiledProgram(compile_reference_phases(v["program"], 3, suppress_bits=True))
    assert prog.run(v["input"])[:2] == (0, v["output"])
    prog.close()
