"""Register actions (`reg@t`, `!reg`, `[reg <- ..]`): the two-phase stage
(transducer phase writing an action stream + action interpreter,
kleenexlang_b200/frontend/actions.py) against the reference's semantics --
the lockstep FST simulation followed by the action interpretation of
src/KMC/Kleenex/Actions.hs -- and the CPU model of the device algorithm."""
import ctypes
import os
import random
import struct

import pytest

from conftest import load_vectors, vec_matches
from action_cases import NAMES, source, gen, rejecting
from act_model import run_model
from kleenexlang_b200.frontend.actions import ESC, run_act_stream, decode_stream
from kleenexlang_b200.frontend.driver import build_ssts, simulate_lockstep, simulate_sst
from kleenexlang_b200.kexprog import compile_kex, MAGIC_ACT
from oracle.sstbin import oracle_run

REG_VECS = [v for v in load_vectors() if v["uses_registers"]]


def random_stream(rng, n, nregs, maxd):
    out = bytearray()
    d = 0
    for _ in range(n):
        x = rng.random()
        if x < 0.5:
            b = rng.choice([ESC, 0, 1, 2, 65, 66, 67, 254])
            out += bytes([ESC, 0]) if b == ESC else bytes([b])
        elif x < 0.65 and d < maxd:
            out += bytes([ESC, 1])
            d += 1
        elif x < 0.8 and d > 0:
            out += bytes([ESC, 2 + 2 * rng.randrange(nregs)])
            d -= 1
        else:
            out += bytes([ESC, 3 + 2 * rng.randrange(nregs)])
    return bytes(out)


@pytest.mark.parametrize("opt", [0, 3])
@pytest.mark.parametrize("v", REG_VECS, ids=[v["name"] for v in REG_VECS])
def test_reference_vectors_with_registers(v, opt):
    # the reference's own golden vectors of programs with register actions
    ssts = build_ssts(v["program"], opt, actions=True)
    try:
        assert vec_matches(v, simulate_sst(ssts, v["input"]))
        st, out, _ = oracle_run(ssts, v["input"])
    except MemoryError:
        pytest.skip("SST too large")
    assert st == 0 and vec_matches(v, out)


@pytest.mark.parametrize("name", NAMES)
def test_two_phase_stage_equals_lockstep_semantics(name):
    src = source(name)
    ssts = build_ssts(src, 3, actions=True)
    for seed, size in [(0, 0), (1, 30), (2, 300), (3, 2000)]:
        data = gen(name, size, seed)
        exp = simulate_lockstep(src, data)
        assert exp is not None
        assert simulate_sst(ssts, data) == exp
        assert oracle_run(ssts, data)[:2] == (0, exp)
        bad = rejecting(name, data)
        if bad is not None:
            assert simulate_lockstep(src, bad) is None
            assert oracle_run(ssts, bad)[0] == 1


def test_action_interpreter_c_oracle_equals_python():
    rng = random.Random(5)
    from oracle import sstbin
    lib = ctypes.CDLL(sstbin.build_lib())
    lib.kex_oracle_act.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_char_p,
                                   ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
    for _ in range(300):
        nregs = rng.randrange(1, 5)
        s = random_stream(rng, rng.randrange(0, 200), nregs, rng.randrange(1, 6))
        if rng.random() < 0.2 and s:
            s = s[:rng.randrange(len(s))]
        buf = ctypes.create_string_buffer(len(s) + 1)
        ol = ctypes.c_size_t()
        assert lib.kex_oracle_act(s, len(s), nregs, buf, len(s), ctypes.byref(ol)) == 0
        assert buf.raw[:ol.value] == run_act_stream(s)


def test_device_algorithm_model():
    # the five passes of csrc/kex_act.cuh, tile by tile, at every seam position
    rng = random.Random(1)
    for it in range(600):
        nregs = rng.randrange(1, 5)
        s = random_stream(rng, rng.randrange(0, 120), nregs, rng.randrange(1, 6))
        if rng.random() < 0.2 and s:
            s = s[:rng.randrange(len(s))]
        assert run_model(s, nregs, rng.choice([1, 2, 3, 5, 8, 16, 64]), rng.choice([1, 2, 3, 8])) == run_act_stream(s)


def test_escape_roundtrip_and_blob():
    assert decode_stream(bytes([65, ESC, 0, ESC, 1, ESC, 2, ESC, 5, ESC])) == [65, ESC, ("push",), ("pop", 0), ("write", 1)]
    blob = compile_kex(source("swap_fields"))
    nph = struct.unpack_from("<I", blob, 8)[0]
    assert nph == 2
    off, ln = struct.unpack_from("<2I", blob, 16 + 8)
    magic, ver, nregs, esc, total = struct.unpack_from("<5I", blob, off)
    assert (magic, ver, nregs, esc, total, ln) == (MAGIC_ACT, 1, 2, ESC, 32, 32)
    with pytest.raises(ValueError):
        compile_kex(source("swap_fields"), actions=False)      # like `--act=false` (Commands.hs:166-168)


BENCH = "/root/reference/bench/kleenex/src"
BENCH_INPUTS = {
    "swap_lines": b"first line\nsecond\n",
    "sort_ab": b"abbabaabbb",
    "worstcase": b"xyzzy",
    "drex_rev-dict": b"a=1;bb=22;;c;",
    "mitm": b"<p><form action=\"http://x/y\" ><input></form><form  action='u'>",
}


@pytest.mark.skipif(not os.path.isdir(BENCH), reason="reference checkout not present")
@pytest.mark.parametrize("name", sorted(BENCH_INPUTS))
def test_reference_bench_programs_with_actions(name):
    src = open(os.path.join(BENCH, name + ".kex"), encoding="utf-8").read()
    data = BENCH_INPUTS[name]
    exp = simulate_lockstep(src, data)
    assert exp is not None
    ssts = build_ssts(src, 3, actions=True)
    assert oracle_run(ssts, data)[:2] == (0, exp)
    compile_kex(src)


def load_action_vectors():
    import base64
    import json
    from conftest import GOLDEN
    vecs = json.load(open(os.path.join(GOLDEN, "action_vectors.json")))
    for v in vecs:
        v["input"] = base64.b64decode(v["input"])
        v["output"] = base64.b64decode(v["output"])
    return vecs


def test_committed_action_vectors():
    # tests/golden/action_vectors.json (scripts/gen_action_vectors.py: lockstep simulation + Actions.hs semantics)
    ssts = {}
    for v in load_action_vectors():
        if v["program"] not in ssts:
            ssts[v["program"]] = build_ssts(source(v["program"]), 3, actions=True)
        st, out, _ = oracle_run(ssts[v["program"]], v["input"])
        assert (st == 0) == v["accept"]
        if v["accept"]:
            assert out == v["output"]


def test_too_many_action_registers_refused_at_compile_time():
    """The device interpreter has 32 slots for builders + registers: a program
    with more registers is refused by compile_kex, not at kex_load."""
    from kleenexlang_b200.kexprog import compile_kex, UnsupportedProgram
    regs = ["r%d" % i for i in range(32)]
    src = "main := " + " ".join("%s@/a/" % r for r in regs) + " " + " ".join("!%s" % r for r in regs) + "\n"
    with pytest.raises(UnsupportedProgram):
        compile_kex(src)


def load_reference_action_vectors():
    import base64
    import json
    from conftest import GOLDEN
    vecs = json.load(open(os.path.join(GOLDEN, "reference_action_vectors.json")))
    for v in vecs:
        v["input"] = base64.b64decode(v["input"])
        v["output"] = base64.b64decode(v["output"])
    return vecs


REF_ACTION_VECS = load_reference_action_vectors()


def c_compilable(sst):
    """`compileAssignment` (SSTCompiler.hs:60-75) assumes that an updated buffer
    is read at most as the head of its own update (`x := x ...`): anything else
    is compiled to `reset(x); ... concat(x, x)` and loses x.  Action SSTs of
    programs that prepend to a register (`acc@(!e !acc)`, drex_rev-dict.kex)
    violate this: the reference's own C back end cannot run them -- the
    reference's bench/Makefile indeed lists only three of its drex programs."""
    for es in sst.edges.values():
        for _, upd, _ in es:
            for v, w in upd.items():
                if any(a == ("v", v) for a in w[1:]):
                    return False
    return True


@pytest.mark.parametrize("v", REF_ACTION_VECS, ids=[v["name"] for v in REF_ACTION_VECS])
def test_reference_register_programs(v):
    """The reference's 12 benchmark programs with register actions
    (bench/kleenex/src: swap_lines, sort_ab, mitm, markdown2html, drex_*, ...)
    on committed inputs; expected output = lockstep simulation + Actions.hs
    (scripts/gen_reference_action_vectors.py).  Two independent routes must
    reproduce it: the action-stream split the CUDA path uses (transducer phase
    + action interpreter, through the C oracle) and the reference's own default
    compilation scheme (oracle SST + action SST with tables) under the SST
    semantics of `SST.run` -- and through the C oracle (= the emitted C over
    crt.c) wherever the reference's C back end can compile the update."""
    from kleenexlang_b200.frontend.driver import build_oracle_action_pipeline
    assert oracle_run(build_ssts(v["program"], 3, actions=True), v["input"])[:2] == (0, v["output"])
    for opt in (0, 3):
        phases = build_oracle_action_pipeline(v["program"], opt)
        assert simulate_sst(phases, v["input"]) == v["output"]
        if all(c_compilable(p) for p in phases):
            assert oracle_run(phases, v["input"])[:2] == (0, v["output"])
        else:
            assert v["name"] in ("drex_rev-dict", "sort_ab", "worstcase", "dna_regex_noalias_2")


def test_reject_inside_an_action_stage_truncates_once():
    """A stage with register actions that rejects: the action stream produced up
    to the failing symbol is interpreted in full (the bottom builder is the
    result) and the 16 KiB flush rule (crt/crt.c:107-159,217-227) applies once,
    to the stage's output -- not to the intermediate stream as well."""
    from kleenexlang_b200.frontend.sst import run_sst
    src = source("swap_fields")
    ssts = build_ssts(src, 3, actions=True)
    good = gen("swap_fields", 60000, 7)
    cut = good.index(b"\n", 50000) + 1
    bad = good[:cut] + b"a line without the separator\n" + good[cut:]
    st, out, cnt = oracle_run(ssts, bad)
    ok, stream, consumed = run_sst(ssts[0], bad)
    assert st == 1 and not ok and cnt == consumed
    assert len(stream) > 3 * 16384 and len(stream) % 16384 != 0
    _, full, _ = oracle_run(ssts[1:], stream)                     # the interpreter alone, nothing truncated
    assert out == full[:len(full) // 16384 * 16384] and len(out) >= 2 * 16384
    # truncating the intermediate stream first would have lost more
    _, twice, _ = oracle_run(ssts[1:], stream[:len(stream) // 16384 * 16384])
    assert len(twice) // 16384 * 16384 <= len(out)


def _run_phases_model(phases, data):
    """Device-table model (tests/gpu_model.py) for SST phases, the action-stream
    interpreter for action-interpreter phases."""
    from gpu_model import run_model
    from kleenexlang_b200.frontend.actions import run_act_stream
    for t in phases:
        if hasattr(t, "nregs"):
            data = run_act_stream(data)
            if data is None:
                return None
        else:
            ok, data, _ = run_model(t, data, 53)
            if not ok:
                return None
    return data


# (markdown2html and dna_regex_noalias_2 take minutes under the Python model of the tables: the GPU test
# tests/test_gpu_actions.py::test_reference_phases_of_register_programs_gpu runs all twelve)
@pytest.mark.parametrize("v", [v for v in REF_ACTION_VECS if v["name"] not in ("markdown2html", "dna_regex_noalias_2")],
                         ids=[v["name"] for v in REF_ACTION_VECS if v["name"] not in ("markdown2html", "dna_regex_noalias_2")])
def test_reference_phases_of_register_programs(v):
    """`kexc compile --phases=reference` on the reference's programs with register
    actions: a stage is an oracle phase + an action-SST phase where the device can
    evaluate the action SST as it is (appends in creation order), else a transducer
    phase + an action-interpreter phase; either way the committed output."""
    from kleenexlang_b200.kexprog import reference_phases
    phases = reference_phases(v["program"], 3, suppress_bits=True)
    assert len(phases) >= 2
    assert _run_phases_model(phases, v["input"]) == v["output"]
