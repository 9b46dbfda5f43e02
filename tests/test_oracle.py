"""The CPU oracle (oracle/kex_oracle.c) pinned against the reference's golden
vectors, the cross-tool digests, and -- when they have been built -- the
reference's own compiled binaries (oracle/_ref: emitted C + verbatim crt.c)."""
import hashlib
import json
import os
import subprocess

import pytest

from conftest import load_vectors, vec_matches, program_source, sample, GOLDEN, ROOT
from kleenexlang_b200.frontend.driver import build_ssts
from kleenexlang_b200 import workloads
from oracle.sstbin import oracle_run

VECS = [v for v in load_vectors() if not v["uses_registers"]]
REF = os.path.join(ROOT, "oracle", "_ref")


@pytest.mark.parametrize("opt", [0, 3])
@pytest.mark.parametrize("v", VECS, ids=[v["name"] for v in VECS])
def test_oracle_golden(v, opt):
    st, out, _ = oracle_run(build_ssts(v["program"], opt), v["input"])
    assert st == 0 and vec_matches(v, out)


def test_oracle_cross_tool():
    cross = json.load(open(os.path.join(GOLDEN, "cross_tool.json")))
    for prog, key in (("iso_datetime_to_json", "perl_sha256"), ("iso_datetime_to_json", "python_sha256"),
                      ("thousand_sep", "perl_sha256"), ("apache_log", "perl_sha256_normalised")):
        st, out, _ = oracle_run(build_ssts(program_source(prog)), sample(cross[prog]["input"]))
        assert st == 0 and hashlib.sha256(out).hexdigest() == cross[prog][key], (prog, key)
    st, out, _ = oracle_run(build_ssts(program_source("csv2json")), sample("csv_sample.csv"))
    assert st == 0 and len(out) == cross["csv2json"]["output_len"]
    assert out.decode().startswith(cross["csv2json"]["first_record"])


def test_csv2json_pinned_by_independent_restatement():
    """csv2json has no golden output in the reference; pin it by a restatement
    of the reference's own Ragel equivalent (bench/ragel/src/csv2json.rl) that
    shares nothing with the restated front end: the reference's 10-row sample
    (committed output tests/golden/csv2json_sample.out, 1878 bytes) and 1 MiB of
    the synthetic CSV the benchmark uses."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from make_csv2json_pin import csv2json_rl
    ssts = build_ssts(program_source("csv2json"))
    d = sample("csv_sample.csv")
    want = open(os.path.join(GOLDEN, "csv2json_sample.out"), "rb").read()
    assert len(want) == 1878 and csv2json_rl(d) == want
    assert oracle_run(ssts, d)[:2] == (0, want)
    if os.path.exists(os.path.join(REF, "csv2json")):
        assert _ref("csv2json", d)[:2] == (0, want)
    big = workloads.gen_csv(1 << 20, seed=21).tobytes()
    assert oracle_run(ssts, big)[:2] == (0, csv2json_rl(big))


def _ref(name, data):
    r = subprocess.run([os.path.join(REF, name)], input=data, capture_output=True)
    return r.returncode, r.stdout, r.stderr


CASES = [("csv2json", "csv2json"), ("iso_datetime_to_json", "iso_datetime_to_json"), ("thousand_sep", "thousand_sep"),
         ("add-commas", "add-commas"), ("fastq2fasta", "fastq2fasta")]


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "csv2json")), reason="oracle/_ref not built")
@pytest.mark.parametrize("prog,gen", CASES)
def test_oracle_vs_reference_binary(prog, gen):
    ssts = build_ssts(program_source(prog))
    data = workloads.GENERATORS[gen](200000, seed=11).tobytes()
    # accept, reject mid-stream (after several 16 KiB flushes), reject at end of input, empty input
    bad = data[:150001] + b"\x01" + data[150001:]
    for d in (data, bad, data[:-1] if prog != "add-commas" else data + b"1", b""):
        st, out, cnt = oracle_run(ssts, d)
        rc, ro, err = _ref(prog, d)
        assert (st, out) == (rc, ro)
        if rc:
            assert ("symbol %d!" % cnt).encode() in err


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "apache_log")), reason="oracle/_ref not built")
def test_oracle_vs_reference_binary_apache():
    d = sample("apache_sample.log")
    assert oracle_run(build_ssts(program_source("apache_log")), d)[:2] == _ref("apache_log", d)[:2]


# ---- the reference's default mode (--act=true): oracle + action program per stage
ALL_VECS = load_vectors()


@pytest.mark.parametrize("v", ALL_VECS, ids=[v["name"] for v in ALL_VECS])
def test_suppressed_bits_golden(v):
    """`--sb=true` (OutputEquivalence.hs): no code for a choice whose arms rejoin
    without output; the action program skips to the last post-dominator.  Same
    transduction, never a longer code stream."""
    from kleenexlang_b200.frontend.driver import build_oracle_action_pipeline
    from kleenexlang_b200.frontend.sst import run_sst
    plain = build_oracle_action_pipeline(v["program"], 3)
    sb = build_oracle_action_pipeline(v["program"], 3, suppress_bits=True)
    st, out, _ = oracle_run(sb, v["input"])
    assert st == 0 and vec_matches(v, out)
    assert len(run_sst(sb[0], v["input"])[1]) <= len(run_sst(plain[0], v["input"])[1])


@pytest.mark.parametrize("opt", [0, 3])
@pytest.mark.parametrize("v", ALL_VECS, ids=[v["name"] for v in ALL_VECS])
def test_oracle_action_pipeline_golden(v, opt):
    """Every golden vector of the reference -- the programs with register
    actions included -- through the restated oracle/action split
    (frontend/oracle_action.py: OracleMachine.hs, ActionMachine.hs,
    ActionSST.hs) evaluated by the C oracle with table atoms."""
    from kleenexlang_b200.frontend.driver import build_oracle_action_pipeline
    phases = build_oracle_action_pipeline(v["program"], opt)
    assert len(phases) % 2 == 0
    st, out, _ = oracle_run(phases, v["input"])
    assert st == 0 and vec_matches(v, out)


def test_oracle_code_stream_shape():
    """One code byte per copied byte of a non-singleton range set and per n-way
    choice, nothing for singleton sets and suppressed reads
    (OracleMachine.hs:52-61); the action program decodes it through tables."""
    from kleenexlang_b200.frontend.driver import build_oracle_action_pipeline
    from kleenexlang_b200.frontend.sst import run_sst
    o, a = build_oracle_action_pipeline('main := (/[a-c]/ | ~/x/ "X")* /;/\n', 0)
    ok, code, _ = run_sst(o, b"bxa;")
    # loop choice (0 = iterate, 1 = leave) and arm choice (0 = class, 1 = x) per item; [a-c] index; ';' is a singleton
    assert ok and code == bytes([0, 0, 1, 0, 1, 0, 0, 0, 1])
    ok, out, _ = run_sst(a, code)
    assert ok and out == b"bXa;"
    assert any(at[0] == "t" for es in o.edges.values() for _, upd, _ in es for w in upd.values() for at in w)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "csv2json.act")), reason="oracle/_ref not built")
@pytest.mark.parametrize("prog,gen", CASES)
def test_default_mode_reference_binary(prog, gen):
    """oracle/_ref/<prog>.act = the reference's default two-process binary
    (emitted oracle + action C programs with tables, verbatim crt.c): same
    output as the direct binary and the oracle on accepted inputs; on a reject
    the oracle phase reports the same `count`."""
    from kleenexlang_b200.frontend.driver import build_oracle_action_pipeline
    src = program_source(prog)
    data = workloads.GENERATORS[gen](300000, seed=12).tobytes()
    rc, out, _ = _ref(prog + ".act", data)
    assert (rc, out) == _ref(prog, data)[:2] == oracle_run(build_ssts(src), data)[:2]
    assert oracle_run(build_oracle_action_pipeline(src, suppress_bits=True), data)[:2] == (0, out)
    bad = data[:150001] + b"\x01" + data[150001:]
    st, _, cnt = oracle_run(build_ssts(src), bad)
    rc, _, err = _ref(prog + ".act", bad)
    if st:
        assert rc == 1 and err.startswith(("Match error at input symbol %d!" % cnt).encode())
