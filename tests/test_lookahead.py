"""Lookahead SSTs (`--la=true`, the reference's default) restated on the
oracle side: longest deterministic prefixes as multi-symbol tests
(src/KMC/SymbolicFST.hs:262-312 `ldp`, `prefixTests`; Determinization.hs
`killTree`, `consumeTreeMany`), their lowering to nested tests with `cmp` and
`readnext(minL, maxL)` (SSTCompiler.hs:37-55,113-156; Classes.hs:56-83), and the
emitted C over the verbatim crt.c.  The CUDA path evaluates the lookahead-free
SSTs; these tests pin that both give the same transduction (as the reference's
own Regression.hs:45-71 does) and record where they differ (reject count, the
`avail < minL` end-of-input rule)."""
import os
import subprocess

import pytest

from conftest import load_vectors, vec_matches, program_source, ROOT
from kleenexlang_b200 import workloads
from kleenexlang_b200.frontend.driver import build_ssts, build_lookahead_ssts, build_oracle_action_pipeline, simulate_sst
from kleenexlang_b200.frontend.il import compile_sst
from kleenexlang_b200.frontend.sst import run_sst
from oracle import build_ref
from oracle.emit_c import render_c

VECS = [v for v in load_vectors() if not v["uses_registers"]]
REF = os.path.join(ROOT, "oracle", "_ref")
needs_crt = pytest.mark.skipif(not build_ref.have_reference(), reason="reference runtime crt.c not present")


@pytest.mark.parametrize("opt", [0, 3])
@pytest.mark.parametrize("v", VECS, ids=[v["name"] for v in VECS])
def test_lookahead_golden(v, opt):
    out = simulate_sst(build_lookahead_ssts(v["program"], opt), v["input"])
    assert out is not None and vec_matches(v, out)


@pytest.mark.parametrize("v", load_vectors()[::3], ids=[v["name"] for v in load_vectors()[::3]])
def test_default_flags_golden(v):
    """--act=true --la=true --sb=true, the reference's default flags: oracle with
    lookahead and suppressed codes + action program."""
    out = simulate_sst(build_oracle_action_pipeline(v["program"], 3, lookahead=True, suppress_bits=True), v["input"])
    assert out is not None and vec_matches(v, out)


def test_multi_symbol_tests_and_cmp():
    src = 'main := (/abc/ "1" | /abd/ "2" | /x[0-9]y/ "3")*\n'
    sst = build_lookahead_ssts(src, 0)[0]
    lens = sorted({len(ps) for es in sst.edges.values() for ps, _, _ in es})
    assert lens[-1] == 3 and sst.la
    assert run_sst(sst, b"abdabcx7y")[:2] == (True, b"abd2abc1x7y3")
    c = render_c([compile_sst(sst)], "/* crt */")
    assert 'cmp(&next[0],(unsigned char *) "' in c            # runs of singleton predicates (Classes.hs:56-83)
    assert "readnext(1, 3)" in c or "readnext(3, 3)" in c     # NextI minL maxL (SSTCompiler.hs:138-146)
    assert "next[1]" in c and "consume(3);" in c


@needs_crt
def test_lookahead_binary_and_eof_rule(tmp_path):
    """`readnext(minL, maxL)` returns 0 as soon as fewer than minL symbols remain
    (crt/crt.c:293-312): a final state whose transitions all test two symbols
    accepts "aba" with the last byte unread under --la=true, while the
    single-symbol program rejects it.  The model (run_sst) and the emitted C
    agree; so do the reject counts (symbols consumed by completed transitions)."""
    kex = tmp_path / "ab.kex"
    kex.write_text("main := (/ab/)*\n")
    la = build_ref.build_one(str(kex), 0, out_dir=str(tmp_path), la=True)
    direct = build_ref.build_one(str(kex), 0, out_dir=str(tmp_path))
    sst = build_lookahead_ssts(kex.read_text(), 0)[0]
    assert {len(ps) for es in sst.edges.values() for ps, _, _ in es} == {2}
    for data, la_exp, direct_rc in ((b"abab", (True, b"abab", 4), 0), (b"aba", (True, b"ab", 2), 1), (b"abb", (True, b"ab", 2), 1),
                                    (b"abxab", (False, b"ab", 2), 1)):
        assert run_sst(sst, data) == la_exp
        r = subprocess.run([la], input=data, capture_output=True)
        assert r.returncode == (0 if la_exp[0] else 1)
        if la_exp[0]:
            assert r.stdout == la_exp[1]
        else:
            assert r.stderr == b"Match error at input symbol %d!\n" % la_exp[2]
        assert subprocess.run([direct], input=data, capture_output=True).returncode == direct_rc


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "csv2json.default")), reason="oracle/_ref variants not built")
@pytest.mark.parametrize("prog", ["csv2json", "iso_datetime_to_json", "thousand_sep", "add-commas", "fastq2fasta"])
def test_reference_binary_variants(prog):
    """oracle/_ref/<prog>{,.la,.act,.default}: the four flag sets of the
    reference's C back end give the same output on accepted input; on a reject
    the lookahead binary's count is the lookahead model's."""
    data = workloads.GENERATORS[prog](400000, seed=13).tobytes()
    base = subprocess.run([os.path.join(REF, prog)], input=data, capture_output=True)
    assert base.returncode == 0
    for variant in (".la", ".act", ".default"):
        r = subprocess.run([os.path.join(REF, prog + variant)], input=data, capture_output=True)
        assert (r.returncode, r.stdout) == (0, base.stdout), variant
    bad = data[:200001] + b"\x01" + data[200002:]
    sst = build_lookahead_ssts(program_source(prog))[0]
    ok, _, cnt = run_sst(sst, bad)
    r = subprocess.run([os.path.join(REF, prog + ".la")], input=bad, capture_output=True)
    assert r.returncode == (0 if ok else 1)
    if not ok:
        assert r.stderr == b"Match error at input symbol %d!\n" % cnt


# ---- the regular-expression flavour (.re): bit-coded parses (compileCoder, Commands.hs:246-275)
RE_CASES = [("(a|b)*c[0-9]+", [b"abbac2016", b"c7", b"bbbbc00"], [b"abx", b"c", b""]),
            # the regex flavour's {n,m} desugars to n copies + m optional ones (Desugaring.hs:110-115): up to n+m
            ("(?:ab|a)(?:bc|c)?x{2,3}", [b"abxx", b"abcxxx", b"acxx", b"abxxxxx"], [b"abx", b"abcxxxxxx"]),
            ("[a-z]+(,[a-z]+)*\\n", [b"ab,c,def\n", b"q\n"], [b"ab,,c\n", b"ab"])]


@pytest.mark.parametrize("re_src,good,bad", RE_CASES)
@pytest.mark.parametrize("sb", [False, True])
@pytest.mark.parametrize("la", [False, True])
def test_regex_coder_round_trip(re_src, good, bad, sb, la):
    """For a regular expression the compiler emits the oracle alone: its output
    is the bit-coded parse (bytes here: Word8 digits).  Decoding it with the
    action machine of the same transducer gives the input back; codes follow
    Coding.hs (fixed width, big-endian, one byte for up to 256 alternatives)."""
    from kleenexlang_b200.frontend.driver import build_coder_ssts, decode_parse
    from oracle.sstbin import oracle_run
    coder = build_coder_ssts(re_src, 3, lookahead=la, suppress_bits=sb)
    for d in good:
        ok, code, _ = run_sst(coder[0], d)
        assert ok and decode_parse(re_src, code, suppress_bits=sb) == d
        if not la:
            assert oracle_run(coder, d)[:2] == (0, code)        # the C oracle has no lookahead
    for d in bad:
        assert not run_sst(coder[0], d)[0]


@needs_crt
def test_regex_coder_binary(tmp_path):
    """The emitted C of a coder (one phase, tables for the classes) over crt.c."""
    from kleenexlang_b200.frontend.driver import build_coder_ssts
    re_src = "(a|b)*c[0-9]+"
    for la in (False, True):
        coder = build_coder_ssts(re_src, 3, lookahead=la, suppress_bits=True)
        ctext = render_c([compile_sst(s) for s in coder], open(build_ref.CRT).read(), info="coder")
        exe = str(tmp_path / ("coder%d" % la))
        r = subprocess.run(["cc", "-O3", "-xc", "-o", exe, "-D FLAG_WORDALIGNED", "-w", "-"], input=ctext.encode(), capture_output=True)
        assert r.returncode == 0, r.stderr
        out = subprocess.run([exe], input=b"abbac2016", capture_output=True)
        assert out.returncode == 0 and out.stdout == run_sst(coder[0], b"abbac2016")[1]
        assert "tbl1[" in ctext
        assert subprocess.run([exe], input=b"abx", capture_output=True).returncode == 1
