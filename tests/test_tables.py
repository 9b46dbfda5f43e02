"""Device-table construction (kexprog) checked on the CPU: the executable
model of the CUDA algorithm (tests/gpu_model.py), run over the same tables the
blob is serialised from, must agree with the C oracle for every chunking."""
import pytest

from conftest import load_vectors, program_source, sample
from gpu_model import run_model
from kleenexlang_b200 import workloads
from kleenexlang_b200.frontend.driver import build_ssts
from kleenexlang_b200.kexprog import build_phase, compile_kex, serialize_pipeline, UnsupportedProgram
from oracle.sstbin import oracle_run

VECS = [v for v in load_vectors() if not v["uses_registers"]]


def _model_pipeline(tabs, data, chunk):
    st, cnt = 0, 0
    for t in tabs:
        ok, data, cnt = run_model(t, data, chunk)
        st = 0 if ok else 1
    return st, data, cnt


@pytest.mark.parametrize("v", VECS, ids=[v["name"] for v in VECS])
def test_model_golden(v):
    ssts = build_ssts(v["program"], 3)
    try:
        tabs = [build_phase(s) for s in ssts]
    except UnsupportedProgram:
        pytest.skip("exceeds device register limit")
    for chunk in (1, 5, 64):
        assert _model_pipeline(tabs, v["input"], chunk) == oracle_run(ssts, v["input"])


@pytest.mark.parametrize("prog", ["csv2json", "iso_datetime_to_json", "thousand_sep", "add-commas", "fastq2fasta"])
@pytest.mark.parametrize("opt", [0, 3])
def test_model_workloads(prog, opt):
    ssts = build_ssts(program_source(prog), opt)
    tabs = [build_phase(s) for s in ssts]
    data = workloads.GENERATORS[prog](3000, seed=5).tobytes()
    bad = data[:1700] + b"\x01" + data[1700:]
    for d in (data, bad, data[:-3], b""):
        for chunk in (3, 32, 4096):
            assert _model_pipeline(tabs, d, chunk) == oracle_run(ssts, d), (prog, opt, chunk, len(d))


def test_model_apache():
    ssts = build_ssts(program_source("apache_log"), 3)
    tabs = [build_phase(s) for s in ssts]
    d = sample("apache_sample.log")[:6000]
    d = d[:d.rfind(b"\n") + 1]
    assert _model_pipeline(tabs, d, 50) == oracle_run(ssts, d)


def test_blob_layout():
    blob = compile_kex(program_source("csv2json"))
    assert blob[:4] == b"KEXL" and len(blob) % 16 == 0
    t = build_phase(build_ssts(program_source("csv2json"))[0])
    assert (t.Q, t.C, t.R) == (27, 5, 5)
    assert len(serialize_pipeline([t])) == len(blob)
    assert build_phase(build_ssts(program_source("fastq2fasta"))[0]).R == 1


# ---- monoid tables of the fast section (kleenexlang_b200/fasttab.py)
from gpu_model import run_fast_model
from kleenexlang_b200 import fasttab


def _fast_pipeline(tabs, data, chunk, sub):
    st, cnt = 0, 0
    for t in tabs:
        ok, data, cnt = run_fast_model(t, fasttab.build_fast(t), data, chunk, sub)
        st = 0 if ok else 1
    return st, data, cnt


@pytest.mark.parametrize("v", VECS, ids=[v["name"] for v in VECS])
def test_fast_model_golden(v):
    ssts = build_ssts(v["program"], 3)
    try:
        tabs = [build_phase(s) for s in ssts]
        for t in tabs:
            fasttab.build_fast(t)
    except (UnsupportedProgram, fasttab.Ineligible):
        pytest.skip("not eligible for the monoid kernels")
    for chunk, sub in ((1, 1), (6, 2), (64, 8)):
        assert _fast_pipeline(tabs, v["input"], chunk, sub) == oracle_run(ssts, v["input"])


@pytest.mark.parametrize("prog", ["csv2json", "iso_datetime_to_json", "thousand_sep", "add-commas", "fastq2fasta"])
@pytest.mark.parametrize("opt", [0, 3])
def test_fast_model_workloads(prog, opt):
    ssts = build_ssts(program_source(prog), opt)
    tabs = [build_phase(s) for s in ssts]
    data = workloads.GENERATORS[prog](3000, seed=5).tobytes()
    bad = data[:1700] + b"\x01" + data[1700:]
    for d in (data, bad, data[:-3], b""):
        for chunk, sub in ((4, 2), (32, 8), (4096, 32)):
            assert _fast_pipeline(tabs, d, chunk, sub) == oracle_run(ssts, d), (prog, opt, chunk, len(d))


def test_fast_section_in_blob():
    import struct
    blob = compile_kex(program_source("csv2json"))
    off = struct.unpack_from("<I", blob, 16)[0]
    fast_off, fast_len = struct.unpack_from("<II", blob, off + 76)
    assert fast_off and blob[off + fast_off:off + fast_off + 4] == b"KEXF"
    f = fasttab.build_fast(build_phase(build_ssts(program_source("csv2json"))[0]))
    assert (f.NM, f.NL, f.NB) == (987, 5, 7)
    # apache_log exceeds the live-set limit and carries no fast section
    blob = compile_kex(program_source("apache_log"))
    off = struct.unpack_from("<I", blob, 16)[0]
    assert struct.unpack_from("<II", blob, off + 76) == (0, 0)


# ---- G-mode tables (fasttab.build_gmode, csrc/kex_v4.cuh)
from gpu_model import run_gmode_model


def _observed_pairs(t, f, data, tile, upto):
    """(state, live set) at the tile ends before `upto`, most frequent first --
    what the host library reads back from the device before it builds the
    G-mode table."""
    import collections
    Q, C, A = t.Q, t.C, t.A
    states, acts = [t.init], []
    for b in data:
        e = f.trans2[states[-1] * C + t.cls[b]]
        states.append(e & 0xFFFF)
        acts.append((e >> 16) & 0xFF)
    n = len(data)
    lam = [0] * (n + 1)
    lam[n] = f.lam_final[states[-1]]
    for i in range(n - 1, -1, -1):
        lam[i] = ((f.BE[lam[i + 1] * A + acts[i]] & 0xFFFC) // 4) // A
    cnt = collections.Counter((states[i], lam[i]) for i in range(tile, upto, tile))
    return [k for k, _ in cnt.most_common()]


@pytest.mark.parametrize("prog,expect_all", [("csv2json", True), ("iso_datetime_to_json", True), ("fastq2fasta", True),
                                             ("thousand_sep", False), ("add-commas", False)])
def test_gmode_model(prog, expect_all):
    """A tile emitted from the G-mode table alone (end live set == G[end state],
    FAIL row never reached) equals the exact evaluation -- asserted inside the
    model for every such tile -- and the whole output equals the oracle's, with
    the static table and with a table learnt from the first third of the input.
    Programs whose live sets follow the state run (nearly) every tile in G-mode
    once the table is learnt; thousand_sep's follow the digit count and do not."""
    ssts = build_ssts(program_source(prog), 3)
    t = build_phase(ssts[0])
    f = fasttab.build_fast(t)
    data = workloads.GENERATORS[prog](60000, seed=8).tobytes()
    exp = oracle_run(ssts, data)
    assert exp[0] == 0
    for tile in (32, 1024):
        static = fasttab.build_gmode(t, f)
        assert static is not None
        out, _, _ = run_gmode_model(t, f, static, data, tile)
        assert out == exp[1]
        learnt = fasttab.build_gmode(t, f, _observed_pairs(t, f, data, tile, 20000))
        out, g, nt = run_gmode_model(t, f, learnt, data, tile)
        assert out == exp[1]
        if expect_all:
            assert g >= nt - 8, (prog, tile, g, nt)          # all but the tiles of the last record
        else:
            assert g < nt // 2


def test_gmode_sections_in_blob():
    import struct
    for prog in ("csv2json", "fastq2fasta"):
        blob = compile_kex(program_source(prog))
        off = struct.unpack_from("<I", blob, 16)[0]
        fast_off = struct.unpack_from("<I", blob, off + 76)[0]
        hdr = struct.unpack_from("<32I", blob, off + fast_off)
        t = build_phase(build_ssts(program_source(prog))[0])
        f = fasttab.build_fast(t)
        G, gtab, min_tpl = fasttab.build_gmode(t, f)
        assert hdr[23] == 1 and hdr[26] == min_tpl
        base = off + fast_off
        assert list(blob[base + hdr[24]:base + hdr[24] + t.Q + 1]) == G
        assert list(struct.unpack_from("<%dI" % len(gtab), blob, base + hdr[25])) == gtab
        # row Q (FAIL) absorbs; every entry is a copy, a template of a valid id, or nothing
        assert all(e == t.Q for e in gtab[t.Q * t.C:])
        for e in gtab:
            fl = e >> 24
            assert (e & 0xFFFF) <= t.Q and not (fl & 0x80 and fl & 0x40)
            assert not (fl & 0x80) or (fl & 0x3F) < f.NT


# ---- table atoms on the device (AppendTblI lowered to constants per byte): the regex coder and the
# reference's oracle/action phase structure
RE_CASES = [("(a|b)*c[0-9]+", [b"abbac2016", b"c7", b"bbbbc00", b"abx", b"c", b""]),
            ("(?:ab|a)(?:bc|c)?x{2,3}", [b"abxx", b"abcxxx", b"acxx", b"abxxxxx", b"abx", b"abcxxxxxx"]),
            ("[a-z]+(,[a-z]+)*\\n", [b"ab,c,def\n", b"q\n", b"ab,,c\n", b"ab"])]


@pytest.mark.parametrize("re_src,inputs", RE_CASES)
@pytest.mark.parametrize("sb", [False, True])
def test_model_regex_coder(re_src, inputs, sb):
    """compileCoder (Commands.hs:246-275) in device form: the coder's tables expanded per byte give the
    same code bytes as the SST with table atoms (C oracle, atom 3), accept and reject alike."""
    from kleenexlang_b200.frontend.driver import build_coder_ssts
    from kleenexlang_b200.frontend.sst import expand_tables, run_sst
    coder = build_coder_ssts(re_src, 3, lookahead=False, suppress_bits=sb)
    assert any(a[0] == "t" for es in coder[0].edges.values() for _, u, _ in es for w in u.values() for a in w) or "[" not in re_src
    flat = [expand_tables(s) for s in coder]
    assert not any(a[0] == "t" for es in flat[0].edges.values() for _, u, _ in es for w in u.values() for a in w)
    tabs = [build_phase(s) for s in flat]
    for d in inputs:
        want = oracle_run(coder, d)
        ok, code, _ = run_sst(flat[0], d)
        assert ok == (want[0] == 0) and (not ok or code == want[1])       # a reject keeps whole 16 KiB flushes only
        for chunk in (1, 4, 64):
            assert _model_pipeline(tabs, d, chunk) == want


@pytest.mark.parametrize("v", VECS, ids=[v["name"] for v in VECS])
@pytest.mark.parametrize("sb", [False, True])
def test_model_reference_phases(v, sb):
    """`--phases=reference`: oracle phase + action phase per stage (tables as constants) transduce like
    the direct SSTs; the stream after the first phase is the reference's code (oracle SST with table atoms)."""
    from kleenexlang_b200.kexprog import reference_phases
    from kleenexlang_b200.frontend.driver import build_oracle_action_pipeline
    try:
        phases = reference_phases(v["program"], 3, suppress_bits=sb)
    except UnsupportedProgram:
        pytest.skip("exceeds device register limit")
    assert not any(hasattr(t, "nregs") for t in phases)
    ref = build_oracle_action_pipeline(v["program"], 3, lookahead=False, suppress_bits=sb)
    assert len(phases) == len(ref)
    want = oracle_run(build_ssts(v["program"], 3), v["input"])
    got = _model_pipeline(phases, v["input"], 7)
    assert got[:2] == want[:2]
    code = oracle_run(ref[:1], v["input"])
    assert _model_pipeline(phases[:1], v["input"], 7)[:2] == code[:2]


# ---- the reference's own register-free benchmark programs on random walks through their own grammar
def _bench_vectors():
    import base64, json, os
    from conftest import GOLDEN
    vs = json.load(open(os.path.join(GOLDEN, "reference_bench_vectors.json")))
    for v in vs:
        v["input"], v["output"] = base64.b64decode(v["input"]), base64.b64decode(v["output"])
    return vs


BENCH = _bench_vectors()
# (syntax and make_danish: 770 / 1039 states, half a minute of determinization each -- the GPU test
# tests/test_gpu_parity.py::test_reference_bench_programs runs them)
BENCH_NAMES = sorted({v["name"] for v in BENCH} - {"syntax", "make_danish"})


@pytest.mark.parametrize("name", BENCH_NAMES)
def test_model_reference_bench_programs(name):
    """bench/kleenex/src/*.kex without register actions (scripts/gen_reference_bench_vectors.py): the device
    tables (with the --opt fallbacks compile_kex takes) under the executable model of the kernels against
    the lockstep simulation the vectors were made with, and the C oracle on the same SSTs."""
    from kleenexlang_b200.kexprog import kex_phases
    vs = [v for v in BENCH if v["name"] == name]
    tabs = kex_phases(vs[0]["program"], 3, actions=False)
    ssts = build_ssts(vs[0]["program"], 3)
    for v in vs:
        assert oracle_run(ssts, v["input"])[:2] == (0, v["output"])
        assert _model_pipeline(tabs, v["input"], 61)[:2] == (0, v["output"])


def test_creation_order_check():
    """Weak constant propagation (--opt 1) can materialise a known register after younger content: the
    update shape stays (old registers)(new material) but the bytes no longer leave in creation order --
    build_phase must refuse (compile_kex then falls back to --opt 0)."""
    from kleenexlang_b200.frontend.driver import build_transducers
    from kleenexlang_b200.frontend.oracle_action import build_oracle_action_ssts
    from kleenexlang_b200.frontend.sst import expand_tables
    from kleenexlang_b200.kexprog import check_chronological
    v = [v for v in load_vectors() if v["name"] == "test_compiled/newlinebug.kex"][0]
    t = build_transducers(v["program"])[0]
    o1 = expand_tables(build_oracle_action_ssts(t, 1, False, False)[0])
    assert check_chronological(o1)
    with pytest.raises(UnsupportedProgram, match="order of their creation"):
        build_phase(o1)
    o0 = expand_tables(build_oracle_action_ssts(t, 0, False, False)[0])
    assert run_model(build_phase(o0), v["input"], 3)[:2] == (True, oracle_run([o0], v["input"])[1])
