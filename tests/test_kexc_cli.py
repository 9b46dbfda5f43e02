"""`kexc` command-line surface (src/kexc.hs:13-27, Options.hs:146-201) and the
process contract of the produced binary (crt/crt.c:372-467).  The CPU part
needs no GPU: flag spellings, artefacts, `-h`/`-i` exit codes, `simulate`.
The GPU part pipes inputs through the launcher as
test/test_compiled/runtest.sh:21-26 does."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT, PROGRAMS, load_vectors, vec_matches

KEXC = [sys.executable, "-m", "kleenexlang_b200.kexc"]
ENV = dict(os.environ, PYTHONPATH=ROOT)


def _compile(tmp_path, name, *flags):
    out = str(tmp_path / name)
    r = subprocess.run(KEXC + ["compile", os.path.join(PROGRAMS, name + ".kex"), "--out", out, *flags],
                       capture_output=True, cwd=ROOT, env=dict(ENV, KEXC_LAUNCHER=LAUNCHER_KIND[0]))
    assert r.returncode == 0, r.stderr
    return out


LAUNCHER_KIND = ["native"]


@pytest.fixture(params=["native", "python"])
def launcher_kind(request):
    """`kexc compile --out bin` installs the native launcher (tools/kexrun.c: the C ABI bound
    from C) by default, the interpreter-based one (launcher.py) with KEXC_LAUNCHER=python."""
    LAUNCHER_KIND[0] = request.param
    yield request.param
    LAUNCHER_KIND[0] = "native"


def test_native_launcher_is_a_binary(tmp_path):
    out = _compile(tmp_path, "add-commas", "--quiet")
    assert open(out, "rb").read(4) == b"\x7fELF"
    assert os.path.exists(out + ".kexprog.lib") and os.path.exists(out + ".kexprog.desc")
    LAUNCHER_KIND[0] = "python"
    try:
        out2 = _compile(tmp_path, "csv2json", "--quiet")
    finally:
        LAUNCHER_KIND[0] = "native"
    assert open(out2, "rb").read(2) == b"#!"


def test_compile_writes_blob_and_launcher(tmp_path):
    out = _compile(tmp_path, "add-commas", "--opt", "0", "--la=false", "--act=false", "--quiet",
                   "--srcout", str(tmp_path / "src.txt"))
    assert os.path.exists(out + ".kexprog") and os.access(out, os.X_OK)
    assert "SST states" in open(tmp_path / "src.txt").read()
    h = subprocess.run([out, "-h"], capture_output=True)
    assert h.returncode == 1 and b"-t" in h.stdout            # RETC_PRINT_USAGE, crt/crt.c:14
    i = subprocess.run([out, "-i"], capture_output=True)
    assert i.returncode == 2 and b"--opt 0" in i.stdout       # RETC_PRINT_INFO, crt/crt.c:15
    assert subprocess.run([out, "--bogus"], capture_output=True).returncode == 1


def test_reference_default_flags_are_accepted(tmp_path):
    # `kexc compile prog.kex --out bin` with the reference's defaults (--la=true --act=true)
    out = str(tmp_path / "t")
    r = subprocess.run(KEXC + ["compile", os.path.join(PROGRAMS, "thousand_sep.kex"), "--out", out],
                       capture_output=True, cwd=ROOT, env=ENV)
    assert r.returncode == 0 and b"--la=false" in r.stderr


def test_usage_without_subcommand():
    r = subprocess.run(KEXC, capture_output=True, cwd=ROOT, env=ENV)
    assert r.returncode == 1 and b"compile" in r.stdout


@pytest.mark.parametrize("sim", ["lockstep", "sst"])
def test_simulate_readme_example(sim):
    # README.md:36-42
    r = subprocess.run(KEXC + ["simulate", "--quiet", "--sim=" + sim, os.path.join(PROGRAMS, "add-commas.kex")],
                       input=b"2016\n", capture_output=True, cwd=ROOT, env=ENV)
    assert r.returncode == 0 and r.stdout == b"2,016\n"


def test_simulate_apache_log_config():
    # BASELINE config 1: apache_log.kex via `kexc simulate` on the bundled log sample (CPU)
    from conftest import sample
    from kleenexlang_b200.frontend.driver import build_ssts
    from oracle.sstbin import oracle_run
    d = sample("apache_sample.log")
    r = subprocess.run(KEXC + ["simulate", "--quiet", os.path.join(PROGRAMS, "apache_log.kex")],
                       input=d, capture_output=True, cwd=ROOT, env=ENV)
    src = open(os.path.join(PROGRAMS, "apache_log.kex"), encoding="utf-8").read()
    est, eout, _ = oracle_run(build_ssts(src), d)
    assert r.returncode == est == 0 and r.stdout == eout


@pytest.mark.gpu
def test_compiled_binary_contract(tmp_path, launcher_kind):
    vecs = [v for v in load_vectors() if not v["uses_registers"]][:12]
    for k, v in enumerate(vecs):
        src = tmp_path / ("p%d.kex" % k)
        src.write_text(v["program"], encoding="utf-8")
        out = str(tmp_path / ("p%d" % k))
        c = subprocess.run(KEXC + ["compile", str(src), "--out", out, "--opt", "0", "--la=false", "--quiet"],
                           capture_output=True, cwd=ROOT, env=dict(ENV, KEXC_LAUNCHER=launcher_kind))
        if c.returncode == 1 and b"exceeds a limit" in c.stderr:
            continue                                          # > 32 simultaneously live registers
        assert c.returncode == 0, c.stderr
        r = subprocess.run([out], input=v["input"], capture_output=True)
        assert r.returncode == 0 and vec_matches(v, r.stdout), v["name"]
    out = _compile(tmp_path, "csv2json", "--quiet")
    from conftest import sample
    d = sample("csv_sample.csv")
    ok = subprocess.run([out, "-t"], input=d, capture_output=True)
    assert ok.returncode == 0 and ok.stderr.startswith(b"time (ms): ")
    ph = subprocess.run([out, "-p", "1"], input=d, capture_output=True)
    assert ph.returncode == 0 and ph.stdout == ok.stdout
    cut = d.index(b"\n", 100) + 2                            # inside the numeric id of a later row
    bad = subprocess.run([out], input=d[:cut] + b"x" + d[cut + 1:], capture_output=True)
    assert bad.returncode == 1 and bad.stderr == b"Match error at input symbol %d!\n" % cut


@pytest.mark.gpu
def test_launcher_streams_large_inputs(tmp_path, launcher_kind):
    """Inputs larger than one block go through kex_stream_* (bounded memory);
    with 1 MiB blocks a 6 MiB input streams and must equal the whole-input run."""
    from kleenexlang_b200 import workloads
    out = _compile(tmp_path, "csv2json", "--quiet")
    d = workloads.gen_csv(6 << 20, seed=8).tobytes()
    whole = subprocess.run([out], input=d, capture_output=True, env=dict(os.environ, KEX_NO_STREAM="1"))
    streamed = subprocess.run([out], input=d, capture_output=True, env=dict(os.environ, KEX_STREAM_BLOCK_MIB="1"))
    assert whole.returncode == streamed.returncode == 0 and whole.stdout == streamed.stdout
    cut = d.index(b"\n", 3 << 20) + 2
    bad = d[:cut] + b"x" + d[cut + 1:]
    w = subprocess.run([out], input=bad, capture_output=True, env=dict(os.environ, KEX_NO_STREAM="1"))
    s2 = subprocess.run([out], input=bad, capture_output=True, env=dict(os.environ, KEX_STREAM_BLOCK_MIB="1"))
    assert w.returncode == s2.returncode == 1 and w.stdout == s2.stdout and w.stderr == s2.stderr


def test_action_program_compile_and_simulate(tmp_path):
    # a program with register actions: refused with --act=false like the reference (Commands.hs:166-168),
    # two phases per stage with the default --act=true; `simulate --sim=sst` runs both
    prog = os.path.join(PROGRAMS, "actions", "swap_fields.kex")
    out = str(tmp_path / "swap")
    r = subprocess.run(KEXC + ["compile", prog, "--out", out, "--act=false", "--quiet"], capture_output=True, cwd=ROOT, env=ENV)
    assert r.returncode == 1 and b"action symbols" in r.stderr
    r = subprocess.run(KEXC + ["compile", prog, "--out", out, "--quiet"], capture_output=True, cwd=ROOT, env=ENV)
    assert r.returncode == 0, r.stderr
    i = subprocess.run([out, "-i"], capture_output=True)
    assert i.returncode == 2 and b"phases: 2" in i.stdout and b"action interpreter" in i.stdout
    for sim in ("sst", "lockstep"):
        s = subprocess.run(KEXC + ["simulate", "--quiet", "--sim=" + sim, prog], input=b"k=v=w\n=\n", capture_output=True,
                           cwd=ROOT, env=ENV)
        assert s.returncode == 0 and s.stdout == b"v=w=k\n=\n"


@pytest.mark.gpu
def test_action_program_binary(tmp_path, launcher_kind):
    out = str(tmp_path / "rev")
    r = subprocess.run(KEXC + ["compile", os.path.join(PROGRAMS, "actions", "reverse_items.kex"), "--out", out, "--quiet"],
                       capture_output=True, cwd=ROOT, env=ENV)
    assert r.returncode == 0, r.stderr
    ok = subprocess.run([out], input=b"a;bb;ccc;", capture_output=True)
    assert ok.returncode == 0 and ok.stdout == b"ccc;bb;a;"
    bad = subprocess.run([out], input=b"a;bb", capture_output=True)
    assert bad.returncode == 1 and bad.stderr.startswith(b"Match error at input symbol")


def test_malformed_programs_exit_with_a_message(tmp_path):
    # parse errors and self-embedding grammars (WellFormedness.hs:205-243) are reported, exit status 1
    for k, (text, msg) in enumerate([('main := /a/ main /b/ | ""\n', b"Strict occurrences"), ("main := /a\n", b"expected"),
                                     ("main := /a/\nmain := /b/\n", b"Multiple declarations")]):
        src = tmp_path / ("m%d.kex" % k)
        src.write_text(text, encoding="utf-8")
        r = subprocess.run(KEXC + ["compile", str(src), "--out", str(tmp_path / "o"), "--quiet"], capture_output=True,
                           cwd=ROOT, env=ENV)
        assert r.returncode == 1 and msg in r.stderr and b"Traceback" not in r.stderr


def test_simulate_default_flags_and_regex_flavour(tmp_path):
    """`simulate --sim=sst` with the reference's default flags runs the oracle +
    action SSTs (tables, lookahead, suppressed bits) on the CPU; a `.re` file is
    the regular-expression flavour: the bit-coded parse."""
    prog = os.path.join(PROGRAMS, "add-commas.kex")
    for flags in (["--sim=sst"], ["--sim=sst", "--la=false"], ["--sim=sst", "--act=false"], ["--sim=sst", "--act=false", "--la=false"],
                  ["--sim=sst", "--sb=false"]):
        r = subprocess.run(KEXC + ["simulate", prog, "--quiet", *flags], input=b"1234567\n", capture_output=True, cwd=ROOT, env=ENV)
        assert (r.returncode, r.stdout) == (0, b"1,234,567\n"), flags
    re_file = tmp_path / "t.re"
    re_file.write_text("(a|b)*c\n")
    r = subprocess.run(KEXC + ["simulate", str(re_file), "--quiet", "--sb=false"], input=b"abc", capture_output=True, cwd=ROOT, env=ENV)
    assert (r.returncode, r.stdout) == (0, bytes([0, 0, 0, 1, 1]))       # iterate, a, iterate, b, leave
    r = subprocess.run(KEXC + ["simulate", str(re_file), "--quiet"], input=b"abx", capture_output=True, cwd=ROOT, env=ENV)
    assert r.returncode == 1


def test_compile_regex_flavour_and_reference_phases(tmp_path):
    """`kexc compile x.re` builds the one-phase coder; `--phases=reference` builds an oracle phase and an
    action phase per stage (what the reference's default `--act=true` compiles)."""
    import struct
    re_file = tmp_path / "t.re"
    re_file.write_text("(a|b)*c[0-9]+\n")
    out = str(tmp_path / "coder")
    r = subprocess.run(KEXC + ["compile", str(re_file), "--out", out, "--quiet"], capture_output=True, cwd=ROOT, env=ENV)
    assert r.returncode == 0, r.stderr
    blob = open(out + ".kexprog", "rb").read()
    assert struct.unpack_from("<I", blob, 8)[0] == 1 and b"coder" in open(out + ".kexprog.info", "rb").read()
    out2 = str(tmp_path / "pair")
    r = subprocess.run(KEXC + ["compile", os.path.join(PROGRAMS, "csv2json.kex"), "--out", out2, "--phases=reference", "--quiet"],
                       capture_output=True, cwd=ROOT, env=ENV)
    assert r.returncode == 0, r.stderr
    assert struct.unpack_from("<I", open(out2 + ".kexprog", "rb").read(), 8)[0] == 2
    assert b"--phases=reference" in open(out2 + ".kexprog.info", "rb").read()


@pytest.mark.gpu
def test_regex_coder_and_reference_phase_binaries(tmp_path, launcher_kind):
    """The binaries: the coder writes the bit-coded parse; with --phases=reference `-p 1` writes the
    reference's oracle code and `-p 2` turns it into the output (crt/crt.c:372-467)."""
    from kleenexlang_b200.frontend.driver import build_coder_ssts, build_oracle_action_pipeline
    from kleenexlang_b200.frontend.sst import run_sst
    env = dict(ENV, KEXC_LAUNCHER=launcher_kind)
    re_file = tmp_path / "t.re"
    re_file.write_text("(a|b)*c[0-9]+\n")
    out = str(tmp_path / "coder")
    assert subprocess.run(KEXC + ["compile", str(re_file), "--out", out, "--quiet"], cwd=ROOT, env=env).returncode == 0
    want = run_sst(build_coder_ssts("(a|b)*c[0-9]+", 3, suppress_bits=True)[0], b"abbac2016")[1]
    r = subprocess.run([out], input=b"abbac2016", capture_output=True)
    assert (r.returncode, r.stdout) == (0, want)
    assert subprocess.run([out], input=b"abx", capture_output=True).returncode == 1
    out2 = str(tmp_path / "pair")
    prog = os.path.join(PROGRAMS, "csv2json.kex")
    assert subprocess.run(KEXC + ["compile", prog, "--out", out2, "--phases=reference", "--quiet"], cwd=ROOT, env=env).returncode == 0
    from conftest import sample
    d = sample("csv_sample.csv")
    ref = build_oracle_action_pipeline(open(prog, encoding="utf-8").read(), 3, lookahead=False, suppress_bits=True)
    code = run_sst(ref[0], d)[1]
    r1 = subprocess.run([out2, "-p", "1"], input=d, capture_output=True)
    assert (r1.returncode, r1.stdout) == (0, code)
    r2 = subprocess.run([out2, "-p", "2"], input=code, capture_output=True)
    whole = subprocess.run([out2], input=d, capture_output=True)
    assert r2.returncode == whole.returncode == 0 and r2.stdout == whole.stdout
    assert whole.stdout == open(os.path.join(ROOT, "tests", "golden", "csv2json_sample.out"), "rb").read()
