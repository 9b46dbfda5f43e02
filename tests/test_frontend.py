"""Front-end restatement against the reference's own golden vectors
(test/test_compiled, test/test_simulated, test/Tests/Regression.hs, README)."""
import pytest

from conftest import load_vectors, vec_matches, program_source, sample
from kleenexlang_b200.frontend.driver import build_ssts, simulate_lockstep, simulate_sst
from kleenexlang_b200.frontend.kleenex import parse_kleenex, KleenexSyntaxError
from kleenexlang_b200.frontend.regex import parse_regex

VECS = load_vectors()


@pytest.mark.parametrize("v", VECS, ids=[v["name"] for v in VECS])
def test_lockstep_simulator(v):
    # `kexc simulate --sim=lockstep` (test_simulated/runtest.sh:21-22); handles register programs too
    assert vec_matches(v, simulate_lockstep(v["program"], v["input"]))


@pytest.mark.parametrize("opt", [0, 1, 3])
@pytest.mark.parametrize("v", [v for v in VECS if not v["uses_registers"]],
                         ids=[v["name"] for v in VECS if not v["uses_registers"]])
def test_sst_simulator(v, opt):
    # `--sim=sst` on the determinized, optimised SSTs
    assert vec_matches(v, simulate_sst(build_ssts(v["program"], opt), v["input"]))


def test_register_program_rejected_for_direct_sst():
    v = [v for v in VECS if v["uses_registers"]][0]
    with pytest.raises(ValueError):
        build_ssts(v["program"])


def test_regex_dialect():
    assert parse_regex("a{2,}") == ("range", ("chr", 97), 2, None)
    assert parse_regex("a{,3}") == ("range", ("chr", 97), 0, 3)
    assert parse_regex("[a-z-]") == ("class", True, [(97, 122), (45, 45)])
    assert parse_regex("[^\\n\\]]") == ("class", False, [(10, 10), (93, 93)])
    assert parse_regex("a*?") == ("lazystar", ("chr", 97))
    assert parse_regex("\\x41\\u00e6")[0] == "concat"
    assert parse_regex("(?:ab)|c")[0] == "branch"


def test_kleenex_surface():
    pl, decls = parse_kleenex('start: a >> b\n// c\na := /x/ "y" /* z */ | ~b*\nb := r@(/q/) !r [r <- "k" r] [r += "z"] 1')
    assert pl == ["a", "b"]
    assert decls[0][1][0] == "sum"
    assert decls[1][1][1][0] == ("redirect", "r", ("re", ("chr", 113)))
    with pytest.raises(KleenexSyntaxError):
        parse_kleenex("main := /a/<2>")
    with pytest.raises(KleenexSyntaxError):
        parse_kleenex("main := ")


def test_csv2json_sample():
    # 608 B in -> 1 878 B out, hand-derived first record
    out = simulate_sst(build_ssts(program_source("csv2json")), sample("csv_sample.csv"))
    assert len(out) == 1878
    assert out.startswith(b'{\n   "id"         : 1,\n   "first_name" : "Louise",\n')


def test_state_counts_stable():
    s = build_ssts(program_source("csv2json"))[0]
    assert (s.nstates, len(s.edges)) == (27, 27)


def test_well_formedness():
    # checkWellFormedness (src/KMC/Kleenex/WellFormedness.hs:205-243)
    from kleenexlang_b200.frontend.driver import build_transducers
    from kleenexlang_b200.frontend.kleenex import KleenexWellFormednessError
    for src, what in [('main := /a/ main /b/ | ""', "Strict occurrences"),
                      ('main := a\na := /x/ b /y/ | ""\nb := a', "Strict occurrences"),
                      ('main := /a/\nmain := /b/', "Multiple declarations"),
                      ('start: foo\nmain := /a/', "Undeclared nonterminals")]:
        with pytest.raises(KleenexWellFormednessError, match=what):
            build_transducers(src)
    # tail recursion and recursion through other nonterminals are fine
    assert len(build_transducers('main := /a/ main | ""')[0].states) == 4
    assert build_transducers('main := a*\na := /x/ b\nb := /y/ | /z/ b')
    # the reference's check lets a tail call inside a starred term through and then builds states forever;
    # here the construction gives up
    with pytest.raises(ValueError, match="self-embedding"):
        build_transducers('main := (/a/ | "b" main)*')


@pytest.mark.parametrize("name", ["csv2json", "iso_datetime_to_json", "thousand_sep", "add-commas", "fastq2fasta", "apache_log"])
def test_bundled_programs_are_well_formed(name):
    from kleenexlang_b200.frontend.kleenex import check_well_formedness
    assert check_well_formedness(*parse_kleenex(program_source(name)))
