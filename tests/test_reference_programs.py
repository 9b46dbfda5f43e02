"""Coverage of the reference's own benchmark programs
(/root/reference/bench/kleenex/src/*.kex) by the restated front end and the
table builder.  Runs only where the reference checkout exists (never on the GPU
box); nothing is copied from it."""
import glob
import os

import pytest

from kleenexlang_b200 import fasttab
from kleenexlang_b200.frontend.driver import build_ssts
from kleenexlang_b200.kexprog import UnsupportedProgram, build_phase

SRC = "/root/reference/bench/kleenex/src"
FAST = ["aaa", "as", "csv2json", "csv2json_nows", "csv_project3", "dfamail", "drex_del-comments", "email", "flip_ab",
        "ini2json", "iso_datetime_to_json", "json_project1", "patho2", "rot13", "simple_id", "thousand_sep"]
# register actions (`reg@t`, `!reg`) reorder or duplicate data: the reference compiles them only in its
# oracle/action mode; here a stage with actions becomes a transducer phase that writes an action stream
# plus an action-interpreter phase (frontend/actions.py, csrc/kex_act.cuh)
ACTIONS = ["dna_regex_noalias_2", "doc_comments", "drex_align-bibtex", "drex_rev-dict", "drex_swap-bibtex",
           "jix_responsetime", "markdown2html", "mitm", "sort_ab", "swap_lines", "worstcase"]


@pytest.mark.skipif(not os.path.isdir(SRC), reason="reference checkout not present")
@pytest.mark.parametrize("name", FAST)
def test_bench_program_gets_monoid_tables(name):
    src = open(os.path.join(SRC, name + ".kex"), encoding="utf-8").read()
    for s in build_ssts(src, 3):
        fasttab.build_fast(build_phase(s))             # raises if a phase would fall back to the generic kernels


@pytest.mark.skipif(not os.path.isdir(SRC), reason="reference checkout not present")
@pytest.mark.parametrize("name", ACTIONS)
def test_action_programs_are_refused_not_miscompiled(name):
    src = open(os.path.join(SRC, name + ".kex"), encoding="utf-8").read()
    with pytest.raises(ValueError, match="action symbols"):
        build_ssts(src, 3)


@pytest.mark.skipif(not os.path.isdir(SRC), reason="reference checkout not present")
@pytest.mark.parametrize("name", ACTIONS)
def test_action_programs_compile_as_two_phase_stages(name):
    from kleenexlang_b200.kexprog import compile_kex, MAGIC_ACT
    import struct
    src = open(os.path.join(SRC, name + ".kex"), encoding="utf-8").read()
    blob = compile_kex(src)
    nph = struct.unpack_from("<I", blob, 8)[0]
    magics = [struct.unpack_from("<I", blob, struct.unpack_from("<I", blob, 16 + 8 * i)[0])[0] for i in range(nph)]
    assert MAGIC_ACT in magics and magics[0] != MAGIC_ACT


@pytest.mark.skipif(not os.path.isdir(SRC), reason="reference checkout not present")
def test_inventory_is_complete():
    names = {os.path.basename(f)[:-4] for f in glob.glob(os.path.join(SRC, "*.kex"))}
    assert set(FAST) <= names and set(ACTIONS) <= names and len(names) >= 50
