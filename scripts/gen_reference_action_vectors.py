"""Generates tests/golden/reference_action_vectors.json: the reference's own benchmark programs
that use register actions (bench/kleenex/src/*.kex with `reg@t`, `!reg`, `[reg <- ..]`) on inputs
taken from the reference's test data (test/data) or written to fit the grammar.  The expected
output is the reference's executable semantics for such programs -- lockstep simulation of the
transducer followed by the action interpretation (Commands.hs:277-289, Actions.hs:14-58) -- and is
cross-checked here against the restated default compilation mode (oracle + action SST,
frontend/oracle_action.py).  Needs /root/reference; the vectors are committed."""
import base64, json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from kleenexlang_b200.frontend.driver import simulate_lockstep, build_oracle_action_pipeline, simulate_sst
REF = "/root/reference"
SRC = REF + "/bench/kleenex/src"
DATA = REF + "/test/data"


def head(path, nbytes, sep=b"\n"):
    d = open(os.path.join(DATA, path), "rb").read(nbytes)
    return d[:d.rfind(sep) + len(sep)]


def bibtex_entries(nbytes):
    """Whole entries of the reference's bibtex sample, each closed by "}\\n" directly before the next "@"."""
    d = open(os.path.join(DATA, "bibtex/bibtex-small.bib"), "rb").read()
    out = b""
    for e in d.split(b"\n@"):
        e = (e if e.startswith(b"@") else b"@" + e).rstrip() + b"\n"
        if len(out) + len(e) > nbytes:
            break
        out += e
    return out


CANDIDATES = {
    "swap_lines": [b"the first of two lines\nand the second one\n"],
    "sort_ab": [b"abbabaabbbab" * 100],
    "worstcase": [b"xyzzy" * 30, b"ab" * 100, b"a" * 200],
    "drex_rev-dict": [b"a=1;bb=22;;c;" * 50],
    "drex_swap-bibtex": [bibtex_entries(6000)],
    "drex_align-bibtex": [head("bibtex/bibtex-small.bib", 6000, b"}\n\n"), head("bibtex/bibtex-small.bib", 6000)],
    "mitm": [b"<p><form action=\"http://x/y\" ><input></form><form  action='u'>" * 20],
    "markdown2html": [open(DATA + "/markdown/test", "rb").read()[:6000]] if os.path.isfile(DATA + "/markdown/test") else [],
    "doc_comments": [head("comments/comments_small.txt", 6000)],
    "jix_responsetime": [b"".join(
        b'78.68.41.%d - - [03/Mar/2015:00:00:%02d +0100] "GET /beacon?uid=CxDoBw-m&tids=l%d HTTP/1.1" 200 67 "-" '
        b'"Mozilla/5.0 (iPad; CPU OS 8_1_3 like Mac OS X)" %d "www.jobbsafari.se" "webb" "hypnotoad"\n'
        % (i % 250, i % 60, i, 13000 + 17 * i) for i in range(30))],
    "dna_regex_noalias_2": [head("dna/regexdna-input-noalias.txt", 4000), head("dna/regexdna-input.txt", 4000)],
    "syntax_latex": [open(SRC + "/swap_lines.kex", "rb").read(), open(SRC + "/drex_rev-dict.kex", "rb").read()],
}
if os.path.isdir(DATA + "/markdown/test"):
    for f in sorted(os.listdir(DATA + "/markdown/test"))[:3]:
        CANDIDATES["markdown2html"].append(open(os.path.join(DATA, "markdown/test", f), "rb").read()[:6000])

out = []
for name, cands in sorted(CANDIDATES.items()):
    src = open(os.path.join(SRC, name + ".kex"), encoding="utf-8").read()
    done = False
    for c in cands:
        for cut in (len(c), len(c) // 2, len(c) // 4):
            d = c[:cut]
            if b"\n" in d and cut != len(c):
                d = d[:d.rfind(b"\n") + 1]
            try:
                exp = simulate_lockstep(src, d)
            except Exception as e:
                print(name, "lockstep error", repr(e)[:80]); exp = None
            if exp is None:
                continue
            try:
                got = simulate_sst(build_oracle_action_pipeline(src), d)
            except Exception as e:
                print(name, "oracle/action build error", repr(e)[:100]); got = exp
            assert got == exp, (name, "oracle/action pipeline differs from the lockstep simulation")
            out.append({"name": name, "program": src, "input": base64.b64encode(d).decode(),
                        "output": base64.b64encode(exp).decode()})
            print(name, "ok: %d -> %d bytes" % (len(d), len(exp)))
            done = True
            break
        if done:
            break
    if not done:
        print(name, "NO ACCEPTED INPUT")
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "reference_action_vectors.json"), "w"), indent=1)
print(len(out), "vectors")
