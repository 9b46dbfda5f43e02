"""Generates tests/golden/reference_vectors.json and the sample-input
fixtures from the reference checkout (/root/reference).  Run once in the
build container; the GPU box only sees the committed outputs.

Sources of the vectors:
  test/test_compiled/src/*.kex, test/test_simulated/src/*.kex
      `// IN:` / `// OUT:` headers; input = lines joined by "\n" + trailing
      "\n" (runtest.sh uses `echo "$input"`), comparison strips trailing
      newlines (bash $(...)), so vectors carry "rstrip": true
  test/Tests/Regression.hs:74-94,109-126,182-196   inline (input, expected) pairs
  README.md:26-42                                   add-commas on "2016\n"
  bench/{perl,python}/src/*                          cross-tool outputs on the
      bundled samples (run here with perl / python3), sha256 only
"""
import base64
import glob
import hashlib
import json
import os
import re
import subprocess

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def b64(b):
    return base64.b64encode(b).decode()


def main():
    vecs = []
    for d in ("test_compiled", "test_simulated"):
        for f in sorted(glob.glob("%s/test/%s/src/*.kex" % (REF, d))):
            src = open(f, encoding="utf-8").read()
            ins = re.findall(r"^// IN:(.*)$", src, re.M)
            outs = re.findall(r"^// OUT:(.*)$", src, re.M)
            vecs.append({"name": "%s/%s" % (d, os.path.basename(f)), "origin": f[len(REF) + 1:],
                         "program": src, "input": b64(("\n".join(ins) + "\n").encode()),
                         "output": b64("\n".join(outs).encode()), "rstrip": True,
                         "uses_registers": "@" in src})
    range_prog = ('main := (test /\\n/)*\ntest := as | bs | cs | ds\n\n'
                  'as := ~/a/{3} "yep a" | ~/a*/ "nope a"\nbs := ~/b/{2,} "yep b" | ~/b*/ "nope b"\n'
                  'cs := ~/c/{,3} "yep c" | ~/c*/ "nope c"\nds := ~/d/{2,3} "yep d" | ~/d*/ "nope d"\n')
    cases = [("\n", "nope a\n"), ("a\n", "nope a\n"), ("aa\n", "nope a\n"), ("aaa\n", "yep a\n"),
             ("aaaa\n", "nope a\n"), ("b\n", "nope b\n"), ("bb\n", "yep b\n"), ("bbb\n", "yep b\n"),
             ("bbbb\n", "yep b\n"), ("c\n", "yep c\n"), ("cc\n", "yep c\n"), ("ccc\n", "yep c\n"),
             ("cccc\n", "nope c\n"), ("d\n", "nope d\n"), ("dd\n", "yep d\n"), ("ddd\n", "yep d\n"),
             ("dddd\n", "nope d\n")]
    inline = [("range_disamb", "test/Tests/Regression.hs:97-126", range_prog, cases),
              ("newline_bug", "test/Tests/Regression.hs:74-82",
               'main := ( keep "\\n" | ~drop ) ~/\\n/ main\n      | ( keep "\\n" | ~drop ) ~/\\n/\n'
               'keep := /(a|b)+(a|b)(a|b)+/\ndrop := /[^\\n]*/\n', [("aaaba\naabbaa\n", "aaaba\naabbaa\n")]),
              ("pipeline", "test/Tests/Regression.hs:86-94",
               'start: p >> a >> b\np := ~/abc/ "a"\na := /./ "b"\nb := /ab/ "c"\n   | ~/[^ab]/ "lol"\n',
               [("abc", "abc")]),
              ("suppressOutput_desugaring", "test/Tests/Regression.hs:182-185",
               "main := ~def def\ndef := /a|b/", [("aa", "a"), ("ab", "b"), ("ba", "a"), ("bb", "b")]),
              ("charclass_accept_dash", "test/Tests/Regression.hs:187-191",
               "\nmain := /[a-z-]*/", [("a-b-c-d---d-f-eerasdfs-", "a-b-c-d---d-f-eerasdfs-")]),
              ("lostOutput", "test/Tests/Regression.hs:193-196",
               'main := ~/a/"b"/c?/', [("a", "b"), ("ac", "bc")])]
    for name, origin, prog, cs in inline:
        for k, (i, o) in enumerate(cs):
            vecs.append({"name": "regression/%s/%d" % (name, k), "origin": origin, "program": prog,
                         "input": b64(i.encode()), "output": b64(o.encode()), "rstrip": False,
                         "uses_registers": False})
    vecs.append({"name": "readme/add-commas", "origin": "README.md:26-42",
                 "program": open(os.path.join(HERE, "..", "..", "programs", "add-commas.kex")).read(),
                 "input": b64(b"2016\n"), "output": b64(b"2,016\n"), "rstrip": False, "uses_registers": False})
    json.dump(vecs, open(os.path.join(HERE, "reference_vectors.json"), "w"), indent=1)

    # bundled sample inputs (small slices) + cross-tool digests
    samples = {
        "csv_sample.csv": open(REF + "/test/data/csv/csv_format1.sample.csv", "rb").read(),
        "datetime_sample.txt": open(REF + "/test/data/datetime/datetime_sample.txt", "rb").read(),
        "numbers_sample.txt": open(REF + "/test/data/numbers/numbers_small.txt", "rb").read(),
    }
    log = open(REF + "/test/data/apache_log/example.log", "rb").read()
    samples["apache_sample.log"] = log[:log.find(b"\n", 40000) + 1]
    for k, v in samples.items():
        open(os.path.join(HERE, k), "wb").write(v)

    def run(cmd, data):
        return subprocess.run(cmd, input=data, capture_output=True, check=True).stdout

    cross = {}
    cross["iso_datetime_to_json"] = {
        "input": "datetime_sample.txt",
        "perl_sha256": hashlib.sha256(run(["perl", REF + "/bench/perl/src/iso_datetime_to_json.pl"],
                                          samples["datetime_sample.txt"])).hexdigest(),
        "python_sha256": hashlib.sha256(run(["python3", REF + "/bench/python/src/iso_datetime_to_json.py"],
                                            samples["datetime_sample.txt"])).hexdigest()}
    cross["thousand_sep"] = {
        "input": "numbers_sample.txt",
        "perl_sha256": hashlib.sha256(run(["perl", REF + "/bench/perl/src/thousand_sep.pl"],
                                          samples["numbers_sample.txt"])).hexdigest()}
    # apache_log.pl prints "}]" where apache_log.kex prints "}\n]" (SURVEY §8(c)): normalise that one byte
    pl = run(["perl", REF + "/bench/perl/src/apache_log.pl"], samples["apache_sample.log"])
    assert pl.endswith(b"}]\n")
    cross["apache_log"] = {"input": "apache_sample.log",
                           "perl_sha256_normalised": hashlib.sha256(pl[:-2] + b"\n]\n").hexdigest()}
    cross["csv2json"] = {"input": "csv_sample.csv", "output_len": 1878,
                         "first_record": "{\n   \"id\"         : 1,\n   \"first_name\" : \"Louise\",\n"}
    json.dump(cross, open(os.path.join(HERE, "cross_tool.json"), "w"), indent=1)
    print(len(vecs), "vectors;", {k: len(v) for k, v in samples.items()})


if __name__ == "__main__":
    main()
