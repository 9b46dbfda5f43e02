"""Generates tests/golden/reference_bench_vectors.json: the reference's own benchmark programs
WITHOUT register actions (bench/kleenex/src/*.kex, 40 of its 52) on inputs produced by a seeded
random walk over the program's own determinized transducer -- every input is accepted by
construction, whatever the grammar -- plus the reference's bundled sample where one exists.
The expected output is the reference's executable semantics, `kexc simulate --sim=lockstep`
(Commands.hs:277-289: lockstep simulation of the FST, then the action interpretation), which shares
nothing with the SST construction, the table builder or the kernels it pins.  Needs
/root/reference; the vectors are committed (the program text travels with its vector)."""
import base64, glob, json, os, random, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from kleenexlang_b200.frontend.driver import build_transducers, build_ssts, simulate_lockstep
from kleenexlang_b200.frontend import byteset as BS
SRC = "/root/reference/bench/kleenex/src"
DATA = "/root/reference/test/data"
SAMPLES = {"apache_log": "apache_log/example.log", "csv2json": "csv/csv_format1.sample.csv",
           "csv2json_nows": "csv/csv_format1.sample.csv", "csv_project3": "csv/csv_format1.sample.csv",
           "dna_regex_noalias": "dna/regexdna-input-noalias.txt", "ini2json": "ini/php.ini", "irc": "irc/irc.txt",
           "url": "url/test_urls.txt", "email": "email/emails_from_apache.txt",
           "iso_datetime_to_json": "datetime/datetime_sample.txt", "thousand_sep": "numbers/numbers_small.txt",
           "issuu_json2sql": "issuu/sample.json", "issuu_fallback": "issuu/sample.json", "aws_json2sql": "json/aws.json"}
NICE = sum(1 << b for b in [9, 10] + list(range(32, 127)))


def walk(sst, rng, steps):
    """Seeded random walk from the initial state: `steps` transitions chosen uniformly among the
    edges of the current state (a byte of the predicate, printable if it has one), then the shortest
    way to a final state."""
    dist = {q: 0 for q in sst.final}
    frontier = list(sst.final)
    rev = {}
    for q, es in sst.edges.items():
        for _, _, q2 in es:
            rev.setdefault(q2, []).append(q)
    while frontier:
        nxt = []
        for q in frontier:
            for p in rev.get(q, ()):
                if p not in dist:
                    dist[p] = dist[q] + 1
                    nxt.append(p)
        frontier = nxt
    out, q = bytearray(), sst.initial

    def take(edge):
        p = edge[0]
        cand = BS.to_list(p & NICE) or BS.to_list(p)
        out.append(rng.choice(cand))
        return edge[2]
    for _ in range(steps):
        es = [e for e in sst.edges.get(q, ()) if e[2] in dist]
        if not es:
            break
        q = take(rng.choice(es))
    while dist.get(q, 0) > 0:
        es = [e for e in sst.edges[q] if dist.get(e[2], 1 << 30) < dist[q]]
        q = take(rng.choice(es))
    assert q in sst.final
    return bytes(out)


out = []
for f in sorted(glob.glob(SRC + "/*.kex")):
    name = os.path.basename(f)[:-4]
    src = open(f, encoding="utf-8").read()
    if any(t.has_actions() for t in build_transducers(src)):
        continue                                  # tests/golden/reference_action_vectors.json
    t0 = time.time()
    first = build_ssts(src, 3)[0]
    rng = random.Random("kex-bench-" + name)
    inputs = [walk(first, rng, n) for n in (40, 400, 1500)]
    if name in SAMPLES and os.path.isfile(os.path.join(DATA, SAMPLES[name])):
        d = open(os.path.join(DATA, SAMPLES[name]), "rb").read(3000)
        inputs.append(d[:d.rfind(b"\n") + 1])
    n = 0
    for d in inputs:
        exp = simulate_lockstep(src, d)
        if exp is None:
            continue                              # a later pipeline stage rejects what the first one wrote
        out.append({"name": name, "program": src, "input": base64.b64encode(d).decode(),
                    "output": base64.b64encode(exp).decode()})
        n += 1
    print("%-24s %d vectors  %.1fs" % (name, n, time.time() - t0))
    if not n:
        print("   no vector for", name)
path = os.path.join(ROOT, "tests", "golden", "reference_bench_vectors.json")
with open(path, "w") as fh:
    json.dump(out, fh, indent=0)
print(len(out), "vectors ->", path, os.path.getsize(path), "bytes")
