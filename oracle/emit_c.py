"""ORACLE / TEST INFRASTRUCTURE -- not part of the product path.

IL `Program` pipeline -> C source in the shape the reference's C back end
emits (src/KMC/Program/Backends/C.hs:41-83 template, :267-311 instructions,
:383-410 expressions, :439-460 constants, :470-493 init/blocks), to be
concatenated with the *verbatim* runtime `crt/crt.c` read from
/root/reference at build time (never copied into this repository) and
compiled with the reference's command line `cc -O3 -xc -D FLAG_WORDALIGNED -`
(C.hs:562-568).  Only `tests/`, `__graft_entry__` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may use this module.
"""


def _cchar(n):
    c = chr(n)
    if 32 < n < 127 and c not in "\\'":
        return "'%s'" % c
    if n == 32:
        return "' '"
    return str(n)


def _expr(e):
    k = e[0]
    if k == "sym":
        return "next[%d]" % e[1]
    if k == "avail":
        return "avail"
    if k == "const":
        return _cchar(e[1])
    if k == "false":
        return "0"
    if k == "true":
        return "1"
    if k == "cmp":
        s = "".join('""\\x%x""' % b for b in e[2])
        return 'cmp(&next[%d],(unsigned char *) "%s",%d)' % (e[1], s, len(e[2]))
    ops = {"lte": "<=", "gte": ">=", "eq": "==", "or": "||", "and": "&&"}
    if k == "gte" and e[1] == ("avail",):
        return "(avail >= %d)" % e[2][1]
    return "(%s %s %s)" % (_expr(e[1]), ops[k], _expr(e[2]))


def _instr(prog, ins, phase, ind):
    pad = " " * ind
    k = ins[0]
    sb = prog.stream_buffer
    if k == "accept":
        return [pad + "goto accept%d;" % phase]
    if k == "fail":
        return [pad + "goto fail%d;" % phase]
    if k == "append":
        n = len(prog.constants[ins[2]]) * prog.out_bits
        if ins[1] == sb:
            return [pad + "outputarray(const_%d_%d,%d);" % (phase, ins[2], n)]
        return [pad + "appendarray(&buf_%d,const_%d_%d,%d);" % (ins[1], phase, ins[2], n)]
    if k == "appendtbl":
        # prettyAppendTbl (C.hs:232-251): one 8-bit cell per table entry (digit size 1)
        arg = "tbl%d[%d][next[%d]]" % (phase, ins[2], ins[3])
        if ins[1] == sb:
            return [pad + "outputconst(%s,8);" % arg]
        return [pad + "append(&buf_%d,%s,8);" % (ins[1], arg)]
    if k == "appendsym":
        if ins[1] == sb:
            return [pad + "outputconst(next[%d],8);" % ins[2]]
        return [pad + "append(&buf_%d,next[%d],8);" % (ins[1], ins[2])]
    if k == "concat":
        if ins[1] == sb:
            return [pad + "output(&buf_%d);" % ins[2]]
        return [pad + "concat(&buf_%d,&buf_%d);" % (ins[1], ins[2])]
    if k == "reset":
        return [pad + "reset(&buf_%d);" % ins[1]]
    if k == "if":
        out = [pad + "if (%s)" % _expr(ins[1]), pad + "{"]
        for j in ins[2]:
            out += _instr(prog, j, phase, ind + 3)
        return out + [pad + "}"]
    if k == "next":
        out = [pad + "if (!readnext(%d, %d))" % (ins[1], ins[2]), pad + "{"]
        for j in ins[3]:
            out += _instr(prog, j, phase, ind + 3)
        return out + [pad + "}"]
    if k == "consume":
        return [pad + "consume(%d);" % ins[1]]
    if k == "goto":
        return [pad + "goto l%d_%d;" % (phase, ins[1])]
    raise AssertionError(ins)


def render_c(programs, crt_text, info="kleenex-b200 oracle build"):
    """`renderCProg . programsToC` for a `Left [Program]` pipeline with
    uint8_t buffer units (C.hs:495-520)."""
    n = len(programs)
    out = ["", "#define NUM_PHASES %d" % n, "#define BUFFER_UNIT_T uint8_t", crt_text, ""]
    for ph, p in enumerate(programs, 1):
        # prettyTableDecl (C.hs:413-430); cells as prettyTableExpr prints them, eight to a line
        if not p.tables:
            out.append("/* no tables */")
            continue
        size = max(len(t) for t in p.tables.values())
        rows = []
        for tid in sorted(p.tables):
            t = p.tables[tid]
            lines = [", ".join("0x%02x" % c for c in t[i:i + 8]) for i in range(0, len(t), 8)]
            rows.append("{" + ",\n".join(lines) + "}")
        out.append("const uint8_t tbl%d[%d][%d] =\n{%s};" % (ph, len(p.tables), size, ",\n".join(rows)))
    bufs = sorted({b for p in programs for b in p.buffers})
    for b in bufs:
        out.append("buffer_t buf_%d;" % b)
    for ph, p in enumerate(programs, 1):
        for cid in sorted(p.constants):
            bs = p.constants[cid]
            out.append("const buffer_unit_t const_%d_%d[%d] = {%s};" % (
                ph, cid, len(bs), ",".join("0x%x" % b for b in bs)))
    out.append("void printCompilationInfo()\n{\n  fprintf(stdout, \"%s\\n\");\n}\n" % info)
    nonstream = sorted({b for p in programs for b in p.buffers if b != p.stream_buffer})
    out.append("void init()\n{")
    for b in nonstream:
        out.append("init_buffer(&buf_%d);" % b)
    out.append("}")
    for ph, p in enumerate(programs, 1):
        out.append("void match%d()\n{\n  int i = 0;" % ph)
        out.append("goto l%d_%d;" % (ph, p.init_block))
        for bid in sorted(p.blocks):
            lines = []
            for ins in p.blocks[bid]:
                lines += _instr(p, ins, ph, 0)
            out.append("l%d_%d: %s" % (ph, bid, lines[0]))
            out.extend(lines[1:])
        out.append("  accept%d:\n    return;\n  fail%d:\n    fprintf(stderr, \"Match error at input symbol %%zu!\\n\", count);\n    exit(1);\n}" % (ph, ph))
    out.append("void match(int phase)\n{\n  switch(phase) {")
    for ph in range(1, n + 1):
        out.append("    case %d: match%d(); break;" % (ph, ph))
    out.append("    default:\n      fprintf(stderr, \"Invalid phase: %d given\\n\", phase);\n      exit(1);\n  }\n}")
    return "\n".join(out) + "\n"
