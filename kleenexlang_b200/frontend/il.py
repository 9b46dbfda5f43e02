"""SST -> intermediate language `Program`.

Restates src/KMC/Program/IL.hs:15-90 (the `Program` record and the `Instr` /
`Expr` sets actually produced) and src/KMC/SSTCompiler.hs:37-200
(`compile`, `compileState`, `compileTransitions`, `kvtree`,
`compileAssignment`, `orderAssignments`) together with the byte-range
predicate lowering of src/KMC/SSTCompiler/Classes.hs:56-83 and the copy
function lowering of :135-145.  This is the data boundary both
back ends consume: the reference-style C emitter under oracle/ and the CUDA
table builder in kleenexlang_b200/kexprog.py.

Programs of the default `--act=true` mode (frontend/oracle_action.py) also use
tables (`AppendTblI`, IL.hs:37-39; built in Classes.hs:95-121): the oracle
writes byte -> code digit, the action program code digit -> byte.

Instructions (tagged tuples):
  ("accept",) ("fail",) ("append", buf, constid) ("appendsym", buf, i) ("appendtbl", buf, tableid, i)
  ("concat", dst, src) ("reset", buf) ("if", expr, block) ("goto", blockid)
  ("next", lo, hi, block) ("consume", k)
Expressions:
  ("sym", i) ("avail",) ("const", n) ("true",) ("false",) ("cmp", i, [bytes])
  ("lte", a, b) ("gte", a, b) ("eq", a, b) ("or", a, b) ("and", a, b)
"""
from . import byteset as BS


class Program:
    def __init__(self):
        self.in_bits = 8
        self.out_bits = 8
        self.tables = {}
        self.constants = {}      # constid -> tuple of bytes
        self.stream_buffer = 0
        self.buffers = []
        self.init_block = 0
        self.blocks = {}         # blockid -> [instr]


def _pred_to_expr(p, j):
    tests = []
    for lo, hi in BS.to_ranges(p):
        if lo == hi:
            tests.append(("eq", ("sym", j), ("const", lo)))
        else:
            tests.append(("and", ("lte", ("const", lo), ("sym", j)), ("lte", ("sym", j), ("const", hi))))
    if not tests:
        return ("false",)
    e = tests[-1]
    for t in reversed(tests[:-1]):
        e = ("or", t, e)
    return e


def pred_list_to_expr(ps, i):
    """predListToExpr for byte range sets (Classes.hs:56-83)."""
    if not ps:
        return ("true",)
    k = 0
    while k < len(ps) and BS.size(ps[k]) == 1:
        k += 1
    eqs = ps[:k]
    m = k
    while m < len(ps) and BS.size(ps[m]) > 1:
        m += 1
    cx = ps[k:m]
    exprs = []
    if len(eqs) == 1:
        exprs.append(_pred_to_expr(eqs[0], i))
    elif len(eqs) > 1:
        exprs.append(("cmp", i, [BS.to_list(p)[0] for p in eqs]))
    for n, p in enumerate(cx):
        exprs.append(_pred_to_expr(p, i + len(eqs) + n))
    exprs.append(pred_list_to_expr(ps[m:], i + m))
    e = exprs[-1]
    for t in reversed(exprs[:-1]):
        e = ("and", t, e)
    return e


def order_assignments(upd):
    """Topological order: an assignment that reads register v runs before the
    assignment that overwrites v (SSTCompiler.hs:85-99)."""
    dag = {k: [a[1] for a in w if a[0] == "v" and a[1] != k] for k, w in upd.items()}
    mark = set()
    acc = []

    def visit(v, temp):
        if v in temp:
            raise ValueError("Not a DAG")
        if v in mark or v not in upd:
            mark.add(v)
            return
        for u in dag.get(v, ()):
            visit(u, temp | {v})
        mark.add(v)
        acc.insert(0, (v, upd[v]))

    for v in sorted(upd):
        visit(v, frozenset())
    return acc


def _compile_assignment(var, atoms, cmap, tmap=None):
    if atoms and atoms[0] == ("v", var):
        rest, pre = atoms[1:], []
    else:
        rest, pre = atoms, [("reset", var)]
    for a in rest:
        if a[0] == "v":
            pre.append(("concat", var, a[1]))
        elif a[0] == "c":
            pre.append(("append", var, cmap[a[1]]))
        elif a[0] == "t":
            pre.append(("appendtbl", var, tmap[a[1]], a[2] if len(a) > 2 else 0))
        else:
            pre.append(("appendsym", var, a[1] if len(a) > 1 else 0))
    return pre


def kvtree(xs):
    """`kvtree` / `lcp` (SSTCompiler.hs:37-55): tests with a common prefix share a
    root; -> (action or None, [(prefix tuple, subtree)]) with the groups in the
    Data.Map order of their first predicate."""
    here = [b for ps, b in xs if not ps]
    if len(here) > 1:
        raise ValueError("Ambiguous transition map")
    groups = {}
    for ps, b in xs:
        if ps:
            groups.setdefault(ps[0], []).append((ps, b))

    def lcp(ys):
        if not ys or any(not ps for ps, _ in ys):
            return (), ys
        h = ys[0][0][0]
        if all(ps[0] == h for ps, _ in ys):
            p, rest = lcp([(ps[1:], b) for ps, b in ys])
            return (h,) + p, rest
        return (), ys

    out = []
    for first in sorted(groups, key=BS.to_ranges):
        # fromListWith (++) prepends: the group is in reverse order of insertion -- irrelevant to lcp
        pre, rest = lcp(groups[first])
        out.append((pre, kvtree(rest)))
    return (here[0] if here else None, out)


def compile_sst(sst) -> Program:
    """`compile` (SSTCompiler.hs:158-200): single-symbol SSTs (direct, oracle,
    action) and lookahead SSTs (`sst.la`: predicate tuples, nested tests)."""
    consts = set()
    for es in sst.edges.values():
        for _, upd, _ in es:
            for w in upd.values():
                consts.update(a[1] for a in w if a[0] == "c")
    for w in sst.final.values():
        consts.update(a[1] for a in w if a[0] == "c")
    cmap = {c: i for i, c in enumerate(sorted(consts))}
    # tables in the Data.Map order of their range sets (`tmap`, SSTCompiler.hs:176-181); a table is
    # identified here by its contents
    tabs = set()
    for es in sst.edges.values():
        for _, upd, _ in es:
            for w in upd.values():
                tabs.update(a[1] for a in w if a[0] == "t")       # ("t", table[, index])
    tmap = {t: i for i, t in enumerate(sorted(tabs))}
    action = getattr(sst, "action", False)
    prog = Program()
    prog.tables = {i: t for t, i in tmap.items()}
    prog.constants = {i: c for c, i in cmap.items()}
    prog.stream_buffer = 0
    prog.buffers = sorted(sst.variables() | {0})
    prog.init_block = sst.initial
    la = getattr(sst, "la", False)

    def transitions(i, tree):
        """compileTransitions (SSTCompiler.hs:113-130): deeper tests first, then this node's own action."""
        action_, tests = tree
        block = []
        for ps, sub in tests:
            test = ("and", ("gte", ("avail",), ("const", i + len(ps))), pred_list_to_expr(list(ps), i))
            block.append(("if", test, transitions(i + len(ps), sub)))
        if action_ is not None:
            upd, q2 = action_
            for v, w in order_assignments(upd):
                block.extend(_compile_assignment(v, w, cmap, tmap))
            block += [("consume", i), ("goto", q2)]
        return block

    for q in range(sst.nstates):
        fin = sst.final.get(q)
        if fin is None:
            eof = [("fail",)]
        else:
            eof = _compile_assignment(0, fin, cmap) + [("accept",)]
        if la:
            es = sst.edges.get(q, ())
            lens = [len(ps) for ps, _, _ in es] or [1]
            block = [("next", min(lens), max(lens), eof)]          # compileState, SSTCompiler.hs:138-156
            block += transitions(0, kvtree([(tuple(ps), (upd, q2)) for ps, upd, q2 in es]))
            block.append(("fail",))
            prog.blocks[q] = block
            continue
        block = [("next", 1, 1, eof)]
        trans = sorted(sst.edges.get(q, ()), key=lambda e: BS.to_ranges(e[0]))
        for p, upd, q2 in trans:
            body = []
            for v, w in order_assignments(upd):
                body.extend(_compile_assignment(v, w, cmap, tmap))
            body += [("consume", 1), ("goto", q2)]
            if action:
                # ConstLab c -> next[0] == c, AnyLab -> no test (Classes.hs:94-101)
                pe = ("true",) if BS.size(p) == 256 else ("eq", ("sym", 0), ("const", BS.to_list(p)[0]))
            else:
                pe = pred_list_to_expr([p], 0)
            test = ("and", ("gte", ("avail",), ("const", 1)), pe)
            block.append(("if", test, body))
        block.append(("fail",))
        prog.blocks[q] = block
    return prog
