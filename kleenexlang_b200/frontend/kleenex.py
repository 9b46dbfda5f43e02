"""Kleenex surface syntax -> AST, and AST -> reduced grammar (RProg).

Restates src/KMC/Kleenex/Parser.hs:33-230 (lexing, term operator table,
`start:` pipelines) and src/KMC/Kleenex/Desugaring.hs:70-230 (regex and term
desugaring into RConst | RRead | RSeq | RSum with hash-consed identifiers,
src/KMC/Kleenex/Syntax.hs:87-102).  Approximate matching (`<k>`) is out of
scope and rejected.

Term AST (tagged tuples):
  ("var", name) ("const", bytes) ("re", regex_ast) ("seq", [t]) ("sum", [t])
  ("star", t) ("plus", t) ("question", t) ("range", m|None, n|None, t)
  ("suppress", t) ("one",) ("update", reg, [("reg", r) | ("const", bytes)])
  ("write", reg) ("redirect", reg, t)

Reduced terms:
  ("const", sym)           sym = int byte | ("push",) | ("pop", r) | ("write", r)
  ("read", byteset, copy)  ("seq", (ids...))  ("sum", (ids...))
"""
from . import byteset as BS
from .regex import parse_regex_at, RegexSyntaxError


class KleenexSyntaxError(Exception):
    pass


_ESC = {"\\": "\\", '"': '"', "n": "\n", "t": "\t", "v": "\v", "r": "\r", "f": "\f"}


class _KP:
    def __init__(self, s):
        self.s = s
        self.i = 0

    def fail(self, msg):
        line = self.s.count("\n", 0, self.i) + 1
        raise KleenexSyntaxError("%s (line %d, offset %d)" % (msg, line, self.i))

    def peek(self, k=0):
        j = self.i + k
        return self.s[j] if j < len(self.s) else ""

    def startswith(self, t):
        return self.s.startswith(t, self.i)

    def ws(self):
        # whiteSpace (Parser.hs:33-39)
        while True:
            c = self.peek()
            if c != "" and c.isspace():
                self.i += 1
            elif self.startswith("//"):
                j = self.s.find("\n", self.i)
                self.i = len(self.s) if j < 0 else j + 1
            elif self.startswith("/*"):
                j = self.s.find("*/", self.i + 2)
                if j < 0:
                    self.fail("unterminated comment")
                self.i = j + 2
            else:
                return

    def symbol(self, t):
        if self.startswith(t):
            self.i += len(t)
            self.ws()
            return True
        return False

    def ident_raw(self, lower_first=False):
        c = self.peek()
        if not c or not c.isalpha() or (lower_first and not c.islower()):
            return None
        j = self.i + 1
        while j < len(self.s) and (self.s[j].isalnum() or self.s[j] in "_-"):
            j += 1
        name = self.s[self.i:j]
        self.i = j
        return name

    def integer(self):
        j = self.i
        while self.peek() != "" and self.peek().isdigit():
            self.i += 1
        if j == self.i:
            return None
        return int(self.s[j:self.i])

    def constant(self):
        # constantP / stringConstant / escapedChar (Parser.hs:73-93,150-153)
        assert self.peek() == '"'
        self.i += 1
        out = []
        while True:
            c = self.peek()
            if c == "":
                self.fail("unterminated string constant")
            if c == '"':
                self.i += 1
                break
            if c == "\\":
                d = self.peek(1)
                if d in _ESC:
                    out.append(_ESC[d])
                    self.i += 2
                elif d == "x":
                    h = self.s[self.i + 2:self.i + 4]
                    try:
                        out.append(chr(int(h, 16)))
                    except ValueError:
                        self.fail("bad hex escape")
                    if len(h) != 2:
                        self.fail("bad hex escape")
                    self.i += 4
                else:
                    self.fail("bad escape sequence")
            else:
                out.append(c)
                self.i += 1
        self.ws()
        return "".join(out).encode("utf-8")

    # --- terms (Parser.hs:154-202)
    def atom(self):
        c = self.peek()
        if c == "1":
            self.i += 1
            self.ws()
            return ("one",)
        if c and c.isalpha():
            save = self.i
            name = self.ident_raw()
            self.ws()
            if self.startswith(":=") or self.startswith("@"):
                self.i = save
                return None
            return ("var", name)
        if c == '"':
            return ("const", self.constant())
        if c == "/":
            self.i += 1
            try:
                e, j = parse_regex_at(self.s, self.i, illegal="/")
            except RegexSyntaxError as ex:
                self.fail("regex: %s" % ex)
            self.i = j
            if self.peek() != "/":
                self.fail("expected closing / of regular expression")
            self.i += 1
            self.ws()
            return ("re", e)
        if c == "!":
            self.i += 1
            r = self.ident_raw(lower_first=True)
            if r is None:
                self.fail("expected register name")
            self.ws()
            return ("write", r)
        if c == "[":
            self.i += 1
            self.ws()
            r = self.ident_raw(lower_first=True)
            if r is None:
                self.fail("expected register name")
            self.ws()
            if self.symbol("<-"):
                atoms = []
            elif self.symbol("+="):
                atoms = [("reg", r)]
            else:
                self.fail("expected <- or +=")
            n0 = len(atoms)
            while True:
                if self.peek() == '"':
                    atoms.append(("const", self.constant()))
                else:
                    save = self.i
                    r2 = self.ident_raw(lower_first=True)
                    if r2 is None:
                        self.i = save
                        break
                    self.ws()
                    atoms.append(("reg", r2))
            if len(atoms) == n0:
                self.fail("expected register update atoms")
            if not self.symbol("]"):
                self.fail("expected ]")
            return ("update", r, atoms)
        if c == "(":
            self.i += 1
            self.ws()
            t = self.term()
            if not self.symbol(")"):
                self.fail("expected )")
            return t
        return None

    def prefixed(self):
        # one prefix operator at most (Parser.hs:158-162)
        if self.peek() == "~":
            self.i += 1
            self.ws()
            a = self.atom()
            if a is None:
                self.fail("expected term after ~")
            return ("suppress", a)
        c = self.peek()
        if c and c.islower():
            save = self.i
            r = self.ident_raw(lower_first=True)
            if r is not None and self.peek() == "@":
                self.i += 1
                a = self.atom()
                if a is None:
                    self.fail("expected term after %s@" % r)
                return ("redirect", r, a)
            self.i = save
        return self.atom()

    def postfixed(self):
        t = self.prefixed()
        if t is None:
            return None
        while True:
            if self.symbol("*"):
                t = ("star", t)
            elif self.symbol("?"):
                t = ("question", t)
            elif self.symbol("+"):
                t = ("plus", t)
            elif self.peek() == "<" and self.peek(1).isdigit():
                self.fail("approximate matching <k> is out of scope for this backend")
            elif self.peek() == "{":
                self.i += 1
                m = self.integer()
                if self.peek() == ",":
                    self.i += 1
                    n = self.integer()
                else:
                    if m is None:
                        self.fail("malformed range")
                    n = m
                if not self.symbol("}"):
                    self.fail("malformed range")
                t = ("range", m, n, t)
            else:
                return t

    def seq(self):
        ts = []
        while True:
            if ts and self.peek() == "|":
                break
            t = self.postfixed()
            if t is None:
                break
            ts.append(t)
        if not ts:
            self.fail("expected term")
        return ts[0] if len(ts) == 1 else ("seq", ts)

    def term(self):
        alts = [self.seq()]
        while self.symbol("|"):
            alts.append(self.seq())
        return alts[0] if len(alts) == 1 else ("sum", alts)

    def prog(self):
        self.ws()
        pipeline = ["main"]
        if self.symbol("start:"):
            pipeline = []
            while True:
                n = self.ident_raw()
                if n is None:
                    self.fail("expected nonterminal in pipeline")
                self.ws()
                pipeline.append(n)
                if not self.symbol(">>"):
                    break
        decls = []
        while self.i < len(self.s):
            name = self.ident_raw()
            if name is None:
                self.fail("expected declaration")
            self.ws()
            if not self.symbol(":="):
                self.fail("expected :=")
            decls.append((name, self.term()))
        if not decls:
            self.fail("expected at least one declaration")
        return pipeline, decls


def parse_kleenex(src: str):
    """-> (pipeline names, [(name, term)])  (Parser.hs:208-230)."""
    return _KP(src).prog()


# ---------------------------------------------------------------- desugaring
def _flatten(tag, t):
    if t[0] == tag:
        out = []
        for u in t[1]:
            out.extend(_flatten(tag, u))
        return out
    return [t]


class _Desugar:
    def __init__(self, names):
        self.idents = {}
        for n in names:
            for out in (True, False):
                self.idents[(n, out)] = len(self.idents)
        self.fresh = len(self.idents) + 1
        self.decls = {}
        self.rev = {}

    def get_fresh(self):
        i = self.fresh
        self.fresh += 1
        return i

    def insert(self, i, t):
        self.decls[i] = t
        self.rev[t] = i
        return i

    def decl(self, t):
        i = self.rev.get(t)
        if i is not None:
            return i
        return self.insert(self.get_fresh(), t)

    def seq(self, ids):
        return self.decl(("seq", tuple(ids)))

    def star(self, ie, lazy=False):
        ieps = self.seq([])
        i = self.get_fresh()
        iloop = self.seq([ie, i])
        return self.insert(i, ("sum", (ieps, iloop) if lazy else (iloop, ieps)))

    # desugarRE (Desugaring.hs:70-123)
    def regex(self, out, e):
        k = e[0]
        if k == "one":
            return self.seq([])
        if k == "dot":
            return self.decl(("read", BS.UNIVERSE, out))
        if k == "chr":
            bs = chr(e[1]).encode("utf-8")
            return self.seq([self.decl(("read", BS.singleton(b), out)) for b in bs])
        if k == "group":
            return self.regex(out, e[1])
        if k == "concat":
            return self.seq([self.regex(out, e[1]), self.regex(out, e[2])])
        if k == "branch":
            return self.decl(("sum", (self.regex(out, e[1]), self.regex(out, e[2]))))
        if k == "class":
            for lo, hi in e[2]:
                if lo > 255 or hi > 255:
                    raise KleenexSyntaxError("character class member outside byte range")
            rs = BS.from_ranges(e[2])
            if not e[1]:
                rs = BS.complement(rs)
            return self.decl(("read", rs, out))
        if k in ("star", "lazystar"):
            return self.star(self.regex(out, e[1]), lazy=(k == "lazystar"))
        if k in ("plus", "lazyplus"):
            ie = self.regex(out, e[1])
            istar = self.regex(out, ("star" if k == "plus" else "lazystar", e[1]))
            return self.seq([ie, istar])
        if k in ("question", "lazyquestion"):
            ie = self.regex(out, e[1])
            ieps = self.seq([])
            return self.decl(("sum", (ie, ieps) if k == "question" else (ieps, ie)))
        if k == "range":
            _, sub, n, m = e
            ie = self.regex(out, sub)
            if m is None:
                istar = self.regex(out, ("star", sub))
                return self.seq([ie] * n + [istar])
            if n == m:
                return self.seq([ie] * n)
            iq = self.regex(out, ("question", sub))
            # the reference appends m (not m-n) optional copies: Desugaring.hs:117-118
            return self.seq([ie] * n + [iq] * m)
        if k == "suppress":
            return self.regex(False, e[1])
        raise KleenexSyntaxError("unsupported regex construct %r" % (k,))

    # desugarTerm (Desugaring.hs:125-177)
    def term(self, out, t):
        k = t[0]
        if k == "var":
            key = (t[1], out)
            if key not in self.idents:
                raise KleenexSyntaxError("undeclared nonterminal %s" % t[1])
            return self.idents[key]
        if k == "const":
            return self.seq([self.decl(("const", b)) for b in (t[1] if out else b"")])
        if k == "re":
            return self.regex(out, t[1])
        if k == "seq":
            return self.seq([self.term(out, u) for u in _flatten("seq", t)])
        if k == "sum":
            return self.decl(("sum", tuple(self.term(out, u) for u in _flatten("sum", t))))
        if k == "star":
            return self.star(self.term(out, t[1]))
        if k == "plus":
            it = self.term(out, t[1])
            return self.seq([it, self.term(out, ("star", t[1]))])
        if k == "question":
            it = self.term(out, t[1])
            return self.decl(("sum", (it, self.seq([]))))
        if k == "range":
            _, m0, n, sub = t
            it = self.term(out, sub)
            m = m0 or 0
            if n is None:
                return self.seq([it] * m + [self.term(out, ("star", sub))])
            if n < m:
                raise KleenexSyntaxError("invalid range {%d,%d}" % (m, n))
            if n == m:
                return self.seq([it] * m)
            iq = self.term(out, ("question", sub))
            return self.seq([it] * m + [iq] * (n - m))
        if k == "suppress":
            return self.term(False, t[1])
        if k == "one":
            return self.seq([])
        if k == "update":
            syms = [("push",)]
            for a in t[2]:
                if a[0] == "reg":
                    syms.append(("write", a[1]))
                else:
                    syms.extend(a[1])
            syms.append(("pop", t[1]))
            return self.seq([self.decl(("const", s)) for s in syms])
        if k == "write":
            return self.decl(("const", ("write", t[1])))
        if k == "redirect":
            it = self.term(out, t[2])
            ipush = self.decl(("const", ("push",)))
            ipop = self.decl(("const", ("pop", t[1])))
            return self.seq([ipush, it, ipop])
        raise KleenexSyntaxError("unknown term %r" % (k,))


# ------------------------------------------------------------ well-formedness
class KleenexWellFormednessError(KleenexSyntaxError):
    pass


_WRAPPERS = {"star": 1, "plus": 1, "question": 1, "suppress": 1, "range": 3, "redirect": 2}


def _term_idents(t):
    """All nonterminals occurring in a term (`termIdents`, WellFormedness.hs:98-112)."""
    k = t[0]
    if k == "var":
        return {t[1]}
    if k in ("seq", "sum"):
        out = set()
        for x in t[1]:
            out |= _term_idents(x)
        return out
    if k in _WRAPPERS:
        return _term_idents(t[_WRAPPERS[k]])
    return set()


def _strict_deps(t):
    """Nonterminals in strict (non-tail) positions (`strictDeps`, WellFormedness.hs:115-128)."""
    k = t[0]
    if k == "seq":
        out = set()
        for x in t[1][:-1]:
            out |= _term_idents(x)
        return out | (_strict_deps(t[1][-1]) if t[1] else set())
    if k == "sum":
        out = set()
        for x in t[1]:
            out |= _strict_deps(x)
        return out
    if k in _WRAPPERS:
        return _strict_deps(t[_WRAPPERS[k]])
    return set()


def _succs(name, t):
    """`succs` (WellFormedness.hs:81-95): star and plus make the declaration self-referential."""
    k = t[0]
    if k == "var":
        return {t[1]}
    if k in ("seq", "sum"):
        out = set()
        for x in t[1]:
            out |= _succs(name, x)
        return out
    if k in ("star", "plus"):
        return {name} | _succs(name, t[1])
    if k in _WRAPPERS:
        return _succs(name, t[_WRAPPERS[k]])
    return set()


def check_well_formedness(pipeline, decls):
    """`checkWellFormedness` (WellFormedness.hs:205-243): no nonterminal declared twice, the pipeline only
    names declared nonterminals, and the grammar is non-self-embedding -- no strongly connected component of
    the dependency graph contains a strict (non-tail) occurrence of one of its own members.  Without the last
    property the transducer's states (continuation stacks, Transducer.hs:57-107) are unbounded."""
    declmap = {}
    for name, t in decls:
        if name in declmap:
            raise KleenexWellFormednessError("Error: Multiple declarations of nonterminal '%s'." % name)
        declmap[name] = t
    if not pipeline:
        raise KleenexWellFormednessError("Error: Empty pipeline")
    undecl = [n for n in pipeline if n not in declmap]
    if undecl:
        raise KleenexWellFormednessError("Error: Undeclared nonterminals in pipeline: " + ", ".join(undecl))
    graph = {n: sorted(x for x in _succs(n, t) if x in declmap) for n, t in declmap.items()}
    # Tarjan's strongly connected components, iterative
    index, low, on, stack, comps = {}, {}, set(), [], []
    for root in sorted(graph):
        if root in index:
            continue
        work = [(root, 0)]
        while work:
            v, i = work.pop()
            if i == 0:
                index[v] = low[v] = len(index)
                stack.append(v)
                on.add(v)
            recurse = False
            for j in range(i, len(graph[v])):
                w = graph[v][j]
                if w not in index:
                    work.append((v, j + 1))
                    work.append((w, 0))
                    recurse = True
                    break
                if w in on:
                    low[v] = min(low[v], index[w])
            if recurse:
                continue
            if low[v] == index[v]:
                comp = set()
                while True:
                    w = stack.pop()
                    on.discard(w)
                    comp.add(w)
                    if w == v:
                        break
                comps.append(comp)
            if work:
                u = work[-1][0]
                low[u] = min(low[u], low[v])
    for comp in comps:
        for name in sorted(comp):
            bad = sorted(_strict_deps(declmap[name]) & comp)
            if bad:
                raise KleenexWellFormednessError(
                    "Error: Strict occurrences in mutually recursive definition of '%s' involving nonterminals %s. "
                    "Occurrences: %s" % (name, ", ".join("'%s'" % c for c in sorted(comp)),
                                         ", ".join("'%s'" % b for b in bad)))
    return comps


def desugar(pipeline, decls):
    """desugarProg (Desugaring.hs:180-212) -> (pipeline ids, {id: reduced term})."""
    names = [n for n, _ in decls]
    d = _Desugar(names)
    for name, t in decls:
        i = d.idents[(name, True)]
        j = d.idents[(name, False)]
        i2 = d.term(True, t)
        j2 = d.term(False, t)
        d.insert(i, ("seq", (i2,)))
        d.insert(j, ("seq", (j2,)))
    pl = []
    for n in pipeline:
        if (n, True) not in d.idents:
            raise KleenexSyntaxError("identifier in pipeline with no declaration: %s" % n)
        pl.append(d.idents[(n, True)])
    return pl, d.decls
