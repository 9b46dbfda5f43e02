"""Reduced grammar -> finite state transducer, and the lockstep FST simulator.

Restates src/KMC/SymbolicFST/Transducer.hs:57-133 (`constructTransducer`,
`follow` contraction, `projectTransducer`) and src/KMC/SymbolicFST.hs:226-380
(`coarsestPredicateSet`, `rightClosure`, `run`).

States are continuation stacks (tuples of reduced-grammar identifiers); the
only final state is the empty stack.  Edge order encodes choice priority.

  sym[q] = [(byteset, func, q')]   func = "copy" | ("const", (out symbols...))
                                   | ("code", byteset): oracle machines only (frontend/oracle_action.py)
  eps[q] = [((out symbols...), q')]

Output symbols are ints (bytes) or action tuples ("push",)/("pop", r)/("write", r).
"""
from . import byteset as BS


class FST:
    def __init__(self, states, sym, eps, initial):
        self.states = states
        self.sym = sym
        self.eps = eps
        self.initial = initial
        self.final = ()

    def is_final(self, q):
        return q == ()

    def has_actions(self):
        for es in self.sym.values():
            for _, f, _ in es:
                if f != "copy" and f[0] == "const" and any(not isinstance(y, int) for y in f[1]):
                    return True
        for es in self.eps.values():
            for ys, _ in es:
                if any(not isinstance(y, int) for y in ys):
                    return True
        return False


MAX_FST_STATES = 2000000
MAX_STACK_DEPTH = 4096          # continuation stacks of well-formed grammars are as deep as the definitions nest


def construct_transducer(decls, initial):
    """constructTransducer (Transducer.hs:57-107)."""

    def follow(st):
        while st:
            d = decls[st[0]]
            if d[0] == "seq":
                st = d[1] + st[1:]
            elif d[0] == "sum" and len(d[1]) == 1:
                st = (d[1][0],) + st[1:]
            else:
                break
        return st

    states = set()
    sym = {}
    eps = {}
    work = [(initial,)]
    while work:
        q = work.pop()
        if q in states:
            continue
        states.add(q)
        if len(states) > MAX_FST_STATES or len(q) > MAX_STACK_DEPTH:
            # only reachable for self-embedding grammars that slip through the (restated) check, e.g. a
            # recursive call in tail position of a starred term: the reference builds states forever there
            raise ValueError("transducer construction does not terminate: the grammar is self-embedding "
                             "(continuation stack deeper than %d or more than %d states)" % (MAX_STACK_DEPTH, MAX_FST_STATES))
        if not q:
            continue
        d = decls[q[0]]
        rest = q[1:]
        if d[0] == "const":
            q2 = follow(rest)
            eps[q] = [((d[1],), q2)]
            work.append(q2)
        elif d[0] == "read":
            q2 = follow(rest)
            sym[q] = [(d[1], "copy" if d[2] else ("const", ()), q2)]
            work.append(q2)
        elif d[0] == "seq":
            q2 = follow(d[1] + rest)
            eps[q] = [((), q2)]
            work.append(q2)
        elif d[0] == "sum":
            qs = [follow((j,) + rest) for j in d[1]]
            eps[q] = [((), q2) for q2 in qs]
            work.extend(qs)
        else:
            raise AssertionError(d)
    return FST(states, sym, eps, (initial,))


def coarsest_predicate_set(fst, qs):
    """coarsestPredicateSet (SymbolicFST.hs:226-238)."""
    ps = set()
    for q in qs:
        for p, _, _ in fst.sym.get(q, ()):
            ps.add(p)
    return BS.coarsest_partition(sorted(ps))


def right_closure(fst, q):
    """Ordered right closure with output (SymbolicFST.hs:241-255)."""
    out = []
    vis = set()

    def go(o, q):
        es = fst.eps.get(q)
        if not es:
            out.append((o, q))
            return
        for w, q2 in es:
            if q2 in vis:
                continue
            vis.add(q2)
            go(o + w, q2)

    go((), q)
    return out


def run_lockstep(fst, data: bytes):
    """NFA lockstep simulation with ordered pruning; returns the output symbol
    list of the highest-priority accepting path, or None on reject
    (`run`, SymbolicFST.hs:361-380).  Outputs are kept as linked cells so a
    step costs O(#live states)."""
    cur = [((o, None), q) for o, q in right_closure(fst, fst.initial)]
    for a in data:
        stepped = []
        for os, q in cur:
            for p, f, q2 in fst.sym.get(q, ()):
                if (p >> a) & 1:
                    o = (a,) if f == "copy" else f[1]
                    stepped.append(((o, os), q2))
        nxt = []
        seen = set()
        for os, q in stepped:
            for o2, q2 in right_closure(fst, q):
                if q2 in seen:
                    continue
                seen.add(q2)
                nxt.append(((o2, os), q2))
        cur = nxt
        if not cur:
            return None
    for os, q in cur:
        if fst.is_final(q):
            parts = []
            while os is not None:
                parts.append(os[0])
                os = os[1]
            out = []
            for p in reversed(parts):
                out.extend(p)
            return out
    return None


def run_actions(symbols):
    """Action semantics (src/KMC/Kleenex/Actions.hs:14-58): a stack of output
    builders plus a register bank."""
    store = {}
    stack = [bytearray()]
    for s in symbols:
        if isinstance(s, int):
            stack[-1].append(s)
        elif s[0] == "push":
            stack.append(bytearray())
        elif s[0] == "pop":
            store[s[1]] = stack.pop()
        elif s[0] == "write":
            stack[-1].extend(store.get(s[1], b""))
            store[s[1]] = bytearray()
    if len(stack) != 1:
        raise ValueError("Malformed action program: non-singleton stack on termination")
    return bytes(stack[0])
