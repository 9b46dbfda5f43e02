"""The reference's default compilation mode (`--act=true`, Options.hs:155): a
transducer is split into an *oracle* machine, which reads the input and writes
one code digit per decision, and an *action* machine, which reads that code and
writes the output (symbols and register actions).

Restates
  src/KMC/Util/Coding.hs:13-62                 bitWidth, codeFixedWidthEnumSized
  src/KMC/SymbolicFST/OracleMachine.hs:44-61   `oracle`
  src/KMC/SymbolicFST/ActionMachine.hs:109-127 `action`
  src/KMC/SymbolicSST/ActionSST.hs:47-129      `actionToSST`, `interp`, `followEps`, `next`
  src/KMC/SymbolicSST.hs:122-136               `composeRegisterUpdate`
  src/KMC/Frontend/Commands.hs:118-157         generateOracleSSTs / generateActionSSTs
  src/KMC/SymbolicFST/OutputEquivalence.hs     post-dominators, `optOracle`, `optAction` (`--sb`)
for byte digits (`Word8`, src/KMC/Frontend.hs:117: base 256, so every code is
one byte as long as a choice or a range set has at most 256 members), with or
without the suppression of codes for output-equivalent choices (`--sb`,
default on) and with or without lookahead in the oracle (`--la`).

Used by the oracle side (oracle/build_ref.py --act: the reference's default
two-process binary as CPU baseline) and by `kexc simulate`; the CUDA path
evaluates the direct SSTs (the transduction is the same).

New pieces of the FST / SST vocabulary of fst.py / sst.py:
  sym func ("code", p)       write the index of the input byte in range set p (CodeArg p)
  SST atom ("t", table)      append table[input byte]; table = tuple of 256 ints
"""
from . import byteset as BS
from .fst import FST
from .sst import SST, normalize_update

BASE = 256
MAXB = (1 << 62)          # `maxBound :: Int`: registers are numbered downwards from it (ActionSST.hs:53)


def bit_width(n, base=BASE):
    """bitWidth (Coding.hs:13-21): digits needed for a domain of n values."""
    w = 0
    while base ** w < n:
        w += 1
    return w


def code_digits(n, ix, base=BASE):
    """codeFixedWidthEnumSized (Coding.hs:53-62): ix among n values, big-endian digits."""
    w = bit_width(n, base)
    out = []
    for k in range(w - 1, -1, -1):
        q, ix = divmod(ix, base ** k)
        assert q < base
        out.append(q)
    return tuple(out)


def code_table(p):
    """Table of CodeArg p (Classes.hs:108-121): byte -> its index in p, 0 outside p."""
    members = BS.to_list(p)
    assert bit_width(len(members)) == 1
    tbl = [0] * 256
    for i, b in enumerate(members):
        tbl[b] = i
    return tuple(tbl)


def decode_table(p):
    """Table of DecodeArg [p] (Classes.hs:95-98): code -> byte, only |p| entries are meaningful."""
    members = BS.to_list(p)
    return tuple(members + [0] * (256 - len(members)))


def post_dominators(fst):
    """postDominators (OutputEquivalence.hs:20-42) on the control-flow graph of the
    transitions without output: PDom(p) = {p} | intersection of PDom(q) over p --> q.
    Sets are bit masks over the states in sorted order."""
    order = sorted(fst.states)
    bit = {q: 1 << i for i, q in enumerate(order)}
    full = (1 << len(order)) - 1
    succs = {}
    for q in order:
        ss = [t for y, t in fst.eps.get(q, ()) if not y]
        ss += [t for _, f, t in fst.sym.get(q, ()) if f != "copy" and f[0] == "const" and not f[1]]
        succs[q] = ss
    dom = {q: full for q in order}
    while True:
        new = {}
        for q in order:
            m = full
            for t in succs[q]:
                m &= dom[t]
            new[q] = bit[q] | (m if succs[q] else 0)
        if new == dom:
            return order, bit, dom
        dom = new


def last_post_dominator(fst):
    """lastPostDominator (OutputEquivalence.hs:44-56): q -> the post-dominator of q
    that post-dominates every other one, if any."""
    order, bit, dom = post_dominators(fst)
    out = {}
    for s in order:
        members = [t for t in order if dom[s] & bit[t]]
        for t in members:
            if all(t2 == t or t2 == s or (dom[t2] & bit[t]) for t2 in members):
                out[s] = t
                break
    return out


def oracle_fst(fst, lpdom=None):
    """`oracle` (OracleMachine.hs:44-61): drop the outputs, write a code for every
    non-deterministic choice and for every copied byte of a non-singleton set.
    With `lpdom` (`--sb`, optOracle, OutputEquivalence.hs:63-67): no code for the
    choices of a state whose alternatives all rejoin at its last post-dominator
    without output."""
    sym, eps = {}, {}
    for q, es in fst.sym.items():
        new = []
        for p, f, q2 in es:
            if f == "copy" and BS.size(p) > 1:
                new.append((p, ("code", p), q2))
            else:
                new.append((p, ("const", ()), q2))
        sym[q] = new
    for q, es in fst.eps.items():
        if len(es) == 1:
            eps[q] = [((), es[0][1])]
        else:
            assert all(not y for y, _ in es), "choice edges carry no output (Transducer.hs:87-91)"
            silent = lpdom is not None and lpdom.get(q, q) != q
            eps[q] = [(() if silent else code_digits(len(es), ix), q2) for ix, (_, q2) in enumerate(es)]
    return FST(fst.states, sym, eps, fst.initial)


class ActionFST:
    """Action machine: deterministic by construction.
    sym[q] = [(label, func, q')], label = ("any",) | ("const", byte),
    func = ("decode", p) | ("const", outs);   eps[q] = [(outs, q')]."""

    def __init__(self, states, sym, eps, initial):
        self.states, self.sym, self.eps, self.initial = states, sym, eps, initial

    def is_final(self, q):
        return q == ()


def action_fst(fst, lpdom=None):
    """`action` (ActionMachine.hs:109-127).  With `lpdom` (`--sb`, optAction,
    OutputEquivalence.hs:58-61) every edge leads to the last post-dominator of its
    target: the silent region the oracle no longer codes is skipped."""
    sym, eps = {}, {}
    for q in fst.states:
        ns, ne = [], []
        for p, f, q2 in fst.sym.get(q, ()):
            if f == "copy" and BS.size(p) > 1:
                assert bit_width(BS.size(p)) == 1
                ns.append((("any",), ("decode", p), q2))
            elif f == "copy":
                ne.append(((BS.to_list(p)[0],), q2))
            else:
                ne.append((tuple(f[1]), q2))
        es = fst.eps.get(q, ())
        if len(es) == 1:
            ne.append((tuple(es[0][0]), es[0][1]))
        elif len(es) > 1:
            for ix, (y, q2) in enumerate(es):
                code = code_digits(len(es), ix)
                assert len(code) == 1
                ns.append((("const", code[0]), ("const", tuple(y)), q2))
        if lpdom is not None:
            ns = [(lab, f, lpdom.get(q2, q2)) for lab, f, q2 in ns]
            ne = [(y, lpdom.get(q2, q2)) for y, q2 in ne]
        if ns:
            sym[q] = ns
        if ne:
            eps[q] = ne
    return ActionFST(fst.states, sym, eps, fst.initial)


def compose_register_update(k1, k2):
    """composeRegisterUpdate k1 k2 (SymbolicSST.hs:122-136): k1 first, then k2."""
    out = {}
    for v in set(k1) | set(k2):
        atoms = []
        for a in k2.get(v, (("v", v),)):
            if a[0] == "v":
                atoms.extend(k1.get(a[1], (("v", a[1]),)))
            else:
                atoms.append(a)
        out[v] = normalize_update(tuple(atoms))
    return out


def _registers(am):
    """`registers` (ActionSST.hs:70-81)."""
    regs = set()

    def scan(ys):
        for y in ys:
            if not isinstance(y, int) and y[0] in ("pop", "write"):
                regs.add(y[1])
    for es in am.sym.values():
        for _, f, _ in es:
            if f[0] == "const":
                scan(f[1])
    for es in am.eps.values():
        for ys, _ in es:
            scan(ys)
    return sorted(regs)


def _interp(regs, h, ys):
    """`interp` (ActionSST.hs:84-104): a sequence of output symbols and register
    actions as a register update; the builder at stack height h is variable h."""
    if not ys:
        return h, {}
    y, rest = ys[0], ys[1:]
    if isinstance(y, int):
        h2, kappa = _interp(regs, h, rest)
        return h2, compose_register_update({h: (("v", h), ("c", (y,)))}, kappa)
    if y[0] == "push":
        return _interp(regs, h + 1, rest)
    if y[0] == "pop":
        h2, kappa = _interp(regs, h - 1, rest)
        return h2, compose_register_update({h: (), regs[y[1]]: (("v", h),)}, kappa)
    if y[0] == "write":
        h2, kappa = _interp(regs, h, rest)
        r = regs[y[1]]
        return h2, compose_register_update({r: (), h: (("v", h), ("v", r))}, kappa)
    raise AssertionError(y)


def _follow_eps(am, q):
    out = ()
    seen = set()
    while True:
        es = am.eps.get(q, ())
        if not es:
            return out, q
        if len(es) > 1 or q in seen:
            raise ValueError("non-deterministic action machine")
        seen.add(q)
        out, q = out + tuple(es[0][0]), es[0][1]


def action_to_sst(am):
    """`actionToSST` (ActionSST.hs:47-68) + enumerateStates / enumerateVariables
    (Commands.hs:144-147).  Transitions read one code byte: a fixed one (choice)
    or any (decoded through a table)."""
    import sys
    regs = {r: MAXB - i for i, r in enumerate(_registers(am))}
    init = (am.initial, 0)
    work, states, edges, outs = [init], set(), [], {}
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))
    while work:
        s = work.pop()
        if s in states:
            continue
        states.add(s)
        q, h = s
        y, q1 = _follow_eps(am, q)
        h1, kappa = _interp(regs, h, y)
        if am.is_final(q1):
            fin = kappa.get(0, (("v", 0),))
            assert all(a[0] in ("v", "c") for a in fin), "input function emitted along epsilon path"
            outs[s] = fin
        for lab, f, q2 in am.sym.get(q1, ()):
            us, q3 = _follow_eps(am, q2)
            if f[0] == "const":
                h2, k2 = _interp(regs, h1, f[1])
            else:
                h2, k2 = h1, {h1: (("v", h1), ("t", decode_table(f[1])))}
            h3, k3 = _interp(regs, h2, us)
            ru = compose_register_update(kappa, compose_register_update(k2, k3))
            s2 = (q3, h3)
            p = BS.complement(0) if lab[0] == "any" else BS.singleton(lab[1])
            edges.append((s, p, ru, s2))
            work.append(s2)
    # enumerateStates: Data.Set order of (state, height); initial state keeps its rank
    order = sorted(states, key=lambda s: (_state_key(s[0]), s[1]))
    sid = {s: i for i, s in enumerate(order)}
    vs = {0}
    for _, _, ru, _ in edges:
        vs.update(ru)
        for w in ru.values():
            vs.update(a[1] for a in w if a[0] == "v")
    for w in outs.values():
        vs.update(a[1] for a in w if a[0] == "v")
    vid = {v: i for i, v in enumerate(sorted(vs))}

    def ren(w):
        return tuple(("v", vid[a[1]]) if a[0] == "v" else a for a in w)

    E = {}
    for s, p, ru, s2 in edges:
        E.setdefault(sid[s], []).append((p, {vid[v]: ren(w) for v, w in ru.items()}, sid[s2]))
    sst = SST(len(order), E, sid[init], {sid[s]: ren(w) for s, w in outs.items()}, len(vid))
    sst.action = True          # predicates are ConstLab / AnyLab (Classes.hs:94-101)
    return sst


def _state_key(q):
    return tuple(q)


def build_oracle_action_ssts(fst, opt=3, lookahead=False, suppress_bits=False):
    """One pipeline stage as the reference's default mode compiles it
    (compileOracleAction, Commands.hs:204-244): (oracle SST, action SST).
    `lookahead` (`--la`) only concerns the oracle (Commands.hs:126: the action
    machine is deterministic and never looks ahead)."""
    from .sst import sst_from_fst, optimize
    lpdom = last_post_dominator(fst) if suppress_bits else None        # `--sb`, Commands.hs:101-110
    o = optimize(sst_from_fst(oracle_fst(fst, lpdom), lookahead=lookahead), opt)
    a = optimize(action_to_sst(action_fst(fst, lpdom)), opt, persistent=True)
    a.action = True
    return o, a
