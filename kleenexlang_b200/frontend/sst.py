"""FST -> streaming string transducer (path-tree determinization), SST
optimisation and the reference SST interpreter.

Restates src/KMC/Determinization.hs (all), src/KMC/TreeWriter.hs (the closure
monad, as explicit tree rebuilding with a shared visited set),
src/KMC/SymbolicSST.hs:72-136 (SST type, update normalisation),
:182-331 (constant propagation `optimize`), :333-367 (enumeration) and
:406-446 (`run`).  Single-symbol mode only (`--la=false`,
Commands.hs:126,171): every transition consumes exactly one byte.

Trees:   ("tip", w, state) | ("fork", w, (children...))
Atoms:   ("v", var) | ("c", (bytes...)) | ("f",)      ("f",) = append input byte
         ("t", table)   append table[input byte] (oracle / action machines, frontend/oracle_action.py)
Vars:    tuples giving the root->leaf path of the tree node that owns the
         register; () is the designated output register.  Tuple order is tree
         pre-order, so an ancestor always sorts before its descendants.
"""
from . import byteset as BS
from .fst import coarsest_predicate_set


# ------------------------------------------------------------ tree plumbing
def _tprepend(w, t):
    if not w:
        return t
    if t[0] == "tip":
        return ("tip", w + t[1], t[2])
    return ("fork", w + t[1], t[2])


def _bind(tree, k):
    """joinTreeWriterT . fmap k (TreeWriter.hs:71-90): run continuation k on
    every tip left to right, pruning failed branches and collapsing forks
    with a single survivor."""
    if tree is None:
        return None
    if tree[0] == "tip":
        r = k(tree[2])
        if r is None:
            return None
        return _tprepend(tree[1], r)
    ts = []
    for c in tree[2]:
        r = _bind(c, k)
        if r is not None:
            ts.append(r)
    if not ts:
        return None
    if len(ts) == 1:
        return _tprepend(tree[1], ts[0])
    return ("fork", tree[1], tuple(ts))


def _reduce(tree):
    """Hoist the longest common atom prefix of a fork's children into the
    fork (Determinization.hs:63-71)."""
    if tree is None or tree[0] == "tip":
        return tree
    ts = [_reduce(c) for c in tree[2]]
    outs = [c[1] for c in ts]
    n = min(len(o) for o in outs)
    k = 0
    while k < n and all(o[k] == outs[0][k] for o in outs):
        k += 1
    if k:
        p = outs[0][:k]
        ts = [(c[0], c[1][k:], c[2]) for c in ts]
        return ("fork", tree[1] + p, tuple(ts))
    return ("fork", tree[1], tuple(ts))


def _tflat(tree):
    if tree is None:
        return []
    if tree[0] == "tip":
        return [tree[2]]
    out = []
    for c in tree[2]:
        out.extend(_tflat(c))
    return out


def _closure(fst, tree):
    """closureTree / closureTreeFunc (Determinization.hs:127-139): follow all
    non-input transitions from every tip, sharing one visited set."""
    vis = set()

    def gen(q):
        es = fst.eps.get(q)
        if not es:
            return ("tip", (), q)
        ts = []
        for out, q2 in es:
            if q2 in vis:
                continue
            vis.add(q2)
            sub = gen(q2)
            if sub is not None:
                ts.append(_tprepend((("c", tuple(out)),), sub))
        if not ts:
            return None
        if len(ts) == 1:
            return ts[0]
        return ("fork", (), tuple(ts))

    return _reduce(_bind(tree, gen))


def _eof(fst, tree):
    vis = set()

    def k(q):
        if fst.is_final(q) and q not in vis:
            vis.add(q)
            return ("tip", (), q)
        return None

    return _reduce(_bind(tree, k))


def _kill(tree, kills):
    """killTree (Determinization.hs:160-163): drop the tips in `kills`."""
    return _reduce(_bind(tree, lambda q: None if q in kills else ("tip", (), q)))


def _consume(fst, p, tree, idx=0):
    """consumeTree (Determinization.hs:152-158); idx = position of the symbol in the
    lookahead window (`inj i`, :221-223): atoms ("f",) / ("t", table) for idx 0,
    ("f", idx) / ("t", table, idx) further in."""
    vis = set()

    def k(q):
        hits = [(f, q2) for (p2, f, q2) in fst.sym.get(q, ()) if BS.is_subset(p, p2)]
        if not hits:
            return None
        if len(hits) > 1:
            raise NotImplementedError(
                "Stepping for FSTs with read-fanout greater than one is not supported yet")
        f, q2 = hits[0]
        if q2 in vis:
            return None
        vis.add(q2)
        if f == "copy":
            atom = ("f",) if idx == 0 else ("f", idx)
        elif f[0] == "code":                     # CodeArg p2 of an oracle machine: index of the byte in p2
            from .oracle_action import code_table
            atom = ("t", code_table(f[1])) if idx == 0 else ("t", code_table(f[1]), idx)
        else:
            atom = ("c", tuple(f[1]))
        return ("tip", (atom,), q2)

    return _reduce(_bind(tree, k))


def _abstract(tree):
    """Name every node by its path; returns (kappa list, abstract tree)
    (Determinization.hs:165-183)."""
    kappa = []

    def go(v, t):
        kappa.append((v, t[1]))
        if t[0] == "tip":
            return ("tip", v, t[2])
        return ("fork", v, tuple(go(v + (m,), c) for m, c in enumerate(t[2])))

    at = go((), tree)
    return kappa, at


def _unabstract_outputs(tree):
    if tree[0] == "tip":
        return ("tip", (("v", tree[1]),), tree[2])
    return ("fork", (("v", tree[1]),), tuple(_unabstract_outputs(c) for c in tree[2]))


def normalize_update(atoms):
    """normalizeUpdateStringFunc (SymbolicSST.hs:107-115)."""
    out = []
    for a in atoms:
        if a[0] == "c":
            if not a[1]:
                continue
            if out and out[-1][0] == "c":
                out[-1] = ("c", out[-1][1] + a[1])
                continue
        out.append(a)
    return tuple(out)


class SST:
    """states: 0..n-1; edges[q] = [(byteset, {var: atoms}, q')];
    final[q] = atoms (only "v"/"c"); variables are ints with 0 = output
    register, numbered in tree pre-order (`enumerateVariables`,
    SymbolicSST.hs:346-367)."""

    def __init__(self, nstates, edges, initial, final, nvars):
        self.nstates = nstates
        self.edges = edges
        self.initial = initial
        self.final = final
        self.nvars = nvars

    def variables(self):
        vs = set()
        for es in self.edges.values():
            for _, upd, _ in es:
                vs.update(upd.keys())
        for atoms in self.final.values():
            vs.update(a[1] for a in atoms if a[0] == "v")
        return vs


def expand_tables(sst):
    """Device form of an SST with table atoms (`AppendTblI`, IL.hs:37-39; the
    `tblP[n][256]` arrays of C.hs:413-430): the device emits literals and the
    input byte, so every transition whose update appends table[input byte] is
    split by the value the tables take -- one transition per distinct tuple of
    table values, the table atoms replaced by those constants (what
    `specialize`, Determinization.hs:190-206, does for singleton predicates,
    applied to every byte of the predicate).  The transduction is unchanged;
    the byte classes of the phase get finer (a code table numbers the bytes of
    its predicate, so its predicate falls apart into single bytes)."""
    if getattr(sst, "la", False):
        raise ValueError("table atoms of lookahead SSTs carry symbol indices: no single-symbol form")
    edges, changed = {}, False
    for q, es in sst.edges.items():
        out = []
        for p, upd, q2 in es:
            tabs = []
            for w in upd.values():
                for a in w:
                    if a[0] == "t" and a[1] not in tabs:
                        tabs.append(a[1])
            if not tabs:
                out.append((p, upd, q2))
                continue
            changed = True
            groups = {}
            for b in BS.to_list(p):
                key = tuple(t[b] for t in tabs)
                groups[key] = groups.get(key, 0) | (1 << b)
            for key, bs in groups.items():
                val = dict(zip(tabs, key))
                out.append((bs, {v: normalize_update(tuple(("c", (val[a[1]],)) if a[0] == "t" else a for a in w))
                                 for v, w in upd.items()}, q2))
        edges[q] = out
    if not changed:
        return sst
    r = SST(sst.nstates, edges, sst.initial, sst.final, sst.nvars)
    for k in ("action",):
        if hasattr(sst, k):
            setattr(r, k, getattr(sst, k))
    return r


# ---- lookahead (`--la=true`): longest deterministic prefixes as multi-symbol tests
def _right_input_closure(fst, q):
    """rightInputClosure (SymbolicFST.hs:282-297): states without epsilon edges reachable over epsilon edges."""
    out, vis = set(), set()

    def go(q):
        es = fst.eps.get(q)
        if not es:
            out.add(q)
            return
        for _, q2 in es:
            if q2 not in vis:
                vis.add(q2)
                go(q2)

    go(q)
    return out


def _step_all(fst, p, ctx):
    """stepAll (SymbolicFST.hs:273-280)."""
    out = set()
    for q in ctx:
        for p2, _, q2 in fst.sym.get(q, ()):
            if BS.is_subset(p, p2):
                out |= _right_input_closure(fst, q2)
    return out


def _ldp(fst, ctx, q, limit=64):
    """ldp (SymbolicFST.hs:262-271): the longest deterministic prefix of state q in
    the context of the state set ctx."""
    out = []
    ctx = set(ctx)
    while len(out) < limit:
        es = fst.sym.get(q, ())
        if len(es) != 1:
            break
        p, _, q2 = es[0]
        if p not in coarsest_predicate_set(fst, sorted(ctx)):
            break
        out.append(p)
        ctx = _step_all(fst, p, ctx)
        q = q2
    return tuple(out)


def prefix_tests(fst, singleton_mode, states):
    """prefixTests (SymbolicFST.hs:299-312) -> [(predicate list, killed states)]."""
    ldps = [] if singleton_mode else [(_ldp(fst, states, q), q) for q in states]
    tests = {(p,) for p in coarsest_predicate_set(fst, states)} | {ps for ps, _ in ldps}

    def entails(t, ps):
        return len(ps) <= len(t) and all(a == b for a, b in zip(t, ps))

    key = lambda t: tuple(BS.to_ranges(p) for p in t)
    return [(t, {q for ps, q in ldps if not entails(t, ps)}) for t in sorted(tests, key=key)]


def sst_from_fst(fst, max_states=200000, lookahead=False):
    """sstFromFST (Determinization.hs:231-257).  `lookahead` = `--la=true`
    (singletonMode = not la, Commands.hs:126,171): transitions may then test
    several symbols (SST.la is set, edges carry predicate tuples)."""
    if lookahead:
        return _sst_from_fst_la(fst, max_states)
    if fst.has_actions():
        raise ValueError("Transducer contains action symbols - direct SST generation not supported")
    init = ("tip", (), fst.initial)
    index = {init: 0}
    order = [init]
    trans = []
    finals = {}
    i = 0
    while i < len(order):
        t = order[i]
        i += 1
        tcl = _closure(fst, _unabstract_outputs(t))
        fin = _eof(fst, tcl)
        if fin is not None and fin[0] == "tip":
            finals[t] = normalize_update(fin[1])
        if tcl is None:
            continue
        for p in coarsest_predicate_set(fst, _tflat(tcl)):
            tr = _closure(fst, _consume(fst, p, _closure(fst, tcl)))
            if tr is None:
                continue
            kappa, t2 = _abstract(tr)
            if BS.size(p) == 1:
                b = BS.to_list(p)[0]
                # `specialize` (Determinization.hs:190-206): a singleton predicate makes every function constant
                kappa = [(v, tuple(("c", (b,)) if a == ("f",) else ("c", (a[1][b],)) if a[0] == "t" else a for a in w))
                         for v, w in kappa]
            upd = {}
            for v, w in kappa:
                assert v not in upd, "Inconsistent register update"
                upd[v] = normalize_update(w)
            if t2 not in index:
                if len(order) >= max_states:
                    raise MemoryError("SST exceeds %d states" % max_states)
                index[t2] = len(order)
                order.append(t2)
            trans.append((t, p, upd, t2))
    # enumerate variables in tuple (pre-order) order
    vs = {()}
    for _, _, upd, _ in trans:
        vs.update(upd.keys())
        for w in upd.values():
            vs.update(a[1] for a in w if a[0] == "v")
    for w in finals.values():
        vs.update(a[1] for a in w if a[0] == "v")
    vid = {v: k for k, v in enumerate(sorted(vs))}

    def ren(w):
        return tuple(("v", vid[a[1]]) if a[0] == "v" else a for a in w)

    edges = {}
    for t, p, upd, t2 in trans:
        edges.setdefault(index[t], []).append(
            (p, {vid[v]: ren(w) for v, w in upd.items()}, index[t2]))
    final = {index[t]: ren(w) for t, w in finals.items()}
    return SST(len(order), edges, 0, final, len(vid))


def _sst_from_fst_la(fst, max_states):
    if fst.has_actions():
        raise ValueError("Transducer contains action symbols - direct SST generation not supported")
    init = ("tip", (), fst.initial)
    index = {init: 0}
    order = [init]
    trans = []
    finals = {}
    i = 0
    while i < len(order):
        t = order[i]
        i += 1
        tcl = _closure(fst, _unabstract_outputs(t))
        fin = _eof(fst, tcl)
        if fin is not None and fin[0] == "tip":
            finals[t] = normalize_update(fin[1])
        if tcl is None:
            continue
        for ps, kills in prefix_tests(fst, False, _tflat(tcl)):
            if not ps:
                continue
            tr = _kill(tcl, kills)
            for k, p in enumerate(ps):          # consumeTreeMany (Determinization.hs:213-229)
                tr = _closure(fst, _consume(fst, p, _closure(fst, tr), k))
            if tr is None:
                continue
            kappa, t2 = _abstract(tr)
            if all(BS.size(p) == 1 for p in ps):
                # `specialize` (Determinization.hs:190-206): a test of singletons makes every function constant
                bs = [BS.to_list(p)[0] for p in ps]

                def const(a):
                    if a[0] == "f":
                        return ("c", (bs[a[1] if len(a) > 1 else 0],))
                    if a[0] == "t":
                        return ("c", (a[1][bs[a[2] if len(a) > 2 else 0]],))
                    return a
                kappa = [(v, tuple(const(a) for a in w)) for v, w in kappa]
            upd = {}
            for v, w in kappa:
                assert v not in upd, "Inconsistent register update"
                upd[v] = normalize_update(w)
            if t2 not in index:
                if len(order) >= max_states:
                    raise MemoryError("SST exceeds %d states" % max_states)
                index[t2] = len(order)
                order.append(t2)
            trans.append((t, ps, upd, t2))
    vs = {()}
    for _, _, upd, _ in trans:
        vs.update(upd.keys())
        for w in upd.values():
            vs.update(a[1] for a in w if a[0] == "v")
    for w in finals.values():
        vs.update(a[1] for a in w if a[0] == "v")
    vid = {v: k for k, v in enumerate(sorted(vs))}

    def ren(w):
        return tuple(("v", vid[a[1]]) if a[0] == "v" else a for a in w)

    edges = {}
    for t, ps, upd, t2 in trans:
        edges.setdefault(index[t], []).append((ps, {vid[v]: ren(w) for v, w in upd.items()}, index[t2]))
    sst = SST(len(order), edges, 0, {index[t]: ren(w) for t, w in finals.items()}, len(vid))
    sst.la = True
    return sst


# ------------------------------------------------------------- optimisation
_AMB = "amb"


def _lift(rho, atoms):
    """liftAbstractValuation (SymbolicSST.hs:208-219): None if some variable
    has no abstract value yet."""
    acc = ()
    amb = False
    for a in atoms:
        if a[0] == "v":
            if a[1] not in rho:
                return None
            v = rho[a[1]]
            if v == _AMB:
                amb = True
            else:
                acc = acc + v
        elif a[0] in ("f", "t"):
            amb = True
        else:
            acc = acc + a[1]
    return _AMB if amb else acc


def _lub(a, b):
    if a == _AMB or b == _AMB or a != b:
        return _AMB
    return a


def optimize(sst, level=3, persistent=False):
    """Constant propagation of registers (`optimize`, SymbolicSST.hs:323-331;
    abstract interpretation :273-321).  level 0 = off; 1-2 = weak variant
    (a register that reads itself is ambiguous); 3 = full.

    `persistent`: for SSTs whose updates leave registers untouched by omission
    (action SSTs, ActionSST.hs: `interp` only names the registers an action
    touches).  The reference's rule deletes every update of a register that is
    statically known in the destination state; if that register later reaches,
    without an update, a state where it is no longer known, its buffer still
    holds the value from before the deleted update (found on
    bench/kleenex/src/drex_align-bibtex.kex: the reset of a popped builder is
    lost, scripts/gen_reference_action_vectors.py).  Path-tree SSTs name every
    live register in every update, so this never happens there; with
    `persistent` the known value is written back on such an edge."""
    if level <= 0:
        return sst
    weak = level < 3
    gamma = {q: {} for q in range(sst.nstates)}
    if persistent:
        # every buffer starts empty (`init`, C.hs:470-484) and a register may be read before any
        # action has assigned it (`!r` before `r@...`): the reference starts from "nothing known",
        # which lets a constant assigned on one path stand for the untouched register of another
        gamma[sst.initial] = {v: () for v in range(1, sst.nvars)}
    states = set(range(sst.nstates))
    while states:
        acc = {}
        for r in states:
            rho_r = gamma[r]
            for _, kappa, s in sst.edges.get(r, ()):
                new = dict(rho_r)
                for k, usf in kappa.items():
                    if weak:
                        tmp = dict(rho_r)
                        tmp[k] = _AMB
                        val = _lift(tmp, usf)
                    else:
                        val = _lift(rho_r, usf)
                    if val is not None:
                        new[k] = val
                if s in acc:
                    old = acc[s]
                    for k, v in new.items():
                        old[k] = _lub(old[k], v) if k in old else v
                else:
                    acc[s] = new
        nxt = set()
        for s, rho2 in acc.items():
            rho_s = gamma[s]
            if all(k in rho_s and (rho_s[k] == _AMB or rho_s[k] == v) for k, v in rho2.items()):
                continue
            merged = dict(rho_s)
            for k, v in rho2.items():
                merged[k] = _lub(merged[k], v) if k in merged else v
            gamma[s] = merged
            nxt.add(s)
        states = nxt

    def apply(rho, atoms):
        return normalize_update(tuple(
            ("c", rho[a[1]]) if a[0] == "v" and a[1] in rho and rho[a[1]] != _AMB else a
            for a in atoms))

    edges = {}
    for q, es in sst.edges.items():
        new = []
        for p, kappa, q2 in es:
            exact = {k for k, v in gamma[q2].items() if v != _AMB}
            upd = {k: apply(gamma[q], w) for k, w in kappa.items() if k not in exact}
            if persistent:
                for k, v in gamma[q].items():
                    if v != _AMB and k not in exact and k not in kappa:
                        upd[k] = (("c", v),) if v else ()
            new.append((p, upd, q2))
        edges[q] = new
    final = {q: apply(gamma[q], w) for q, w in sst.final.items()}
    out = SST(sst.nstates, edges, sst.initial, final, sst.nvars)
    for flag in ("la", "action"):
        if getattr(sst, flag, False):
            setattr(out, flag, True)
    return out


# ---------------------------------------------------------------- simulator
def run_sst(sst, data: bytes):
    """Sequential SST semantics with *persistent* registers, i.e. what the
    emitted C program computes (registers absent from an update keep their
    value).  Returns (accepted, output, consumed)."""
    if getattr(sst, "la", False):
        return _run_sst_la(sst, data)
    regs = {v: b"" for v in range(sst.nvars)}
    out = bytearray()
    q = sst.initial
    for i, b in enumerate(data):
        hit = None
        for p, upd, q2 in sst.edges.get(q, ()):
            if (p >> b) & 1:
                hit = (upd, q2)
                break
        if hit is None:
            return False, bytes(out), i
        upd, q = hit
        new = {}
        for v, atoms in upd.items():
            buf = bytearray()
            for a in atoms:
                if a[0] == "v":
                    buf += regs[a[1]]
                elif a[0] == "c":
                    buf += bytes(a[1])
                elif a[0] == "t":
                    buf.append(a[1][b])
                else:
                    buf.append(b)
            new[v] = bytes(buf)
        regs.update(new)
        out += regs[0]
        regs[0] = b""
    if q not in sst.final:
        return False, bytes(out), len(data)
    for a in sst.final[q]:
        out += regs[a[1]] if a[0] == "v" else bytes(a[1])
    return True, bytes(out), len(data)


def _run_sst_la(sst, data: bytes):
    """Lookahead SSTs as the emitted C runs them: the longest matching test wins
    (`findTrans`, SymbolicSST.hs:441-446 = the nesting of `kvtree`,
    SSTCompiler.hs:37-55), and the end-of-input branch is taken as soon as fewer
    than minL symbols remain (`readnext(minL, maxL)`, crt/crt.c:293-312;
    SSTCompiler.hs:138-146) -- a final state then accepts with unread input."""
    regs = {v: b"" for v in range(sst.nvars)}
    out = bytearray()
    q = sst.initial
    i, n = 0, len(data)
    while True:
        es = sst.edges.get(q, ())
        min_l = min((len(ps) for ps, _, _ in es), default=1)
        if n - i < min_l:
            break
        best = None
        for ps, upd, q2 in es:
            k = len(ps)
            if k <= n - i and all((p >> data[i + j]) & 1 for j, p in enumerate(ps)) and (best is None or k > best[0]):
                best = (k, upd, q2)
        if best is None:
            return False, bytes(out), i
        k, upd, q2 = best
        new = {}
        for v, atoms in upd.items():
            buf = bytearray()
            for a in atoms:
                if a[0] == "v":
                    buf += regs[a[1]]
                elif a[0] == "c":
                    buf += bytes(a[1])
                elif a[0] == "t":
                    buf.append(a[1][data[i + (a[2] if len(a) > 2 else 0)]])
                else:
                    buf.append(data[i + (a[1] if len(a) > 1 else 0)])
            new[v] = bytes(buf)
        regs.update(new)
        out += regs[0]
        regs[0] = b""
        q = q2
        i += k
    if q not in sst.final:
        return False, bytes(out), i
    for a in sst.final[q]:
        out += regs[a[1]] if a[0] == "v" else bytes(a[1])
    return True, bytes(out), i
