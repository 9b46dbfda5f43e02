"""SST pipeline -> `kexprog` blob: the serialised transition and
register-action tables the CUDA back end loads (`kex_load`, include/kexcuda.h).

This takes the role `compileProgram` has in the reference
(src/KMC/Program/Backends/C.hs:529-540): it consumes the determinized,
optimised SSTs of a pipeline and produces the executable
artefact.  Where the C back end prints one goto-labelled block per state, the
blob holds dense tables:

  cls[256]            byte -> byte class (coarsest partition of all predicates,
                      src/KMC/Theories.hs:58-80)
  trans[(Q+1)*C]      (next state | action id << 16); state Q is the FAIL sink
  final[Q+1]          action applied at end of input, or -1 = reject
  acts[A]             per action: kind, piece range, bytes appended to out
  fate[A*R]           where the *old* content of register r goes: a register
                      index, 0 = flushed to the output stream, 0xFF = dropped
  addlen[A*R]         bytes newly appended to each register
  pieces[]            (target register, const|sym, length, const offset) in
                      output order
  consts[]            literal bytes

Register updates of path-tree SSTs have the shape  x := y1 .. yk new1 .. newm
(old registers first, in ancestor order, then new material;
src/KMC/Determinization.hs:165-183), so every byte that reaches the output
does so in the order it was created.  `check_chronological` verifies the
shape; the CUDA back end relies on it to let the chunk that *creates* a byte
write it.
"""
import struct

from .frontend import byteset as BS

MAGIC_PHASE = 0x5058454B      # "KEXP"
MAGIC_PIPE = 0x4C58454B       # "KEXL"
MAGIC_ACT = 0x4158454B        # "KEXA": action-interpreter phase (frontend/actions.py, csrc/kex_act.cuh)
VERSION = 1
DEAD = 0xFF
MAX_REGS = 32

KIND_NOP, KIND_OUTSYM, KIND_OUTONLY, KIND_GENERAL = 0, 1, 2, 3
PIECE_CONST, PIECE_SYM = 0, 1


DEVICE_ACT_SLOTS = 32       # csrc/kex_act.cuh ACT_NSLOT


def check_device_limits(t):
    """The limits kex_load enforces on a phase (csrc/kexcuda.cu load_phase): 16-bit transition indices and
    the shared-memory footprint of the generic kernels (transition table + action headers next to the chunk
    buffers, 200 KiB) -- checked here so that `kexc compile` refuses instead of writing a binary that cannot load."""
    q1c = (t.Q + 1) * t.C
    if q1c >= 65536:
        raise UnsupportedProgram("%d states x %d byte classes exceed the 16-bit transition index of the device" % (t.Q, t.C))
    mask = 0 if t.R == 1 else (1 if t.R <= 8 else 4)
    chunk, stage, nt = 4096, 20480, 128                 # KEX_CHUNK, KEX_STAGE, KEX_NT
    emit = chunk + stage + 32 + 2 * chunk + chunk * mask + (2 * nt * 32 if mask else 0) + q1c * 4 + t.A * 4 + 256
    walk = 256 + q1c * 4 + t.A * 4
    if max(emit, walk) > 200 * 1024:
        raise UnsupportedProgram("%d states x %d byte classes, %d actions: the tables exceed the device's shared memory"
                                 % (t.Q, t.C, t.A))


class UnsupportedProgram(Exception):
    pass


def check_chronological(sst):
    def ok(atoms):
        new = False
        for a in atoms:
            if a[0] == "v":
                if new:
                    return False
            else:
                new = True
        return True
    for es in sst.edges.values():
        for _, upd, _ in es:
            if not all(ok(w) for w in upd.values()):
                return False
    return all(ok(w) for w in sst.final.values())


def check_creation_order(sst, live):
    """The device never moves register contents: the step that creates a byte
    writes it, at the prefix sum of the surviving bytes -- sound iff the order
    of the bytes in the output is the order of their creation.  Per update
    that is `check_chronological`; across registers it is a property of
    path-tree SSTs (a parent's content is older than its children's,
    Determinization.hs:165-183) that constant propagation can break: a
    register whose content is statically known is materialised as a literal
    only where it is used, i.e. *later* than the contents it precedes
    (SymbolicSST.hs:273-331).  This is the static check: a must-relation
    older[q] over the live registers of state q (y, z) = "every byte of y was
    created before every byte of z, or in the same step with y's pieces first
    (pieces of a step are laid out by register number)", propagated to a
    fixpoint; every concatenation `.. y z ..` needs (y, z), and whatever is
    appended to the output stream must be older than everything still pending
    in a live register.  Register 0 is the stream."""
    NEW = -1

    def atoms_of(upd, x):
        return upd[x] if x in upd else (("v", x),)

    def older(rel, a, xa, b, xb):
        # a in the update of register xa, b in that of xb (xa != xb), both of the same step
        ra = a[1] if a[0] == "v" else NEW
        rb = b[1] if b[0] == "v" else NEW
        if ra != NEW and rb != NEW:
            return ra == 0 or (ra, rb) in rel
        if rb == NEW:
            return ra != NEW or xa < xb
        return False

    def transfer(rel, upd, regs2):
        out = set()
        lists = {x: atoms_of(upd, x) for x in regs2}
        for x in regs2:
            for y in regs2:
                if x != y and all(older(rel, a, x, b, y) for a in lists[x] for b in lists[y]):
                    out.add((x, y))
        return out

    regs = {q: sorted(live[q] | {0}) for q in range(sst.nstates)}
    rel = {sst.initial: {(x, y) for x in regs[sst.initial] for y in regs[sst.initial] if x != y}}
    work = [sst.initial]
    while work:
        q = work.pop()
        for _, upd, q2 in sst.edges.get(q, ()):
            r2 = transfer(rel[q], upd, regs[q2])
            if q2 not in rel:
                rel[q2] = r2
                work.append(q2)
            elif not rel[q2] <= r2:
                rel[q2] &= r2
                work.append(q2)
    for q, es in sst.edges.items():
        if q not in rel:
            continue
        for _, upd, q2 in es:
            for x in regs[q2]:
                w = atoms_of(upd, x)
                for a, b in zip(w, w[1:]):
                    if a[0] == "v" and b[0] == "v" and a[1] != 0 and (a[1], b[1]) not in rel[q]:
                        return False
            after = transfer(rel[q], upd, regs[q2])
            if any((0, x) not in after for x in regs[q2] if x != 0):
                return False
    for q, w in sst.final.items():
        if q in rel:
            for a, b in zip(w, w[1:]):
                if a[0] == "v" and b[0] == "v" and (a[1], b[1]) not in rel[q]:
                    return False
    return True


def liveness(sst):
    """Backward dataflow: live[q] = registers whose content at state q can
    still reach the output."""
    live = {q: set() for q in range(sst.nstates)}
    for q, w in sst.final.items():
        live[q].update(a[1] for a in w if a[0] == "v")
    changed = True
    while changed:
        changed = False
        for q, es in sst.edges.items():
            cur = live[q]
            n0 = len(cur)
            for _, upd, q2 in es:
                l2 = live[q2]
                for r2, w in upd.items():
                    if r2 == 0 or r2 in l2:
                        cur.update(a[1] for a in w if a[0] == "v")
                for r in l2:
                    if r not in upd:
                        cur.add(r)
            if len(cur) != n0:
                changed = True
    for s in live.values():
        s.discard(0)
    return live


class PhaseTables:
    """Plain-Python view of one phase's tables (also used by the tests'
    algorithm model)."""
    pass


def build_phase(sst) -> PhaseTables:
    if not check_chronological(sst):
        raise UnsupportedProgram("register update not in (old registers)(new material) form")
    live = liveness(sst)
    if not check_creation_order(sst, live):
        raise UnsupportedProgram("bytes would not reach the output in the order of their creation "
                                 "(a propagated constant is materialised after younger content)")
    Q = sst.nstates
    if Q >= 0xFFFF:
        raise UnsupportedProgram("too many states for 16-bit state ids")
    R = 1 + max([len(s) for s in live.values()] + [0])
    if R > MAX_REGS:
        raise UnsupportedProgram("%d simultaneously live registers exceed the device limit of %d" % (R, MAX_REGS))
    # Per-state register allocation: a register keeps the slot of the register
    # whose content it inherits whenever that slot is free, so that most
    # transitions stay identity moves.  Slot 0 is always the output stream.
    alloc = {sst.initial: {}}
    order = [sst.initial]
    i = 0
    while i < len(order):
        q = order[i]
        i += 1
        for _, upd, q2 in sst.edges.get(q, ()):
            if q2 in alloc:
                continue
            src = alloc[q]
            want = {}
            for v in sorted(live[q2]):
                if v not in upd:
                    want[v] = src.get(v)
                else:
                    w = upd[v]
                    want[v] = src.get(w[0][1]) if w and w[0][0] == "v" else None
            slots = {}
            taken = set()
            for v in sorted(live[q2]):
                if want[v] is not None and want[v] not in taken:
                    slots[v] = want[v]
                    taken.add(want[v])
            free = [k for k in range(1, R) if k not in taken]
            for v in sorted(live[q2]):
                if v not in slots:
                    slots[v] = free.pop(0)
            alloc[q2] = slots
            order.append(q2)
    for q in range(Q):
        alloc.setdefault(q, {v: k + 1 for k, v in enumerate(sorted(live[q]))})

    preds = sorted({p for es in sst.edges.values() for p, _, _ in es})
    parts = BS.coarsest_partition(preds)
    covered = 0
    for p in parts:
        covered |= p
    if BS.complement(covered):
        parts.append(BS.complement(covered))
    C = len(parts)
    cls = [0] * 256
    for ci, p in enumerate(parts):
        for b in BS.to_list(p):
            cls[b] = ci

    consts = bytearray()
    const_off = {}

    def intern_const(bs):
        bs = bytes(bs)
        if bs not in const_off:
            const_off[bs] = len(consts)
            consts.extend(bs)
        return const_off[bs]

    actions = []      # (kind, fate tuple, addlen tuple, pieces tuple)
    action_id = {}

    def make_action(upd, live_src, live_dst, asrc, adst):
        fate = [DEAD] * R
        fate[0] = 0
        pieces = []
        addlen = [0] * R
        seen = set()
        assigned = set()
        for v in sorted(upd):
            if v != 0 and v not in live_dst:
                continue
            assigned.add(v)
            tgt = 0 if v == 0 else adst[v]
            for a in upd[v]:
                if a[0] == "v":
                    if a[1] == 0:
                        if v != 0:
                            raise UnsupportedProgram("output register read into a register")
                        continue
                    if a[1] in seen:
                        raise UnsupportedProgram("register update is not copyless")
                    seen.add(a[1])
                    if a[1] in asrc:
                        fate[asrc[a[1]]] = tgt
                elif a[0] == "c":
                    pieces.append((tgt, PIECE_CONST, len(a[1]), intern_const(a[1])))
                    addlen[tgt] += len(a[1])
                else:
                    pieces.append((tgt, PIECE_SYM, 1, 0))
                    addlen[tgt] += 1
        if 0 in upd and (not upd[0] or upd[0][0] != ("v", 0)):
            raise UnsupportedProgram("output register is reset by an update")
        for v in live_dst:
            if v not in assigned:
                if v in seen:
                    raise UnsupportedProgram("register update is not copyless")
                fate[asrc[v]] = adst[v]
        ident = all(fate[asrc[v]] == asrc[v] for v in live_src)
        if not pieces and ident:
            kind = KIND_NOP
        elif ident and len(pieces) == 1 and pieces[0][:2] == (0, PIECE_SYM):
            kind = KIND_OUTSYM
        elif ident and all(p[0] == 0 for p in pieces):
            kind = KIND_OUTONLY
        else:
            kind = KIND_GENERAL
        if kind != KIND_GENERAL:
            # content of registers that are not live is irrelevant; make the
            # fate row a clean identity so consumers can skip it
            fate = list(range(R))
        key = (kind, tuple(fate), tuple(addlen), tuple(pieces))
        if key not in action_id:
            action_id[key] = len(actions)
            actions.append(key)
        return action_id[key]

    nop = make_action({}, set(), set(), {}, {})
    assert nop == 0
    FAIL = Q
    trans = [(FAIL | (nop << 16))] * ((Q + 1) * C)
    for q, es in sst.edges.items():
        for p, upd, q2 in es:
            a = make_action(upd, live[q], live[q2], alloc[q], alloc[q2])
            for ci, part in enumerate(parts):
                if part & p:
                    assert BS.is_subset(part, p)
                    assert trans[q * C + ci] == (FAIL | (nop << 16)), "non-deterministic SST"
                    trans[q * C + ci] = q2 | (a << 16)
    final = [-1] * (Q + 1)
    for q, w in sst.final.items():
        final[q] = make_action({0: (("v", 0),) + tuple(w)}, live[q], set(), alloc[q], {})
    if len(actions) >= 0xFFFF:
        raise UnsupportedProgram("too many distinct actions")

    t = PhaseTables()
    t.Q, t.C, t.R, t.A = Q, C, R, len(actions)
    t.init = sst.initial
    t.cls = cls
    t.trans = trans
    t.final = final
    t.kind = [a[0] for a in actions]
    t.fate = [list(a[1]) for a in actions]
    t.addlen = [list(a[2]) for a in actions]
    t.pieces = [list(a[3]) for a in actions]
    t.consts = bytes(consts)
    t.max_out_per_byte = max([sum(a[2]) for a in actions] + [1])
    check_device_limits(t)
    return t


def serialize_phase(t: PhaseTables, with_fast: bool = True) -> bytes:
    piece_off = []
    flat = []
    for ps in t.pieces:
        piece_off.append(len(flat))
        flat.extend(ps)
    sections = []
    sections.append(bytes(t.cls))
    sections.append(struct.pack("<%dI" % len(t.trans), *t.trans))
    sections.append(struct.pack("<%di" % len(t.final), *t.final))
    acts = bytearray()
    for a in range(t.A):
        flush = 0
        for r in range(1, t.R):
            if t.fate[a][r] == 0:
                flush |= 1 << r
        acts += struct.pack("<IIIIII", t.kind[a], len(t.pieces[a]), piece_off[a],
                            t.addlen[a][0], flush, sum(t.addlen[a]))
    sections.append(bytes(acts))
    sections.append(bytes(b for row in t.fate for b in row))
    sections.append(struct.pack("<%dI" % (t.A * t.R), *[x for row in t.addlen for x in row]))
    sections.append(b"".join(struct.pack("<BBHI", p[0], p[1], p[2], p[3]) for p in flat))
    sections.append(t.consts)
    nhdr = 24
    off = nhdr * 4
    offs = []
    body = bytearray()
    for s in sections:
        pad = (-off) % 16
        body += b"\0" * pad
        off += pad
        offs.append(off)
        body += s
        off += len(s)
    pad = (-off) % 16
    body += b"\0" * pad
    off += pad
    # optional fast section (monoid tables, fasttab.py); programs that exceed
    # its limits run on the generic kernels
    fast_off, fast_len = 0, 0
    if with_fast:
        from . import fasttab
        try:
            fb = fasttab.serialize_fast(t, fasttab.build_fast(t))
            fast_off, fast_len = off, len(fb)
            body += fb
            off += len(fb)
        except fasttab.Ineligible:
            pass
    hdr = [MAGIC_PHASE, VERSION, t.Q, t.C, t.R, t.A, len(flat), len(t.consts), t.init,
           t.max_out_per_byte] + offs + [off, fast_off, fast_len]
    hdr += [0] * (nhdr - len(hdr))
    return struct.pack("<%dI" % nhdr, *hdr) + bytes(body)


def serialize_act(stage) -> bytes:
    """Action-interpreter phase: nothing but the register count and the escape byte."""
    from .frontend.actions import ESC
    return struct.pack("<8I", MAGIC_ACT, VERSION, stage.nregs, ESC, 32, 0, 0, 0)


def serialize_pipeline(phases, with_fast: bool = True) -> bytes:
    blobs = [serialize_act(t) if hasattr(t, "nregs") else serialize_phase(t, with_fast) for t in phases]
    n = len(blobs)
    hdr_words = 4 + 2 * n
    hdr_size = (hdr_words * 4 + 15) // 16 * 16
    off = hdr_size
    table = []
    for b in blobs:
        table += [off, len(b)]
        off += len(b)
    hdr = struct.pack("<%dI" % hdr_words, MAGIC_PIPE, VERSION, n, off, *table)
    hdr += b"\0" * (hdr_size - len(hdr))
    return hdr + b"".join(blobs)


def compile_ssts(ssts, with_fast: bool = True) -> bytes:
    """Pipeline of ready-made SSTs -> kexprog blob; table atoms (`AppendTblI`)
    are lowered to constants per byte first (frontend/sst.py expand_tables)."""
    from .frontend.sst import expand_tables
    return serialize_pipeline([build_phase(expand_tables(s)) for s in ssts], with_fast)


def compile_re(re_src: str, opt: int = 3, suppress_bits: bool = False, with_fast: bool = True) -> bytes:
    """The regular-expression flavour on the device (`kexc compile x.re`,
    compileCoder, Commands.hs:246-275): the oracle SST of the regex alone, one
    phase that writes the bit-coded parse of its input."""
    from .frontend.driver import build_coder_ssts
    return compile_ssts(build_coder_ssts(re_src, opt, lookahead=False, suppress_bits=suppress_bits), with_fast)


def reference_phases(src: str, opt: int = 3, suppress_bits: bool = True):
    """`kexc compile --act=true --la=false` with the reference's phase structure
    (compileOracleAction, Commands.hs:204-244; C.hs:507-510): every pipeline
    stage becomes an oracle phase (input -> code bytes, tables lowered to
    constants) and an action phase (code -> output; the action machine as an
    SST, ActionSST.hs:47-129), so that `-p N` selects the same phases as in the
    reference's default build and the stream between them is the reference's
    code.  A stage whose action SST is outside the device's update shape
    (registers that are prepended to, reordered ...) keeps the split of
    compile_kex: transducer phase + action-interpreter phase.
    -> [PhaseTables | ActStage]"""
    from .frontend.driver import build_transducers
    from .frontend.oracle_action import build_oracle_action_ssts
    from .frontend.sst import expand_tables, optimize, sst_from_fst
    from .frontend.actions import action_stream_fst
    phases = []

    def pair_at(t, levels):
        # as in compile_kex: where full constant propagation breaks the chronological shape of an update,
        # the weaker levels (`--opt 1`, then none) are tried; I/O behaviour does not depend on --opt
        built, out = {}, []
        for k in (0, 1):
            for o in levels:
                if o not in built:
                    built[o] = build_oracle_action_ssts(t, o, lookahead=False, suppress_bits=suppress_bits)
                try:
                    out.append(build_phase(expand_tables(built[o][k])))
                    break
                except UnsupportedProgram as e:
                    err = e
            else:
                raise err
        return out

    def direct_at(fst):
        raw = sst_from_fst(fst)
        for o in [opt] + [o for o in (1, 0) if o < opt]:
            try:
                return build_phase(optimize(raw, o))
            except UnsupportedProgram as e:
                err = e
        raise err

    for t in build_transducers(src):
        try:
            pair = pair_at(t, [opt] + [o for o in (1, 0) if o < opt])
        except UnsupportedProgram:
            if t.has_actions():
                t2, act = action_stream_fst(t)
                if act.nregs + 1 > DEVICE_ACT_SLOTS:
                    raise UnsupportedProgram("%d action registers exceed the %d slots of the device action interpreter"
                                             % (act.nregs, DEVICE_ACT_SLOTS))
                pair = [direct_at(t2), act]
            else:
                pair = [direct_at(t)]
        phases += pair
    return phases


def compile_reference_phases(src: str, opt: int = 3, suppress_bits: bool = True, with_fast: bool = True) -> bytes:
    """-> kexprog blob of reference_phases (`kexc compile --phases=reference`)."""
    return serialize_pipeline(reference_phases(src, opt, suppress_bits), with_fast)


def compile_kex(src: str, opt: int = 3, with_fast: bool = True, actions: bool = True) -> bytes:
    """`.kex` source -> kexprog blob (the CUDA counterpart of
    `kexc compile --act=false --la=false`).  A stage with register actions
    (`reg@t`, `!reg`), which the reference only compiles in its oracle/action
    mode, becomes two phases here as well: a transducer phase that writes the
    action stream and an action-interpreter phase (frontend/actions.py);
    `actions=False` refuses such programs like `--act=false` does."""
    return serialize_pipeline(kex_phases(src, opt, actions), with_fast)


def kex_phases(src: str, opt: int = 3, actions: bool = True):
    """The phases compile_kex serialises: [PhaseTables | ActStage]."""
    from .frontend.driver import build_ssts
    phases = []
    weaker = {}                      # SSTs at weaker optimisation levels, built on demand
    for s in build_ssts(src, opt, actions=actions):
        if hasattr(s, "nregs"):
            # the device interpreter shares ACT_NSLOT = 32 slots between the builders (one per stack
            # height, at least the bottom one) and the registers (csrc/kex_act.cuh; kex_load rejects
            # nregs + 1 > 32): refuse here, not with a binary that always fails at load
            if s.nregs + 1 > DEVICE_ACT_SLOTS:
                raise UnsupportedProgram("%d action registers exceed the %d slots of the device action interpreter"
                                         % (s.nregs, DEVICE_ACT_SLOTS))
            phases.append(s)
            continue
        try:
            phases.append(build_phase(s))
        except UnsupportedProgram as first:
            # full constant propagation may move a literal in front of an older register (the
            # update is then no longer chronological), the unoptimised SST may keep more registers
            # live than the device holds: try the weak propagation (`--opt 1`, SymbolicSST.hs:323-331),
            # then none.  I/O behaviour does not depend on --opt.
            for o in (1, 0):
                if o >= opt:
                    continue
                if o not in weaker:
                    weaker[o] = build_ssts(src, o, actions=actions)
                try:
                    phases.append(build_phase(weaker[o][len(phases)]))
                    break
                except UnsupportedProgram:
                    pass
            else:
                raise first
    return phases
