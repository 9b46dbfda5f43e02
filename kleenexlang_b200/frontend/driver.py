"""Front-end driver: .kex source -> pipeline of SSTs.

Mirrors `createProgram` / `buildTransducers` / `generateDirectSSTs`
(src/KMC/Frontend/Commands.hs:50-115,159-180) for the `--act=false --la=false`
configuration: parse, desugar, one transducer per pipeline stage, path-tree
determinization, optional constant-propagation optimisation.
"""
from .kleenex import parse_kleenex, desugar, check_well_formedness
from .fst import construct_transducer, run_lockstep, run_actions
from .sst import sst_from_fst, optimize, run_sst
from .actions import ActStage, action_stream_fst, run_act_stream


def build_transducers(src: str):
    pipeline, decls = parse_kleenex(src)
    check_well_formedness(pipeline, decls)
    pl, rdecls = desugar(pipeline, decls)
    return [construct_transducer(rdecls, ident) for ident in pl]


def build_ssts(src: str, opt: int = 3, actions: bool = False):
    """-> [SST] (one per pipeline stage), as `kexc compile --act=false
    --la=false --opt <opt>` would determinize them.  Like the reference
    (Commands.hs:166-168) this refuses transducers with register actions unless
    `actions` is set: such a stage then becomes two phases, an SST that writes
    the action stream and an `ActStage` that interprets it (frontend/actions.py)."""
    out = []
    for t in build_transducers(src):
        if actions and t.has_actions():
            t2, act = action_stream_fst(t)
            out += [optimize(sst_from_fst(t2), opt), act]
        else:
            out.append(optimize(sst_from_fst(t), opt))
    return out


def build_oracle_action_pipeline(src: str, opt: int = 3, lookahead: bool = False, suppress_bits: bool = False):
    """-> [oracle SST, action SST, oracle SST, action SST, ...]: the phases of
    the reference's default `kexc compile` (`--act=true`; two per pipeline
    stage, C.hs:507-510); `lookahead` = `--la`, `suppress_bits` = `--sb` (both on
    by default in the reference, Options.hs:146-167)."""
    from .oracle_action import build_oracle_action_ssts
    out = []
    for t in build_transducers(src):
        out.extend(build_oracle_action_ssts(t, opt, lookahead, suppress_bits))
    return out


def build_regex_transducer(re_src: str):
    """The regular-expression flavour (`.re` files / `--re`, Commands.hs:68-79):
    `desugarRegex` (Desugaring.hs:231-239) = the regex as a one-stage program
    whose reads are copied."""
    from .regex import parse_regex
    pl, rdecls = desugar(["main"], [("main", ("re", parse_regex(re_src.rstrip("\n"))))])
    return construct_transducer(rdecls, pl[0])


def build_coder_ssts(re_src: str, opt: int = 3, lookahead: bool = False, suppress_bits: bool = False):
    """`compileCoder` (Commands.hs:246-275): for a regular expression only the
    oracle program is compiled -- it writes the bit-coded parse of the input
    (one code byte per choice and per byte of a non-singleton class)."""
    from .oracle_action import oracle_fst, last_post_dominator
    t = build_regex_transducer(re_src)
    lpdom = last_post_dominator(t) if suppress_bits else None
    return [optimize(sst_from_fst(oracle_fst(t, lpdom), lookahead=lookahead), opt)]


def decode_parse(re_src: str, code: bytes, suppress_bits: bool = False):
    """Inverse of the coder: the action machine of the same transducer turns the
    code back into the matched string (ActionMachine.hs:109-127)."""
    from .oracle_action import action_fst, action_to_sst, last_post_dominator
    t = build_regex_transducer(re_src)
    lpdom = last_post_dominator(t) if suppress_bits else None
    ok, out, _ = run_sst(action_to_sst(action_fst(t, lpdom)), code)
    return out if ok else None


def build_lookahead_ssts(src: str, opt: int = 3):
    """`kexc compile --act=false --la=true`: direct SSTs whose transitions may
    test several symbols (longest deterministic prefixes, SymbolicFST.hs:262-312)."""
    return [optimize(sst_from_fst(t, lookahead=True), opt) for t in build_transducers(src)]


def simulate_lockstep(src: str, data: bytes):
    """`kexc simulate --sim=lockstep` (Commands.hs:277-289): returns output
    bytes or None on reject."""
    for t in build_transducers(src):
        syms = run_lockstep(t, data)
        if syms is None:
            return None
        data = run_actions(syms)
    return data


def simulate_sst(ssts, data: bytes):
    """`kexc simulate --sim=sst` over a pipeline; None on reject."""
    for s in ssts:
        if isinstance(s, ActStage):
            data = run_act_stream(data)
            continue
        ok, data, _ = run_sst(s, data)
        if not ok:
            return None
    return data
