"""ORACLE / TEST INFRASTRUCTURE -- not part of the product path.

Builds `oracle/_ref/<program>`: the reference's own compiled C binary for a
workload, i.e. C text emitted in the reference back end's shape
(oracle/emit_c.py) + the verbatim `crt/crt.c` taken from /root/reference
where it lies, compiled with the reference's command line
(src/KMC/Program/Backends/C.hs:562-568).  Outputs go only to oracle/_ref/
(git-ignored, but shipped to the GPU box).  Requires /root/reference; on a
box without it the prebuilt binaries are used as they are.

With --act the programs are compiled in the reference's default mode
(`--act=true`: an oracle and an action program per pipeline stage, two
processes per stage, tables; frontend/oracle_action.py) as
`oracle/_ref/<program>.act` (`--la=false`) and `<program>.default` (`--la=true`, the
reference's default flag set); `<program>.la` is `--act=false --la=true`.

usage: python oracle/build_ref.py [--opt N] [--act] [prog.kex ...]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
CRT = "/root/reference/crt/crt.c"
REF_DIR = os.path.join(HERE, "_ref")


def build_one(kex_path, opt=3, out_dir=REF_DIR, name=None, keep_c=False, act=False, la=False):
    """Variants: <prog> (--act=false --la=false), <prog>.la (--act=false --la=true),
    <prog>.act (--act=true --sb=true --la=false), <prog>.default (--act=true --sb=true
    --la=true: the reference's default flags, Options.hs:146-167)."""
    from kleenexlang_b200.frontend.driver import build_ssts, build_oracle_action_pipeline, build_lookahead_ssts
    from kleenexlang_b200.frontend.il import compile_sst
    from oracle.emit_c import render_c
    src = open(kex_path, encoding="utf-8").read()
    if act:
        ssts = build_oracle_action_pipeline(src, opt, lookahead=la, suppress_bits=True)
    else:
        ssts = build_lookahead_ssts(src, opt) if la else build_ssts(src, opt)
    progs = [compile_sst(s) for s in ssts]
    ctext = render_c(progs, open(CRT).read(), info="%s --opt %d --la=%s --act=%s%s" % (
        os.path.basename(kex_path), opt, "true" if la else "false", "true" if act else "false",
        " --sb=true" if act else ""))
    os.makedirs(out_dir, exist_ok=True)
    suffix = {(False, False): "", (False, True): ".la", (True, False): ".act", (True, True): ".default"}[(act, la)]
    name = name or os.path.splitext(os.path.basename(kex_path))[0] + suffix
    out = os.path.join(out_dir, name)
    if keep_c:
        open(out + ".c", "w").write(ctext)
    r = subprocess.run(["cc", "-O3", "-xc", "-o", out, "-D FLAG_WORDALIGNED", "-w", "-"],
                       input=ctext.encode(), capture_output=True)
    if r.returncode != 0:
        raise RuntimeError("cc failed for %s:\n%s" % (kex_path, r.stderr.decode()[-2000:]))
    return out


def have_reference():
    return os.path.exists(CRT)


def main(argv):
    opt = 3
    paths = []
    act = None
    i = 0
    while i < len(argv):
        if argv[i] == "--opt":
            opt = int(argv[i + 1])
            i += 2
        elif argv[i] == "--act":
            act = True
            i += 1
        else:
            paths.append(argv[i])
            i += 1
    if not have_reference():
        print("reference runtime %s not present; keeping prebuilt oracle/_ref" % CRT)
        return 0
    if not paths:
        pdir = os.path.join(ROOT, "programs")
        paths = sorted(os.path.join(pdir, f) for f in os.listdir(pdir) if f.endswith(".kex"))
    for p in paths:
        if act is None or not act:
            print("built", build_one(p, opt))
        if act is None or act:
            # the default-mode binaries as well (programs with register actions only exist in this mode),
            # and the lookahead variants
            for a_, l_ in ((True, False), (False, True), (True, True)):
                try:
                    print("built", build_one(p, opt, act=a_, la=l_))
                except (ValueError, AssertionError, NotImplementedError, MemoryError) as e:
                    print("skipped %s --act=%s --la=%s: %s" % (os.path.basename(p), a_, l_, e))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
