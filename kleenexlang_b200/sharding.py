"""Host-side seam stitching for sharded evaluation (one shard per GPU).

The transducer run is a prefix computation over the SST's transition monoid
(src/KMC/SymbolicSST.hs:122-136): forward over the state maps, backward over
which registers still reach the output stream.  Ranks exchange two small
summaries with one all-gather each; everything else stays local: the rank
that creates a byte writes it, so no data crosses the seam and the output
stays sharded in rank order.  The register-side summaries are opaque bytes
that libkexcuda stitches itself (`kex_stitch_live`).
"""


def stitch_states(state_maps, init_state):
    """state_maps[r][q] = state after shard r when started in q.
    -> true start state of every shard."""
    starts = []
    s = init_state
    for m in state_maps:
        starts.append(s)
        s = m[s]
    return starts


def stitch_live(prog, seams, final_code):
    """seams[r] = seam summary of shard r (from `shard_walk`), final_code =
    seam code of the end-of-input action (0 when the run rejects).
    -> seam code at the end of every shard."""
    return prog.stitch_live(list(seams), final_code)


def stitch_codes(seams, final_code):
    """Pure-Python statement of `kex_stitch_live` for monoid-kernel programs:
    seams[r][code] = live-set index at the start of shard r given the index at
    its end."""
    codes = [0] * len(seams)
    code = final_code
    for r in range(len(seams) - 1, -1, -1):
        codes[r] = code
        code = seams[r][code]
    return codes
