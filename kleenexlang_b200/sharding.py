"""Host-side seam stitching for sharded evaluation (one shard per GPU).

The transducer run is a prefix computation over (state map, register fate
map) -- the monoid of src/KMC/SymbolicSST.hs:122-136 restricted to what the
seams need.  Ranks exchange two small summaries with one all-gather each
(`gather_summaries` does that over torch.distributed / NCCL); everything else
stays local: the rank that creates a byte writes it, so no data crosses the
seam and the output stays sharded in rank order.
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


def pullback(fate, live_mask):
    """Registers whose content goes, through `fate`, to the stream or to a
    register that is live afterwards."""
    m = live_mask | 1
    out = 0
    for r in range(1, len(fate)):
        d = fate[r]
        if d != 0xFF and (m >> d) & 1:
            out |= 1 << r
    return out


def stitch_live(fate_maps, final_mask):
    """fate_maps[r] = register fate map over shard r (from its true start
    state); final_mask = registers flushed by the end-of-input action.
    -> live mask at the end of every shard."""
    lives = [0] * len(fate_maps)
    live = final_mask
    for r in range(len(fate_maps) - 1, -1, -1):
        lives[r] = live
        live = pullback(fate_maps[r], live)
    return lives
