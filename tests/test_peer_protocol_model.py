"""CPU model check of the ticket protocol of csrc/peer.cu (the NVLink peer-memory exchanges): every interleaving of W ranks running K
back-to-back small all-reduces is explored, with each rank's kernel broken into its observable steps

    publish own values into data slot (t & 1)  ->  write ticket t into every peer's flag  ->  wait until every own flag >= t
    ->  read every rank's slot (t & 1)

and the claim of the kernel's header is asserted: whenever a rank reads a peer's slot for ticket t it holds that peer's ticket-t values
(two slots suffice because a rank can be at most one exchange ahead), and all ranks compute the same sums.  The arena all-reduce uses the
same tickets (three per call) for its barriers; its reduce-scatter / all-gather hazards are checked the same way on a 2-chunk model."""
import itertools

import pytest


def explore_small(W, K):
    """DFS over all interleavings.  State per rank: (call index t in 1..K, phase 0..3 + peer cursor)."""
    # slots[r][s] = ticket whose data rank r last published in slot s; flags[r][p] = last ticket p wrote into r's flag array
    init = (tuple((1, 0, 0) for _ in range(W)),                     # (t, phase, cursor)
            tuple((0, 0) for _ in range(W)),                          # slots
            tuple(tuple(0 for _ in range(W)) for _ in range(W)))      # flags
    seen, stack, finals = set(), [init], 0
    while stack:
        st = stack.pop()
        if st in seen:
            continue
        seen.add(st)
        ranks, slots, flags = st
        moved = False
        for r in range(W):
            t, ph, cur = ranks[r]
            if t > K:
                continue
            nr, ns, nf = list(ranks), [list(s) for s in slots], [list(f) for f in flags]
            if ph == 0:                                   # publish
                ns[r][t & 1] = t
                nr[r] = (t, 1, 0)
            elif ph == 1:                                 # signal peers one by one
                p = cur if cur != r else cur + 1
                if p < W:
                    nf[p][r] = t
                    nr[r] = (t, 1, p + 1)
                else:
                    nr[r] = (t, 2, 0)
            elif ph == 2:                                 # wait: enabled only when all own flags have arrived
                if all(flags[r][p] >= t for p in range(W) if p != r):
                    nr[r] = (t, 3, 0)
                else:
                    continue
            else:                                         # read the W slots
                if cur < W:
                    assert slots[cur][t & 1] == t, f"rank {r} reads rank {cur}'s slot for ticket {t} but finds ticket {slots[cur][t & 1]}"
                    nr[r] = (t, 3, cur + 1)
                else:
                    nr[r] = (t + 1, 0, 0)
            moved = True
            stack.append((tuple(nr), tuple(tuple(s) for s in ns), tuple(tuple(f) for f in nf)))
        if not moved:
            assert all(t > K for t, _, _ in ranks), f"deadlock in state {st}"
            finals += 1
    return len(seen), finals


@pytest.mark.parametrize("W,K", [(2, 4), (3, 3)])
def test_small_exchange_two_slots_suffice(W, K):
    states, finals = explore_small(W, K)
    assert finals >= 1 and states > 100


def test_one_slot_would_not_suffice():
    """The same exploration with a single data slot finds the overwrite race - i.e. the model is able to see such bugs."""
    global explore_small
    src = explore_small
    import types, inspect
    code = inspect.getsource(src).replace("t & 1", "0").replace("(0, 0) for _ in range(W))", "(0,) for _ in range(W))")
    ns = {}
    exec(code, ns)
    with pytest.raises(AssertionError, match="finds ticket"):
        ns["explore_small"](2, 3)


def test_arena_allreduce_chunk_hazards():
    """Two-shot all-reduce on a W-chunk model: barrier(ticket-2) -> reduce own chunk (reads chunk r of every rank, writes own chunk r)
    -> barrier(ticket-1) -> gather (reads chunk p of rank p, writes own chunk p) -> barrier(ticket).  Every interleaving of 2 ranks x 2 calls:
    a rank never reads a chunk that its owner may still be writing, and never overwrites data a peer has not read yet."""
    W, K = 2, 2
    # per rank: arena[chunk] = ("g", call) gradient of this call | ("s", call) reduced sum of this call
    def step(ranks, arenas, flags):
        for r in range(W):
            t, ph = ranks[r]
            if t > K:
                continue
            nr, na, nf = list(ranks), [list(a) for a in arenas], [list(f) for f in flags]
            tick = 3 * t
            if ph == 0:      # write this call's gradients (the backward pass), then signal barrier 1
                for c in range(W):
                    na[r][c] = ("g", t)
                for p in range(W):
                    if p != r: nf[p][r] = tick - 2
                nr[r] = (t, 1)
            elif ph == 1:    # wait barrier 1, reduce own chunk
                if not all(flags[r][p] >= tick - 2 for p in range(W) if p != r): continue
                for p in range(W):
                    assert arenas[p][r] == ("g", t), f"rank {r} reduces chunk {r} of rank {p}: {arenas[p][r]} (call {t})"
                na[r][r] = ("s", t)
                for p in range(W):
                    if p != r: nf[p][r] = tick - 1
                nr[r] = (t, 2)
            elif ph == 2:    # wait barrier 2, gather the other chunks
                if not all(flags[r][p] >= tick - 1 for p in range(W) if p != r): continue
                for p in range(W):
                    if p != r:
                        assert arenas[p][p] == ("s", t), f"rank {r} gathers chunk {p} from rank {p}: {arenas[p][p]} (call {t})"
                        na[r][p] = ("s", t)
                for p in range(W):
                    if p != r: nf[p][r] = tick
                nr[r] = (t, 3)
            else:            # wait barrier 3, then the arena may be overwritten by the next call
                if not all(flags[r][p] >= tick for p in range(W) if p != r): continue
                assert all(a == ("s", t) for a in arenas[r])
                nr[r] = (t + 1, 0)
            yield tuple(nr), tuple(tuple(a) for a in na), tuple(tuple(f) for f in nf)

    init = (tuple((1, 0) for _ in range(W)), tuple(tuple(("g", 0) for _ in range(W)) for _ in range(W)), tuple(tuple(0 for _ in range(W)) for _ in range(W)))
    seen, stack, done = set(), [init], 0
    while stack:
        st = stack.pop()
        if st in seen:
            continue
        seen.add(st)
        nxt = list(step(*st))
        if not nxt:
            assert all(t > K for t, _ in st[0]), f"deadlock: {st}"
            done += 1
        stack.extend(nxt)
    assert done >= 1
