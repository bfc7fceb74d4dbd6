"""Model check of scan3.cu's stager -> scanner hand-over with mbarriers (round 2): ONE stager role and W scanner warps
that no longer wait for each other at FULL.  FULL(b) is an mbarrier per ring buffer (the stagers arrive, every scanner
warp waits on the buffer's phase parity, tracked in a bit mask exactly as the kernel does); EMPTY(b) stays a named
barrier that completes when ALL scanner warps have arrived and the stager syncs.  Random interleavings check: no
deadlock; a scanner warp reads buffer b at stage gs only when it holds stage gs (and the previous buffer stage gs - 1);
the stager never rewrites a buffer some warp still reads; scanner warps drift apart by at most two stages; a FULL
phase is never completed twice before every warp has consumed it (the parity bit would alias)."""
import random

import pytest


def stager(ntiles, nch, st):
    b = gs = 0
    for n in range(ntiles):
        for ch in range(nch + 1):
            if gs >= 2:
                yield ("sync_empty", b)
            assert st["readers"][b] == 0, f"buffer {b} rewritten while read (stage {gs})"
            # every warp must have consumed the previous phase of this buffer before it completes again
            assert all(c >= st["full_done"][b] for c in st["consumed"][b]), f"FULL({b}) phase would alias at stage {gs}"
            st["stage"][b] = gs
            st["full_done"][b] += 1                       # mbarrier phase completes (all stager warps arrived)
            yield ("arrived_full", b)
            b = 0 if b == 2 else b + 1
            gs += 1
    yield ("exit",)


def scanner(w, ntiles, nch, st):
    b = gs = 0
    ph = 0
    for n in range(ntiles):
        next_exists = n + 1 < ntiles
        for it in range(nch + 1):
            while True:                                   # mbarrier.try_wait.parity
                parity = (ph >> b) & 1
                # phase k (k = 0, 1, ...) of the barrier has parity k & 1; waiting on parity p succeeds once the
                # phase with that parity has completed, i.e. completed phases > consumed phases
                if st["full_done"][b] > st["consumed"][b][w]:
                    assert (st["consumed"][b][w] & 1) == parity, "parity mask out of step with the phases consumed"
                    break
                yield ("wait_full", b)
            st["consumed"][b][w] += 1
            ph ^= 1 << b
            bp = 2 if b == 0 else b - 1
            assert st["stage"][b] == gs, f"warp {w} stage {gs}: buffer {b} holds {st['stage'][b]}"
            assert st["stage"][bp] == gs - 1, f"warp {w} stage {gs}: previous buffer holds {st['stage'][bp]}"
            st["readers"][b] += 1; st["readers"][bp] += 1
            st["pos"][w] = gs
            run_pos = [p for p in st["pos"] if p < (1 << 30)]
            assert max(run_pos) - min(run_pos) <= 2, f"scanner warps drifted {st['pos']}"
            yield ("reading",)
            st["readers"][b] -= 1; st["readers"][bp] -= 1
            if it + 2 <= nch or next_exists:
                st["empty_arrivals"][bp] += 1
                yield ("arrived_empty", bp)
            b = 0 if b == 2 else b + 1
            gs += 1
    st["pos"][w] = 1 << 30                                # done: no longer bounds the drift
    yield ("exit",)


def run(ntiles, nch, W, seed):
    rng = random.Random(seed)
    st = {"stage": {0: None, 1: None, 2: -1}, "readers": {0: 0, 1: 0, 2: 0}, "full_done": {0: 0, 1: 0, 2: 0},
          "consumed": {b: [0] * W for b in range(3)}, "empty_arrivals": {0: 0, 1: 0, 2: 0}, "pos": [0] * W}
    roles = {"S": stager(ntiles, nch, st)}
    roles.update({f"C{w}": scanner(w, ntiles, nch, st) for w in range(W)})
    blocked_empty = None              # buffer the stager syncs on
    live = set(roles)
    steps = 0
    while live:
        steps += 1
        assert steps < 2_000_000, "no progress"
        runnable = [r for r in live if not (r == "S" and blocked_empty is not None)]
        if blocked_empty is not None and st["empty_arrivals"][blocked_empty] >= W:
            st["empty_arrivals"][blocked_empty] -= W
            blocked_empty = None
            continue
        if not runnable:
            raise AssertionError("deadlock: stager waits on EMPTY and no scanner can move")
        r = rng.choice(runnable)
        ev = next(roles[r])
        if ev[0] == "exit":
            live.discard(r)
        elif ev[0] == "sync_empty":
            blocked_empty = ev[1]
        elif ev[0] == "wait_full" and len(runnable) == 1 and blocked_empty is not None and \
                st["empty_arrivals"][blocked_empty] < W:
            raise AssertionError("deadlock: a scanner waits on FULL while the stager waits on EMPTY")
    assert all(v == 0 for v in st["readers"].values())


@pytest.mark.parametrize("ntiles,nch", [(1, 1), (1, 12), (2, 1), (3, 2), (5, 12), (9, 3)])
@pytest.mark.parametrize("W", [1, 2, 8])
def test_mbarrier_full_handover(ntiles, nch, W):
    for seed in range(12):
        run(ntiles, nch, W, seed)
