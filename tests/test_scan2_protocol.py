"""Model check of the streaming scan kernel's hand-over protocol (lancedb_b200/csrc/scan2.cu): the builder and
scanner role loops are transcribed as event generators (named-barrier arrive / sync, ring-buffer writes and
reads, tile-descriptor slot writes and reads) and run under random interleavings.  Checked: no deadlock, every
barrier phase is met by exactly one arrive-side and one sync-side event, a ring buffer is never rewritten
while a scanner stage still reads it, scanners only read buffers holding the stages they expect, and a
descriptor slot is never rewritten while a role still needs the tile it holds.  This does not run the kernel;
the GPU parity tests do.  It pins the protocol's index arithmetic (stage counter across tiles, `next_exists`,
the 4-slot descriptor ring) so a change to it fails here, on the CPU."""
import random

import pytest

FULL, EMPTY, PROD = "FULL", "EMPTY", "PROD"


def producer(ntiles, nch, st):
    """builder role: yields ('sync'|'arrive', barrier, buffer) or ('wait',) when blocked"""
    b, gs = 0, 0
    st["slots"][0] = 0 if ntiles > 0 else None          # prologue: tiles 0 and 1
    st["slots"][1] = 1 if ntiles > 1 else None
    yield ("prologue_done",)
    n = 0
    while True:
        tile = st["slots"][n & 3]
        if tile is None:
            break
        assert tile == n, f"builder read slot {n & 3} holding tile {tile}, wanted {n}"
        for ch in range(nch + 1):
            if gs >= 2:
                yield ("sync", EMPTY, b)
            # build stage gs into buffer b
            assert st["buf_readers"][b] == 0, f"buffer {b} rewritten while scanners read it (stage {gs})"
            st["buf_stage"][b] = gs
            if ch == 1:                                   # publish tile n+2 before FULL(stage 1)
                slot = (n + 2) & 3
                assert st["slot_users"][slot] == 0, f"descriptor slot {slot} rewritten while in use"
                st["slots"][slot] = n + 2 if n + 2 < ntiles else None
            yield ("arrive", FULL, b)
            yield ("sync", PROD, 0)
            b = 0 if b == 2 else b + 1
            gs += 1
        n += 1
    yield ("exit",)


def consumer(ntiles, nch, st):
    b, gs = 0, 0
    yield ("prologue_wait",)
    n = 0
    while True:
        tile = st["slots"][n & 3]
        if tile is None:
            break
        assert tile == n, f"scanner read slot {n & 3} holding tile {tile}, wanted {n}"
        st["slot_users"][n & 3] += 1
        next_exists = st["slots"][(n + 1) & 3] is not None
        assert next_exists == (n + 1 < ntiles)
        for it in range(nch + 1):
            yield ("sync", FULL, b)
            bp = 2 if b == 0 else b - 1
            assert st["buf_stage"][b] == gs, f"scanner stage {gs}: buffer {b} holds stage {st['buf_stage'][b]}"
            assert st["buf_stage"][bp] == gs - 1, f"scanner stage {gs}: prev buffer holds {st['buf_stage'][bp]}"
            st["buf_readers"][b] += 1; st["buf_readers"][bp] += 1
            yield ("reading",)
            st["buf_readers"][b] -= 1; st["buf_readers"][bp] -= 1
            if it + 2 <= nch or next_exists:
                yield ("arrive", EMPTY, bp)
            b = 0 if b == 2 else b + 1
            gs += 1
        st["slot_users"][n & 3] -= 1                      # epilogue done (reads T->out from the slot)
        n += 1
    yield ("exit",)


def run(ntiles, nch, seed):
    rng = random.Random(seed)
    st = {"slots": [None] * 4, "slot_users": [0] * 4, "buf_stage": {0: None, 1: None, 2: -1}, "buf_readers": {0: 0, 1: 0, 2: 0}}
    roles = {"P": producer(ntiles, nch, st), "C": consumer(ntiles, nch, st)}
    pending = {}                      # (barrier, id) -> set of roles that arrived / are waiting
    blocked = {}                      # role -> (barrier, id) it syncs on
    done, prologue = set(), False
    steps = 0
    while len(done) < 2:
        steps += 1
        assert steps < 200000, "no progress"
        runnable = [r for r in roles if r not in done and r not in blocked]
        if not prologue and "C" in runnable and "C_started" not in st:
            runnable = [r for r in runnable if r != "C"] or []
        assert runnable, f"deadlock: blocked={blocked} pending={pending}"
        r = rng.choice(runnable)
        ev = next(roles[r])
        if ev[0] == "prologue_done":
            prologue = True
        elif ev[0] == "prologue_wait":
            st["C_started"] = True
            if not prologue:                              # __syncthreads: scanners start after the prologue
                blocked["C"] = ("PROLOGUE", 0)
        elif ev[0] == "exit":
            done.add(r)
        elif ev[0] in ("sync", "arrive"):
            key = (ev[1], ev[2])
            if ev[1] == PROD:                             # builder-only barrier: completes at once in this model
                continue
            got = pending.setdefault(key, {})
            assert r not in got, f"{r} hit {key} twice in one phase"
            got[r] = ev[0]
            if len(got) == 2:                             # one arrive side + one sync side
                assert sorted(got.values()) == ["arrive", "sync"], f"{key}: {got}"
                del pending[key]
                for rr in list(blocked):
                    if blocked[rr] == key:
                        del blocked[rr]
            elif ev[0] == "sync":
                blocked[r] = key
        if prologue and blocked.get("C") == ("PROLOGUE", 0):
            del blocked["C"]
    assert not pending, f"unbalanced barriers at exit: {pending}"
    assert st["buf_readers"] == {0: 0, 1: 0, 2: 0}


@pytest.mark.parametrize("nch", [1, 2, 3, 12])
@pytest.mark.parametrize("ntiles", [0, 1, 2, 3, 4, 5, 9])
def test_scan2_handover_protocol(ntiles, nch):
    for seed in range(25):
        run(ntiles, nch, seed)


# ---------------------------------------------------------------- codebook staging ring (cp.async groups)
def _staging_run(ntask, depth, nch, tiles, seed):
    """One builder warp's cp.async ring as scan2.cu::build_chunk / stage_start drive it.  Copies complete in
    commit order but at arbitrary times unless a wait_group forces them; a slot read must find its task
    landed, and a slot may only be refilled after its previous task was read back into registers."""
    rng = random.Random(seed)
    nslots = depth
    groups = []                                # committed groups, oldest first: [task or None, slot, landed]
    slot_task = [None] * nslots                # task whose bytes are (being) written to the slot
    slot_read = [True] * nslots                # has that task been read back?

    def commit(task, slot):
        if task is not None:
            assert slot_read[slot], f"slot {slot} refilled before task {slot_task[slot]} was read"
            slot_task[slot], slot_read[slot] = task, False
        groups.append([task, slot, False])

    def wait_group(n):                         # all but the newest n groups are complete
        for g in groups[:max(0, len(groups) - n)]:
            g[2] = True
        for g in groups:                       # in-order completion at random times
            if not g[2] and rng.random() < 0.3:
                g[2] = True
            elif not g[2]:
                break

    def read(task, slot):
        assert slot_task[slot] == task, f"slot {slot} holds task {slot_task[slot]}, wanted {task}"
        landed = [g for g in groups if g[0] == task]
        assert landed and landed[-1][2], f"task {task} read before its copy landed"
        slot_read[slot] = True

    for t in range(tiles):
        base = (t, 0, 0)
        # stage_start: tasks 0..D-1 of chunk 0, ring phase reset
        for k in range(depth):
            commit((t, 0, k), k)
        so = 0
        for ch in range(nch):
            has_next = ch + 1 < nch
            if ch == 0:
                wait_group(depth - 1)
                read((t, 0, 0), so)
            for k in range(ntask):
                wait_group(depth - 2)
                if k + depth < ntask:
                    commit((t, ch, k + depth), so)
                elif has_next:
                    commit((t, ch + 1, k + depth - ntask), so)
                else:
                    commit(None, so)
                so_next = 0 if so + 1 == nslots else so + 1
                if k + 1 < ntask:
                    read((t, ch, k + 1), so_next)
                elif has_next:
                    read((t, ch + 1, 0), so_next)
                so = so_next
        # every real copy of the tile has been consumed before the next stage_start overwrites the ring
        assert all(slot_read), f"tile {t} left unread slots: {slot_task}"
        for g in groups:
            g[2] = True


@pytest.mark.parametrize("ntask,depth", [(16, 6), (8, 3)])
@pytest.mark.parametrize("nch", [1, 2, 3, 12])
def test_scan2_codebook_staging_ring(ntask, depth, nch):
    for seed in range(20):
        _staging_run(ntask, depth, nch, tiles=3, seed=seed)
