"""Barrier-protocol model of conv_tail_kernel (facerec_b200/csrc/conv_tail.cu), one CTA, no GPU needed.

The kernel's seven roles (three epilogue groups, the G1 and M2 issuers, the TMA producer, the store warp) talk through mbarriers whose waits tell
phases apart BY PARITY ONLY: a waiter that gets two phases ahead of a barrier takes the phase before last for its own and
either corrupts a buffer or hangs.  This script replays the protocol under random interleavings of the roles and of the
asynchronous completions (TMA loads, tcgen05.commit arrivals, store drains) and checks at every wait that the barrier is
in exactly the phase the waiter means, and that nothing deadlocks.  It found the two bugs of the first version
(groups taking slab slots independently; e2 items sharing the "slab written" phases of a slot) that only showed on the GPU
with more than one tile per CTA and 4 slots.

    python tools/conv_tail_protocol_sim.py            # the shipped protocol over a grid of (N2/64, slots, ring, tiles)
    python tools/conv_tail_protocol_sim.py --no-turn  # without in-order slot taking: reports the aliasing
"""
from __future__ import annotations

import random
import sys


class Bar:
    """mbarrier: `count` arrivals complete a phase; test(parity) is mbarrier.test_wait.parity."""

    def __init__(self, count):
        self.count = count
        self.pending = count
        self.phase = 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.pending = self.count
            self.phase += 1

    def test(self, parity):
        return (self.phase & 1) != parity


class ProtocolError(Exception):
    pass


def run(my_tiles, ns2, slots, ring, kb1=1, groups=3, seed=0, in_order=True, lag=1, nskip=4):
    """Returns (ok, detail).  ns2 = N2 / 64 (e2 items per tile); slots / ring as in ConvTailArgs."""
    rnd = random.Random(seed)
    ipt = 4 + ns2
    full = [Bar(1) for _ in range(8)]
    empty = [Bar(1) for _ in range(8)]
    slabw = [Bar(1) for _ in range(8)]
    sfree = [Bar(2) for _ in range(8)]
    swr = [Bar(1) for _ in range(8)]
    turn = [Bar(1) for _ in range(groups)]
    acc1, acc2 = [Bar(1), Bar(1)], [Bar(1), Bar(1)]
    afree = [Bar(4 * ns2), Bar(4 * ns2)]
    pending = []                               # asynchronous completions, in issue order

    def later(fn):
        pending.append(fn)

    def check(cond, *what):
        if not cond:
            raise ProtocolError(what)

    def epilogue(g):
        q = g
        while q < my_tiles * ipt:
            t, j = divmod(q, ipt)
            slot, use, buf, aph = q % slots, q // slots, t & 1, (t >> 1) & 1
            second = j >= 4
            if in_order and q >= 1:
                pg, pn = (g + groups - 1) % groups, (q - 1) // groups
                while not turn[pg].test(pn & 1):
                    yield ("turn", pg, q)
                check(turn[pg].phase == pn + 1, "turn barrier aliased", q)
            while not sfree[slot].test((use & 1) ^ 1):
                yield ("slot free", slot, q)
            check(sfree[slot].phase == use, "slot-free barrier aliased", q, slot, use, sfree[slot].phase)
            if in_order:
                turn[g].arrive()
            b = acc2[buf] if second else acc1[buf]
            while not b.test(aph):
                yield ("accumulator", buf, q)
            check(b.phase == (t >> 1) + 1, "accumulator barrier aliased", q)
            if second:
                for _ in range(4):
                    afree[buf].arrive()
            yield ("work", q)
            if not second:
                slabw[slot].arrive()
            swr[slot].arrive()
            q += groups

    def storer():
        """The store warp: items in order; TMA store of a written slab, the slot's "free" arrival(s) once the store has
        drained, at most `lag` stores behind."""
        inflight = []
        for q in range(my_tiles * ipt):
            slot, use, second = q % slots, q // slots, (q % ipt) >= 4
            while not swr[slot].test(use & 1):
                yield ("slab staged", slot, q)
            check(swr[slot].phase == use + 1, "slab-staged barrier aliased", q)
            done = [False]
            later(lambda d=done: d.__setitem__(0, True))
            inflight.append((slot, second, done))
            while len(inflight) > lag:
                pslot, psecond, pdone = inflight.pop(0)
                while not pdone[0]:
                    yield ("store drain", pslot)
                sfree[pslot].arrive()
                if psecond:
                    sfree[pslot].arrive()
        for pslot, psecond, pdone in inflight:
            while not pdone[0]:
                yield ("store drain", pslot)
            sfree[pslot].arrive()
            if psecond:
                sfree[pslot].arrive()

    def issuer(which):                         # G1 and M2 are issued by a warp each
        st = {"stage": 0, "phase": 0, "slot": 0}
        e1n = [0] * 8                          # e1 slabs consumed per slot

        def g1(t):
            buf = t & 1
            while not afree[buf].test(((t >> 1) & 1) ^ 1):
                yield ("buffer free", buf, t)
            check(afree[buf].phase == (t >> 1), "buffer-free barrier aliased", t)
            for i in range(kb1 + nskip):
                s = st["stage"]
                while not full[s].test(st["phase"]):
                    yield ("ring full", s)
                later(lambda s=s: empty[s].arrive())
                if i == kb1 + nskip - 1:
                    later(lambda b=buf: acc1[b].arrive())
                st["stage"] += 1
                if st["stage"] == ring:
                    st["stage"], st["phase"] = 0, st["phase"] ^ 1

        def m2(t):
            buf = t & 1
            for s in range(4):
                x = st["slot"]
                while not slabw[x].test(e1n[x] & 1):
                    yield ("slab written", x, t, s)
                check(slabw[x].phase == e1n[x] + 1, "slab-written barrier aliased", t, s)
                if s == 0 and ns2 == 2:
                    x1 = 0 if x + 1 == slots else x + 1
                    while not slabw[x1].test(e1n[x1] & 1):
                        yield ("slab written", x1, t, 1)
                    check(slabw[x1].phase == e1n[x1] + 1, "slab-written barrier aliased", t, 1)
                later(lambda x=x: sfree[x].arrive())
                if s == 3:
                    later(lambda b=buf: acc2[b].arrive())
                e1n[x] += 1
                st["slot"] = 0 if x + 1 == slots else x + 1
            for _ in range(ns2):
                st["slot"] = 0 if st["slot"] + 1 == slots else st["slot"] + 1

        for t in range(my_tiles):
            yield from (g1(t) if which == "g1" else m2(t))

    def producer():
        stage = phase = 0
        for _ in range(my_tiles):
            for _ in range(kb1 + nskip):
                while not empty[stage].test(phase ^ 1):
                    yield ("ring empty", stage)
                later(lambda s=stage: full[s].arrive())
                stage += 1
                if stage == ring:
                    stage, phase = 0, phase ^ 1

    agents = {f"epilogue {g}": epilogue(g) for g in range(groups)}
    agents["G1 issuer"] = issuer("g1")
    agents["M2 issuer"] = issuer("m2")
    agents["store warp"] = storer()
    agents["producer"] = producer()
    state, idle = {}, 0
    try:
        while agents:
            progressed = False
            names = list(agents)
            rnd.shuffle(names)
            for n in names:
                try:
                    w = next(agents[n])
                    if state.get(n) != w or w[0] == "work":
                        progressed = True
                    state[n] = w
                except StopIteration:
                    del agents[n]
                    state.pop(n, None)
                    progressed = True
            if pending:
                k = rnd.randint(1, len(pending))
                for fn in pending[:k]:
                    fn()
                del pending[:k]
                progressed = True
            idle = 0 if progressed else idle + 1
            if idle > 3:
                return False, ("deadlock", dict(state))
    except ProtocolError as e:
        return False, e.args[0]
    return True, None


CONFIGS = ((2, 2, 5), (1, 2, 7), (1, 5, 4), (1, 6, 3), (1, 3, 3), (1, 4, 5), (1, 8, 8), (2, 4, 3), (2, 3, 3), (2, 5, 3), (2, 6, 3), (2, 7, 3), (2, 8, 4))


def sweep(in_order=True, seeds=40, tiles=(1, 2, 3, 5, 12, 57)):
    failures = []
    for ns2, slots, ring in CONFIGS:
        for n in tiles:
            for seed in range(seeds):
                ok, detail = run(n, ns2, slots, ring, seed=seed, in_order=in_order, lag=min(seed % 3, slots - 1), nskip=4 if seed % 2 else 1)
                if not ok:
                    failures.append(((ns2, slots, ring, n, seed), detail))
                    break
    return failures


# ----------------------------------------------------------------------------------------------------------------------
# conv_tail2_kernel (layer2: 512 output channels of conv3 as two halves of 256, both weight matrices streamed)
def run2(my_tiles, slots, gring, wring, groups=3, seed=0, lag=1):
    """Per tile: halves h = 0, 1.  G1(t, h): 10 G-ring units (per k-block: a, W3 rows 0-127, W3 rows 128-255; then 4 skip
    slabs) into buffer X_t = t & 1 (h = 0) or Y_t = 1 - X_t (h = 1).  Items per tile: e1(h0) x 4, e1(h1) x 4, e2 x 2.
    M2(t, h): 4 slabs x one W-ring unit each, accumulating into columns [0, 128) of X_t.  Buffer Y_t is free once e1(t, h1)
    has drained it (16 warp arrivals), X_t once e2(t) has (8 warp arrivals)."""
    rnd = random.Random(seed)
    ipt = 10
    gfull = [Bar(1) for _ in range(8)]
    gempty = [Bar(1) for _ in range(8)]
    wfull = [Bar(1) for _ in range(8)]
    wempty = [Bar(1) for _ in range(8)]
    slabw = [Bar(1) for _ in range(8)]
    sfree = [Bar(2) for _ in range(8)]
    swr = [Bar(1) for _ in range(8)]
    turn = [Bar(1) for _ in range(groups)]
    acc1 = [Bar(1), Bar(1)]
    acc2 = [Bar(1), Bar(1)]
    free_e1 = [Bar(16), Bar(16)]               # buffer drained by the four e1 items of a second half
    free_e2 = [Bar(8), Bar(8)]                 # buffer drained by the two e2 items
    pending = []

    def later(fn):
        pending.append(fn)

    def check(cond, *what):
        if not cond:
            raise ProtocolError(what)

    def epilogue(g):
        q = g
        while q < my_tiles * ipt:
            t, j = divmod(q, ipt)
            slot, use = q % slots, q // slots
            X = t & 1
            second = j >= 8
            h = 0 if j < 4 else 1
            buf = X if (second or h == 0) else 1 - X
            if q >= 1:
                pg, pn = (g + groups - 1) % groups, (q - 1) // groups
                while not turn[pg].test(pn & 1):
                    yield ("turn", pg, q)
                check(turn[pg].phase == pn + 1, "turn barrier aliased", q)
            while not sfree[slot].test((use & 1) ^ 1):
                yield ("slot free", slot, q)
            check(sfree[slot].phase == use, "slot-free barrier aliased", q)
            turn[g].arrive()
            if second:
                b, want = acc2[buf], (t >> 1)
            else:
                b, want = acc1[buf], t                     # every buffer is filled once per tile
            while not b.test(want & 1):
                yield ("accumulator", buf, q)
            check(b.phase == want + 1, "accumulator barrier aliased", q, b.phase, want)
            if second:
                for _ in range(4):
                    free_e2[buf].arrive()
            elif h == 1:
                for _ in range(4):
                    free_e1[buf].arrive()
            yield ("work", q)
            if not second:
                slabw[slot].arrive()
            swr[slot].arrive()
            q += groups

    def storer():
        inflight = []
        for q in range(my_tiles * ipt):
            slot, use, second = q % slots, q // slots, (q % ipt) >= 8
            while not swr[slot].test(use & 1):
                yield ("slab staged", slot, q)
            check(swr[slot].phase == use + 1, "slab-staged barrier aliased", q)
            done = [False]
            later(lambda d=done: d.__setitem__(0, True))
            inflight.append((slot, second, done))
            while len(inflight) > lag:
                pslot, psecond, pdone = inflight.pop(0)
                while not pdone[0]:
                    yield ("store drain", pslot)
                sfree[pslot].arrive()
                if psecond:
                    sfree[pslot].arrive()
        for pslot, psecond, pdone in inflight:
            while not pdone[0]:
                yield ("store drain", pslot)
            sfree[pslot].arrive()
            if psecond:
                sfree[pslot].arrive()

    def issuer(which):                         # the G1 issuer and the M2 issuer are warps of their own
        st = {"g": 0, "gp": 0, "w": 0, "wp": 0, "slot": 0}
        e1n = [0] * 8

        def g1(t, h):
            X = t & 1
            buf = X if h == 0 else 1 - X
            if t >= 1:
                k = (t - 1) >> 1
                b = free_e1[buf] if h == 0 else free_e2[buf]
                while not b.test(k & 1):
                    yield ("buffer free", buf, t, h)
                check(b.phase == k + 1, "buffer-free barrier aliased", t, h, b.phase, k)
            for i in range(10):
                s = st["g"]
                while not gfull[s].test(st["gp"]):
                    yield ("G ring full", s)
                later(lambda s=s: gempty[s].arrive())
                if i == 9:
                    later(lambda b=buf: acc1[b].arrive())
                st["g"] += 1
                if st["g"] == gring:
                    st["g"], st["gp"] = 0, st["gp"] ^ 1

        def m2(t, h):
            X = t & 1
            for s in range(4):
                x = st["slot"]
                while not slabw[x].test(e1n[x] & 1):
                    yield ("slab written", x, t, h, s)
                check(slabw[x].phase == e1n[x] + 1, "slab-written barrier aliased", t, h, s)
                if s == 0 and h == 0:
                    x1 = 0 if x + 1 == slots else x + 1
                    while not slabw[x1].test(e1n[x1] & 1):
                        yield ("slab written", x1, t, h, 1)
                    check(slabw[x1].phase == e1n[x1] + 1, "slab-written barrier aliased", t, h, 1)
                ws = st["w"]
                while not wfull[ws].test(st["wp"]):
                    yield ("W ring full", ws)
                later(lambda ws=ws: wempty[ws].arrive())
                st["w"] += 1
                if st["w"] == wring:
                    st["w"], st["wp"] = 0, st["wp"] ^ 1
                later(lambda x=x: sfree[x].arrive())
                if s == 3 and h == 1:
                    later(lambda b=X: acc2[b].arrive())
                e1n[x] += 1
                st["slot"] = 0 if x + 1 == slots else x + 1
            if h == 1:
                for _ in range(2):
                    st["slot"] = 0 if st["slot"] + 1 == slots else st["slot"] + 1

        if which == "g1":
            for t in range(my_tiles):
                yield from g1(t, 0)
                yield from g1(t, 1)
        else:
            for t in range(my_tiles):
                yield from m2(t, 0)
                yield from m2(t, 1)

    def producer(full, empty, depth, per_tile):
        stage = phase = 0
        for _ in range(my_tiles * per_tile):
            while not empty[stage].test(phase ^ 1):
                yield ("ring empty", stage)
            later(lambda s=stage: full[s].arrive())
            stage += 1
            if stage == depth:
                stage, phase = 0, phase ^ 1

    agents = {f"epilogue {g}": epilogue(g) for g in range(groups)}
    agents["G1 issuer"] = issuer("g1")
    agents["M2 issuer"] = issuer("m2")
    agents["store warp"] = storer()
    agents["G producer"] = producer(gfull, gempty, gring, 20)
    agents["W producer"] = producer(wfull, wempty, wring, 8)
    state, idle = {}, 0
    try:
        while agents:
            progressed = False
            names = list(agents)
            rnd.shuffle(names)
            for n in names:
                try:
                    w = next(agents[n])
                    if state.get(n) != w or w[0] == "work":
                        progressed = True
                    state[n] = w
                except StopIteration:
                    del agents[n]
                    state.pop(n, None)
                    progressed = True
            if pending:
                k = rnd.randint(1, len(pending))
                for fn in pending[:k]:
                    fn()
                del pending[:k]
                progressed = True
            idle = 0 if progressed else idle + 1
            if idle > 3:
                return False, ("deadlock", dict(state))
    except ProtocolError as e:
        return False, e.args[0]
    return True, None


def sweep2(seeds=30, tiles=(1, 2, 3, 4, 7, 30)):
    failures = []
    for slots, gring, wring in ((3, 8, 2), (2, 8, 3), (4, 7, 2), (4, 6, 3), (4, 5, 3), (5, 5, 3), (6, 4, 2), (8, 3, 2)):
        for n in tiles:
            for seed in range(seeds):
                ok, detail = run2(n, slots, gring, wring, seed=seed, lag=min(seed % 3, slots - 1))
                if not ok:
                    failures.append(((slots, gring, wring, n, seed), detail))
                    break
    return failures


if __name__ == "__main__":
    bad = sweep(in_order="--no-turn" not in sys.argv) + sweep2()
    for cfg, detail in bad:
        print("FAIL (N2/64, slots, ring, tiles, seed) =", cfg, detail)
    print(f"{len(bad)} failing configurations")
    sys.exit(1 if bad else 0)
