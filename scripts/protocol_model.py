#!/usr/bin/env python
"""CPU-only randomized model of the persistent kernel's barrier protocol (csrc/fa_fwd_sm100_persist.cuh).

The four roles (scheduler/producer, UMMA issuer, two softmax warpgroups) are Python generators that mirror the kernel's
control flow statement by statement -- same barrier table, same parity bookkeeping, same hoist / release rules -- and
yield at every mbarrier wait.  A random scheduler interleaves them; TMA loads and tensor-core work complete after random
delays through in-order queues.  Shadow state catches what a deadlock-only test would miss:

  * a ring slot / Q buffer overwritten by TMA while an issued-but-unfinished MMA still reads it,
  * an MMA that reads a slot holding the wrong (item, tile) -- i.e. a release that came too early or a parity slip,
  * S_t / P_t / O_t in TMEM overwritten before their consumer is done (softmax read, PV read, epilogue read),
  * a barrier completing a phase its waiter never observed (phase overrun),
  * every item's every tile gets exactly its n KV tiles, in order, and every output tile is stored once.

Run:  python scripts/protocol_model.py [trials]      (exit code 0 = no violation found)
"""
import random
import sys


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0
        self.completions = 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: too many arrivals"
        if self.pending == 0:
            self.pending = self.count
            self.phase ^= 1
            self.completions += 1

    def done(self, parity):          # try_wait.parity
        return self.phase != parity


class Sim:
    def __init__(self, items, nstage, rng, hoist=True, two=False, handoffs=2):
        self.handoffs = handoffs      # TFA_P_HANDOFFS: 2 (persistent kernel) or 3 (one-CTA-per-item kernel)
        self.items, self.N, self.rng, self.hoist, self.two = items, nstage, rng, hoist, two     # items: list of (n0, n1)
        B = lambda n, c: Bar(n, c)
        self.q_full = [B(f"q_full{t}", 1) for t in range(2)]
        self.q_empty = [B(f"q_empty{t}", 1) for t in range(2)]
        self.kv_full = [B(f"kv_full{i}", 1) for i in range(nstage)]
        self.kv_empty = [B(f"kv_empty{i}", 2 if two else 1) for i in range(nstage)]
        self.s_full = [B(f"s_full{t}", 1) for t in range(2)]
        self.p_half = [B(f"p_half{t}", 1) for t in range(2)]      # (the 4 warps of a warpgroup arrive together: count 1 here)
        self.p_3q = [B(f"p_3q{t}", 1) for t in range(2)]
        self.p_full = [B(f"p_full{t}", 1) for t in range(2)]
        self.o_full = [B(f"o_full{t}", 1) for t in range(2)]
        self.sched_full = [B(f"sched_full{i}", 1) for i in range(2)]
        self.sched_empty = [B(f"sched_empty{i}", 4 if two else 3) for i in range(2)]   # issuer(s) + 2 warpgroups
        self.sched_ring = [None, None]
        # shadow state
        self.q_buf = [None, None]            # (item, t) resident (after TMA completion)
        self.kv_buf = [None] * nstage        # (item, j, 'K'|'V')
        self.q_readers = [0, 0]              # issued, unfinished MMAs reading the buffer
        self.kv_readers = [0] * nstage
        self.tmem_S = [None, None]           # ('S', item, j) after the S MMA completed; ('P', item, j, stage) while being written
        self.S_unread = [False, False]       # S produced, not yet loaded by the softmax warpgroup
        self.P_readers = [0, 0]              # issued, unfinished PV MMAs reading P_t
        self.O_state = [None, None]          # (item, tiles accumulated) | None
        self.O_unread = [False, False]       # final O waiting for the epilogue
        self.tma_q = []                      # in-flight loads: [remaining, fn]
        self.mma_q = []                      # in-order tensor queue: [remaining, fn]
        self.stored = {}
        self.counter = 0
        self.log = []

    # ---- async engines ----
    def tma(self, fn):
        self.tma_q.append([self.rng.randint(1, 12), fn])

    def mma(self, fn_start, fn_done):
        """tcgen05.mma: reads happen while it executes (in order), completion effects at the end."""
        self.mma_q.append([self.rng.randint(1, 6), fn_start, fn_done, False])

    def tick(self):
        for e in self.tma_q:
            e[0] -= 1
        for e in [e for e in self.tma_q if e[0] <= 0]:
            self.tma_q.remove(e)
            e[1]()
        if self.mma_q:
            e = self.mma_q[0]
            if not e[3]:
                e[1]()
                e[3] = True
            e[0] -= 1
            if e[0] <= 0:
                self.mma_q.pop(0)
                e[2]()

    # ---- roles ----
    def fetch(self):
        while True:
            i = self.counter
            self.counter += 1
            if i >= len(self.items):
                return len(self.items)
            if max(self.items[i]) > 0:
                return i

    def producer(self):
        total, N = len(self.items), self.N
        ent, qpar, k = 0, [0, 0], 0

        def publish(kk, item):
            yield lambda: self.sched_empty[kk & 1].done(((kk >> 1) & 1) ^ 1)
            self.sched_ring[kk & 1] = item
            self.sched_full[kk & 1].arrive()

        cur = self.fetch()
        yield from publish(0, cur)
        while cur < total:
            n = self.items[cur]
            nmax = max(n)

            def load_q(t, cur=cur):
                if n[t] > 0:
                    yield lambda: self.q_empty[t].done(qpar[t] ^ 1)
                    qpar[t] ^= 1
                    assert self.q_readers[t] == 0, f"TMA overwrites Q{t} while an MMA reads it (item {cur})"
                    self.q_buf[t] = None

                    def land(t=t, cur=cur):
                        self.q_buf[t] = (cur, t)
                        self.q_full[t].arrive()
                    self.tma(land)

            def load_kv(j, kind, cur=cur):
                nonlocal ent
                slot, par = ent % N, (ent // N) & 1
                yield lambda: self.kv_empty[slot].done(par ^ 1)
                assert self.kv_readers[slot] == 0, f"TMA overwrites ring slot {slot} while an MMA reads it (item {cur} {kind}{j})"
                self.kv_buf[slot] = None

                def land(slot=slot, j=j, kind=kind, cur=cur):
                    self.kv_buf[slot] = (cur, j, kind)
                    self.kv_full[slot].arrive()
                self.tma(land)
                ent += 1

            yield from load_q(0)
            yield from load_kv(0, 'K')
            yield from load_kv(0, 'V')
            yield from load_q(1)
            for j in range(1, nmax):
                yield from load_kv(j, 'K')
                yield from load_kv(j, 'V')
            nxt = self.fetch()                 # the next item is drawn as late as possible: after ALL loads of this item
            yield from publish(k + 1, nxt)
            cur = nxt
            k += 1

    def sched_get(self, k):
        yield lambda: self.sched_full[k & 1].done((k >> 1) & 1)
        item = self.sched_ring[k & 1]
        self.sched_empty[k & 1].arrive()
        return item

    def issuer(self):
        total, N = len(self.items), self.N
        ent_base, k = 0, 0
        qfull_par, p_par, hoisted = [0, 0], [0, 0], [False, False]
        slot_of = lambda e: e % N
        par_of = lambda e: (e // N) & 1

        def issue_S(item, t, j, kslot, release_kv, release_q):
            self.q_readers[t] += 1
            self.kv_readers[kslot] += 1

            def start():
                assert self.q_buf[t] == (item, t), f"S{t}(item {item}, j {j}) reads Q buffer holding {self.q_buf[t]}"
                assert self.kv_buf[kslot] == (item, j, 'K'), f"S{t}(item {item}, j {j}) reads slot {kslot} holding {self.kv_buf[kslot]}"
                assert not self.S_unread[t], f"S{t}(item {item}, j {j}) overwrites an S tile the softmax has not read"
                assert self.P_readers[t] == 0 or True    # in-order pipe: earlier PV_t finished before this starts
                self.tmem_S[t] = None

            def done():
                self.q_readers[t] -= 1
                self.kv_readers[kslot] -= 1
                self.tmem_S[t] = ('S', item, j)
                self.S_unread[t] = True
                self.s_full[t].arrive()
                if release_kv:
                    self.kv_empty[kslot].arrive()
                if release_q:
                    self.q_empty[t].arrive()
            self.mma(start, done)

        def issue_PV(item, t, j, vslot, part, release_kv, done_bar):
            self.kv_readers[vslot] += 1
            self.P_readers[t] += 1

            def start():
                assert self.kv_buf[vslot] == (item, j, 'V'), f"PV{t}(item {item}, j {j}) reads slot {vslot} holding {self.kv_buf[vslot]}"
                ts = self.tmem_S[t]
                assert ts is not None and ts[0] == 'P' and ts[1:3] == (item, j) and ts[3] >= part, \
                    f"PV{t}(item {item}, j {j}, part {part}) reads P = {ts}"
                if part == 1:
                    if j == 0:
                        assert not self.O_unread[t], f"PV{t}(item {item}) overwrites an O tile the epilogue has not read"
                        self.O_state[t] = (item, 0)
                    assert self.O_state[t] == (item, j), f"PV{t}(item {item}, j {j}) accumulates onto {self.O_state[t]}"

            def done():
                self.kv_readers[vslot] -= 1
                self.P_readers[t] -= 1
                if part == 3:
                    self.O_state[t] = (item, j + 1)
                if release_kv:
                    self.kv_empty[vslot].arrive()
                if done_bar:
                    self.O_unread[t] = True
                    self.o_full[t].arrive()
            self.mma(start, done)

        cur = yield from self.sched_get(0)
        while cur < total:
            n = self.items[cur]
            nmax = max(n)
            ent_next = ent_base + 2 * nmax

            def first_S(item, t, x_nt, other_done, e0):
                yield lambda: self.q_full[t].done(qfull_par[t])
                qfull_par[t] ^= 1
                yield lambda: self.kv_full[slot_of(e0)].done(par_of(e0))
                issue_S(item, t, 0, slot_of(e0), other_done, x_nt == 1)

            for t in range(2):
                nt, no = n[t], n[t ^ 1]
                if nt > 0 and not hoisted[t]:
                    other_done = (no == 0) or hoisted[t ^ 1] or (t == 1)
                    yield from first_S(cur, t, nt, other_done, ent_base)
                    hoisted[t] = True
            hoisted = [False, False]
            nxt = None                                   # picked up lazily (non-blocking) at the hoist probes
            nn = (0, 0)

            kv_confirmed = False
            for j in range(nmax):
                ev, ek = ent_base + 2 * j + 1, ent_base + 2 * j + 2
                vslot, kslot = slot_of(ev), slot_of(ek)
                if not kv_confirmed:
                    yield lambda: self.kv_full[vslot].done(par_of(ev))
                    if j + 1 < nmax:
                        yield lambda: self.kv_full[kslot].done(par_of(ek))
                kv_confirmed = False
                for t in range(2):
                    nt, no = n[t], n[t ^ 1]
                    active = j < nt
                    last_v_user = (t == 1) or (j >= no)
                    last_k_user = (t == 1) or (j + 1 >= no)
                    has_next = j + 1 < nt
                    if active:
                        ppar = p_par[t]
                        yield lambda: self.p_half[t].done(ppar)
                        issue_PV(cur, t, j, vslot, 1, False, False)
                        if t == 1 and j + 1 < nmax:
                            yield lambda: self.kv_full[slot_of(ev + 2)].done(par_of(ev + 2))
                            if j + 2 < nmax:
                                yield lambda: self.kv_full[slot_of(ek + 2)].done(par_of(ek + 2))
                            kv_confirmed = True
                        if self.handoffs == 3:
                            yield lambda: self.p_3q[t].done(ppar)
                            issue_PV(cur, t, j, vslot, 2, False, False)
                        yield lambda: self.p_full[t].done(ppar)
                        p_par[t] ^= 1
                        issue_PV(cur, t, j, vslot, 3, last_v_user, not has_next)
                    if active and has_next:
                        issue_S(cur, t, j + 1, kslot, last_k_user, j + 2 == nt)
                    else:
                        if self.hoist and not hoisted[t] and nxt is None and self.sched_full[(k + 1) & 1].done(((k + 1) >> 1) & 1):
                            nxt = yield from self.sched_get(k + 1)
                            nn = self.items[nxt] if nxt < total else (0, 0)
                        has_nxt = nxt is not None and nxt < total
                        nnt, nno = nn[t], nn[t ^ 1]
                        if (self.hoist and has_nxt and nnt > 0 and not hoisted[t] and self.q_full[t].done(qfull_par[t])
                                and self.kv_full[slot_of(ent_next)].done(par_of(ent_next))):
                            other_done = (nno == 0) or hoisted[t ^ 1]
                            qfull_par[t] ^= 1
                            issue_S(nxt, t, 0, slot_of(ent_next), other_done, nnt == 1)
                            hoisted[t] = True
            if nxt is None:
                nxt = yield from self.sched_get(k + 1)
            ent_base = ent_next
            cur = nxt
            k += 1

    def issuer2(self, t):
        """The two-issuer variant (one issuer per tile, K/V entries released by count) that was measured 28 % slower on B200 and
        removed from the kernel in r02; its protocol is kept here because it is the simpler one to reason about."""
        total, N = len(self.items), self.N
        slot_of = lambda e: e % N
        par_of = lambda e: (e // N) & 1
        ent_base, k, qpar, ppar = 0, 0, 0, 0

        def S(item, j, e, release_q):
            kslot = slot_of(e)
            self.q_readers[t] += 1
            self.kv_readers[kslot] += 1

            def start():
                assert self.q_buf[t] == (item, t), f"S{t}(item {item}, j {j}) reads Q buffer holding {self.q_buf[t]}"
                assert self.kv_buf[kslot] == (item, j, 'K'), f"S{t}(item {item}, j {j}) reads slot {kslot} holding {self.kv_buf[kslot]}"
                assert not self.S_unread[t], f"S{t}(item {item}, j {j}) overwrites an unread S tile"
                self.tmem_S[t] = None

            def done():
                self.q_readers[t] -= 1
                self.kv_readers[kslot] -= 1
                self.tmem_S[t] = ('S', item, j)
                self.S_unread[t] = True
                self.s_full[t].arrive()
                self.kv_empty[kslot].arrive()
                if release_q:
                    self.q_empty[t].arrive()
            self.mma(start, done)

        def PV(item, j, e, part, release, done_bar):
            vslot = slot_of(e)
            self.kv_readers[vslot] += 1
            self.P_readers[t] += 1

            def start():
                assert self.kv_buf[vslot] == (item, j, 'V'), f"PV{t}(item {item}, j {j}) reads slot {vslot} holding {self.kv_buf[vslot]}"
                ts = self.tmem_S[t]
                assert ts is not None and ts[0] == 'P' and ts[1:3] == (item, j) and ts[3] >= part, f"PV{t}(item {item}, j {j}, part {part}) reads P = {ts}"
                if part == 1:
                    if j == 0:
                        assert not self.O_unread[t], f"PV{t}(item {item}) overwrites an unread O tile"
                        self.O_state[t] = (item, 0)
                    assert self.O_state[t] == (item, j)

            def done():
                self.kv_readers[vslot] -= 1
                self.P_readers[t] -= 1
                if part == 3:
                    self.O_state[t] = (item, j + 1)
                if release:
                    self.kv_empty[vslot].arrive()
                if done_bar:
                    self.O_unread[t] = True
                    self.o_full[t].arrive()
            self.mma(start, done)

        def pass_entry(e):
            yield lambda: self.kv_full[slot_of(e)].done(par_of(e))
            self.kv_empty[slot_of(e)].arrive()

        cur = yield from self.sched_get(0)
        while cur < total:
            n = self.items[cur]
            nt, nmax = n[t], max(n)
            if nt > 0:
                yield lambda: self.q_full[t].done(qpar)
                qpar ^= 1
                yield lambda: self.kv_full[slot_of(ent_base)].done(par_of(ent_base))
                S(cur, 0, ent_base, nt == 1)
            else:
                yield from pass_entry(ent_base)
            for j in range(nmax):
                ev, ek = ent_base + 2 * j + 1, ent_base + 2 * j + 2
                k_exists = j + 1 < nmax
                if j < nt:
                    has_next = j + 1 < nt
                    yield lambda: self.kv_full[slot_of(ev)].done(par_of(ev))
                    if has_next:
                        yield lambda: self.kv_full[slot_of(ek)].done(par_of(ek))
                    yield lambda: self.p_half[t].done(ppar)
                    PV(cur, j, ev, 1, False, False)
                    yield lambda: self.p_3q[t].done(ppar)
                    PV(cur, j, ev, 2, False, False)
                    yield lambda: self.p_full[t].done(ppar)
                    ppar ^= 1
                    PV(cur, j, ev, 3, True, not has_next)
                    if has_next:
                        S(cur, j + 1, ek, j + 2 == nt)
                    elif k_exists:
                        yield from pass_entry(ek)
                else:
                    yield from pass_entry(ev)
                    if k_exists:
                        yield from pass_entry(ek)
            ent_base += 2 * nmax
            cur = yield from self.sched_get(k + 1)
            k += 1

    def softmax(self, t):
        total = len(self.items)
        scnt = ocnt = 0
        k = 0
        while True:
            item = yield from self.sched_get(k)
            k += 1
            if item >= total:
                break
            n = self.items[item][t]
            if n == 0:
                continue
            for j in range(n):
                yield lambda: self.s_full[t].done(scnt & 1)
                assert self.tmem_S[t] == ('S', item, j), f"softmax{t} expected S(item {item}, j {j}), TMEM holds {self.tmem_S[t]}"
                self.S_unread[t] = False                       # row in registers
                if j > 0:
                    assert self.O_state[t] == (item, j), f"softmax{t} rescales O in state {self.O_state[t]} at (item {item}, j {j})"
                for stage, barl in ((1, self.p_half), (2, self.p_3q), (3, self.p_full)):
                    yield lambda: True                          # (compute; lets other roles interleave)
                    assert self.P_readers[t] == 0 or stage > 1, f"softmax{t} overwrites P while PV reads it"
                    self.tmem_S[t] = ('P', item, j, stage)
                    if stage != 2 or self.handoffs == 3 or self.two:
                        barl[t].arrive()
                scnt += 1
            yield lambda: self.o_full[t].done(ocnt & 1)
            ocnt += 1
            assert self.O_state[t] == (item, n), f"epilogue{t} of item {item}: O holds {self.O_state[t]}, want {n} tiles"
            self.O_unread[t] = False
            key = (item, t)
            assert key not in self.stored, f"tile {key} stored twice"
            self.stored[key] = n

    def run(self, max_steps=2_000_000):
        if self.two:
            roles = {"producer": self.producer(), "issuer0": self.issuer2(0), "issuer1": self.issuer2(1),
                     "sm0": self.softmax(0), "sm1": self.softmax(1)}
        else:
            roles = {"producer": self.producer(), "issuer": self.issuer(), "sm0": self.softmax(0), "sm1": self.softmax(1)}
        waiting = {}
        for name, g in roles.items():
            try:
                waiting[name] = next(g)
            except StopIteration:
                waiting[name] = None
        steps = idle = 0
        while any(w is not None for w in waiting.values()):
            steps += 1
            assert steps < max_steps, "did not terminate"
            names = [nm for nm, w in waiting.items() if w is not None]
            self.rng.shuffle(names)
            progressed = False
            for nm in names[:self.rng.randint(1, len(names))]:
                if waiting[nm]():
                    progressed = True
                    try:
                        waiting[nm] = next(roles[nm])
                    except StopIteration:
                        waiting[nm] = None
            if self.rng.random() < 0.7:
                had = bool(self.tma_q or self.mma_q)
                self.tick()
                progressed = progressed or had
            idle = 0 if progressed else idle + 1
            if idle > 2000 and not (self.tma_q or self.mma_q):
                blocked = [nm for nm, w in waiting.items() if w is not None]
                raise AssertionError(f"DEADLOCK: blocked roles {blocked}")
        while self.tma_q or self.mma_q:
            self.tick()
        want = {(i, t): n[t] for i, n in enumerate(self.items) for t in range(2) if n[t] > 0}
        assert self.stored == want, f"stored tiles {len(self.stored)} != expected {len(want)}"


def random_items(rng):
    kind = rng.choice(["causal", "noncausal", "ragged", "split", "mixed"])
    n_items = rng.randint(1, 9)
    items = []
    for _ in range(n_items):
        if kind == "causal":
            a = rng.randint(1, 6)
            items.append((a, a + 1))
        elif kind == "noncausal":
            a = rng.randint(1, 6)
            items.append((a, a))
        elif kind == "ragged":
            a = rng.randint(1, 5)
            items.append((a, rng.choice([0, a])))
        elif kind == "split":
            items.append(rng.choice([(0, 0), (0, 1), (1, 2), (2, 2), (3, 3), (1, 1), (0, 2)]))
        else:
            items.append((rng.randint(0, 4), rng.randint(0, 5)))
    return items


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    for trial in range(trials):
        rng = random.Random(seed0 * 1000003 + trial)
        items = random_items(rng)
        nstage = rng.choice([4, 8])
        try:
            Sim(items, nstage, rng, hoist=rng.random() < 0.8, two=(trial % 2 == 1)).run()
        except AssertionError as e:
            print(f"VIOLATION trial {trial} seed {seed0} nstage {nstage} items {items}: {e}")
            sys.exit(1)
    print(f"protocol model: {trials} random trials, no violation")


if __name__ == "__main__":
    main()
