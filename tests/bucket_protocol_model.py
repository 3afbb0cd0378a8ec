"""Executable model of the LOCK-FREE protocol inside the bucket kernel (genome-assembler_b200/csrc/ga_superkmer.cu,
`sk_bucket_body`): the shared-memory table K[slot] / A[slot], the state word that only moves forward
(0 -> FIRST -> REPEAT -> PENDING -> SOLID), the note queue and the candidate stamps.

Test infrastructure, pure Python, no code shared with the product.  superkmer_model.py states WHAT a bucket pass
must compute; this file states HOW the kernel's lanes compute it together and lets a scheduler interleave them at
every shared-memory access, so that the result can be checked for EVERY order in which lanes meet -- the evidence
compute-sanitizer would give on the GPU (closed on the pool, DESIGN.md "Races and determinism"), on the protocol
instead of on the binary.

A lane is a generator over its occurrences; every `yield` is a point where another lane may run.  What happens
between two yields is one indivisible hardware operation of the kernel: one load, one store, one CAS, one atomicAdd.
Line numbers name the statement of `sk_bucket_body` a step stands for.
"""
from __future__ import annotations

NONE = (1 << 64) - 1
FIRST, REPEAT, SOLID = 1 << 30, 2 << 30, 3 << 30
PAYLOAD = (1 << 30) - 1
PENDING = SOLID | PAYLOAD


class Occurrence:
    """One window occurrence as sk_for_each_window hands it to the walk: key, next symbol, ordinal, whether a next
    symbol exists, and `where` = (record tag + 1) << 5 | window number -- what finds it again from a FIRST state."""
    __slots__ = ("key", "c", "ord", "follows", "where")

    def __init__(self, key, c, ordinal, follows, where):
        self.key, self.c, self.ord, self.follows, self.where = key, c, ordinal, follows, where


class Bucket:
    """The shared state of one pass."""

    def __init__(self, cap: int, threshold: int, by_where: dict):
        assert cap % 2 == 0
        self.cap, self.threshold = cap, threshold
        self.K = [NONE] * cap
        self.A = [0] * cap
        self.queue = []                  # (first, slot, payload); ctl.n_q is its length
        self.skeys = []
        self.stamps = []                 # 4 per solid window
        self.by_where = by_where         # where -> Occurrence: the record re-read of phase D
        self.n_solid = 0

    @staticmethod
    def slot_hash(key: int) -> int:
        h = ((key & 0xFFFFFFFF) * 0x9E3779B1 ^ (key >> 32) * 0x85EBCA77) & 0xFFFFFFFF
        h ^= h >> 15
        return (h * 0xC2B2AE3D) & 0xFFFFFFFF

    # ---- the stamp: compare, then CAS only when the stored ordinal is larger (lambda `stamp`)
    def stamp(self, idx, c, ordinal):
        at = 4 * idx + c
        cur = self.stamps[at]                                   # stamp_ld
        yield
        while ordinal < cur:
            old = self.stamps[at]                               # stamp_cas
            if old == cur:
                self.stamps[at] = ordinal
            yield
            if old == cur:
                break
            cur = old

    def note(self, first, slot, payload):
        self.queue.append((first, slot, payload))               # atomicAdd(&ctl.n_q) + q_st: the entry is private
        yield                                                   # until the barrier that ends the walk

    # ---- one occurrence of the walk (the lambda given to sk_for_each_window)
    def occurrence(self, o: Occurrence):
        key, cap, F = o.key, self.cap, self.threshold
        s0 = ((self.slot_hash(key) * (cap >> 1)) >> 32) << 1
        while True:                                             # probe in slot pairs
            K0, K1 = self.K[s0], self.K[s0 + 1]                 # ld_k2 (one 128-bit load)
            yield
            if K0 == key:
                s = s0
                break
            if K1 == key:
                s = s0 + 1
                break
            if K0 == NONE or K1 == NONE:
                s = s0 if K0 == NONE else s0 + 1
                old = self.K[s]                                 # cas_k
                if old == NONE:
                    self.K[s] = key
                yield
                if old == NONE or old == key:
                    break
                continue                                        # somebody else's key landed there: look again
            s0 = 0 if s0 + 2 >= cap else s0 + 2
        mine = (o.c << 47) | o.ord
        a = self.A[s]                                           # ld_a
        yield
        while True:
            if a >= SOLID:
                if o.follows:
                    if a != PENDING:
                        yield from self.stamp(a & PAYLOAD, o.c, o.ord)
                    else:
                        yield from self.note(False, s, mine)
                return
            cnt = 0 if a == 0 else (1 if a < REPEAT else a & PAYLOAD)
            if cnt + 1 > F:                                     # this occurrence takes the window above the threshold
                old = self.A[s]                                 # cas_a(s, a, PENDING)
                if old == a:
                    self.A[s] = PENDING
                yield
                if old != a:
                    a = old
                    continue
                at = self.n_solid                               # atomicAdd(&ctl.n_solid)
                self.n_solid += 1
                self.skeys.append(None)
                self.stamps.extend([None] * 4)
                yield
                self.skeys[at] = key                            # skey_st
                yield
                for q in range(4):                              # stamp_st x 4
                    self.stamps[4 * at + q] = o.ord if (o.follows and q == o.c) else NONE
                    yield
                self.A[s] = SOLID | at                          # __threadfence_block + st_a
                yield
                if FIRST <= a < REPEAT and (a & PAYLOAD):
                    yield from self.note(True, s, a & PAYLOAD)
                return
            if a == 0:                                          # first occurrence: where to find it again
                old = self.A[s]                                 # cas_a(s, 0, FIRST | where)
                if old == 0:
                    self.A[s] = FIRST | (o.where if o.follows else 0)
                yield
                if old == 0:
                    return
                a = old
                continue
            old = self.A[s]                                     # cas_a(s, a, REPEAT | cnt + 1)
            if old == a:
                self.A[s] = REPEAT | (cnt + 1)
            yield
            if old != a:
                a = old
                continue
            if a < REPEAT and (a & PAYLOAD):
                yield from self.note(True, s, a & PAYLOAD)      # the first occurrence moves to the queue
            if o.follows:
                yield from self.note(False, s, mine)
            return

    def lane(self, occurrences):
        for o in occurrences:
            yield from self.occurrence(o)

    # ---- phase D: after the barrier, the notes whose slot ended SOLID fold their stamp in
    def note_lane(self, notes):
        for first, s, payload in notes:
            a = self.A[s]
            yield
            if a < SOLID:
                continue
            assert a != PENDING, "a window is still PENDING after the walk"
            idx = a & PAYLOAD
            if not first:
                yield from self.stamp(idx, (payload >> 47) & 3, payload & ((1 << 47) - 1))
            else:
                o = self.by_where[payload]                      # one 32-byte record load
                assert o.follows
                yield from self.stamp(idx, o.c, o.ord)

    def result(self):
        """({solid key}, {(key, c): min ordinal}) as superkmer_model.bucket_pass states them."""
        cand = {}
        for idx, key in enumerate(self.skeys):
            for c in range(4):
                if self.stamps[4 * idx + c] != NONE:
                    cand[(key, c)] = self.stamps[4 * idx + c]
        return set(self.skeys), cand


def run(generators, choose):
    """Drive the lanes to completion; `choose(n)` picks which of the n live lanes takes the next step."""
    live = list(generators)
    steps = 0
    while live:
        i = choose(len(live))
        try:
            next(live[i])
        except StopIteration:
            live.pop(i)
        steps += 1
    return steps


def bucket_pass(occurrences, cap, threshold, n_lanes, choose, deal=None):
    """The whole pass: walk (lanes over disjoint shares of the occurrences, interleaved by `choose`), barrier, notes
    (interleaved again).  Returns (solid keys, candidate stamps, statistics)."""
    by_where = {o.where: o for o in occurrences if o.follows}
    bucket = Bucket(cap, threshold, by_where)
    shares = [[] for _ in range(n_lanes)]
    for i, o in enumerate(occurrences):
        shares[(deal(i) if deal else i) % n_lanes].append(o)
    steps = run([bucket.lane(share) for share in shares], choose)
    notes = list(bucket.queue)                                  # __syncthreads(): the queue is complete and visible
    note_shares = [notes[i::n_lanes] for i in range(n_lanes)]
    steps += run([bucket.note_lane(share) for share in note_shares], choose)
    solid, cand = bucket.result()
    return solid, cand, {"steps": steps, "notes": len(notes), "solid": len(solid)}


def sequential(occurrences, threshold):
    """What the pass must compute (superkmer_model.bucket_pass, step 3)."""
    tally = {}
    for o in occurrences:
        tally[o.key] = tally.get(o.key, 0) + 1
    solid = {key for key, n in tally.items() if n > threshold}
    cand = {}
    for o in occurrences:
        if o.follows and o.key in solid:
            cand[(o.key, o.c)] = min(cand.get((o.key, o.c), NONE), o.ord)
    return solid, cand
