"""The bucket kernel's lock-free protocol (state word, note queue, candidate stamps) under arbitrary interleavings:
whatever order the lanes meet in, a pass must give the sequential answer.  CPU stand-in for a race checker on
`sk_bucket_body` (csrc/ga_superkmer.cu); the GPU side of the same claim is test_bucket_path_is_deterministic and the
parity tests over several bucket geometries."""
import random

import pytest

import bucket_protocol_model as bp


def _bucket(seed: int, n_sites: int = 12, coverage=(4, 30), n_errors: int = 40):
    """Occurrences of one bucket: a few genomic windows seen many times with one or two different next symbols
    (some at a read end: no next symbol), plus error windows seen once to three times."""
    rng = random.Random(seed)
    out = []
    tag = 0

    def add(key, copies, symbols):
        nonlocal tag
        for _ in range(copies):
            follows = rng.random() < 0.9
            ordinal = rng.randrange(1 << 40)
            out.append(bp.Occurrence(key, rng.choice(symbols), ordinal, follows, ((tag + 1) << 5) | rng.randrange(32)))
            tag += 1

    for _ in range(n_sites):
        add(rng.getrandbits(60), rng.randint(*coverage), rng.sample(range(4), rng.choice((1, 1, 2))))
    for _ in range(n_errors):
        add(rng.getrandbits(60), rng.choice((1, 1, 1, 2, 3)), [rng.randrange(4)])
    rng.shuffle(out)
    return out


def _schedulers(seed: int):
    rng = random.Random(seed)
    turn = [0]

    def round_robin(n):
        turn[0] += 1
        return turn[0] % n

    return {
        "random": lambda n: rng.randrange(n),
        "round-robin": round_robin,                           # every lane one step at a time: maximal overlap
        "lifo": lambda n: n - 1,                               # the last lane runs to completion first
        "fifo": lambda n: 0,                                   # no overlap at all: the sequential order
        "bursty": lambda n: 0 if rng.random() < 0.8 else rng.randrange(n),
    }


@pytest.mark.parametrize("threshold", [0, 1, 2, 3, 7])
def test_every_interleaving_gives_the_sequential_answer(threshold):
    for seed in range(6):
        occ = _bucket(seed)
        want = bp.sequential(occ, threshold)
        distinct = len({o.key for o in occ})
        for cap in (2 * ((distinct * 5 // 4 + 1) // 2) + 2, 256):        # crowded table (long probe chains) and roomy one
            for lanes in (2, 32, 96):
                for name, choose in _schedulers(seed * 31 + lanes).items():
                    solid, cand, stats = bp.bucket_pass(occ, cap, threshold, lanes, choose)
                    assert (solid, cand) == want, (threshold, seed, cap, lanes, name)
                    assert stats["solid"] == len(want[0])


def test_lanes_meeting_on_one_window():
    """The worst case for the state word: every lane holds an occurrence of the SAME window at the same moment
    (occurrences sorted by key, dealt round-robin, lanes stepped in lock step), for every threshold around the count."""
    rng = random.Random(5)
    key = rng.getrandbits(60)
    for copies in (1, 2, 3, 4, 5, 33):
        occ = [bp.Occurrence(key, i % 3, 1000 - i, i % 5 != 0, ((i + 1) << 5) | 1) for i in range(copies)]
        occ += [bp.Occurrence(key ^ 1, 2, 7, True, ((copies + 1) << 5) | 2)]          # a neighbour in the same slot pair
        for threshold in range(0, copies + 2):
            want = bp.sequential(occ, threshold)
            for lanes in (copies + 1, 4):
                for name, choose in _schedulers(copies * 7 + threshold).items():
                    got = bp.bucket_pass(occ, 64, threshold, lanes, choose)
                    assert got[:2] == want, (copies, threshold, lanes, name)


def test_notes_are_left_only_by_the_first_threshold_occurrences_and_pending_races():
    """Bookkeeping the kernel's sizing relies on: without overlap (fifo) a window leaves at most `threshold` notes
    (the occurrences before it turns solid; the first one only when a second one comes); overlap adds at most the
    occurrences that met the window while its stamp slots were being set up (PENDING)."""
    for seed in range(4):
        occ = _bucket(100 + seed)
        for threshold in (1, 3):
            tally = {}
            for o in occ:
                if o.follows:
                    tally[o.key] = tally.get(o.key, 0) + 1
            calm = bp.bucket_pass(occ, 256, threshold, 8, lambda n: 0)[2]["notes"]
            assert calm <= sum(min(n, threshold) for n in tally.values())
            wild = bp.bucket_pass(occ, 256, threshold, 64, _schedulers(seed)["round-robin"])[2]["notes"]
            assert wild <= sum(tally.values())


def test_the_model_notices_a_broken_protocol():
    """The check has teeth: drop the hand-over of the FIRST state's back-reference to the queue and the earliest
    occurrence of some window is lost under some interleaving."""
    class Broken(bp.Bucket):
        def note(self, first, slot, payload):
            if not first:
                self.queue.append((first, slot, payload))
            yield

    occ = [bp.Occurrence(12345, 1, 1 if i == 0 else 100 + i, True, ((i + 1) << 5) | 3) for i in range(6)]
    want = bp.sequential(occ, 3)
    assert want[1] == {(12345, 1): 1}
    by_where = {o.where: o for o in occ}
    bucket = Broken(64, 3, by_where)
    bp.run([bucket.lane(occ)], lambda n: 0)
    bp.run([bucket.note_lane(list(bucket.queue))], lambda n: 0)
    solid, cand = bucket.result()
    assert solid == want[0] and cand == {(12345, 1): 101}
    assert bp.bucket_pass(occ, 64, 3, 1, lambda n: 0)[:2] == want


def _all_schedules(make_and_run):
    """Every interleaving, depth first: `make_and_run(choose)` builds fresh state and runs it; a schedule is replayed
    from its prefix and continued with lane 0, and every later step with more than one live lane branches."""
    stack, seen = [[]], 0
    while stack:
        prefix = stack.pop()
        trace = []                       # (choice, live lanes) per step

        def choose(n, prefix=prefix, trace=trace):
            pick = prefix[len(trace)] if len(trace) < len(prefix) else 0
            trace.append((pick, n))
            return pick

        make_and_run(choose)
        seen += 1
        for i in range(len(prefix), len(trace)):
            for alt in range(1, trace[i][1]):
                stack.append([c for c, _ in trace[:i]] + [alt])
    return seen


@pytest.mark.parametrize("threshold", [0, 1, 2, 3])
def test_two_lanes_exhaustively(threshold):
    """EVERY interleaving of two lanes that meet on one window, from every state the window can be in: `before`
    earlier occurrences have already been counted (0 .. threshold + 1), then two lanes bring one occurrence each."""
    total = 0
    for before in range(0, threshold + 2):
        for follows in ((True, True), (True, False), (False, True)):
            occ = [bp.Occurrence(77, i % 2, 500 - 7 * i, True, ((i + 1) << 5) | 2) for i in range(before)]
            pair = [bp.Occurrence(77, 1, 40, follows[0], (30 << 5) | 4), bp.Occurrence(77, 1, 30, follows[1], (31 << 5) | 5)]
            want = bp.sequential(occ + pair, threshold)
            by_where = {o.where: o for o in occ + pair if o.follows}

            def make_and_run(choose):
                bucket = bp.Bucket(8, threshold, by_where)
                bp.run([bucket.lane(occ)], lambda n: 0)                      # the window's history, in order
                bp.run([bucket.lane(pair[:1]), bucket.lane(pair[1:])], choose)
                notes = list(bucket.queue)
                bp.run([bucket.note_lane(notes[0::2]), bucket.note_lane(notes[1::2])], lambda n: n - 1)
                assert bucket.result() == want, (threshold, before, follows)

            total += _all_schedules(make_and_run)
    assert total > 1000
