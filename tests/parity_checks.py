"""Parity checks shared by the CPU (hostsim) and GPU (C ABI) test suites.

`engine` is any object with the batched interface of gbwt_rs_b200.GBWT / tests.hostsim_build.HostSim;
`g` is the oracle index (oracle.oracle.GBWT) loaded from the same image. Everything is compared
bit-exactly: SearchState / BidirectionalState ranges, positions and extracted paths.
"""
from __future__ import annotations

import numpy as np

from oracle import oracle as orc

U64MAX = np.uint64(2**64 - 1)


def states_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def all_sequences(g):
    return [g.sequence(i) for i in range(g.sequences())]


def check_find_all_nodes(engine, g):
    """src/gbwt/tests.rs:268-292: find() for every id in 0..=alphabet_size (plus far-out ids)."""
    nodes = np.concatenate([np.arange(g.alphabet_size() + 2, dtype=np.uint64),
                            np.array([2**32 - 1, 2**32, 2**32 + 24, 2**63, 2**64 - 1], dtype=np.uint64)])
    assert states_equal(engine.find(nodes), g.find_batch(nodes))


def check_extend_all_pairs(engine, g):
    """src/gbwt/tests.rs:300-311: every (node, next) pair, full ranges and every sub-range."""
    first, nxt, st = [], [], []
    states = g.find_batch(np.arange(g.alphabet_size() + 1, dtype=np.uint64))
    for s in states:
        if s["end"] <= s["start"]:
            continue
        ranges = [(int(s["start"]), int(s["end"]))]
        n = int(s["end"])
        if n <= 6:
            ranges += [(a, b) for a in range(n + 1) for b in range(a, n + 2)]
        for a, b in ranges:
            for i in range(g.alphabet_size() + 1):
                st.append((int(s["node"]), a, b)); nxt.append(i)
    st = np.array(st, dtype=orc.STATE_DTYPE)
    nxt = np.array(nxt, dtype=np.uint64)
    assert states_equal(engine.extend(st, nxt), g.extend_batch(st, nxt))
    # hand-made states: node below the offset, endmarker node, out-of-range nodes and ranges
    weird = np.array([(0, 0, 5), (g.alphabet_offset(), 0, 3), (1, 0, 1), (g.alphabet_size() + 7, 0, 1),
                      (2**40, 0, 1), (g.first_node(), 5, 2), (g.first_node(), 0, 2**40), (g.first_node(), 2**33, 2**34)],
                     dtype=orc.STATE_DTYPE)
    for node in [0, 1, g.first_node(), g.first_node() + 1, g.alphabet_size() - 1, 2**35]:
        nn = np.full(len(weird), node, dtype=np.uint64)
        assert states_equal(engine.extend(weird, nn), g.extend_batch(weird, nn))


def subpath_patterns(seqs, k):
    return [p[j:j + k] for p in seqs for j in range(len(p) - k + 1)]


def check_find_extend_subpaths(engine, g, ks=(1, 2, 3, 4)):
    seqs = all_sequences(g)
    for k in ks:
        pats = subpath_patterns(seqs, k)
        if not pats:
            continue
        pats = np.array(pats, dtype=np.uint64)
        want = g.find_extend_batch(pats)
        assert states_equal(engine.find_extend(pats), want)
        assert np.all(want["end"] > want["start"])
    # ragged: every suffix of every sequence, plus an empty pattern
    nodes, offsets = [], [0]
    for p in seqs:
        for j in range(len(p)):
            nodes += p[j:]; offsets.append(len(nodes))
    offsets.append(len(nodes))
    want = g.find_extend_ragged(np.array(nodes, dtype=np.uint64), np.array(offsets, dtype=np.uint64))
    got = engine.find_extend_ragged(np.array(nodes, dtype=np.uint64), np.array(offsets, dtype=np.uint64))
    assert states_equal(got, want)


def check_find_extend_random(engine, g, n=4000, k=5, seed=0):
    """Mostly non-matching random patterns: exercises every None path."""
    rng = np.random.default_rng(seed)
    lo, hi = max(0, g.alphabet_offset() - 2), g.alphabet_size() + 3
    pats = rng.integers(lo, hi, size=(n, k), dtype=np.uint64)
    seqs = [p for p in all_sequences(g) if len(p) >= k]
    for i in range(0, n, 3):  # a third: real subpaths with one random mutation
        if seqs:
            p = seqs[rng.integers(len(seqs))]
            j = rng.integers(len(p) - k + 1)
            pats[i] = p[j:j + k]
            if i % 2:
                pats[i, rng.integers(k)] = rng.integers(lo, hi)
    assert states_equal(engine.find_extend(pats), g.find_extend_batch(pats))


def bd_triples(seqs):
    nodes, offsets, first, start, end = [], [0], [], [], []
    for p in seqs:
        for f in range(len(p)):
            for s in range(f + 1):
                for e in range(f + 1, len(p) + 1):
                    nodes += p; offsets.append(len(nodes))
                    first.append(f); start.append(s); end.append(e)
    return tuple(np.array(x, dtype=np.uint64) for x in (nodes, offsets, first, start, end))


def check_bd(engine, g):
    """src/gbwt/tests.rs:365-462: bd_find for every id; every (first, start, end) of every sequence."""
    nodes = np.arange(g.alphabet_size() + 2, dtype=np.uint64)
    states = g.bd_find_batch(nodes)
    assert states_equal(engine.bd_find(nodes), states)
    # all single-node extensions of every initial state, both directions (tests.rs:400-417)
    live = states[states["forward"]["end"] > states["forward"]["start"]]
    rep = np.repeat(live, g.alphabet_size() + 1)
    nxt = np.tile(np.arange(g.alphabet_size() + 1, dtype=np.uint64), len(live))
    assert states_equal(engine.extend_forward(rep, nxt), g.bd_extend_batch(rep, nxt, backward=False))
    assert states_equal(engine.extend_backward(rep, nxt), g.bd_extend_batch(rep, nxt, backward=True))
    seqs = all_sequences(g)
    args = bd_triples(seqs)
    want = g.bd_search_batch(*args)
    got = engine.bd_search(*args)
    assert states_equal(got, want)
    assert np.all(want["forward"]["end"] > want["forward"]["start"])
    # invalid (first, start, end) combinations are None
    if len(args[2]):
        bad = [a.copy() for a in args]
        bad[3][0] = bad[2][0] + 1  # start > first
        assert states_equal(engine.bd_search(*bad)[:1], np.zeros(1, orc.BDSTATE_DTYPE))
    return len(args[2])


def check_navigation(engine, g):
    """src/gbwt/tests.rs:164-238: start / forward / sequence for every sequence, plus invalid inputs."""
    ids = np.arange(g.sequences() + 3, dtype=np.uint64)
    starts = engine.start(ids)
    for i in ids:
        want = g.start(int(i)) or (0, 0)
        assert (int(starts[i]["node"]), int(starts[i]["offset"])) == want
    # forward at every (node, offset) up to len + 1
    pos = []
    for node in range(g.alphabet_size() + 2):
        st = g.find(node)
        n = (st[2] if st else 0) + 2
        pos += [(node, i) for i in range(n)]
    pos += [(2**33, 0), (5, 2**40)]
    pos = np.array(pos, dtype=orc.POS_DTYPE)
    assert states_equal(engine.forward(pos), g.forward_batch(pos))
    if g.is_bidirectional():
        # src/gbwt/tests.rs:191-214: backward() at every position (and past-the-end offsets)
        want = np.zeros(len(pos), dtype=orc.POS_DTYPE)
        for j, (node, off) in enumerate(pos):
            want[j] = g.backward((int(node), int(off))) or (0, 0)
        assert states_equal(engine.backward(pos), want)
    lengths = engine.sequence_lengths(ids)
    assert np.array_equal(lengths, g.sequence_lengths(ids))
    offsets, nodes, got = engine.extract(ids)
    for i in range(g.sequences()):
        assert list(nodes[int(offsets[i]):int(offsets[i + 1])]) == g.sequence(i)
    assert np.all(got[g.sequences():] == U64MAX)


def check_follow(engine, g, max_rounds=6):
    """src/gbz/tests.rs:100-168 (check_states): breadth-first over the states reachable from every node, all
    forward and backward extensions of every state against the oracle, and the oracle's follow() against
    extend_forward / extend_backward over the successors (the reference test's own truth)."""
    nodes = np.arange(g.alphabet_size() + 2, dtype=np.uint64)
    frontier = g.bd_find_batch(nodes)
    seen = set()
    total = 0
    for _ in range(max_rounds):
        if len(frontier) == 0:
            break
        nxt = []
        for backward in (False, True):
            offsets, want, counts = g.follow_batch(frontier, backward=backward)
            got_offsets, got, got_counts = engine.follow(frontier, backward)
            assert np.array_equal(got_counts, counts) and np.array_equal(got_offsets, offsets)
            assert states_equal(got, want)
            total += len(want)
            nxt.append(want)
        # the reference's own truth on a sample: follow == non-empty extend_* over every node id
        for st in frontier[:: max(1, len(frontier) // 8)]:
            tup = ((int(st["forward"]["node"]), int(st["forward"]["start"]), int(st["forward"]["end"])),
                   (int(st["reverse"]["node"]), int(st["reverse"]["start"]), int(st["reverse"]["end"])))
            if tup[0][2] <= tup[0][1]:
                continue
            found = g.follow(tup)
            if found is None:
                continue
            truth = [x for x in (g.extend_forward(tup, int(v)) for v in range(g.alphabet_size())) if x is not None]
            assert sorted(found) == sorted(truth)
        new = np.concatenate(nxt) if nxt else np.zeros(0, orc.BDSTATE_DTYPE)
        keep = []
        for row in new:
            key = row.tobytes()
            if key not in seen:
                seen.add(key)
                keep.append(row)
        frontier = np.array(keep, dtype=orc.BDSTATE_DTYPE) if keep else np.zeros(0, orc.BDSTATE_DTYPE)
        if len(frontier) > 3000:
            frontier = frontier[:3000]
    return total


def check_everything(engine, g):
    check_find_all_nodes(engine, g)
    check_extend_all_pairs(engine, g)
    check_find_extend_subpaths(engine, g)
    check_find_extend_random(engine, g)
    if g.is_bidirectional():
        check_bd(engine, g)
        check_follow(engine, g)
    check_navigation(engine, g)
