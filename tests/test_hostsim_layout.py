"""Checks the product's device layout (K0, layout_builder.cpp) and the one-lane scan code
(record_scan.cuh) on the CPU against the oracle. The same checks run on the GPU through the C ABI in
test_gpu_parity.py. CPU only; tests/hostsim is test infrastructure, not a fallback."""
import os
import random

import numpy as np
import pytest

import gbwt_builder as gb
import golden_vectors as gv
import parity_checks as pc
from hostsim_build import HostSim
from oracle import oracle as orc
from synth import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = ["example.gbwt", "with-empty.gbwt", "translation.gbwt", "example.gbz", "translation.gbz",
            "example-v1.gbz", "translation-v1.gbz"]


def image_of(b, bidirectional=True):
    flags = 4 | (1 if bidirectional else 0)
    return synth.gbwt_image(b["sequences"], b["size"], b["offset"], b["alphabet_size"], flags, b["starts"], b["data"])


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("name", FIXTURES)
def test_fixtures(name, layout):
    raw = open(os.path.join(GOLDEN, name), "rb").read()
    g = orc.GBWT.load(raw)
    e = HostSim(raw, layout)
    pc.check_everything(e, g)
    if layout == 1:
        assert e.format_counts()[2] == 0  # no dense records under the runs-only policy


def test_config2_anchor_count():
    raw = open(os.path.join(GOLDEN, "translation.gbz"), "rb").read()
    assert pc.check_bd(HostSim(raw), orc.GBWT.load(raw)) == 504


def random_paths(rng, n_nodes, n_paths, max_len, reverse_prob=0.2):
    paths = []
    for _ in range(n_paths):
        length = rng.randint(0, max_len)
        p = []
        for _ in range(length):
            node = rng.randint(1, n_nodes)
            p.append(2 * node + (1 if rng.random() < reverse_prob else 0))
        paths.append(p)
    return paths


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("seed", range(6))
def test_random_graphs(seed, layout):
    # Random walks over few nodes: high outdegrees (sigma > 2 -> external edge lists), both orientations of
    # a node as successors of one record (the FlipSet special cases), empty paths, empty records.
    rng = random.Random(seed)
    paths = random_paths(rng, n_nodes=rng.choice([2, 3, 6, 12]), n_paths=rng.choice([3, 10, 40]), max_len=rng.choice([3, 8, 20]))
    paths = [p for p in paths] or [[2]]
    if not any(paths):
        paths.append([2, 4])
    b = gb.build_bwt(gb.bidirectional_sequences(paths))
    img = image_of(b)
    g = orc.GBWT.load(img)
    assert [g.sequence(2 * i) for i in range(len(paths))] == paths
    pc.check_everything(HostSim(img, layout), g)


@pytest.mark.parametrize("layout", [0, 1])
def test_long_runs_and_unidirectional(layout):
    # Identical paths -> long runs (escape bytes in the reference encoding, split runs in RUN8 / RUN32),
    # a unidirectional index, and bd_* on it are exercised at the ABI level (GPU tests).
    paths = [[3, 5, 7, 9]] * 700 + [[3, 6, 7, 9]] * 300 + [[3, 5, 8]] * 5 + [[4, 5, 7]]
    b = gb.build_bwt([list(p) for p in paths])
    img = image_of(b, bidirectional=False)
    g = orc.GBWT.load(img)
    e = HostSim(img, layout)
    pc.check_find_all_nodes(e, g)
    pc.check_find_extend_subpaths(e, g)
    pc.check_find_extend_random(e, g)
    pc.check_navigation(e, g)


def records_image(edges, runs, sequences, size, offset, bidirectional=False):
    g = orc.GBWT.from_records(edges, runs, sequences=sequences, size=size, offset=offset)
    return synth.gbwt_image(sequences, size, offset, len(edges) + offset, 4 | (1 if bidirectional else 0),
                            g.record_starts(), g.bwt_data()), g


@pytest.mark.parametrize("layout", [0, 1])
def test_paper_and_bidirectional_examples(layout):
    for edges, runs, bidir in [(gv.PAPER_EDGES, gv.PAPER_RUNS, False), (gv.BIDIR_EDGES, gv.BIDIR_RUNS, True)]:
        img, _ = records_image(edges, runs, sequences=3 if not bidir else 6, size=17, offset=0 if not bidir else 1,
                               bidirectional=bidir)
        g = orc.GBWT.load(img)
        pc.check_everything(HostSim(img, layout), g)


def wide_record_index(sigma, rng, run_len_bits):
    """One hub record with `sigma` successors (each a single-edge record back to the endmarker)."""
    succ = list(range(2, 2 + sigma))
    runs = []
    counts = [0] * sigma
    for _ in range(3 * sigma):
        v = rng.randrange(sigma)
        l = rng.getrandbits(rng.choice(run_len_bits)) + 1
        if runs and runs[-1][0] == v:
            continue
        runs.append((v, l)); counts[v] += l
    for v in range(sigma):
        if counts[v] == 0:
            runs.append((v, 1)); counts[v] = 1
    total = sum(counts)
    edges = [[(1, 0)], [(s, 0) for s in succ]]
    rr = [[(0, total)], runs]
    for v in range(sigma):
        edges.append([(0, 0)]); rr.append([(0, counts[v])])
    return edges, rr, total


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("sigma,bits", [(2, (1, 3)), (2, (9, 12)), (3, (1, 4, 10)), (7, (2, 8)), (64, (1, 5)), (200, (1, 3, 9)),
                                        (254, (1, 2)), (255, (1, 9)), (256, (1, 4)), (300, (1, 12)), (1000, (1, 3))])
def test_wide_records(sigma, bits, layout):
    # Every body format and the two-varint reference encoding (sigma >= 255, src/support.rs:1415-1417).
    rng = random.Random(sigma * 31 + len(bits))
    edges, runs, total = wide_record_index(sigma, rng, bits)
    img, gsrc = records_image(edges, runs, sequences=total, size=3 * total, offset=0)
    g = orc.GBWT.load(img)
    e = HostSim(img, layout)
    assert g.record_len(1) == total
    # follow from the hub for every successor over assorted ranges, and lf at assorted offsets
    st, nx = [], []
    cuts = sorted(set([0, 1, 2, total // 3, total // 2, total - 1, total, total + 5] + [rng.randrange(total + 1) for _ in range(12)]))
    for a in cuts:
        for b_ in cuts:
            if a < b_:
                for node in list(range(0, sigma + 4)):
                    st.append((1, a, b_)); nx.append(node)
    st = np.array(st, dtype=orc.STATE_DTYPE); nx = np.array(nx, dtype=np.uint64)
    assert pc.states_equal(e.extend(st, nx), g.extend_batch(st, nx))
    pos = np.array([(1, i) for i in sorted(set(cuts + [rng.randrange(total) for _ in range(300)]))], dtype=orc.POS_DTYPE)
    assert pc.states_equal(e.forward(pos), g.forward_batch(pos))
    pc.check_find_all_nodes(e, g)
    ids = np.array(sorted(set(rng.randrange(total) for _ in range(50))), dtype=np.uint64)
    assert np.array_equal(e.sequence_lengths(ids), g.sequence_lengths(ids))


@pytest.mark.parametrize("layout", [0, 1])
def test_bubble_chain(layout):
    for S, H, seed in [(50, 16, 1), (20, 300, 2), (9, 1100, 3)]:
        img = synth.bubble_chain(S, H, seed)
        g = orc.GBWT.load(img.array)
        e = HostSim(img.array, layout)
        pats = synth.patterns(S, H, seed, n=3000, k=min(32, 2 * S + 1))
        assert pc.states_equal(e.find_extend(pats), g.find_extend_batch(pats))
        pc.check_find_extend_random(e, g, n=2000, k=6, seed=S)
        ids = np.arange(2 * H, dtype=np.uint64)
        offsets, nodes, lengths = e.extract(ids)
        assert np.all(lengths == 2 * S + 1)
        for i in range(0, 2 * H, max(1, H // 8)):
            assert np.array_equal(nodes[int(offsets[i]):int(offsets[i + 1])], synth.sequence(S, H, seed, i))
        nodes_, offs, first, start, end = pc.bd_triples([[int(x) for x in synth.sequence(S, H, seed, 1)][:7]])
        assert pc.states_equal(e.bd_search(nodes_, offs, first, start, end), g.bd_search_batch(nodes_, offs, first, start, end))
        if layout == 0:
            assert e.format_counts()[2] > 0  # anchors are dense under the auto policy


def test_div_magic_is_exact():
    for sigma in range(1, 257):
        magic = 65536 // sigma + 1
        for b in range(256):
            assert (b * magic) >> 16 == b // sigma


def test_load_errors():
    data = bytearray(open(os.path.join(GOLDEN, "example.gbwt"), "rb").read())
    for mutate in (lambda d: d.__setitem__(0, d[0] ^ 0xFF), lambda d: d.__setitem__(4, 4),
                   lambda d: d.__setitem__(40, d[40] & ~4 & 0xFF), lambda d: d.__setitem__(40, d[40] & ~2 & 0xFF)):
        bad = bytearray(data)
        mutate(bad)
        with pytest.raises(IOError):
            HostSim(bytes(bad))
    with pytest.raises(IOError):
        HostSim(bytes(data[:200]))


def check_shortcuts(e, g):
    """K0 pass 3 (IndexView::skips) against the oracle's records: a shortcut over edge b of record u exists exactly
    when the successor is a single-edge record that does not end the path, and then LF(LF(u, i)) = (n_b, o_b + rank)."""
    skips = e.skips()
    offset = g.alphabet_offset()
    checked = 0
    for rec in range(e.records()):
        edges = g.record_edges(rec)
        if edges is None or len(edges) > 2:
            assert not skips[rec].any()
            continue
        for b, (node, off) in enumerate(edges):
            n_b, o_b = int(skips[rec][2 * b]), int(skips[rec][2 * b + 1])
            succ = g.record_edges(node - offset) if node != 0 else None
            if succ is not None and len(succ) == 1 and succ[0][0] != 0:
                assert (n_b, o_b) == (succ[0][0], off + succ[0][1])
                # spot check with LF: the first position of u that maps to edge b
                for i in range(g.record_len(rec)):
                    first = g.record_lf(rec, i)
                    if first is not None and first[0] == node:
                        second = g.record_lf(node - offset, first[1])
                        assert second == (n_b, o_b + (first[1] - off))
                        checked += 1
                        break
            else:
                assert (n_b, o_b) == (0, 0)
        for b in range(len(edges), 2):
            assert not skips[rec][2 * b:2 * b + 2].any()
    return checked


@pytest.mark.parametrize("layout", [0, 1])
def test_path_walk_shortcuts(layout):
    assert HostSim(open(os.path.join(GOLDEN, "example.gbz"), "rb").read(), layout).edges_valid()
    total = 0
    for name in FIXTURES:
        raw = open(os.path.join(GOLDEN, name), "rb").read()
        e = HostSim(raw, layout)
        assert e.edges_valid()
        total += check_shortcuts(e, orc.GBWT.load(raw))
    img = synth.bubble_chain(40, 9, 5)
    total += check_shortcuts(HostSim(img.array, layout), orc.GBWT.load(img.array))
    rng = random.Random(77)
    for _ in range(4):
        paths = random_paths(rng, n_nodes=rng.choice([3, 6, 12]), n_paths=rng.choice([3, 10, 40]), max_len=rng.choice([3, 8, 20]))
        img = image_of(gb.build_bwt(gb.bidirectional_sequences(paths + [[2, 4]])))
        total += check_shortcuts(HostSim(img, layout), orc.GBWT.load(img))
    assert total > 100


def test_invalid_edge_targets_are_flagged():
    # an edge to a node beyond the alphabet: queries still answer None like the reference, and the layout says so
    # (the extraction kernels then use their bounds-checked variant)
    img, _ = records_image([[(0, 0), (40, 0)], [(0, 0)], [(0, 0)]], [[(0, 1), (1, 1)], [(0, 1)], [(0, 1)]], sequences=2, size=4, offset=0)
    e = HostSim(img)
    assert not e.edges_valid() and not e.skips().any()


@pytest.mark.parametrize("layout", [0, 1])
def test_records_with_edges_but_no_runs(layout):
    # Only a damaged file has them (found by fuzzing the loader): Record::len() is 0, every query on them is None, and
    # nothing may read a body that does not exist.
    edges = [[(1, 0)], [(2, 0), (3, 0)], [(1, 0), (3, 0)], [(0, 0)]]
    runs = [[(0, 3)], [(0, 2), (1, 1)], [], [(0, 1)]]
    img, _ = records_image(edges, runs, sequences=3, size=8, offset=0, bidirectional=True)
    g, e = orc.GBWT.load(img), HostSim(img, layout)
    assert g.record_len(2) == 0 and e.record_format(2) not in (2,)   # never a dense body of zero blocks
    states = np.array([(2, 0, 2), (2, 0, 0), (2, 1, 5)], dtype=orc.STATE_DTYPE)
    for node in (1, 3, 0, 7):
        nodes = np.full(len(states), node, dtype=np.uint64)
        assert pc.states_equal(e.extend(states, nodes), g.extend_batch(states, nodes))
    bd = np.zeros(2, dtype=e.bd_find(np.array([1], dtype=np.uint64)).dtype)
    bd["forward"]["node"] = 3; bd["forward"]["end"] = 2; bd["reverse"]["node"] = 2; bd["reverse"]["end"] = 2
    for backward in (False, True):
        offsets, want, counts = g.follow_batch(bd, backward=backward)
        got_offsets, got, got_counts = e.follow(bd, backward)
        assert np.array_equal(got_counts, counts) and np.array_equal(got_offsets, offsets) and pc.states_equal(got, want)
    # find() on the record without runs: the reference returns Some(empty range), which the ABI folds into None
    nodes = np.arange(g.alphabet_size() + 2, dtype=np.uint64)
    want = g.find_batch(nodes).copy()
    want[want["end"] <= want["start"]] = (0, 0, 0)
    assert pc.states_equal(e.find(nodes), want)
    pos = np.array([(2, 0), (2, 1), (1, 0), (1, 2)], dtype=orc.POS_DTYPE)
    assert pc.states_equal(e.forward(pos), g.forward_batch(pos))
