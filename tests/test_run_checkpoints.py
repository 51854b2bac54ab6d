"""Checkpointed run-length bodies (layout.h, record_scan.cuh: load_checkpoint / scan_runs_to): a rank on a long RUN8 /
RUN32 / RUN64 body is one table entry plus a scan of one interval instead of the reference's scan from the start of
the record (src/bwt.rs:603-613 over RLEIter, src/support.rs:1413-1430). The results must not change: everything is
compared with the oracle, with the tables forced onto short bodies (knobs read by the layout builder) and at their
default thresholds. CPU only (tests/hostsim runs the product's one-lane code); the GPU runs the same checks in
tests/test_gpu_runs.py."""
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
from test_hostsim_layout import FIXTURES, GOLDEN, image_of, random_paths, records_image, wide_record_index


@pytest.fixture
def forced(monkeypatch):
    """Checkpoints on every run body with more than one run, an interval of about two runs."""
    monkeypatch.setenv("GBWT_B200_CKPT_MIN_RUNS", "1")
    monkeypatch.setenv("GBWT_B200_CKPT_INTERVAL_RUNS", "2")


@pytest.mark.parametrize("name", FIXTURES)
def test_fixtures_with_forced_checkpoints(name, forced):
    raw = open(os.path.join(GOLDEN, name), "rb").read()
    g = orc.GBWT.load(raw)
    e = HostSim(raw, 1)  # runs-only policy: every record with two or more edges has a run body
    assert e.checkpointed_records() > 0
    pc.check_everything(e, g)
    # the tables are not part of the index: the file written back is the one that was read
    assert HostSim(e.serialize(), 1).serialize() == e.serialize()


@pytest.mark.parametrize("seed", range(6))
def test_random_graphs_with_forced_checkpoints(seed, forced):
    rng = random.Random(100 + seed)
    paths = random_paths(rng, n_nodes=rng.choice([2, 3, 6]), n_paths=rng.choice([10, 40, 120]), max_len=rng.choice([8, 20]))
    if not any(paths):
        paths.append([2, 4])
    b = gb.build_bwt(gb.bidirectional_sequences(paths))
    img = image_of(b)
    g = orc.GBWT.load(img)
    e = HostSim(img, 1)
    assert e.checkpointed_records() > 0
    pc.check_everything(e, g)


@pytest.mark.parametrize("knobs", [("1", "2"), ("1", "7"), (None, None)])
@pytest.mark.parametrize("sigma,bits", [(2, (1, 3)), (2, (9, 12)), (3, (1, 4, 10)), (4, (1, 2)), (7, (2, 8)), (64, (1, 5)), (65, (1, 5)),
                                        (200, (1, 3, 9)), (300, (1, 12))])
def test_wide_records(sigma, bits, knobs, monkeypatch):
    """One hub record per format (RUN8, RUN32 for long runs, RUN64 above 256 edges; tables up to 64 edges): follow over
    every pair of cut positions for every successor, lf at assorted offsets."""
    if knobs[0] is not None:
        monkeypatch.setenv("GBWT_B200_CKPT_MIN_RUNS", knobs[0])
        monkeypatch.setenv("GBWT_B200_CKPT_INTERVAL_RUNS", knobs[1])
    rng = random.Random(sigma * 131 + len(bits))
    edges, runs_all, total = wide_record_index(sigma, rng, bits)
    if knobs[0] is None:
        # default thresholds: make the hub long enough to get a table (well over 32 runs per 16-byte unit of table entry)
        runs = list(runs_all[1])
        while len(runs) <= 60 * max(1, sigma // 4):
            v = rng.randrange(sigma)
            if runs[-1][0] != v:
                runs.append((v, rng.getrandbits(rng.choice(bits)) + 1))
        counts = [0] * sigma
        for v, l in runs:
            counts[v] += l
        total = sum(counts)
        edges = [[(1, 0)], [(s, 0) for s in range(2, 2 + sigma)]] + [[(0, 0)] for _ in range(sigma)]
        runs_all = [[(0, total)], runs] + [[(0, counts[v])] for v in range(sigma)]
    img, _ = records_image(edges, runs_all, sequences=total, size=3 * total, offset=0)
    g = orc.GBWT.load(img)
    e = HostSim(img, 1)
    if sigma > 64:
        assert e.checkpointed_records() == 0
    elif knobs[1] != "7":
        assert e.checkpointed_records() > 0
    st, nx = [], []
    cuts = sorted(set([0, 1, 2, total // 3, total // 2, total - 1, total, total + 5] + [rng.randrange(total + 1) for _ in range(14)]))
    for a in cuts:
        for b_ in cuts:
            if a < b_:
                for node in range(0, sigma + 4):
                    st.append((1, a, b_)); nx.append(node)
    st = np.array(st, dtype=orc.STATE_DTYPE); nx = np.array(nx, dtype=np.uint64)
    assert pc.states_equal(e.extend(st, nx), g.extend_batch(st, nx))
    pos = np.array([(1, i) for i in sorted(set(cuts + [rng.randrange(total) for _ in range(400)]))], dtype=orc.POS_DTYPE)
    assert pc.states_equal(e.forward(pos), g.forward_batch(pos))
    pc.check_find_all_nodes(e, g)


def test_bidirectional_hub_with_forced_checkpoints(forced):
    """bd_follow's flipped count (src/bwt.rs:646-648) from the tables: both orientations of a node among the successors."""
    rng = random.Random(5)
    paths = []
    for _ in range(300):
        a = rng.choice([4, 5, 6, 7, 8, 9, 10])
        paths.append([2, a, rng.choice([12, 13, 14])])
    b = gb.build_bwt(gb.bidirectional_sequences(paths))
    img = image_of(b)
    g = orc.GBWT.load(img)
    e = HostSim(img, 1)
    assert e.checkpointed_records() > 0
    pc.check_everything(e, g)


@pytest.mark.parametrize("layout", [0, 1])
def test_variant_bubble_chain_default_thresholds(layout):
    """Low-frequency alleles and tri-allelic sites (synth variant model): anchors with three edges keep run bodies under
    both policies, and with 1100 haplotypes those bodies are long enough for a table at the default thresholds."""
    S, H, seed, ppm, tri = 40, 1100, 3, 50_000, 3
    img = synth.bubble_chain(S, H, seed, alt_ppm=ppm, tri_mod=tri)
    g = orc.GBWT.load(img.array)
    e = HostSim(img.array, layout)
    # (under the default policy those anchors are dense, two bits per position; the run bodies are the runs-only policy's)
    assert e.checkpointed_records() > 0 if layout == 1 else e.format_counts()[6] > 0
    pats = synth.patterns(S, H, seed, n=4000, k=32, alt_ppm=ppm, tri_mod=tri)
    assert pc.states_equal(e.find_extend(pats), g.find_extend_batch(pats))
    pc.check_find_extend_random(e, g, n=2000, k=6, seed=S)
    ids = np.arange(0, 2 * H, 37, dtype=np.uint64)
    offsets, nodes, lengths = e.extract(ids)
    assert np.all(lengths == 2 * S + 1)
    for j, i in enumerate(ids):
        assert np.array_equal(nodes[int(offsets[j]):int(offsets[j + 1])], synth.sequence(S, H, seed, int(i), ppm, tri))
    nodes_, offs, first, start, end = pc.bd_triples([[int(x) for x in synth.sequence(S, H, seed, 1, ppm, tri)][:9]])
    assert pc.states_equal(e.bd_search(nodes_, offs, first, start, end), g.bd_search_batch(nodes_, offs, first, start, end))


# ---- FMT_DENSE4: records with three or four edges as two bits per position -------------------------------------------

@pytest.mark.parametrize("sigma,bits", [(3, (1, 2)), (3, (1, 2, 3)), (4, (1, 2)), (4, (1, 3))])
def test_dense4_hub(sigma, bits):
    rng = random.Random(sigma * 977 + len(bits))
    edges, runs_all, total = wide_record_index(sigma, rng, bits)
    runs = list(runs_all[1])
    while len(runs) < 400:  # several blocks of 64 positions, and past the sizes where runs would be smaller
        v = rng.randrange(sigma)
        if runs[-1][0] != v:
            runs.append((v, rng.getrandbits(rng.choice(bits)) + 1))
    counts = [0] * sigma
    for v, l in runs:
        counts[v] += l
    total = sum(counts)
    edges = [[(1, 0)], [(s, 0) for s in range(2, 2 + sigma)]] + [[(0, 0)] for _ in range(sigma)]
    rr = [[(0, total)], runs] + [[(0, counts[v])] for v in range(sigma)]
    img, _ = records_image(edges, rr, sequences=total, size=3 * total, offset=0)
    g = orc.GBWT.load(img)
    e = HostSim(img, 0)
    assert e.format_counts()[6] == 1 and e.record_format(1) == 6
    st, nx = [], []
    cuts = sorted(set([0, 1, 2, 63, 64, 65, 127, 128, total // 3, total // 2, total - 1, total, total + 5] + [rng.randrange(total + 1) for _ in range(14)]))
    for a in cuts:
        for b_ in cuts:
            if a < b_:
                for node in range(0, sigma + 4):
                    st.append((1, a, b_)); nx.append(node)
    st = np.array(st, dtype=orc.STATE_DTYPE); nx = np.array(nx, dtype=np.uint64)
    assert pc.states_equal(e.extend(st, nx), g.extend_batch(st, nx))
    pos = np.array([(1, i) for i in sorted(set(cuts + [rng.randrange(total) for _ in range(500)]))], dtype=orc.POS_DTYPE)
    assert pc.states_equal(e.forward(pos), g.forward_batch(pos))
    pc.check_find_all_nodes(e, g)
    # written back: the same record bytes as the reference encoding of the same runs
    assert HostSim(e.serialize(), 1).serialize() == e.serialize()
    assert orc.GBWT.load(e.serialize()).bwt_data() == g.bwt_data()


def sparse_paths(rng, n_nodes, n_paths):
    """Walks towards higher node identifiers with three kinds of steps (next node forward, next node reversed, the node
    after next): most records have three or four successors, both orientations of a node among them."""
    paths = []
    for _ in range(n_paths):
        v = rng.randint(1, 2)
        p = [2 * v]
        while True:
            c = rng.random()
            v, o = (v + 1, 0) if c < 0.5 else ((v + 1, 1) if c < 0.75 else (v + 2, 0))
            if v > n_nodes:
                break
            p.append(2 * v + o)
        paths.append(p)
    return paths


@pytest.mark.parametrize("seed", range(8))
def test_dense4_random_bidirectional_graphs(seed):
    """Records with three and four successors, both orientations of a node among them (the FlipSet cases of bd_follow),
    become DENSE4 under the default policy; everything against the oracle, backward navigation included."""
    rng = random.Random(500 + seed)
    paths = sparse_paths(rng, rng.choice([6, 10]), rng.choice([150, 400, 900]))
    img = image_of(gb.build_bwt(gb.bidirectional_sequences(paths)))
    g = orc.GBWT.load(img)
    e = HostSim(img, 0)
    assert e.format_counts()[6] > 0
    pc.check_everything(e, g)
    assert orc.GBWT.load(e.serialize()).bwt_data() == g.bwt_data()


def test_variant_bubble_chain_is_dense4_by_default():
    S, H, seed, ppm, tri = 30, 1024, 9, 50_000, 2
    img = synth.bubble_chain(S, H, seed, alt_ppm=ppm, tri_mod=tri)
    e = HostSim(img.array, 0)
    counts = e.format_counts()
    assert counts[6] > 0  # tri-allelic anchors
    g = orc.GBWT.load(img.array)
    pats = synth.patterns(S, H, seed, n=3000, k=32, alt_ppm=ppm, tri_mod=tri)
    assert pc.states_equal(e.find_extend(pats), g.find_extend_batch(pats))
