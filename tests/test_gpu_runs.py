"""Checkpointed run-length bodies on the GPU, through the C ABI, against the oracle (the same checks as
tests/test_run_checkpoints.py runs on the CPU through tests/hostsim), plus the run-length benchmark workload at test
size: low-frequency alleles and tri-allelic sites, every kernel that can meet a run body."""
import os
import random

import numpy as np
import pytest

import gbwt_builder as gb
import parity_checks as pc
from oracle import oracle as orc
from synth import synth
from test_hostsim_layout import FIXTURES, GOLDEN, image_of, random_paths, records_image, wide_record_index

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b200():
    import gbwt_rs_b200
    return gbwt_rs_b200


@pytest.fixture
def forced(monkeypatch):
    monkeypatch.setenv("GBWT_B200_CKPT_MIN_RUNS", "1")
    monkeypatch.setenv("GBWT_B200_CKPT_INTERVAL_RUNS", "2")


@pytest.mark.parametrize("name", FIXTURES)
def test_fixtures_with_forced_checkpoints(b200, name, forced):
    raw = open(os.path.join(GOLDEN, name), "rb").read()
    g, e = orc.GBWT.load(raw), b200.GBWT.from_bytes(raw, layout="runs")
    assert e.device_bytes()["records_run_checkpointed"] > 0
    pc.check_everything(e, g)


@pytest.mark.parametrize("seed", range(4))
def test_random_graphs_with_forced_checkpoints(b200, seed, forced):
    rng = random.Random(100 + seed)
    paths = random_paths(rng, n_nodes=rng.choice([2, 3, 6]), n_paths=rng.choice([10, 40, 120]), max_len=rng.choice([8, 20]))
    if not any(paths):
        paths.append([2, 4])
    img = image_of(gb.build_bwt(gb.bidirectional_sequences(paths)))
    g, e = orc.GBWT.load(img), b200.GBWT.from_bytes(img, layout="runs")
    assert e.device_bytes()["records_run_checkpointed"] > 0
    pc.check_everything(e, g)


@pytest.mark.parametrize("sigma,bits", [(2, (9, 12)), (3, (1, 4, 10)), (4, (1, 2)), (7, (2, 8)), (64, (1, 5)), (65, (1, 5)), (300, (1, 12))])
def test_wide_records_default_thresholds(b200, sigma, bits):
    rng = random.Random(sigma * 131 + len(bits))
    runs = []
    while len(runs) <= 60 * max(1, sigma // 4):
        v = rng.randrange(sigma)
        if not runs or runs[-1][0] != v:
            runs.append((v, rng.getrandbits(rng.choice(bits)) + 1))
    for v in range(sigma):
        runs.append((v, 1))
    counts = [0] * sigma
    for v, l in runs:
        counts[v] += l
    total = sum(counts)
    edges = [[(1, 0)], [(s, 0) for s in range(2, 2 + sigma)]] + [[(0, 0)] for _ in range(sigma)]
    rr = [[(0, total)], runs] + [[(0, counts[v])] for v in range(sigma)]
    img, _ = records_image(edges, rr, sequences=total, size=3 * total, offset=0)
    g, e = orc.GBWT.load(img), b200.GBWT.from_bytes(img, layout="runs")
    assert (e.device_bytes()["records_run_checkpointed"] > 0) == (sigma <= 64)
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


@pytest.mark.parametrize("layout", ["auto", "runs"])
@pytest.mark.parametrize("lean", [None, "0", "3", "window"])
def test_run_length_workload(b200, layout, lean, monkeypatch):
    """The benchmark's run-length workload (bench.py --workload find-runs) at test size, through the default dispatch,
    the general kernels (GBWT_B200_FIND_LEAN=0), the lean loop with the out-of-line step (=3), and with the record-window
    kernel forced (wide records: DENSE4 and byte-per-run bodies answered from shared memory)."""
    if lean == "window":
        monkeypatch.setenv("GBWT_B200_FIND_WINDOW", "2")
    elif lean is not None:
        monkeypatch.setenv("GBWT_B200_FIND_LEAN", lean)
    monkeypatch.setenv("GBWT_B200_LOCALITY", "1")
    S, H, seed, ppm, tri = 3000, 1024, 42, 50_000, 10
    img = synth.bubble_chain(S, H, seed, alt_ppm=ppm, tri_mod=tri)
    g, e = orc.GBWT.load(img.array), b200.GBWT.from_bytes(img.array, layout=layout)
    stats = e.device_bytes()
    if layout == "runs":
        assert stats["records_run8"] + stats["records_run32"] > 0 and stats["records_run_checkpointed"] > 0
    else:
        assert stats["records_dense4"] > 0
    pats = synth.patterns(S, H, seed, n=100_000, k=32, alt_ppm=ppm, tri_mod=tri)
    got = e.find_extend(pats)
    assert pc.states_equal(got, g.find_extend_batch(pats))
    assert np.all(got["end"] > got["start"])
    assert pc.states_equal(e.find_extend_u32(pats.astype(np.uint32)), got)
    pc.check_find_extend_random(e, g, n=5000, k=6, seed=S)
    ids = np.arange(0, 2 * H, 61, dtype=np.uint64)
    offsets, nodes, lengths = e.extract(ids)
    o_off, o_nodes = g.extract_batch(ids)
    assert np.array_equal(offsets, o_off) and np.array_equal(nodes, o_nodes)
    seq = [int(x) for x in synth.sequence(S, H, seed, 1, ppm, tri)][:40]
    nodes_, offs, first, start, end = pc.bd_triples([seq[:12]])
    assert pc.states_equal(e.bd_search(nodes_, offs, first, start, end), g.bd_search_batch(nodes_, offs, first, start, end))


@pytest.mark.parametrize("seed", range(6))
def test_window_kernel_on_wide_records(b200, seed, monkeypatch):
    """Graphs whose records mostly have three or four successors in both orientations, record-window kernel forced, with
    and without forced checkpoint tables on the byte-per-run bodies: find/extend of every subpath against the oracle."""
    from test_run_checkpoints import sparse_paths
    monkeypatch.setenv("GBWT_B200_FIND_WINDOW", "2")
    monkeypatch.setenv("GBWT_B200_LOCALITY", "1")
    if seed % 2 == 1:
        monkeypatch.setenv("GBWT_B200_CKPT_MIN_RUNS", "1")
        monkeypatch.setenv("GBWT_B200_CKPT_INTERVAL_RUNS", "2")
    rng = random.Random(700 + seed)
    paths = sparse_paths(rng, rng.choice([6, 10, 40]), rng.choice([150, 400, 900]))
    img = image_of(gb.build_bwt(gb.bidirectional_sequences(paths)))
    g, e = orc.GBWT.load(img), b200.GBWT.from_bytes(img)
    w0 = e.window_info()
    pc.check_find_extend_subpaths(e, g)
    pc.check_find_extend_random(e, g, n=20000, k=7, seed=seed)
    seqs = [g.sequence(i) for i in range(g.sequences())]
    for k in (2, 3, 5, 16, 17, 33):
        pats = pc.subpath_patterns(seqs, k)
        if pats:
            pats = np.array(pats, dtype=np.uint64)
            assert pc.states_equal(e.find_extend(pats), g.find_extend_batch(pats))
            assert pc.states_equal(e.find_extend_u32(pats.astype(np.uint32)), g.find_extend_batch(pats))
