"""Full-size parity (BASELINE.json configs[3] and [4]): the 10 M-node x 1024-haplotype index is built once and the CUDA
paths are compared with the CPU oracle on >= 1 M sampled find/extend queries (64- and 32-bit patterns), >= 100 k
bidirectional searches, and 16 whole haplotype paths extracted every way the library can (checkpointed segments, one
chain per path, two chains per path), plus the size-independent properties on everything that was computed."""
import numpy as np
import pytest

import parity_checks as pc
from oracle import oracle as orc
from synth import synth

pytestmark = pytest.mark.gpu

S, H, SEED = 3_333_333, 1024, 42


@pytest.fixture(scope="module")
def world():
    import gbwt_rs_b200 as b200
    img = synth.bubble_chain(S, H, SEED)
    engine = b200.GBWT.from_bytes(img.array, checkpoints=True)
    oracle = orc.GBWT.load(img.array, native=True)
    yield b200, img, engine, oracle
    engine.close()


def test_full_size_find_extend_against_oracle(world):
    b200, img, e, g = world
    assert e.window_info()["default"] == 1 and e.device_bytes()["records_dense"] > 6_000_000
    n = 1 << 20
    pats = synth.patterns(S, H, SEED, n=n, k=32, seed_q=7, q0=123_456_789)
    want = g.find_extend_batch(pats, threads=orc.max_threads())
    got = e.find_extend(pats)
    assert pc.states_equal(got, want)
    assert np.all(got["end"] > got["start"]) and np.array_equal(got["node"], pats[:, -1])
    assert pc.states_equal(e.find_extend_u32(pats.astype(np.uint32)), want)
    # damaged patterns at full size: the other strand of one node, a node that is no node
    rng = np.random.default_rng(5)
    bad = pats[: 1 << 18].copy()
    rows, cols = np.arange(len(bad)), rng.integers(0, 32, len(bad))
    bad[rows, cols] ^= np.uint64(1)
    bad[::7, 5] = np.uint64(2**40)
    assert pc.states_equal(e.find_extend(bad), g.find_extend_batch(bad, threads=orc.max_threads()))


def test_full_size_bd_search_against_oracle(world):
    b200, img, e, g = world
    n = 120_000
    pats = synth.patterns(S, H, SEED, n=n, k=32, seed_q=11)
    rng = np.random.default_rng(3)
    first = rng.integers(0, 32, size=n).astype(np.uint64)
    start = (first * rng.random(n)).astype(np.uint64)
    end = (first + 1 + ((32 - first - 1) * rng.random(n)).astype(np.uint64)).astype(np.uint64)
    offs = np.arange(n + 1, dtype=np.uint64) * 32
    flat = pats.reshape(-1)
    got = e.bd_search(flat, offs, first, start, end)
    assert pc.states_equal(got, g.bd_search_batch(flat, offs, first, start, end))
    assert np.all(got["forward"]["end"] > got["forward"]["start"])
    assert np.array_equal(got["forward"]["end"] - got["forward"]["start"], got["reverse"]["end"] - got["reverse"]["start"])


def test_full_size_extraction_against_oracle(world, monkeypatch):
    b200, img, e, g = world
    ids = np.array(sorted(set(int(x) for x in np.linspace(0, 2 * H - 1, 16))), dtype=np.uint64)   # both strands
    o_off, o_nodes = g.extract_batch(ids, threads=len(ids))
    assert np.all(np.diff(o_off.astype(np.int64)) == 2 * S + 1)
    offsets, nodes, lengths = e.extract(ids)                       # checkpointed segments, one lane each from global memory
    assert e.checkpoint_info()["present"]
    assert np.array_equal(offsets, o_off) and np.array_equal(nodes, o_nodes)
    monkeypatch.setenv("GBWT_B200_EXTRACT_WINDOW", "2")            # checkpointed segments from staged record windows
    offsets, nodes, lengths = e.extract(ids)
    assert np.array_equal(offsets, o_off) and np.array_equal(nodes, o_nodes)
    # ... and the way a large batch takes by default (256 sequences: one CTA per segment): identical to the one-lane walks,
    # four of the paths compared with the oracle
    many = np.arange(0, 2 * H, 8, dtype=np.uint64)
    monkeypatch.delenv("GBWT_B200_EXTRACT_WINDOW")
    offs_w, nodes_w, _ = e.extract(many)
    monkeypatch.setenv("GBWT_B200_EXTRACT_WINDOW", "0")
    offs_0, nodes_0, _ = e.extract(many)
    monkeypatch.delenv("GBWT_B200_EXTRACT_WINDOW")
    assert np.array_equal(offs_w, offs_0) and np.array_equal(nodes_w, nodes_0)
    del nodes_0
    picks = np.array([0, 85, 170, 255])
    p_off, p_nodes = g.extract_batch(many[picks], threads=4)
    for t, k in enumerate(picks):
        assert np.array_equal(nodes_w[int(offs_w[k]):int(offs_w[k + 1])], p_nodes[int(p_off[t]):int(p_off[t + 1])])
    del nodes_w
    monkeypatch.setenv("GBWT_B200_EXTRACT_CHECKPOINTS", "0")       # two chains per path (lengths are known to the index)
    offsets, nodes, lengths = e.extract(ids)
    assert np.array_equal(offsets, o_off) and np.array_equal(nodes, o_nodes)
    monkeypatch.setenv("GBWT_B200_EXTRACT_SPLIT", "0")             # one chain per path
    offsets, nodes, lengths = e.extract(ids)
    assert np.array_equal(offsets, o_off) and np.array_equal(nodes, o_nodes)
    for j, i in enumerate(ids[:3]):
        assert np.array_equal(nodes[int(offsets[j]):int(offsets[j + 1])], synth.sequence(S, H, SEED, int(i)))


def test_full_size_run_length_workload_against_oracle():
    """The run-length workload of bench.py (10 M node ids, alternative alleles of frequency 0.05, every tenth site
    tri-allelic) at full size: DENSE4 anchors and wide records in the window kernel, 1 M queries against the oracle."""
    import gbwt_rs_b200 as b200
    sites, ppm, tri = 2_500_000, 50_000, 10
    img = synth.bubble_chain(sites, H, SEED, alt_ppm=ppm, tri_mod=tri)
    e = b200.GBWT.from_bytes(img.array, checkpoints=False)
    g = orc.GBWT.load(img.array, native=True)
    stats = e.device_bytes()
    assert stats["records_dense4"] > 400_000 and e.window_info()["default"] == 1
    n = 1 << 20
    pats = synth.patterns(sites, H, SEED, n=n, k=32, seed_q=7, q0=987_654_321, alt_ppm=ppm, tri_mod=tri)
    want = g.find_extend_batch(pats, threads=orc.max_threads())
    got = e.find_extend(pats)
    assert pc.states_equal(got, want)
    assert np.all(got["end"] > got["start"]) and np.array_equal(got["node"], pats[:, -1])
    assert pc.states_equal(e.find_extend_u32(pats.astype(np.uint32)), want)
    bad = pats[: 1 << 18].copy()
    rng = np.random.default_rng(6)
    bad[np.arange(len(bad)), rng.integers(0, 32, len(bad))] ^= np.uint64(1)
    assert pc.states_equal(e.find_extend(bad), g.find_extend_batch(bad, threads=orc.max_threads()))
    ids = np.array([0, 1, 2 * H - 1], dtype=np.uint64)
    offsets, nodes, lengths = e.extract(ids)
    o_off, o_nodes = g.extract_batch(ids, threads=3)
    assert np.array_equal(offsets, o_off) and np.array_equal(nodes, o_nodes)
    e.close()
