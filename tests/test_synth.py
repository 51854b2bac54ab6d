"""Pins the test-only generators.

1. tests/gbwt_builder.py (brute-force GBWT definition) reproduces the BWT bytes and record starts of the
   reference's C++-built fixtures exactly -> pins encoder, ordering rule and edge offsets together.
2. synth/gbwt_synth.c's Simple-SDS writer reproduces the serialized BWT section of the fixtures.
3. The closed-form bubble-chain generator equals the brute-force builder on random small instances,
   including multi-chunk instances, and its images load in the oracle with the right paths / queries.
CPU only.
"""
import os

import numpy as np
import pytest

import gbwt_builder as gb
import golden_vectors as gv
from oracle import oracle as orc
from synth import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name,paths", [
    ("example.gbwt", gv.true_paths(False)),
    ("with-empty.gbwt", gv.true_paths(True)),
    ("translation.gbwt", gv.TRANSLATION_PATHS),
])
def test_builder_reproduces_fixture(name, paths):
    g = orc.GBWT.load(os.path.join(GOLDEN, name))
    b = gb.build_bwt(gb.bidirectional_sequences(paths))
    assert b["data"] == g.bwt_data()
    assert b["starts"] == [int(x) for x in g.record_starts()]
    assert (b["offset"], b["alphabet_size"], b["sequences"], b["size"]) == \
           (g.alphabet_offset(), g.alphabet_size(), g.sequences(), g.len())


@pytest.mark.parametrize("name", ["example.gbwt", "with-empty.gbwt", "translation.gbwt"])
def test_writer_reproduces_fixture_bwt_section(name):
    raw = open(os.path.join(GOLDEN, name), "rb").read()
    g = orc.GBWT.load(raw)
    section = synth.bwt_section(g.record_starts(), g.bwt_data())
    assert raw.find(section) > 0  # the serialized SparseVector + Vec<u8> appear verbatim in the file
    # and a full re-assembled image loads to the same index
    img = synth.gbwt_image(g.sequences(), g.len(), g.alphabet_offset(), g.alphabet_size(), g.flags() & ~2,
                           g.record_starts(), g.bwt_data())
    g2 = orc.GBWT.load(img)
    assert g2.bwt_data() == g.bwt_data() and np.array_equal(g2.record_starts(), g.record_starts())
    assert [g2.sequence(i) for i in range(g2.sequences())] == [g.sequence(i) for i in range(g.sequences())]


def chain_paths(S, H, seed):
    return [[int(x) for x in synth.sequence(S, H, seed, 2 * h)] for h in range(H)]


@pytest.mark.parametrize("S,H,seed", [(1, 1, 0), (1, 3, 1), (5, 2, 2), (7, 9, 3), (12, 33, 42), (40, 5, 5), (3, 300, 6)])
def test_bubble_chain_matches_brute_force(S, H, seed):
    paths = chain_paths(S, H, seed)
    for h, p in enumerate(paths):
        assert len(p) == 2 * S + 1 and p[0] == 2 and p[-1] == 2 * (3 * S + 1)
        assert [int(x) for x in synth.sequence(S, H, seed, 2 * h + 1)] == gv.reverse_path(p)
    b = gb.build_bwt(gb.bidirectional_sequences(paths))
    img = synth.bubble_chain(S, H, seed)
    g = orc.GBWT.load(img.array)
    assert g.bwt_data() == b["data"]
    assert [int(x) for x in g.record_starts()] == b["starts"]
    assert (g.alphabet_offset(), g.alphabet_size(), g.sequences(), g.len()) == \
           (b["offset"], b["alphabet_size"], b["sequences"], b["size"])
    assert g.is_bidirectional()
    for i in range(2 * H):
        assert g.sequence(i) == [int(x) for x in synth.sequence(S, H, seed, i)]


def test_bubble_chain_chunk_boundaries_are_exact():
    # 2048 sites per generator chunk: S = 5000 spans three chunks per strand. With H = 3 almost every
    # window of 64 alleles ties, which exercises the exact tie-breaking path of order_at().
    for S, H, seed in [(5000, 3, 11), (4100, 40, 12)]:
        img = synth.bubble_chain(S, H, seed, threads=4)
        one = synth.bubble_chain(S, H, seed, threads=1)
        assert img.tobytes() == one.tobytes()
        g = orc.GBWT.load(img.array)
        ids = np.arange(2 * H, dtype=np.uint64)
        offs, nodes = g.extract_batch(ids)
        for i in range(2 * H):
            assert np.array_equal(nodes[int(offs[i]):int(offs[i + 1])], synth.sequence(S, H, seed, i))


def test_patterns_are_subpaths_and_match():
    S, H, seed = 300, 16, 42
    img = synth.bubble_chain(S, H, seed)
    g = orc.GBWT.load(img.array)
    pats = synth.patterns(S, H, seed, n=2000, k=32)
    paths = [synth.sequence(S, H, seed, i) for i in range(2 * H)]
    for q in range(0, 2000, 97):
        h = synth.lib().synth_mix64(7 + 3 * q) % H
        o = synth.lib().synth_mix64(7 + 3 * q + 1) & 1
        t = synth.lib().synth_mix64(7 + 3 * q + 2) % (2 * S + 1 - 31)
        assert np.array_equal(pats[q], paths[2 * h + o][t:t + 32])
    out = g.find_extend_batch(pats)
    assert np.all(out["end"] > out["start"])          # every sampled pattern occurs
    assert np.array_equal(out["node"], pats[:, -1])
    # occurrence counts against brute force on a few
    for q in range(0, 2000, 211):
        p = list(pats[q])
        count = sum(1 for path in paths for i in range(len(path) - 31) if list(path[i:i + 32]) == p)
        assert int(out["end"][q] - out["start"][q]) == count
    # q0 offsetting is consistent (used for sharding across ranks)
    part = synth.patterns(S, H, seed, n=500, k=32, q0=1000)
    assert np.array_equal(part, pats[1000:1500])
