"""The way back: the device layout written out as a Simple-SDS GBWT image
(gbwt-rs_b200/csrc/layout_writer.cpp, run on the CPU through tests/hostsim). The image must load in the oracle
(= the reference's loader restated), give the same records, and -- for inputs whose runs are maximal, which is what
the reference's builders write -- reproduce the BWT section byte for byte (serialize::test of src/bwt/tests.rs:298-340
and src/gbwt/tests.rs:72-84 checks exactly that round trip)."""
import os
import random

import numpy as np
import pytest

import gbwt_builder as gb
import parity_checks as pc
from hostsim_build import HostSim
from oracle import oracle as orc
from synth import synth
from test_hostsim_layout import FIXTURES, image_of, random_paths, records_image, wide_record_index

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def same_index(a, b):
    assert (a.len(), a.sequences(), a.alphabet_size(), a.alphabet_offset(), a.is_bidirectional()) == \
           (b.len(), b.sequences(), b.alphabet_size(), b.alphabet_offset(), b.is_bidirectional())
    for rec in range(a.alphabet_size() - a.alphabet_offset()):
        assert a.record_edges(rec) == b.record_edges(rec)
        assert a.record_decompress(rec) == b.record_decompress(rec)


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("name", FIXTURES)
def test_fixtures_round_trip(name, layout):
    raw = open(os.path.join(GOLDEN, name), "rb").read()
    image = HostSim(raw, layout).serialize()
    src, back = orc.GBWT.load(raw), orc.GBWT.load(image)
    same_index(src, back)
    # files written by the reference have maximal runs: the BWT comes back byte for byte
    assert back.bwt_data() == src.bwt_data() and np.array_equal(back.record_starts(), src.record_starts())
    assert not (back.flags() & 2) and (back.flags() & 4)            # no metadata, Simple-SDS
    pc.check_everything(HostSim(image, layout), src)                 # and the product loads its own output
    assert HostSim(image, layout).serialize() == image               # idempotent


@pytest.mark.parametrize("layout", [0, 1])
def test_bubble_chain_round_trip(layout):
    for S, H, seed in [(60, 9, 1), (15, 700, 2)]:
        img = synth.bubble_chain(S, H, seed)
        back = HostSim(img.array, layout).serialize()
        src, out = orc.GBWT.load(img.array), orc.GBWT.load(back)
        assert out.bwt_data() == src.bwt_data() and np.array_equal(out.record_starts(), src.record_starts())
        assert (out.len(), out.sequences(), out.alphabet_size(), out.alphabet_offset(), out.flags()) == \
               (src.len(), src.sequences(), src.alphabet_size(), src.alphabet_offset(), src.flags())
        assert len(back) == img.nbytes       # same container, same size (the packed `source` tag value differs)


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("sigma,bits", [(2, (9, 12)), (3, (1, 4, 10)), (64, (1, 5)), (200, (1, 3, 9)), (254, (1, 2)), (255, (1, 9)),
                                        (256, (1, 4)), (300, (1, 12))])
def test_wide_records_round_trip(sigma, bits, layout):
    # every body format, escape-coded long runs (len >= 256 / sigma) and the two-varint encoding (sigma >= 255)
    rng = random.Random(sigma * 7 + len(bits))
    edges, runs, total = wide_record_index(sigma, rng, bits)
    img, _ = records_image(edges, runs, sequences=total, size=3 * total, offset=0)
    back = HostSim(img, layout).serialize()
    src, out = orc.GBWT.load(img), orc.GBWT.load(back)
    same_index(src, out)
    assert out.bwt_data() == src.bwt_data()   # wide_record_index never repeats a value in adjacent runs


@pytest.mark.parametrize("seed", range(4))
def test_random_graphs_round_trip(seed):
    rng = random.Random(40 + seed)
    paths = random_paths(rng, n_nodes=rng.choice([3, 6, 12]), n_paths=rng.choice([3, 10, 40]), max_len=rng.choice([3, 8, 20]))
    img = image_of(gb.build_bwt(gb.bidirectional_sequences(paths + [[2, 4]])))
    for layout in (0, 1):
        back = HostSim(img, layout).serialize()
        src, out = orc.GBWT.load(img), orc.GBWT.load(back)
        same_index(src, out)
        assert out.bwt_data() == src.bwt_data()


def test_split_runs_are_merged():
    # a source whose runs are NOT maximal (adjacent runs of one value): same records, canonical bytes
    edges = [[(1, 0)], [(2, 0), (3, 0)], [(0, 0)], [(0, 0)]]
    runs = [[(0, 7)], [(0, 2), (0, 1), (1, 1), (1, 2), (0, 1)], [(0, 4)], [(0, 3)]]
    img, _ = records_image(edges, runs, sequences=7, size=21, offset=0)
    canonical, _ = records_image(edges, [[(0, 7)], [(0, 3), (1, 3), (0, 1)], [(0, 4)], [(0, 3)]], sequences=7, size=21, offset=0)
    for layout in (0, 1):
        back = HostSim(img, layout).serialize()
        same_index(orc.GBWT.load(img), orc.GBWT.load(back))
        assert orc.GBWT.load(back).bwt_data() == orc.GBWT.load(canonical).bwt_data()
