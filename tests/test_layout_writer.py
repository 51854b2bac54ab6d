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


# ---- a minimal reader of the Simple-SDS container, to cut an image into its sections ----------------

class Words:
    def __init__(self, raw, at=0):
        self.raw, self.at = raw, at

    def word(self):
        v = int.from_bytes(self.raw[self.at:self.at + 8], "little")
        self.at += 8
        return v

    def skip(self, words):
        self.at += 8 * words

    def raw_vector(self):
        self.word(); self.skip(self.word())

    def int_vector(self):
        self.word(); self.word(); self.raw_vector()

    def sparse_vector(self):
        self.word(); self.word(); self.raw_vector()
        for _ in range(3):
            self.skip(self.word())
        self.int_vector()

    def string_array(self):
        """Returns the strings of a packed StringArray and leaves the cursor behind it."""
        start = self.at
        universe = self.word(); ones = self.word(); hbits = self.word(); hwords = self.word()
        high = int.from_bytes(self.raw[self.at:self.at + 8 * hwords], "little"); self.skip(hwords)
        for _ in range(3):
            self.skip(self.word())
        n, width, lbits, lwords = self.word(), self.word(), self.word(), self.word()
        low = int.from_bytes(self.raw[self.at:self.at + 8 * lwords], "little"); self.skip(lwords)
        starts, j, pos = [], 0, 0
        while j < ones:
            if (high >> pos) & 1:
                starts.append(((pos - j) << width) | ((low >> (j * width)) & ((1 << width) - 1)))
                j += 1
            pos += 1
        alen = self.word()
        alphabet = self.raw[self.at:self.at + alen]; self.skip((alen + 7) // 8)
        total, cw, cbits, cwords = self.word(), self.word(), self.word(), self.word()
        packed = int.from_bytes(self.raw[self.at:self.at + 8 * cwords], "little"); self.skip(cwords)
        text = bytes(alphabet[(packed >> (i * cw)) & ((1 << cw) - 1)] for i in range(total))
        ends = starts[1:] + [total]
        return [text[a:b].decode() for a, b in zip(starts, ends)]


def gbwt_sections(raw, at=0):
    """(header bytes, tags dict, bytes from the BWT to the end of the GBWT, offset behind the GBWT)."""
    w = Words(raw, at)
    w.skip(6)
    header = raw[at:w.at]
    strings = w.string_array()
    tags = dict(zip(strings[0::2], strings[1::2]))
    rest_at = w.at
    w.sparse_vector()                       # BWT index
    n = w.word(); w.skip((n + 7) // 8)      # BWT data
    w.skip(w.word())                        # DA samples
    w.skip(w.word())                        # Option<Metadata>
    return header, tags, raw[rest_at:w.at], w.at


def gbz_sections(raw):
    w = Words(raw)
    w.skip(2)
    strings = w.string_array()
    tags = dict(zip(strings[0::2], strings[1::2]))
    header, gbwt_tags, rest, end = gbwt_sections(raw, w.at)
    return raw[:16], tags, header, gbwt_tags, rest, raw[end:]


def expected_tags(tags):
    out = {k.lower(): v for k, v in tags.items()}
    out["source"] = "jltsiren/gbwt-rs"      # Tags::insert(SOURCE_KEY, SOURCE_VALUE) at load, src/gbwt.rs:404-405
    return out


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("name", FIXTURES)
def test_fixtures_round_trip(name, layout):
    raw = open(os.path.join(GOLDEN, name), "rb").read()
    image = HostSim(raw, layout).serialize()
    src, back = orc.GBWT.load(raw), orc.GBWT.load(image)
    same_index(src, back)
    # files written by the reference have maximal runs: the BWT comes back byte for byte
    assert back.bwt_data() == src.bwt_data() and np.array_equal(back.record_starts(), src.record_starts())
    assert back.flags() == src.flags()                               # the metadata travels with the index now
    pc.check_everything(HostSim(image, layout), src)                 # and the product loads its own output
    assert HostSim(image, layout).serialize() == image               # idempotent


@pytest.mark.parametrize("name", FIXTURES)
def test_whole_object_round_trip(name):
    """serialize::test of the reference is a whole-object round trip (src/gbwt/tests.rs:72-84): tags, BWT, DA samples and
    metadata. Everything behind the tags comes back byte for byte; the tags are the loaded ones with `source` replaced,
    in key order, which is what the reference writes after a load."""
    raw = open(os.path.join(GOLDEN, name), "rb").read()
    if name.endswith(".gbz"):
        _, _, header, tags, rest, _ = gbz_sections(raw)
    else:
        header, tags, rest, end = gbwt_sections(raw)
        assert end == len(raw)
    image = HostSim(raw, 0).serialize()
    out_header, out_tags, out_rest, out_end = gbwt_sections(image)
    assert out_end == len(image)
    assert out_header == header and out_rest == rest and out_tags == expected_tags(tags)
    assert list(out_tags) == sorted(out_tags)
    if tags == expected_tags(tags) and list(tags) == sorted(tags) and not name.endswith(".gbz"):
        assert image == raw                                          # a file the reference wrote: identical as a whole


@pytest.mark.parametrize("name", [n for n in FIXTURES if n.endswith(".gbz")])
def test_gbz_round_trip(name):
    """GBZ::serialize (src/gbz.rs:662-671): header, tags, GBWT, Graph. The Graph section is the loaded one."""
    raw = open(os.path.join(GOLDEN, name), "rb").read()
    sim = HostSim(raw, 0)
    image = sim.serialize(gbz=True)
    gh, gtags, header, tags, rest, graph = gbz_sections(raw)
    oh, otags, oheader, otags_gbwt, orest, ograph = gbz_sections(image)
    assert oh[:4] == gh[:4] and int.from_bytes(oh[4:8], "little") == 2 and oh[8:] == gh[8:]   # written as version 2
    assert otags == expected_tags(gtags) and otags_gbwt == expected_tags(tags)
    assert oheader == header and orest == rest and ograph == graph
    # the oracle (= the reference's loader restated) and the product load it: same index, same node labels
    src, back = orc.GBWT.load(raw), orc.GBWT.load(image)
    same_index(src, back)
    again = HostSim(image, 0)
    assert same_labels(again.labels(), sim.labels()) and again.serialize(gbz=True) == image
    if gtags == expected_tags(gtags) and tags == expected_tags(tags) and int.from_bytes(gh[4:8], "little") == 2:
        assert image == raw


def same_labels(a, b):
    return a is not None and b is not None and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("name", ["example.gbz", "translation-v1.gbz", None])
def test_gbz_written_from_attached_labels(name):
    """Labels that did not come with a Graph section (gbwt_b200_index_attach_graph) are written as a version-3 Graph:
    plain StringArray, no translation. The product's loader and the oracle read it back."""
    if name is None:
        S, H, seed = 40, 6, 3
        starts, labels = synth.node_labels(3 * S + 1, seed=5)
        raw = synth.gbz_image(synth.bubble_chain(S, H, seed), starts, labels, 4)
    else:
        raw = open(os.path.join(GOLDEN, name), "rb").read()
    sim = HostSim(raw, 0)
    want = sim.labels()
    sim.drop_graph_section()
    image = sim.serialize(gbz=True)
    back = HostSim(image, 0)
    assert same_labels(back.labels(), want)
    same_index(orc.GBWT.load(raw), orc.GBWT.load(image))
    w = Words(image); w.skip(2); w.string_array()
    _, _, _, end = gbwt_sections(image, w.at)
    assert int.from_bytes(image[end + 4:end + 8], "little") == 3 and int.from_bytes(image[end + 16:end + 24], "little") == 2


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
