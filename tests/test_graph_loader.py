"""Node labels of GBZ files as parsed by the product's loader (gbwt-rs_b200/csrc/sds_loader.cpp, run on the CPU
through tests/hostsim) and by the oracle, against the reference's literals (src/graph/tests.rs:21-34, 66-79), and
the synthetic GBZ writer (synth/) through both loaders in both label encodings (Graph versions 3 and 4)."""
import os
import struct

import numpy as np
import pytest

import gbwt_builder as gb
import golden_vectors as gv
from hostsim_build import HostSim
from oracle import oracle as orc
from synth import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GRAPH_TAG = 0x6B3764AF


def labels_of(starts, data):
    return [data[int(starts[i]):int(starts[i + 1])].tobytes() for i in range(len(starts) - 1)]


@pytest.mark.parametrize("name,truth", [("example.gbz", gv.GRAPH_SEQUENCES), ("example-v1.gbz", gv.GRAPH_SEQUENCES),
                                        ("translation.gbz", gv.GRAPH_SEQUENCES_TRANSLATION),
                                        ("translation-v1.gbz", gv.GRAPH_SEQUENCES_TRANSLATION)])
def test_product_loader_reads_fixture_labels(name, truth):
    raw = open(os.path.join(GOLDEN, name), "rb").read()
    starts, data = HostSim(raw).labels()
    assert labels_of(starts, data) == [s.encode() for s in truth]


def test_plain_gbwt_has_no_labels():
    raw = open(os.path.join(GOLDEN, "example.gbwt"), "rb").read()
    assert HostSim(raw).labels() is None


def small_gbz(version, max_anchor=40):
    S, H, seed = 40, 5, 8
    img = synth.bubble_chain(S, H, seed)
    starts, data = synth.node_labels(3 * S + 1, seed=2, max_anchor=max_anchor)
    return synth.gbz_image(img, starts, data, version), starts, data


@pytest.mark.parametrize("version", [3, 4])
def test_synthetic_gbz_through_both_loaders(version):
    gbz, starts, data = small_gbz(version)
    got_starts, got_data = HostSim(gbz).labels()
    assert np.array_equal(got_starts, starts) and np.array_equal(got_data, data)
    g = orc.GBWT.load(gbz)
    assert g.has_graph() and g.graph_sequences() == len(starts) - 1
    truth = labels_of(starts, data)
    for i, label in enumerate(truth):
        got = g.node_sequence(1 + i)
        assert got is None or got == label   # None: an allele no haplotype uses
    assert sum(g.node_sequence(1 + i) is not None for i in range(len(truth))) > len(truth) * 0.9


def test_versions_spell_the_same_dna():
    a, b = orc.GBWT.load(small_gbz(3)[0]), orc.GBWT.load(small_gbz(4)[0])
    for i in range(a.sequences()):
        assert a.extract_dna(i) == b.extract_dna(i)


def graph_header_at(image):
    words = np.frombuffer(image[:len(image) // 8 * 8], dtype="<u8")
    hits = [int(i) * 8 for i in np.nonzero((words & np.uint64(0xFFFFFFFF)) == np.uint64(GRAPH_TAG))[0]]
    assert hits
    return hits[-1]


@pytest.mark.parametrize("version", [3, 4])
def test_graph_load_errors(version):
    gbz, starts, data = small_gbz(version)
    at = graph_header_at(gbz)
    cases = {
        "GraphHeader: Invalid tag": gbz[:at] + struct.pack("<I", 0x12345678) + gbz[at + 4:],
        "GraphHeader: Invalid version": gbz[:at + 4] + struct.pack("<I", 9) + gbz[at + 8:],
        "GraphHeader: Invalid flags": gbz[:at + 16] + struct.pack("<Q", 0x12) + gbz[at + 24:],
        "GraphHeader: SDSL format is not supported": gbz[:at + 16] + struct.pack("<Q", 0) + gbz[at + 24:],
        "Graph: Translation flag does not match the presence of segment names": gbz[:at + 16] + struct.pack("<Q", 3) + gbz[at + 24:],
    }
    for message, image in cases.items():
        with pytest.raises(IOError, match=message):
            HostSim(image)
        with pytest.raises(IOError, match=message):
            orc.GBWT.load(image)
    for cut in (at + 8, at + 40, len(gbz) - 200):
        with pytest.raises(IOError):
            HostSim(gbz[:cut])
        with pytest.raises(IOError):
            orc.GBWT.load(gbz[:cut])


def test_sequence_count_must_match_the_alphabet():
    # GBZ::load, src/gbz.rs:690-694
    S, H = 40, 5
    img = synth.bubble_chain(S, H, 8)
    starts, data = synth.node_labels(3 * S, seed=2)   # one label short
    for version in (3, 4):
        gbz = synth.gbz_image(img, starts, data, version)
        for load in (HostSim, orc.GBWT.load):
            with pytest.raises(IOError, match="Mismatch between GBWT alphabet size and Graph sequence count"):
                load(gbz)


def test_corrupt_zstd_frame_is_rejected():
    gbz, starts, data = small_gbz(4, max_anchor=200)
    at = graph_header_at(gbz)
    broken = bytearray(gbz)
    # the frame sits after the sparse vector of starts; flipping bytes well inside the graph section breaks it
    for i in range(len(gbz) - 700, len(gbz) - 500):
        broken[i] ^= 0x5A
    assert at < len(gbz) - 700
    for load in (HostSim, orc.GBWT.load):
        with pytest.raises(IOError):
            load(bytes(broken))
