"""Oracle node sequences and DNA-level extraction against the reference's literals:
src/graph/tests.rs:21-34, 66-79, src/support/tests.rs:13-42, src/gbz/tests.rs (example.gbz / example-v1.gbz hold
the same graph in the version 4 zstd and the version 3 packed StringArray encodings)."""
import os

import numpy as np
import pytest

import golden_vectors as gv
from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return orc.GBWT.load(os.path.join(GOLDEN, name))


@pytest.mark.parametrize("name,truth,first", [("example.gbz", gv.GRAPH_SEQUENCES, 11), ("example-v1.gbz", gv.GRAPH_SEQUENCES, 11),
                                              ("translation.gbz", gv.GRAPH_SEQUENCES_TRANSLATION, 1)])
def test_node_sequences(name, truth, first):
    g = load(name)
    assert g.has_graph() and g.graph_sequences() == len(truth)
    for i, label in enumerate(truth):
        node_id = first + i
        got = g.node_sequence(node_id)
        # GBZ::sequence is None for identifiers without a node (empty records; GBZ::has_node, src/gbz.rs:286-298);
        # in these fixtures those are exactly the identifiers with an empty label.
        assert got == (label.encode() if label else None)
    assert g.node_sequence(first - 1) is None and g.node_sequence(first + len(truth)) is None


def test_reverse_complement_literals():
    for seq, truth in gv.REVERSE_COMPLEMENTS:
        assert orc.reverse_complement(seq) == truth
    every = bytes(range(256))
    assert orc.reverse_complement(every) == every[::-1].translate(gv.complement_table())


@pytest.mark.parametrize("name", ["example.gbz", "example-v1.gbz", "translation.gbz"])
def test_extract_dna(name):
    g = load(name)
    for seq_id in range(g.sequences()):
        path = [int(x) for x in g.sequence(seq_id)]
        for endmarker in (0, ord("$")):
            truth = gv.true_dna(g.node_sequence, path, bytes([endmarker]))
            assert g.extract_dna(seq_id, endmarker) == truth
    assert g.extract_dna(g.sequences()) is None
    # forward and reverse orientations of a path are reverse complements of each other
    for path_id in range(g.sequences() // 2):
        fw, rv = g.extract_dna(2 * path_id)[:-1], g.extract_dna(2 * path_id + 1)[:-1]
        assert orc.reverse_complement(fw) == rv


def test_example_dna_literals():
    # paths of the example graph (src/gbwt/tests.rs true_paths) spelled over GRAPH_SEQUENCES
    g = load("example.gbz")
    paths = gv.true_paths(False)
    labels = {11 + i: s.encode() for i, s in enumerate(gv.GRAPH_SEQUENCES)}
    for path_id, path in enumerate(paths):
        assert g.extract_dna(2 * path_id, ord("$")) == gv.true_dna(lambda n: labels[n], path, b"$")
        assert g.extract_dna(2 * path_id + 1, ord("$")) == gv.true_dna(lambda n: labels[n], gv.reverse_path(path), b"$")
    assert g.extract_dna(0, ord("$")) == b"GATAA$"


def test_v1_and_v2_files_agree():
    a, b = load("example.gbz"), load("example-v1.gbz")
    ids = np.arange(a.sequences(), dtype=np.uint64)
    oa, da, la = a.extract_dna_batch(ids)
    ob, db, lb = b.extract_dna_batch(ids)
    assert np.array_equal(oa, ob) and np.array_equal(da, db) and np.array_equal(la, lb)
    assert all(da[int(oa[i]):int(oa[i + 1])].tobytes() == a.extract_dna(i) for i in range(len(ids)))


def test_gbwt_file_has_no_graph():
    g = load("example.gbwt")
    assert not g.has_graph() and g.extract_dna(0) is None and g.node_sequence(11) is None


@pytest.mark.parametrize("name,nodes", [("example.gbz", gv.GBZ_NODES), ("example-v1.gbz", gv.GBZ_NODES),
                                        ("translation.gbz", gv.GBZ_NODES_TRANSLATION), ("translation-v1.gbz", gv.GBZ_NODES_TRANSLATION)])
def test_check_nodes_like_reference(name, nodes):
    # check_nodes of src/gbz/tests.rs:11-34 (random access part): has_node / sequence for every id up to max_node + 1
    g = load(name)
    truth = dict(nodes)
    for node_id in range(max(truth) + 2):
        got = g.node_sequence(node_id)
        assert got == (truth[node_id].encode() if node_id in truth else None)
