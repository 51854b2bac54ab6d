"""GPU parity for node sequences and DNA-level path extraction: the CUDA path through the
C ABI against the oracle and the reference's literals (src/graph/tests.rs:21-34, 66-79; src/support/tests.rs:13-42;
extract_sequence of src/bin/gbz-extract.rs:173-189), bit-exact."""
import os
import random

import numpy as np
import pytest

import gbwt_builder as gb
import golden_vectors as gv
from oracle import oracle as orc
from synth import synth

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
U64MAX = np.uint64(2**64 - 1)


@pytest.fixture(scope="module")
def b200():
    import gbwt_rs_b200
    return gbwt_rs_b200


def check_dna(e, g, endmarker=0, ids=None):
    """extract_dna of the product against the oracle for the given sequence ids (default: all + one past the end)."""
    if ids is None:
        ids = np.arange(g.sequences() + 1, dtype=np.uint64)
    want_offsets, want, want_lengths = g.extract_dna_batch(ids, endmarker)
    lengths = e.dna_lengths(ids)
    assert np.array_equal(lengths, want_lengths)
    offsets, data, got_lengths = e.extract_dna(ids, endmarker)
    assert np.array_equal(offsets, want_offsets) and np.array_equal(got_lengths, want_lengths)
    assert np.array_equal(data, want)
    return int(offsets[-1])


@pytest.mark.parametrize("layout", ["auto", "runs"])
@pytest.mark.parametrize("name,truth,first", [("example.gbz", gv.GRAPH_SEQUENCES, 11), ("example-v1.gbz", gv.GRAPH_SEQUENCES, 11),
                                              ("translation.gbz", gv.GRAPH_SEQUENCES_TRANSLATION, 1),
                                              ("translation-v1.gbz", gv.GRAPH_SEQUENCES_TRANSLATION, 1)])
def test_fixture_node_sequences(b200, name, truth, first, layout):
    path = os.path.join(GOLDEN, name)
    e, g = b200.GBWT.load(path, layout=layout), orc.GBWT.load(path)
    assert e.has_graph() and e.graph_sequences() == len(truth)
    ids = np.arange(0, first + len(truth) + 3, dtype=np.uint64)
    offsets, data, lengths = e.node_sequences(ids)
    for i, node_id in enumerate(ids):
        k = int(node_id) - first
        label = truth[k] if 0 <= k < len(truth) else ""
        if label:
            assert data[int(offsets[i]):int(offsets[i + 1])].tobytes() == label.encode() and lengths[i] == len(label)
            assert e.node_sequence(int(node_id)) == label.encode() == g.node_sequence(int(node_id))
            assert e.sequence_len(int(node_id)) == len(label)
        else:
            assert lengths[i] == U64MAX and e.node_sequence(int(node_id)) is None and g.node_sequence(int(node_id)) is None
    check_dna(e, g)
    check_dna(e, g, endmarker=ord("$"))


@pytest.mark.parametrize("name,nodes", [("example.gbz", gv.GBZ_NODES), ("example-v1.gbz", gv.GBZ_NODES),
                                        ("translation.gbz", gv.GBZ_NODES_TRANSLATION), ("translation-v1.gbz", gv.GBZ_NODES_TRANSLATION)])
def test_check_nodes_like_reference(b200, name, nodes):
    # check_nodes of src/gbz/tests.rs:11-34 (random access part) in one batch: sequence / sequence_len of every id
    e = b200.GBWT.load(os.path.join(GOLDEN, name))
    truth = dict(nodes)
    ids = np.arange(max(truth) + 2, dtype=np.uint64)
    offsets, data, lengths = e.node_sequences(ids)
    for node_id in range(len(ids)):
        if node_id in truth:
            assert data[int(offsets[node_id]):int(offsets[node_id + 1])].tobytes() == truth[node_id].encode()
            assert lengths[node_id] == len(truth[node_id]) == e.sequence_len(node_id)
        else:
            assert lengths[node_id] == U64MAX and e.sequence_len(node_id) is None and e.node_sequence(node_id) is None


def test_example_dna_literals(b200):
    e = b200.GBWT.load(os.path.join(GOLDEN, "example.gbz"))
    labels = {11 + i: s.encode() for i, s in enumerate(gv.GRAPH_SEQUENCES)}
    for path_id, path in enumerate(gv.true_paths(False)):
        assert e.path_dna(2 * path_id, ord("$")) == gv.true_dna(lambda n: labels[n], path, b"$")
        assert e.path_dna(2 * path_id + 1, ord("$")) == gv.true_dna(lambda n: labels[n], gv.reverse_path(path), b"$")
    assert e.path_dna(0, ord("$")) == b"GATAA$" and e.path_dna(e.sequences()) is None


def test_reverse_complement_table(b200):
    # one node whose label is every byte value, walked in both orientations: the reverse path spells
    # support::reverse_complement of the label (src/support.rs:87-110, literals of src/support/tests.rs:13-42)
    img = synth.gbwt_image(**image_args(gb.build_bwt(gb.bidirectional_sequences([[2], [2, 4], [4, 2]]))))
    labels = [bytes(range(256)), b"GATTACAT"]
    starts = np.cumsum([0] + [len(x) for x in labels]).astype(np.uint64)
    data = np.frombuffer(b"".join(labels), dtype=np.uint8)
    for version in (3, 4):
        gbz = synth.gbz_image(img, starts, data, version)
        e, g = b200.GBWT.from_bytes(gbz), orc.GBWT.load(gbz)
        assert e.path_dna(0)[:-1] == labels[0] and e.path_dna(1)[:-1] == labels[0][::-1].translate(gv.complement_table())
        assert e.path_dna(3)[:-1] == b"ATGTAATC" + labels[0][::-1].translate(gv.complement_table())
        check_dna(e, g)


def image_args(b, bidirectional=True):
    return dict(sequences=b["sequences"], size=b["size"], offset=b["offset"], alphabet_size=b["alphabet_size"],
                flags=4 | (1 if bidirectional else 0), rec_starts=b["starts"], data=b["data"])


def random_labels(rng, n, max_len):
    labels = [bytes(rng.choice(b"ACGTacgtN") for _ in range(rng.choice([0, 1, 1, 2, 5, max_len]))) for _ in range(n)]
    starts = np.cumsum([0] + [len(x) for x in labels]).astype(np.uint64)
    return starts, np.frombuffer(b"".join(labels) or b"", dtype=np.uint8)


@pytest.mark.parametrize("layout", ["auto", "runs"])
@pytest.mark.parametrize("seed", range(6))
def test_random_graphs_with_labels(b200, seed, layout):
    from test_hostsim_layout import random_paths
    rng = random.Random(100 + seed)
    paths = random_paths(rng, n_nodes=rng.choice([2, 3, 6, 12]), n_paths=rng.choice([3, 10, 40]), max_len=rng.choice([3, 8, 20, 90]))
    if not any(paths):
        paths.append([2, 4])
    b = gb.build_bwt(gb.bidirectional_sequences(paths))
    img = synth.gbwt_image(**image_args(b))
    n_labels = (b["alphabet_size"] - (b["offset"] + 1)) // 2
    starts, data = random_labels(rng, n_labels, rng.choice([3, 40, 700, 3000]))
    gbz = synth.gbz_image(img, starts, data, rng.choice([3, 4]))
    e, g = b200.GBWT.from_bytes(gbz, layout=layout), orc.GBWT.load(gbz)
    check_dna(e, g, endmarker=rng.choice([0, ord("$"), 255]))
    # the same graph attached to an index built from parts
    e2 = b200.GBWT.from_parts(g.sequences(), g.len(), g.alphabet_offset(), g.alphabet_size(), g.flags(), g.bwt_data(),
                              g.record_starts(), layout=layout).attach_graph(starts, data)
    check_dna(e2, g)


@pytest.mark.parametrize("layout", ["auto", "runs"])
def test_bubble_chain_dna(b200, layout):
    # long paths (4001 nodes: many 32-node groups per walk), all haplotypes in both orientations
    S, H, seed = 2000, 48, 11
    img = synth.bubble_chain(S, H, seed)
    starts, data = synth.node_labels(3 * S + 1, seed=5, max_anchor=48)
    gbz = synth.gbz_image(img, starts, data, 4)
    e, g = b200.GBWT.from_bytes(gbz, layout=layout), orc.GBWT.load(gbz)
    total = check_dna(e, g, endmarker=ord("\n"))
    assert total > 2 * H * 2 * S
    # forward and reverse orientations of a path are reverse complements of each other
    offsets, out, _ = e.extract_dna(np.arange(2 * H, dtype=np.uint64))
    for p in range(0, H, 7):
        fw = out[int(offsets[2 * p]):int(offsets[2 * p + 1]) - 1].tobytes()
        rv = out[int(offsets[2 * p + 1]):int(offsets[2 * p + 2]) - 1].tobytes()
        assert rv == fw[::-1].translate(gv.complement_table())


def test_truncated_slots_and_device_pointers(b200):
    import torch
    S, H, seed = 500, 16, 4
    img = synth.bubble_chain(S, H, seed)
    starts, data = synth.node_labels(3 * S + 1, seed=9)
    e = b200.GBWT.from_bytes(img.array).attach_graph(starts, data)
    g = orc.GBWT.load(synth.gbz_image(img, starts, data, 3))
    ids = np.arange(2 * H, dtype=np.uint64)
    want_offsets, want, want_lengths = g.extract_dna_batch(ids, 7)
    # slots shorter than the results: each result is cut to its slot, lengths still report the full size
    cut = np.minimum(want_lengths, np.uint64(1000)) - np.arange(2 * H, dtype=np.uint64) % np.uint64(3)
    offsets = np.zeros(2 * H + 1, np.uint64)
    np.cumsum(cut, out=offsets[1:])
    d_ids = torch.from_numpy(ids.view(np.int64)).cuda()
    d_offsets = torch.from_numpy(offsets.view(np.int64)).cuda()
    d_bytes = torch.full((int(offsets[-1]) + 64,), 0xEE, dtype=torch.uint8, device="cuda")
    d_lengths = torch.zeros(2 * H, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    before = b200.kernel_launches()
    e.extract_dna_device(d_ids.data_ptr(), 2 * H, 7, d_offsets.data_ptr(), d_bytes.data_ptr(), d_lengths.data_ptr(), stream)
    torch.cuda.synchronize()
    assert b200.kernel_launches() == before + 1
    got = d_bytes.cpu().numpy()
    assert np.array_equal(d_lengths.cpu().numpy().view(np.uint64), want_lengths)
    for i in range(2 * H):
        lo, hi = int(offsets[i]), int(offsets[i + 1])
        assert np.array_equal(got[lo:hi], want[int(want_offsets[i]):int(want_offsets[i]) + hi - lo])
    assert np.all(got[int(offsets[-1]):] == 0xEE)  # nothing written past the last slot
    e.dna_lengths_device(d_ids.data_ptr(), 2 * H, d_lengths.data_ptr(), stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_lengths.cpu().numpy().view(np.uint64), want_lengths)


def test_no_graph_is_an_error(b200):
    e = b200.GBWT.load(os.path.join(GOLDEN, "example.gbwt"))
    assert not e.has_graph() and e.graph_sequences() == 0
    for call in (lambda: e.dna_lengths([0]), lambda: e.extract_dna([0]), lambda: e.node_sequence_lengths([11])):
        with pytest.raises(b200.GBWTError) as err:
            call()
        assert err.value.code == 8
    # GBZ::load checks (src/gbz.rs:686-694)
    with pytest.raises(IOError, match="Mismatch between GBWT alphabet size and Graph sequence count"):
        e.attach_graph(np.array([0, 1], dtype=np.uint64), b"A")


def test_two_ended_dna_extraction(b200, monkeypatch):
    # Once the node count and DNA length of a sequence are known (dna_lengths), it is spelled from both ends by two
    # warps: identical bytes to the one-ended walk and the oracle, also for slots shorter than the results.
    import ctypes as C
    S, H, seed = 900, 20, 13
    img = synth.bubble_chain(S, H, seed)
    starts, data = synth.node_labels(3 * S + 1, seed=21, max_anchor=70)
    gbz = synth.gbz_image(img, starts, data, 3)
    e, g = b200.GBWT.from_bytes(gbz), orc.GBWT.load(gbz)
    ids = np.arange(2 * H + 1, dtype=np.uint64)
    first = e.extract_dna(ids, ord("#"))     # lengths measured inside -> two-ended
    again = e.extract_dna(ids, ord("#"))
    monkeypatch.setenv("GBWT_B200_EXTRACT_SPLIT", "0")
    plain = e.extract_dna(ids, ord("#"))
    monkeypatch.delenv("GBWT_B200_EXTRACT_SPLIT")
    want = g.extract_dna_batch(ids, ord("#"))
    for got in (first, again, plain):
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
    lib = b200.library()
    for slot in (100, 4001, int(want[2][0]) - 1, int(want[2][0]) // 2, 0):
        offsets = np.arange(len(ids) + 1, dtype=np.uint64) * np.uint64(slot)
        out = np.full(int(offsets[-1]) + 16, 0xEE, dtype=np.uint8)
        lengths = np.zeros(len(ids), dtype=np.uint64)
        rc = lib.gbwt_b200_extract_dna(e._h, ids.ctypes.data_as(C.c_void_p), len(ids), ord("#"), offsets.ctypes.data_as(C.c_void_p),
                                       out.ctypes.data_as(C.c_void_p), lengths.ctypes.data_as(C.c_void_p))
        assert rc == 0 and np.array_equal(lengths, want[2])
        for i in range(2 * H):
            full = want[1][int(want[0][i]):int(want[0][i + 1])]
            assert np.array_equal(out[i * slot:i * slot + min(slot, len(full))], full[:slot])
        assert np.all(out[int(offsets[-1]):] == 0xEE)


def test_two_ended_dna_needs_mirror_strands(b200):
    # flagged bidirectional, but the odd sequences are unrelated paths: the halves do not meet / do not add up, and
    # every sequence is spelled from the front like the reference does
    rng = random.Random(6)
    paths = [[2 * rng.randint(1, 40) + rng.randint(0, 1) for _ in range(rng.choice([150, 200, 333]))] for _ in range(8)]
    # ... and pairs that ARE each other's reverse strand except for one node far from the middle (the halves of a
    # two-ended walk would meet on equal nodes), next to a true pair: only the signatures of whole walks tell them apart
    for kind in ("one node", "mirror", "one node"):
        p = [2 * rng.randint(1, 40) + rng.randint(0, 1) for _ in range(rng.choice([180, 333]))]
        q = [x ^ 1 for x in reversed(p)]
        if kind == "one node":
            q[len(q) // 6] ^= 2 if q[len(q) // 6] >= 4 else 4
        paths += [p, q]
    b = gb.build_bwt(paths)
    img = synth.gbwt_image(**image_args(b))
    n_labels = (b["alphabet_size"] - (b["offset"] + 1)) // 2
    starts, data = random_labels(rng, n_labels, 9)
    gbz = synth.gbz_image(img, starts, data, 4)
    e, g = b200.GBWT.from_bytes(gbz), orc.GBWT.load(gbz)
    for _ in range(3):
        check_dna(e, g, endmarker=ord("$"), ids=np.arange(len(paths), dtype=np.uint64))


def test_whole_object_serialization_through_the_c_abi(b200, tmp_path):
    # GBWT::serialize / GBZ::serialize (src/gbwt.rs:388-400, src/gbz.rs:662-671) on the GPU-resident index: tags, BWT,
    # DA samples, metadata and the Graph come back; the image loads again and answers like the original
    from test_layout_writer import expected_tags, gbwt_sections, gbz_sections
    for name in ("example.gbz", "translation.gbz", "translation-v1.gbz"):
        raw = open(os.path.join(GOLDEN, name), "rb").read()
        e = b200.GBWT.from_bytes(raw)
        image = e.serialize(gbz=True)
        gh, gtags, header, tags, rest, graph = gbz_sections(raw)
        oh, otags, oheader, otags_gbwt, orest, ograph = gbz_sections(image)
        assert otags == expected_tags(gtags) and otags_gbwt == expected_tags(tags)
        assert oheader == header and orest == rest and ograph == graph
        path = tmp_path / name
        e.save(path, gbz=True)
        again = b200.GBWT.load(path)
        g = orc.GBWT.load(raw)
        ids = np.arange(g.sequences(), dtype=np.uint64)
        a, b = e.extract_dna(ids, endmarker=ord("$")), again.extract_dna(ids, endmarker=ord("$"))
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        plain = e.serialize()
        h2, t2, r2, end = gbwt_sections(plain)
        assert end == len(plain) and h2 == header and r2 == rest and t2 == expected_tags(tags)
    # labels attached by hand: written as a version-3 Graph, read back by the product and by the oracle
    S, H, seed = 60, 8, 3
    img = synth.bubble_chain(S, H, seed)
    starts, labels = synth.node_labels(3 * S + 1, seed=5)
    e = b200.GBWT.from_bytes(img.array)
    with pytest.raises(b200.GBWTError):
        e.serialize(gbz=True)                       # GBWT_B200_E_NO_GRAPH
    e.attach_graph(starts, labels)
    image = e.serialize(gbz=True)
    again = b200.GBWT.from_bytes(image)
    ids = np.arange(2 * H, dtype=np.uint64)
    a, b = e.extract_dna(ids), again.extract_dna(ids)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    g = orc.GBWT.load(image)
    w_off, w_bytes, _ = g.extract_dna_batch(ids, 0)
    assert np.array_equal(a[0], w_off) and np.array_equal(a[1], w_bytes)


@pytest.mark.parametrize("shift", [6, 8])
def test_checkpointed_dna_extraction(b200, shift, monkeypatch):
    # With path checkpoints AND node labels the index also knows where the DNA of every sequence stands at each checkpoint
    # (one count pass when both are there), and a sequence is spelled as independent segments, one lane each
    # (k_extract_dna_checkpointed): identical bytes to the whole-sequence walks and the oracle.
    import torch
    monkeypatch.setenv("GBWT_B200_CHECKPOINTS", "1")
    monkeypatch.setenv("GBWT_B200_CHECKPOINT_SHIFT", str(shift))
    S, H, seed = 2000, 48, 11   # 4001 nodes per path: 63 segments of 64 nodes, 16 of 256
    img = synth.bubble_chain(S, H, seed)
    starts, data = synth.node_labels(3 * S + 1, seed=5, max_anchor=48)
    gbz = synth.gbz_image(img, starts, data, 4)
    e, g = b200.GBWT.from_bytes(gbz), orc.GBWT.load(gbz)
    info = e.checkpoint_info()
    assert info["present"] == 1 and info["max_segments"] >= 4001 >> shift
    ids = np.arange(2 * H + 1, dtype=np.uint64)   # all sequences and one past the end (None)
    want_offsets, want, want_lengths = g.extract_dna_batch(ids, ord("\n"))
    before = b200.kernel_launches()
    assert np.array_equal(e.dna_lengths(ids), want_lengths)       # a table lookup now
    offsets, got, lengths = e.extract_dna(ids, ord("\n"))
    assert np.array_equal(offsets, want_offsets) and np.array_equal(lengths, want_lengths) and np.array_equal(got, want)
    assert b200.kernel_launches() == before + 3   # lengths; lengths again inside extract_dna; the segments
    # slots shorter than the results: each result is cut to its slot, lengths still report the full size
    n = 2 * H
    cut = np.minimum(want_lengths[:n], np.uint64(1500)) - np.arange(n, dtype=np.uint64) % np.uint64(3)
    slots = np.zeros(n + 1, np.uint64)
    np.cumsum(cut, out=slots[1:])
    d_ids = torch.from_numpy(ids[:n].view(np.int64)).cuda()
    d_offsets = torch.from_numpy(slots.view(np.int64)).cuda()
    d_bytes = torch.full((int(slots[-1]) + 64,), 0xEE, dtype=torch.uint8, device="cuda")
    d_lengths = torch.zeros(n, dtype=torch.int64, device="cuda")
    e.extract_dna_device(d_ids.data_ptr(), n, 7, d_offsets.data_ptr(), d_bytes.data_ptr(), d_lengths.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = d_bytes.cpu().numpy()
    assert np.array_equal(d_lengths.cpu().numpy().view(np.uint64), want_lengths[:n])
    for i in range(n):
        lo, hi = int(slots[i]), int(slots[i + 1])
        ref = want[int(want_offsets[i]):int(want_offsets[i]) + hi - lo].copy()
        if hi - lo == int(want_lengths[i]):
            ref[-1] = 7   # (the endmarker of this call)
        assert np.array_equal(out[lo:hi], ref), i
    assert np.all(out[int(slots[-1]):] == 0xEE)   # nothing written past the last slot
    # the same index without the table walks whole sequences: same bytes
    monkeypatch.setenv("GBWT_B200_DNA_CHECKPOINTS", "0")
    _, plain, _ = e.extract_dna(ids, ord("\n"))
    assert np.array_equal(plain, want)
    monkeypatch.delenv("GBWT_B200_DNA_CHECKPOINTS")
    # labels attached to an index that already has its checkpoints; other labels replace the table
    e2 = b200.GBWT.from_bytes(img.array).attach_graph(starts, data)
    _, got2, lengths2 = e2.extract_dna(ids, ord("\n"))
    assert np.array_equal(got2, want) and np.array_equal(lengths2, want_lengths)
    starts3, data3 = synth.node_labels(3 * S + 1, seed=6, max_anchor=20)
    g3 = orc.GBWT.load(synth.gbz_image(img, starts3, data3, 3))
    e2.attach_graph(starts3, data3)
    check_dna(e2, g3, endmarker=ord("$"))


@pytest.mark.parametrize("seed", range(4))
def test_checkpointed_dna_on_random_graphs_and_fixtures(b200, seed, monkeypatch):
    # short and empty paths, labels of 0 to 700 bytes, nodes on both strands: one segment per path or a few
    from test_hostsim_layout import random_paths
    monkeypatch.setenv("GBWT_B200_CHECKPOINTS", "1")
    monkeypatch.setenv("GBWT_B200_CHECKPOINT_SHIFT", "6")
    rng = random.Random(300 + seed)
    paths = random_paths(rng, n_nodes=rng.choice([3, 6, 12]), n_paths=rng.choice([3, 10, 40]), max_len=rng.choice([8, 90, 300]))
    paths.append([])
    paths.append([2, 4])
    b = gb.build_bwt(gb.bidirectional_sequences(paths))
    img = synth.gbwt_image(**image_args(b))
    n_labels = (b["alphabet_size"] - (b["offset"] + 1)) // 2
    starts, data = random_labels(rng, n_labels, rng.choice([3, 40, 700]))
    gbz = synth.gbz_image(img, starts, data, rng.choice([3, 4]))
    e, g = b200.GBWT.from_bytes(gbz), orc.GBWT.load(gbz)
    assert e.checkpoint_info()["present"] == 1
    check_dna(e, g, endmarker=rng.choice([0, ord("$"), 255]))
    if seed == 0:
        for name in ("example.gbz", "example-v1.gbz", "translation.gbz", "translation-v1.gbz"):
            path = os.path.join(GOLDEN, name)
            check_dna(b200.GBWT.load(path), orc.GBWT.load(path), endmarker=ord("$"))
