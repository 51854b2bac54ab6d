"""Pins the CPU oracle against the reference's codec and BWT known-answer tests.

Mirrors gbwt-rs src/support/tests.rs:362-513 and src/bwt/tests.rs (check_records, check_lf,
check_follow, negative_offset_to, check_predecessor_at) on the literal examples of
src/bwt/tests.rs:10-87. CPU only.
"""
import random

import pytest

import golden_vectors as gv
from oracle import oracle as orc

ENDMARKER = 0


# ---- codecs -----------------------------------------------------------------------------

def test_bytecode_doc_vector():
    assert orc.bytecode_encode(gv.BYTECODE_VALUES) == gv.BYTECODE_BYTES
    assert orc.bytecode_decode(gv.BYTECODE_BYTES) == gv.BYTECODE_VALUES


def test_bytecode_random_roundtrip():
    # src/support/tests.rs:362-384: 647 values, widths geometric in blocks of 4 bits
    rng = random.Random(647)
    values = []
    for _ in range(647):
        w = 4
        while rng.random() < 0.5 and w < 64:
            w += 4
        values.append(rng.getrandbits(w))
    values += [0, 127, 128, 2**63, 2**64 - 1]
    data = orc.bytecode_encode(values)
    assert len(data) >= len(values)
    assert orc.bytecode_decode(data) == values


def test_rle_doc_vector():
    assert orc.rle_encode(4, gv.RLE4_RUNS) == gv.RLE4_BYTES
    assert orc.rle_decode(4, gv.RLE4_BYTES) == gv.RLE4_RUNS


def _random_runs(rng, n, sigma, w=4):
    eff = sigma if sigma else 2**63
    runs = []
    for _ in range(n):
        width = w
        while rng.random() < 0.5 and width < 32:
            width += w
        runs.append((rng.randrange(eff), rng.getrandbits(width) + 1))
    return runs


@pytest.mark.parametrize("n,sigma", gv.RLE_ROUNDTRIP_SIGMAS)
def test_rle_roundtrip(n, sigma):
    rng = random.Random(n * 1000003 + sigma)
    runs = _random_runs(rng, n, sigma)
    data = orc.rle_encode(sigma, runs)
    assert len(data) >= len(runs)
    assert orc.rle_decode(sigma, data) == runs


@pytest.mark.parametrize("sigma", gv.RLE_THRESHOLD_SIGMAS)
def test_rle_thresholds(sigma):
    # src/support/tests.rs:439-450: a run of threshold-1 takes 1 byte, a run of threshold takes 2.
    threshold = 256 // sigma
    truth, data = [], b""
    if threshold > 1:
        piece = orc.rle_encode(sigma, [(sigma - 1, threshold - 1)])
        assert len(piece) == 1
        data += piece
        truth.append((sigma - 1, threshold - 1))
    piece = orc.rle_encode(sigma, [(sigma - 1, threshold)])
    assert len(piece) == 2
    data += piece
    truth.append((sigma - 1, threshold))
    assert orc.rle_decode(sigma, data) == truth


def test_rle_all_single_byte_codes_decode_like_divmod():
    # RLEIter::next, support.rs:1421-1424, for every sigma < 255 and every non-escape head byte.
    for sigma in range(1, 255):
        threshold = 256 // sigma
        for b in range(256):
            v, l = b % sigma, b // sigma + 1
            if l == threshold:
                continue
            assert orc.rle_decode(sigma, bytes([b])) == [(v, l)]


def test_hand_assembled_record():
    # src/support/tests.rs:471-513
    rng = random.Random(8)
    runs = _random_runs(rng, 8, gv.RECORD_SIGMA)
    header = [gv.RECORD_SIGMA]
    prev = 0
    for node, offset in gv.RECORD_EDGES:
        header += [node - prev, offset]
        prev = node
    data = orc.bytecode_encode(header) + orc.rle_encode(gv.RECORD_SIGMA, runs)
    g = orc.GBWT.from_records([gv.RECORD_EDGES], [runs])
    assert g.bwt_data() == data
    assert g.record_edges(0) == gv.RECORD_EDGES
    assert g.record_len(0) == sum(l for _, l in runs)


# ---- BWT examples -----------------------------------------------------------------------

def _build(edges, runs):
    return orc.GBWT.from_records(edges, runs)


def _check_records(g, edges):
    assert g.bwt_records() == len(edges)
    for i, cur in enumerate(edges):
        got = g.record_edges(i)
        assert (got is None) == (len(cur) == 0), f"record {i} existence"
        if got is not None:
            assert got == cur


def _check_lf(g, edges, runs):
    # src/bwt/tests.rs:159-184
    for i in range(g.bwt_records()):
        if g.record_edges(i) is None:
            continue
        cur = [list(e) for e in edges[i]]
        decompressed = g.record_decompress(i)
        assert len(decompressed) == g.record_len(i)
        offset = 0
        for value, length in runs[i]:
            for _ in range(length):
                edge = tuple(cur[value])
                expected = None if edge[0] == ENDMARKER else edge
                assert g.record_lf(i, offset) == expected
                assert decompressed[offset] == edge
                expected = None if edge[0] == ENDMARKER else offset
                assert g.record_offset_to(i, edge) == expected
                offset += 1
                cur[value][1] += 1
        assert g.record_len(i) == offset
        assert g.record_lf(i, offset) is None


def _check_follow(g, invalid_node):
    # src/bwt/tests.rs:189-237: every (start, limit) x every successor, with lf() as the truth
    for i in range(g.bwt_records()):
        edges = g.record_edges(i)
        if edges is None:
            continue
        n = g.record_len(i)
        for start in range(n + 1):
            for limit in range(start, n + 1):
                assert g.record_follow(i, start, limit, ENDMARKER) is None
                assert g.record_bd_follow(i, start, limit, ENDMARKER) is None
                for successor, _ in edges:
                    if successor == ENDMARKER:
                        continue
                    result = g.record_follow(i, start, limit, successor)
                    bd = g.record_bd_follow(i, start, limit, successor)
                    if result is not None:
                        found = [result[0], result[0]]
                        for j in range(start, limit):
                            pos = g.record_lf(i, j)
                            if pos is not None and pos[0] == successor and pos[1] == found[1]:
                                found[1] += 1
                        assert tuple(found) == result
                        assert bd is not None and bd[0] == result
                    else:
                        for j in range(start, limit):
                            pos = g.record_lf(i, j)
                            if pos is not None:
                                assert pos[0] != successor
                        assert bd is None
                assert g.record_follow(i, start, limit, invalid_node) is None
                assert g.record_bd_follow(i, start, limit, invalid_node) is None


def _negative_offset_to(g, invalid_node):
    # src/bwt/tests.rs:239-258
    for i in range(g.bwt_records()):
        edges = g.record_edges(i)
        if edges is None:
            continue
        assert g.record_offset_to(i, (ENDMARKER, 0)) is None
        assert g.record_offset_to(i, (invalid_node, 0)) is None
        for successor, offset in edges:
            if successor == ENDMARKER:
                continue
            if offset > 0:
                assert g.record_offset_to(i, (successor, offset - 1)) is None
            r = g.record_follow(i, 0, g.record_len(i), successor)
            assert g.record_offset_to(i, (successor, offset + (r[1] - r[0]))) is None


def _check_predecessor_at(g):
    # src/bwt/tests.rs:261-283
    starting = set()
    for i in range(g.record_len(ENDMARKER)):
        starting.add(g.record_lf(ENDMARKER, i))
    for rid in range(1, g.bwt_records()):
        if g.record_edges(rid) is None:
            continue
        reverse_id = ((rid + 1) ^ 1) - 1
        n = g.record_len(rid)
        for i in range(n):
            pred = g.record_predecessor_at(reverse_id, i)
            if (rid + 1, i) in starting:
                assert pred is None
            else:
                assert pred is not None
        assert g.record_predecessor_at(reverse_id, n) is None


def test_empty_bwt():
    g = _build([], [])
    assert g.bwt_records() == 0
    _check_records(g, [])


def test_paper_example():
    g = _build(gv.PAPER_EDGES, gv.PAPER_RUNS)
    _check_records(g, gv.PAPER_EDGES)
    _check_lf(g, gv.PAPER_EDGES, gv.PAPER_RUNS)
    _check_follow(g, gv.PAPER_INVALID_NODE)
    _negative_offset_to(g, gv.PAPER_INVALID_NODE)
    d = gv.PAPER_DOC  # src/bwt.rs:22-35
    assert g.bwt_records() == d["len"]
    assert g.record_outdegree(d["rec"]) == d["outdegree"]
    assert g.record_edges(d["rec"])[1] == (d["successor1"], d["offset1"])
    assert g.record_len(d["rec"]) == d["rec_len"]
    assert g.record_lf(d["rec"], 1) == d["lf1"]
    (rng, node, expected) = d["follow"]
    assert g.record_follow(d["rec"], rng[0], rng[1], node) == expected
    assert sum(g.record_len(i) for i in range(g.bwt_records())) == d["total_len"]


def test_empty_records():
    edges = [list(e) for e in gv.PAPER_EDGES]
    runs = [list(r) for r in gv.PAPER_RUNS]
    for i in gv.EMPTY_RECORDS:
        edges[i], runs[i] = [], []
    g = _build(edges, runs)
    _check_records(g, edges)
    _check_lf(g, edges, runs)
    _check_follow(g, gv.PAPER_INVALID_NODE)
    _negative_offset_to(g, gv.PAPER_INVALID_NODE)


def test_bidirectional_example():
    g = _build(gv.BIDIR_EDGES, gv.BIDIR_RUNS)
    _check_records(g, gv.BIDIR_EDGES)
    _check_lf(g, gv.BIDIR_EDGES, gv.BIDIR_RUNS)
    _check_follow(g, gv.BIDIR_INVALID_NODE)
    _negative_offset_to(g, gv.BIDIR_INVALID_NODE)
    _check_predecessor_at(g)
