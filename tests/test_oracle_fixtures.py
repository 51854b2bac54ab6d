"""Pins the CPU oracle against the reference's GBWT-level tests on the shipped fixtures.

Mirrors gbwt-rs src/gbwt/tests.rs:44-56 (statistics), 164-238 (extract / backward / sequence),
268-350 (find / extend vs brute-force count_occurrences), 365-462 (bd_find / bd_extend), the
doc-tests of src/gbwt.rs:55-83, 546-548 and the raw record bytes / config anchors of
SURVEY.md App. B.7-8. CPU only.
"""
import os

import numpy as np
import pytest

import golden_vectors as gv
from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [("example.gbwt", False), ("with-empty.gbwt", True)]


def load(name):
    return orc.GBWT.load(os.path.join(GOLDEN, name))


@pytest.mark.parametrize("name", ["example.gbwt", "with-empty.gbwt"])
def test_statistics(name):
    g = load(name)
    s = gv.STATS[name]
    assert g.len() == s["len"] and g.sequences() == s["sequences"]
    assert g.alphabet_size() == s["alphabet_size"] and g.alphabet_offset() == s["alphabet_offset"]
    assert g.effective_size() == s["alphabet_size"] - s["alphabet_offset"]
    assert g.first_node() == s["alphabet_offset"] + 1
    assert g.is_bidirectional()
    for i in range(g.first_node()):
        assert not g.has_node(i)
    for i in range(g.first_node(), g.alphabet_size()):
        assert g.has_node(i)
    assert not g.has_node(g.alphabet_size())


def extract_sequence(g, seq_id):
    out, pos = [], g.start(seq_id)
    while pos is not None:
        out.append(pos[0])
        pos = g.forward(pos)
    return out


@pytest.mark.parametrize("name,with_empty", CASES)
def test_extract(name, with_empty):
    g = load(name)
    truth = gv.true_paths(with_empty)
    for i in range(g.sequences() // 2):
        fwd = extract_sequence(g, 2 * i)
        assert fwd == truth[i]
        assert extract_sequence(g, 2 * i + 1) == gv.reverse_path(fwd)


@pytest.mark.parametrize("name,with_empty", CASES)
def test_backward(name, with_empty):
    g = load(name)
    for i in range(g.sequences()):
        fwd = extract_sequence(g, i)
        last, pos = None, g.start(i)
        while pos is not None:
            last, pos = pos, g.forward(pos)
        rev, pos = [], last
        while pos is not None:
            rev.append(pos[0])
            pos = g.backward(pos)
        assert rev == fwd[::-1]


@pytest.mark.parametrize("name,with_empty", CASES)
def test_sequence(name, with_empty):
    g = load(name)
    for i in range(g.sequences()):
        assert g.sequence(i) == extract_sequence(g, i)
    assert g.sequence(g.sequences()) is None
    if not with_empty:
        assert g.sequence(7) == gv.SEQ7  # src/gbwt.rs:546-548


@pytest.mark.parametrize("name,with_empty", CASES)
def test_find(name, with_empty):
    g = load(name)
    nodes = gv.true_nodes()
    for i in range(g.alphabet_size() + 1):
        st = g.find(i)
        if st is not None:
            assert i in nodes and st[0] == i and st[2] > st[1]
        else:
            assert i not in nodes


@pytest.mark.parametrize("name,with_empty", CASES)
def test_extend(name, with_empty):
    g = load(name)
    paths = gv.true_paths(with_empty)
    for first in gv.true_nodes():
        start = g.find(first)
        for i in range(g.alphabet_size() + 1):
            count = gv.count_occurrences(paths, [first, i])
            st = g.extend(start, i)
            assert (st[2] - st[1] if st else 0) == count
    for path in paths:
        for j in range(len(path)):
            fwd = g.find(path[j])
            for k in range(j + 1, len(path)):
                fwd = g.extend(fwd, path[k])
                assert fwd is not None and fwd[2] - fwd[1] == gv.count_occurrences(paths, path[j:k + 1])
            bwd = g.find(gv.flip_node(path[j]))
            for k in range(j - 1, -1, -1):
                bwd = g.extend(bwd, gv.flip_node(path[k]))
                assert bwd is not None and bwd[2] - bwd[1] == gv.count_occurrences(paths, path[k:j + 1])


def bd_search(g, path, first, start, end):
    """src/gbwt/tests.rs:352-361."""
    state = g.bd_find(path[first])
    if state is None:
        return None
    for i in range(first + 1, end):
        state = g.extend_forward(state, path[i])
        if state is None:
            return None
    for i in range(first - 1, start - 1, -1):
        state = g.extend_backward(state, path[i])
        if state is None:
            return None
    return state


@pytest.mark.parametrize("name,with_empty", CASES)
def test_bd_find(name, with_empty):
    g = load(name)
    nodes = gv.true_nodes()
    for i in range(g.alphabet_size() + 1):
        st = g.bd_find(i)
        if st is not None:
            fwd, rev = st
            assert i in nodes and fwd[0] == i and fwd[2] > fwd[1]
            assert rev[0] == gv.flip_node(i) and rev[2] - rev[1] == fwd[2] - fwd[1]
        else:
            assert i not in nodes


@pytest.mark.parametrize("name,with_empty", CASES)
def test_bd_extend(name, with_empty):
    g = load(name)
    paths = gv.true_paths(with_empty)
    for first in gv.true_nodes():
        start = g.bd_find(first)
        for i in range(g.alphabet_size() + 1):
            st = g.extend_forward(start, i)
            assert (st[0][2] - st[0][1] if st else 0) == gv.count_occurrences(paths, [first, i])
            st = g.extend_backward(start, i)
            assert (st[0][2] - st[0][1] if st else 0) == gv.count_occurrences(paths, [i, first])
    n_searches = 0
    for path in paths:
        for p in (path, gv.reverse_path(path)):
            for first in range(len(p)):
                for start in range(first + 1):
                    for end in range(first + 1, len(p) + 1):
                        st = bd_search(g, p, first, start, end)
                        assert st is not None
                        fwd, rev = st
                        count = gv.count_occurrences(paths, p[start:end])
                        assert fwd[2] - fwd[1] == count and rev[2] - rev[1] == count
                        assert fwd[0] == p[end - 1] and rev[0] == gv.flip_node(p[start])
                        # SURVEY.md 8(d) invariant: the reverse state is the unidirectional search of
                        # the reversed pattern.
                        rp = gv.reverse_path(p[start:end])
                        uni = g.find(rp[0])
                        for x in rp[1:]:
                            uni = g.extend(uni, x)
                        assert uni == rev
                        n_searches += 1
    if not with_empty:
        assert n_searches == 360


def test_doc_examples():
    g = load("example.gbwt")
    d = gv.DOC_FIND  # src/gbwt.rs:70-75
    st = g.find(d["nodes"][0])
    for x in d["nodes"][1:]:
        st = g.extend(st, x)
    assert st[0] == d["final_node"] and st[2] - st[1] == d["len"]
    assert st == (30, 0, 2)
    b = gv.DOC_BD  # src/gbwt.rs:77-83
    st = g.bd_find(b["first"])
    assert st == ((28, 0, 3), (29, 0, 3))
    st = g.extend_backward(st, b["back"])
    assert st == ((28, 0, 2), (25, 0, 2))
    st = g.extend_forward(st, b["fwd"])
    assert st == ((30, 0, 2), (25, 0, 2))
    assert st[0][0] == b["forward_node"] and st[1][0] == b["reverse_node"]
    # src/gbwt.rs:60-68: second-to-last node of path 2 forward via backward()
    pos, last = g.start(4), None
    while pos is not None:
        last, pos = pos, g.forward(pos)
    assert g.backward(last)[0] == gv.encode_node(15)


def test_follow_doc_example():
    # StateIter doc-test, src/gbz.rs:1189-1208, and the probe values of SURVEY.md App. B.6
    g = load("example.gbz")
    d = gv.DOC_STATEITER
    state = g.bd_find(gv.encode_node(d["node"]))
    assert state[0][2] - state[0][1] == d["len"]
    successors = g.follow(state)
    assert len(successors) == d["successors"]
    assert successors == [((30, 0, 2), (29, 0, 2)), ((32, 0, 1), (29, 2, 3))]
    predecessors = []
    for s in successors:
        for p in g.follow(s, backward=True):
            frm, to = p[1][0] ^ 1, p[0][0]
            predecessors.append(((frm // 2, bool(frm & 1)), (to // 2, bool(to & 1)), p[0][2] - p[0][1]))
    assert predecessors == d["predecessors"]
    assert g.follow(((36, 0, 1), (37, 0, 1))) is None      # node 18 does not exist in example.gbz
    assert g.follow(((2**40, 0, 1), (2**40 + 1, 0, 1))) is None


def test_check_states_like_reference():
    # src/gbz/tests.rs:100-168 on both GBZ fixtures: follow() == extend_*() over the graph successors
    for name in ("example.gbz", "translation.gbz"):
        g = load(name)
        stack = [s for s in (g.bd_find(v) for v in range(g.alphabet_size() + 2)) if s is not None]
        visited = set(stack)
        while stack:
            state = stack.pop()
            for backward in (False, True):
                found = g.follow(state, backward=backward)
                assert found is not None
                ext = g.extend_backward if backward else g.extend_forward
                truth = {x for x in (ext(state, v) for v in range(g.alphabet_size())) if x is not None}
                assert set(found) == truth and len(found) == len(truth)
                for s in found:
                    if s not in visited:
                        visited.add(s); stack.append(s)
        assert len(visited) > 20


def test_raw_record_bytes():
    # SURVEY.md App. B.7 (bytes of the C++-built fixture)
    g = load("example.gbwt")
    data = g.bwt_data()
    assert len(data) == 141 and g.bwt_records() == 31
    s, e = g.record_bytes(0)
    assert data[s:e].hex() == "0416000d000700090000010203000100010a03"
    assert g.record_edges(0) == [(22, 0), (35, 0), (42, 0), (51, 0)]
    s, e = g.record_bytes(22 - 21)
    assert data[s:e].hex() == "02180002000201"
    assert g.record_edges(1) == [(24, 0), (26, 0)]
    s, e = g.record_bytes(23 - 21)
    assert data[s:e].hex() == "01000002"
    for node in range(36, 42):
        s, e = g.record_bytes(node - 21)
        assert data[s:e] == b"\x00"
        assert g.record_edges(node - 21) is None
    g2 = load("with-empty.gbwt")
    s, e = g2.record_bytes(0)
    assert g2.bwt_data()[s:e].hex() == "05000016000d00070009000102030401020102050d04"


def test_gbz_embeds_the_same_gbwt():
    for gbz, gbwt in [("example.gbz", "example.gbwt"), ("example-v1.gbz", "example.gbwt"),
                      ("translation.gbz", "translation.gbwt"), ("translation-v1.gbz", "translation.gbwt")]:
        a, b = load(gbz), load(gbwt)
        assert a.bwt_data() == b.bwt_data()
        assert (a.len(), a.sequences(), a.alphabet_size(), a.alphabet_offset()) == \
               (b.len(), b.sequences(), b.alphabet_size(), b.alphabet_offset())
        assert np.array_equal(a.record_starts(), b.record_starts())


def test_translation_paths_and_config2_anchor():
    g = load("translation.gbz")
    assert (g.sequences(), g.len(), g.alphabet_offset(), g.alphabet_size()) == (6, 48, 1, 24)
    paths = [g.sequence(2 * i) for i in range(3)]
    assert paths == gv.TRANSLATION_PATHS
    n = 0
    for i in range(g.sequences()):
        p = g.sequence(i)
        for first in range(len(p)):
            for start in range(first + 1):
                for end in range(first + 1, len(p) + 1):
                    st = bd_search(g, p, first, start, end)
                    assert st is not None
                    assert st[0][2] - st[0][1] == gv.count_occurrences(paths, p[start:end])
                    n += 1
    assert n == 504


def test_config1_anchor():
    # every length-4 window of every stored sequence: 20 queries, 32 occurrences
    g = load("example.gbz")
    pats = []
    for i in range(g.sequences()):
        p = g.sequence(i)
        pats += [p[j:j + 4] for j in range(len(p) - 3)]
    assert len(pats) == 20
    out = g.find_extend_batch(np.array(pats, dtype=np.uint64))
    assert int((out["end"] - out["start"]).sum()) == 32
    paths = gv.true_paths(False)
    for p, st in zip(pats, out):
        assert int(st["end"] - st["start"]) == gv.count_occurrences(paths, list(p))
        assert int(st["node"]) == p[-1]


def test_load_errors():
    data = bytearray(open(os.path.join(GOLDEN, "example.gbwt"), "rb").read())
    bad = bytearray(data); bad[0] ^= 0xFF
    with pytest.raises(IOError):
        orc.GBWT.load(bytes(bad))
    bad = bytearray(data); bad[4] = 4  # version
    with pytest.raises(IOError):
        orc.GBWT.load(bytes(bad))
    bad = bytearray(data); bad[40] &= ~4 & 0xFF  # drop the simple-sds flag (headers.rs:226-231)
    with pytest.raises(IOError):
        orc.GBWT.load(bytes(bad))
    bad = bytearray(data); bad[40] &= ~2 & 0xFF  # metadata flag mismatch (gbwt.rs:421-423)
    with pytest.raises(IOError):
        orc.GBWT.load(bytes(bad))
    with pytest.raises(IOError):
        orc.GBWT.load(bytes(data[:200]))


def test_batch_matches_scalar():
    g = load("example.gbwt")
    rng = np.random.default_rng(1)
    pats = rng.integers(20, 54, size=(500, 3), dtype=np.uint64)
    out = g.find_extend_batch(pats, threads=2)
    for p, st in zip(pats, out):
        s = g.find(int(p[0]))
        for x in p[1:]:
            s = g.extend(s, int(x)) if s else None
        assert (s or (0, 0, 0)) == (int(st["node"]), int(st["start"]), int(st["end"]))
    offs, nodes = g.extract_batch(np.arange(g.sequences() + 1))
    for i in range(g.sequences()):
        assert list(nodes[int(offs[i]):int(offs[i + 1])]) == g.sequence(i)
