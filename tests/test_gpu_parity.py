"""GPU parity tests: the CUDA kernels, called through the C ABI (ctypes -> libgbwt_b200.so), against the CPU
oracle on the same inputs, bit-exact. Mirrors the reference's own tests (src/bwt/tests.rs, src/gbwt/tests.rs)
via tests/parity_checks.py and adds the synthetic configs of BASELINE.json."""
import os
import random

import numpy as np
import pytest

import gbwt_builder as gb
import golden_vectors as gv
import parity_checks as pc
from oracle import oracle as orc
from synth import synth

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = ["example.gbwt", "with-empty.gbwt", "translation.gbwt", "example.gbz", "translation.gbz",
            "example-v1.gbz", "translation-v1.gbz"]


@pytest.fixture(scope="module")
def b200():
    import gbwt_rs_b200
    return gbwt_rs_b200


def image_of(b, bidirectional=True):
    flags = 4 | (1 if bidirectional else 0)
    return synth.gbwt_image(b["sequences"], b["size"], b["offset"], b["alphabet_size"], flags, b["starts"], b["data"])


@pytest.mark.parametrize("layout", ["auto", "runs"])
@pytest.mark.parametrize("name", FIXTURES)
def test_fixtures(b200, name, layout):
    raw = open(os.path.join(GOLDEN, name), "rb").read()
    g = orc.GBWT.load(raw)
    e = b200.GBWT.from_bytes(raw, layout=layout)
    s = gv.STATS.get(name)
    if s:
        assert (e.len(), e.sequences(), e.alphabet_size(), e.alphabet_offset()) == \
               (s["len"], s["sequences"], s["alphabet_size"], s["alphabet_offset"])
        assert e.effective_size() == s["alphabet_size"] - s["alphabet_offset"] and e.first_node() == s["alphabet_offset"] + 1
        assert e.is_bidirectional() and not e.has_node(e.alphabet_offset()) and e.has_node(e.first_node())
    pc.check_everything(e, g)


def test_config1_example_gbz(b200):
    # BASELINE.json configs[0]: find() of every length-4 subpath of every stored path + full path extraction
    path = os.path.join(GOLDEN, "example.gbz")
    e, g = b200.GBWT.load(path), orc.GBWT.load(path)
    seqs = [list(e.sequence(i)) for i in range(e.sequences())]
    assert seqs[:12:2] == gv.true_paths(False) and seqs[7] == gv.SEQ7
    pats = np.array(pc.subpath_patterns(seqs, 4), dtype=np.uint64)
    out = e.find_extend(pats)
    assert len(pats) == 20 and int((out["end"] - out["start"]).sum()) == 32
    assert pc.states_equal(out, g.find_extend_batch(pats))
    assert e.sequence(e.sequences()) is None


def test_config2_translation_bidirectional(b200):
    # BASELINE.json configs[1]: bd_find + extend_forward/backward over all (first, start, end) of all paths
    path = os.path.join(GOLDEN, "translation.gbz")
    e, g = b200.GBWT.load(path), orc.GBWT.load(path)
    assert pc.check_bd(e, g) == 504


def test_scalar_reference_api(b200):
    # the doc-test of src/gbwt.rs:55-83 through the scalar mirror (a batch of one per call)
    e = b200.GBWT.load(os.path.join(GOLDEN, "example.gbwt"))
    st = e.find(24)
    st = e.extend(st, 28)
    st = e.extend(st, 30)
    assert st == b200.SearchState(30, range(0, 2)) and st.len() == 2
    bd = e.bd_find(28)
    bd = e.extend_backward(bd, 24)
    bd = e.extend_forward(bd, 30)
    assert bd.forward.node == 30 and bd.reverse.node == 25 and bd.len() == 2
    assert bd.from_() == (12, False) and bd.to() == (15, False)
    assert e.find(0) is None and e.find(36) is None and e.extend(st, 22) is None
    pos, last = e.start(4), None
    while pos is not None:
        last, pos = pos, e.forward(pos)
    assert e.backward(last).node == 30
    assert e.start(12) is None and e.forward(b200.Pos(0, 0)) is None


def test_unidirectional_index_rejects_bd(b200):
    paths = [[3, 5, 7, 9]] * 700 + [[3, 6, 7, 9]] * 300 + [[3, 5, 8]] * 5 + [[4, 5, 7]]
    img = image_of(gb.build_bwt([list(p) for p in paths]), bidirectional=False)
    for layout in ("auto", "runs"):
        e, g = b200.GBWT.from_bytes(img, layout=layout), orc.GBWT.load(img)
        assert not e.is_bidirectional()
        pc.check_find_all_nodes(e, g)
        pc.check_find_extend_subpaths(e, g)
        pc.check_find_extend_random(e, g)
        pc.check_navigation(e, g)
        for call in (lambda: e.bd_find(3), lambda: e.bd_find(np.array([3], dtype=np.uint64)),
                     lambda: e.backward(b200.Pos(5, 0)),
                     lambda: e.extend_forward(np.zeros(1, b200.BDSTATE_DTYPE), np.array([3], dtype=np.uint64))):
            with pytest.raises(AssertionError):
                call()


@pytest.mark.parametrize("layout", ["auto", "runs"])
@pytest.mark.parametrize("seed", range(6))
def test_random_graphs(b200, seed, layout):
    from test_hostsim_layout import random_paths
    rng = random.Random(seed)
    paths = random_paths(rng, n_nodes=rng.choice([2, 3, 6, 12]), n_paths=rng.choice([3, 10, 40]), max_len=rng.choice([3, 8, 20]))
    if not any(paths):
        paths.append([2, 4])
    img = image_of(gb.build_bwt(gb.bidirectional_sequences(paths)))
    pc.check_everything(b200.GBWT.from_bytes(img, layout=layout), orc.GBWT.load(img))


@pytest.mark.parametrize("layout", ["auto", "runs"])
def test_paper_and_bidirectional_examples(b200, layout):
    from test_hostsim_layout import records_image
    for edges, runs, bidir in [(gv.PAPER_EDGES, gv.PAPER_RUNS, False), (gv.BIDIR_EDGES, gv.BIDIR_RUNS, True)]:
        img, _ = records_image(edges, runs, sequences=3 if not bidir else 6, size=17, offset=0 if not bidir else 1,
                               bidirectional=bidir)
        pc.check_everything(b200.GBWT.from_bytes(img, layout=layout), orc.GBWT.load(img))


@pytest.mark.parametrize("layout", ["auto", "runs"])
@pytest.mark.parametrize("sigma,bits", [(2, (1, 3)), (2, (9, 12)), (3, (1, 4, 10)), (7, (2, 8)), (64, (1, 5)), (200, (1, 3, 9)),
                                        (254, (1, 2)), (255, (1, 9)), (256, (1, 4)), (300, (1, 12)), (1000, (1, 3))])
def test_wide_records(b200, sigma, bits, layout):
    from test_hostsim_layout import records_image, wide_record_index
    rng = random.Random(sigma * 31 + len(bits))
    edges, runs, total = wide_record_index(sigma, rng, bits)
    img, _ = records_image(edges, runs, sequences=total, size=3 * total, offset=0)
    g, e = orc.GBWT.load(img), b200.GBWT.from_bytes(img, layout=layout)
    st, nx = [], []
    cuts = sorted(set([0, 1, 2, total // 3, total // 2, total - 1, total, total + 5] + [rng.randrange(total + 1) for _ in range(12)]))
    for a in cuts:
        for b_ in cuts:
            if a < b_:
                for node in range(0, sigma + 4):
                    st.append((1, a, b_)); nx.append(node)
    st = np.array(st, dtype=orc.STATE_DTYPE); nx = np.array(nx, dtype=np.uint64)
    assert pc.states_equal(e.extend(st, nx), g.extend_batch(st, nx))
    pos = np.array([(1, i) for i in sorted(set(cuts + [rng.randrange(total) for _ in range(300)]))], dtype=orc.POS_DTYPE)
    assert pc.states_equal(e.forward(pos), g.forward_batch(pos))
    pc.check_find_all_nodes(e, g)
    ids = np.array(sorted(set(rng.randrange(total) for _ in range(50))), dtype=np.uint64)
    assert np.array_equal(e.sequence_lengths(ids), g.sequence_lengths(ids))


@pytest.mark.parametrize("layout", ["auto", "runs"])
def test_locality_schedule_on_small_inputs(b200, layout, monkeypatch):
    # The bucket pre-pass and the permuted kernel with tiny batches, non-matching patterns, out-of-range first
    # nodes, batches that are not a multiple of anything, and more buckets than queries.
    monkeypatch.setenv("GBWT_B200_LOCALITY", "1")
    for name in ("example.gbwt", "with-empty.gbwt", "translation.gbz"):
        raw = open(os.path.join(GOLDEN, name), "rb").read()
        g, e = orc.GBWT.load(raw), b200.GBWT.from_bytes(raw, layout=layout)
        pc.check_find_extend_subpaths(e, g)
        pc.check_find_extend_random(e, g, n=4001, k=4, seed=11)
        pats = np.array([[2**40, 22], [0, 0], [22, 2**63]], dtype=np.uint64)
        assert pc.states_equal(e.find_extend(pats), g.find_extend_batch(pats))
    S, H, seed = 700, 300, 5
    img = synth.bubble_chain(S, H, seed)
    g, e = orc.GBWT.load(img.array), b200.GBWT.from_bytes(img.array, layout=layout)
    for n in (1, 2, 255, 257, 70_001):
        pats = synth.patterns(S, H, seed, n=n, k=32, q0=n)
        assert pc.states_equal(e.find_extend(pats), g.find_extend_batch(pats))
    pc.check_find_extend_random(e, g, n=30_000, k=7, seed=2)


def from_parts_of(b200, g, layout="auto"):
    return b200.GBWT.from_parts(g.sequences(), g.len(), g.alphabet_offset(), g.alphabet_size(), g.flags(),
                                g.bwt_data(), g.record_starts(), layout=layout)


def test_from_parts_matches_from_bytes(b200):
    raw = open(os.path.join(GOLDEN, "example.gbwt"), "rb").read()
    g = orc.GBWT.load(raw)
    pc.check_everything(from_parts_of(b200, g), g)


@pytest.mark.parametrize("locality", [None, "1"])
@pytest.mark.parametrize("layout", ["auto", "runs"])
def test_config3_shape_small(b200, layout, locality, monkeypatch):
    # bubble chain like configs[2] (fewer sites so the oracle finishes in seconds): every query matches.
    # locality = "1" forces the bucketed schedule (counting sort by first record + permuted kernel), which
    # the library otherwise only uses for indexes that do not fit L2.
    if locality:
        monkeypatch.setenv("GBWT_B200_LOCALITY", locality)
    monkeypatch.setenv("GBWT_B200_HOST_CHUNK_MB", "64")   # 262144 patterns per chunk: the batch below needs two
    S, H, seed = 4000, 64, 42
    img = synth.bubble_chain(S, H, seed)
    g, e = orc.GBWT.load(img.array), b200.GBWT.from_bytes(img.array, layout=layout)
    pats = synth.patterns(S, H, seed, n=300_000, k=32)   # crosses the host pipeline's chunk boundary
    out = e.find_extend(pats)
    assert pc.states_equal(out, g.find_extend_batch(pats))
    assert np.all(out["end"] > out["start"]) and np.array_equal(out["node"], pats[:, -1])
    pc.check_find_extend_random(e, g, n=20000, k=8, seed=5)
    # ragged view of the same patterns with varying lengths
    rng = np.random.default_rng(3)
    lens = rng.integers(0, 33, size=50_000)
    offsets = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    nodes = np.concatenate([pats[i, :lens[i]] for i in range(len(lens))]) if offsets[-1] else np.zeros(0, np.uint64)
    assert pc.states_equal(e.find_extend_ragged(nodes, offsets), g.find_extend_ragged(nodes, offsets))
    # bidirectional searches sampled from the haplotypes
    first = rng.integers(0, 32, size=20_000).astype(np.uint64)
    start = (first * rng.random(20_000)).astype(np.uint64)
    end = (first + 1 + ((32 - first - 1) * rng.random(20_000)).astype(np.uint64)).astype(np.uint64)
    offs = (np.arange(20_001, dtype=np.uint64) * 32)
    flat = pats[:20_000].reshape(-1)
    got = e.bd_search(flat, offs, first, start, end)
    assert pc.states_equal(got, g.bd_search_batch(flat, offs, first, start, end))
    assert np.all(got["forward"]["end"] > got["forward"]["start"])
    # extraction of every haplotype, both strands
    ids = np.arange(2 * H, dtype=np.uint64)
    offsets, nodes, lengths = e.extract(ids)
    assert np.all(lengths == 2 * S + 1)
    for i in range(2 * H):
        assert np.array_equal(nodes[int(offsets[i]):int(offsets[i + 1])], synth.sequence(S, H, seed, i))


def test_config4_shape_small(b200):
    # 1024 haplotypes like configs[3]/[4]: long anchor records (dense under "auto", ~512 runs under "runs")
    S, H, seed = 600, 1024, 42
    img = synth.bubble_chain(S, H, seed)
    g = orc.GBWT.load(img.array)
    pats = synth.patterns(S, H, seed, n=100_000, k=32)
    want = g.find_extend_batch(pats)
    for layout in ("auto", "runs"):
        e = b200.GBWT.from_bytes(img.array, layout=layout)
        assert pc.states_equal(e.find_extend(pats), want)
        stats = e.device_bytes()
        assert (stats["records_dense"] > 0) == (layout == "auto")
        ids = np.arange(0, 2 * H, 37, dtype=np.uint64)
        offsets, nodes, lengths = e.extract(ids)
        for j, i in enumerate(ids):
            assert np.array_equal(nodes[int(offsets[j]):int(offsets[j + 1])], synth.sequence(S, H, seed, int(i)))


def test_full_size_config3_properties(b200):
    # BASELINE.json configs[2] at full size (100k nodes x 64 haplotypes, 10M patterns): size-independent
    # properties + exact comparison with the oracle on a sample.
    S, H, seed = 33333, 64, 42
    img = synth.bubble_chain(S, H, seed)
    e = b200.GBWT.from_bytes(img.array)
    n = 10_000_000
    pats = synth.patterns(S, H, seed, n=n, k=32)
    out = e.find_extend(pats)
    assert np.all(out["end"] > out["start"])                 # every sampled pattern occurs at least once
    assert np.array_equal(out["node"], pats[:, -1])          # SearchState.node is the last pattern node
    occ = (out["end"] - out["start"]).astype(np.int64)
    assert occ.max() <= H and 1.0 <= occ.mean() < 1.01
    g = orc.GBWT.load(img.array)
    idx = np.random.default_rng(0).choice(n, size=200_000, replace=False)
    assert pc.states_equal(out[idx], g.find_extend_batch(pats[idx]))
    # idempotence: the same batch again gives the same bytes; a permuted batch gives the permuted result
    perm = np.random.default_rng(1).permutation(1_000_000)
    assert pc.states_equal(e.find_extend(pats[:1_000_000][perm]), out[:1_000_000][perm])


def test_device_pointer_entry_points(b200):
    import torch
    S, H, seed = 2000, 64, 42
    img = synth.bubble_chain(S, H, seed)
    g, e = orc.GBWT.load(img.array), b200.GBWT.from_bytes(img.array)
    pats = synth.patterns(S, H, seed, n=50_000, k=32)
    d_pats = torch.from_numpy(pats.view(np.int64)).cuda()
    d_out = torch.zeros((len(pats), 3), dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    before = b200.kernel_launches()
    e.find_extend_device(d_pats.data_ptr(), len(pats), 32, d_out.data_ptr(), stream)
    torch.cuda.synchronize()
    assert b200.kernel_launches() == before + 1
    got = d_out.cpu().numpy().view(np.uint64).reshape(-1, 3)
    want = g.find_extend_batch(pats)
    assert np.array_equal(got, want.view(np.uint64).reshape(-1, 3))
    ids = torch.arange(2 * H, dtype=torch.int64, device="cuda")
    lens = torch.zeros(2 * H, dtype=torch.int64, device="cuda")
    e.sequence_lengths_device(ids.data_ptr(), 2 * H, lens.data_ptr(), stream)
    torch.cuda.synchronize()
    assert torch.all(lens == 2 * S + 1)
    offs = torch.arange(2 * H + 1, dtype=torch.int64, device="cuda") * (2 * S + 1)
    nodes = torch.zeros(2 * H * (2 * S + 1), dtype=torch.int64, device="cuda")
    e.extract_device(ids.data_ptr(), 2 * H, offs.data_ptr(), nodes.data_ptr(), lens.data_ptr(), stream)
    torch.cuda.synchronize()
    host = nodes.cpu().numpy().view(np.uint64).reshape(2 * H, 2 * S + 1)
    for i in range(0, 2 * H, 9):
        assert np.array_equal(host[i], synth.sequence(S, H, seed, i))


def test_pinned_host_buffers_and_threads(b200):
    import ctypes
    import threading
    S, H, seed = 2000, 64, 42
    img = synth.bubble_chain(S, H, seed)
    g, e = orc.GBWT.load(img.array), b200.GBWT.from_bytes(img.array)
    n = 400_000
    lib = b200.library()
    p_in, p_out = lib.gbwt_b200_host_alloc(n * 32 * 8), lib.gbwt_b200_host_alloc(n * 24)
    assert p_in and p_out
    pats = np.ctypeslib.as_array((ctypes.c_uint64 * (n * 32)).from_address(p_in)).reshape(n, 32)
    synth.patterns(S, H, seed, n=n, k=32, out=pats)
    out = np.ctypeslib.as_array((ctypes.c_uint64 * (n * 3)).from_address(p_out)).reshape(n, 3)
    assert lib.gbwt_b200_find_extend(e._h, p_in, n, 32, p_out) == 0
    want = g.find_extend_batch(pats).view(np.uint64).reshape(-1, 3)
    assert np.array_equal(out, want)
    # the handle is usable from several host threads at once (the reference type is Sync)
    results = [None] * 4
    def work(t):
        results[t] = e.find_extend(pats[t * 50_000:(t + 1) * 50_000])
    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in threads]; [t.join() for t in threads]
    for t in range(4):
        assert np.array_equal(results[t].view(np.uint64).reshape(-1, 3), want[t * 50_000:(t + 1) * 50_000])
    del pats, out
    lib.gbwt_b200_host_free(p_in); lib.gbwt_b200_host_free(p_out)


def test_load_errors_match_reference(b200):
    data = bytearray(open(os.path.join(GOLDEN, "example.gbwt"), "rb").read())
    for mutate in (lambda d: d.__setitem__(0, d[0] ^ 0xFF), lambda d: d.__setitem__(4, 4),
                   lambda d: d.__setitem__(40, d[40] & ~4 & 0xFF), lambda d: d.__setitem__(40, d[40] & ~2 & 0xFF)):
        bad = bytearray(data)
        mutate(bad)
        with pytest.raises(IOError):
            b200.GBWT.from_bytes(bytes(bad))
    with pytest.raises(IOError):
        b200.GBWT.from_bytes(bytes(data[:200]))
    with pytest.raises(IOError):
        b200.GBWT.load("/nonexistent/file.gbwt")
    empty = b200.GBWT.from_bytes(synth.gbwt_image(0, 0, 0, 0, 4, [], b""))
    assert empty.find(np.arange(4, dtype=np.uint64))["end"].sum() == 0 and empty.start(0) is None


def test_extraction_with_invalid_edge_targets(b200):
    # An edge to a node beyond the alphabet: GBWT::forward returns None there, so the sequence ends after that node.
    # The layout flags such an index (edges_valid = false) and extraction runs its bounds-checked kernel.
    edges = [[(1, 0)], [(2, 0), (40, 0)], [(0, 0)]]
    runs = [[(0, 2)], [(1, 1), (0, 1)], [(0, 1)]]
    g0 = orc.GBWT.from_records(edges, runs, sequences=2, size=6, offset=0)
    img = synth.gbwt_image(2, 6, 0, 3, 4, g0.record_starts(), g0.bwt_data())
    g, e = orc.GBWT.load(img), b200.GBWT.from_bytes(img)
    ids = np.arange(3, dtype=np.uint64)
    offsets, nodes, lengths = e.extract(ids)
    for i in range(2):
        want = [int(x) for x in g.sequence(i)]
        assert [int(x) for x in nodes[int(offsets[i]):int(offsets[i + 1])]] == want
    assert sorted(len(list(g.sequence(i))) for i in range(2)) == [2, 2] and lengths[2] == np.uint64(2**64 - 1)
    assert {tuple(int(x) for x in nodes[int(offsets[i]):int(offsets[i + 1])]) for i in range(2)} == {(1, 40), (1, 2)}


def test_two_ended_extraction(b200, monkeypatch):
    # Long sequences of a bidirectional index are walked from both ends once their length is known (k_extract_split):
    # same result as the one-ended walk and the oracle, for exact, short and over-long output slots.
    S, H, seed = 700, 24, 9
    img = synth.bubble_chain(S, H, seed)
    g, e = orc.GBWT.load(img.array), b200.GBWT.from_bytes(img.array)
    ids = np.arange(2 * H + 1, dtype=np.uint64)
    first = e.extract(ids)                      # lengths measured here (sequence_lengths) -> two-ended walks
    again = e.extract(ids)
    monkeypatch.setenv("GBWT_B200_EXTRACT_SPLIT", "0")
    plain = e.extract(ids)
    for a, b in zip(first, plain):
        assert np.array_equal(a, b)
    for a, b in zip(again, plain):
        assert np.array_equal(a, b)
    monkeypatch.delenv("GBWT_B200_EXTRACT_SPLIT")
    for i in range(0, 2 * H, 5):
        assert np.array_equal(first[1][int(first[0][i]):int(first[0][i + 1])], np.array(list(g.sequence(i)), dtype=np.uint64))
    # slots that are too short / too long: every result is cut to its slot, lengths report the full size
    import ctypes as C
    lib = b200.library()
    L = 2 * S + 1
    for slot in (L - 1, L // 2, L // 2 + 1, 3, 0, L + 7):
        offsets = (np.arange(len(ids) + 1, dtype=np.uint64) * np.uint64(slot))
        nodes = np.full(int(offsets[-1]) + 8, 0xABCD, dtype=np.uint64)
        lengths = np.zeros(len(ids), dtype=np.uint64)
        rc = lib.gbwt_b200_extract(e._h, ids.ctypes.data_as(C.c_void_p), len(ids), offsets.ctypes.data_as(C.c_void_p),
                                   nodes.ctypes.data_as(C.c_void_p), lengths.ctypes.data_as(C.c_void_p))
        assert rc == 0
        assert np.all(lengths[:-1] == L) and lengths[-1] == np.uint64(2**64 - 1)
        for i in range(2 * H):
            want = plain[1][int(plain[0][i]):int(plain[0][i + 1])][:slot]
            got = nodes[i * slot:(i + 1) * slot]
            assert np.array_equal(got[:len(want)], want)   # (the rest of an over-long slot is unspecified)
        assert np.all(nodes[int(offsets[-1]):] == 0xABCD)


def test_two_ended_extraction_needs_mirror_image_strands(b200, monkeypatch):
    # An index flagged bidirectional whose odd sequences are NOT the reverse strands of the even ones. Sequences are only
    # walked from both ends when the signatures of both strands (taken by earlier whole walks) say that they are mirror
    # images; everything else is walked from the front, like the reference walks it. Three kinds of pairs:
    # unrelated strands, strands of equal length that agree around the middle (where the halves would meet) but differ
    # in ONE node elsewhere, and true mirror images.
    rng = random.Random(6)

    def mirror(p):
        return [x ^ 1 for x in reversed(p)]

    paths = []
    for kind in ("unrelated", "one node", "one node", "mirror", "one node near an end", "mirror"):
        p = [2 * rng.randint(1, 40) + rng.randint(0, 1) for _ in range(rng.choice([150, 200, 333]))]
        q = mirror(p)
        if kind == "unrelated":
            q = [2 * rng.randint(1, 40) + rng.randint(0, 1) for _ in range(len(p))]
        elif kind == "one node":
            at = len(q) // 5
            q[at] = q[at] ^ 2 if (q[at] ^ 2) >= 2 else q[at] + 2
        elif kind == "one node near an end":
            q[-2] = q[-2] + 2
        paths += [p, q]
    for i in range(2, 10, 2):   # the one-node pairs agree where the halves would meet
        p, q = paths[i], paths[i + 1]
        if i != 6:
            assert q[len(p) - 1 - len(p) // 2] ^ 1 == p[len(p) // 2] and q != mirror(p)
    img = image_of(gb.build_bwt(paths), bidirectional=True)
    g = orc.GBWT.load(img)
    ids = np.arange(len(paths), dtype=np.uint64)
    for checkpoints in (False, True):
        if checkpoints:
            monkeypatch.setenv("GBWT_B200_EXTRACT_CHECKPOINTS", "0")   # built (signatures of every sequence), but walk the chains
        e = b200.GBWT.from_bytes(img, checkpoints=checkpoints)
        for _ in range(3):          # lengths + signatures, then (where allowed) two-ended, and again
            offsets, nodes, lengths = e.extract(ids)
            for i, p in enumerate(paths):
                assert [int(x) for x in nodes[int(offsets[i]):int(offsets[i + 1])]] == p == [int(x) for x in g.sequence(i)]


def test_lean_find_kernel_edge_cases(b200, monkeypatch):
    # A dense-only index takes the lean find/extend kernel: every way a pattern can fail, pattern lengths around the
    # 4-node sectors the reader works in, and the same answers from the general kernel.
    S, H, seed = 300, 40, 17
    img = synth.bubble_chain(S, H, seed)
    g, e = orc.GBWT.load(img.array), b200.GBWT.from_bytes(img.array)
    assert e.device_bytes()["records_run8"] == 0
    rng = np.random.default_rng(8)
    for k in (1, 2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 64):
        pats = synth.patterns(S, H, seed, n=3000, k=k) if k <= 2 * S + 1 else None
        bad = pats.copy()
        rows = np.arange(len(bad))
        cols = rng.integers(0, k, len(bad))
        kind = rng.integers(0, 6, len(bad))
        values = np.array([0, 1, 2**32, 2**63, g.alphabet_size(), g.alphabet_size() + 5], dtype=np.uint64)
        bad[rows, cols] = values[kind]
        swapped = pats.copy()
        swapped[rows, cols] ^= np.uint64(1)          # the other strand of one node
        skipped = pats.copy()
        skipped[:, k // 2:] = np.roll(skipped[:, k // 2:], 1, axis=1)   # a valid node in the wrong place
        for batch in (pats, bad, swapped, skipped):
            want = g.find_extend_batch(batch)
            assert pc.states_equal(e.find_extend(batch), want)
            # the same rows as ragged patterns cut to assorted lengths (unaligned starts for the sector reader)
            lens = rng.integers(0, k + 1, len(batch))
            offsets = np.zeros(len(batch) + 1, dtype=np.uint64)
            np.cumsum(lens, out=offsets[1:])
            flat = np.concatenate([batch[r, :lens[r]] for r in range(len(batch))]) if offsets[-1] else np.zeros(0, np.uint64)
            want_ragged = g.find_extend_ragged(flat, offsets)
            assert pc.states_equal(e.find_extend_ragged(flat, offsets), want_ragged)
            monkeypatch.setenv("GBWT_B200_FIND_LEAN", "0")
            assert pc.states_equal(e.find_extend(batch), want)
            assert pc.states_equal(e.find_extend_ragged(flat, offsets), want_ragged)
            monkeypatch.delenv("GBWT_B200_FIND_LEAN")


@pytest.mark.parametrize("layout", ["auto", "runs"])
def test_serialize_round_trip(b200, layout, tmp_path):
    # GBWT::serialize through the C ABI: the layout comes back from HBM and is re-encoded on the host. Same image as
    # the CPU run of the writer (tests/test_layout_writer.py checks that one against the reference's files), loads in
    # the oracle and in the product, and gives the same answers.
    from hostsim_build import HostSim
    for name in ("example.gbwt", "with-empty.gbwt", "translation.gbz"):
        raw = open(os.path.join(GOLDEN, name), "rb").read()
        e = b200.GBWT.from_bytes(raw, layout=layout)
        image = e.serialize()
        assert image == HostSim(raw, 0 if layout == "auto" else 1).serialize()
        src, back = orc.GBWT.load(raw), orc.GBWT.load(image)
        assert back.bwt_data() == src.bwt_data() and np.array_equal(back.record_starts(), src.record_starts())
        path = str(tmp_path / (name + ".out"))
        e.save(path)
        assert open(path, "rb").read() == image
        pc.check_everything(b200.GBWT.load(path, layout=layout), src)
    S, H, seed = 400, 100, 6
    img = synth.bubble_chain(S, H, seed)
    e = b200.GBWT.from_bytes(img.array, layout=layout)
    back = orc.GBWT.load(e.serialize())
    assert back.bwt_data() == orc.GBWT.load(img.array).bwt_data()


def test_extraction_of_many_short_sequences(b200, monkeypatch):
    # more sequences than warps worth running: one sequence per thread (k_extract_lanes), and the same batch with a
    # forced thread stride; repeated ids are fine
    rng = random.Random(3)
    from test_hostsim_layout import random_paths
    paths = random_paths(rng, n_nodes=9, n_paths=60, max_len=14)
    img = image_of(gb.build_bwt(gb.bidirectional_sequences(paths + [[2, 4]])))
    g, e = orc.GBWT.load(img), b200.GBWT.from_bytes(img)
    ids = np.array([rng.randrange(g.sequences() + 1) for _ in range(70_000)], dtype=np.uint64)
    want_offsets, want_nodes = g.extract_batch(ids)
    offsets, nodes, lengths = e.extract(ids)
    assert np.array_equal(offsets, want_offsets) and np.array_equal(nodes, want_nodes)
    assert np.array_equal(lengths == np.uint64(2**64 - 1), ids >= g.sequences())
    for stride in ("1", "4", "32"):
        monkeypatch.setenv("GBWT_B200_EXTRACT_STRIDE", stride)
        offsets, nodes, _ = e.extract(ids[:5000])
        monkeypatch.delenv("GBWT_B200_EXTRACT_STRIDE")
        w_off, w_nodes = g.extract_batch(ids[:5000])
        assert np.array_equal(offsets, w_off) and np.array_equal(nodes, w_nodes)


@pytest.mark.parametrize("force", ["3", "2"])
def test_forced_find_kernels_on_mixed_indexes(b200, force, monkeypatch):
    # GBWT_B200_FIND_LEAN=3 forces the lean loop with its out-of-line step for run-length / high-outdegree records
    # (find_mixed.cu), 2 the general kernel: both must agree with the oracle on indexes full of such records.
    monkeypatch.setenv("GBWT_B200_FIND_LEAN", force)
    from test_hostsim_layout import random_paths, wide_record_index, records_image
    for name in FIXTURES:
        raw = open(os.path.join(GOLDEN, name), "rb").read()
        g, e = orc.GBWT.load(raw), b200.GBWT.from_bytes(raw)
        pc.check_find_all_nodes(e, g)
        pc.check_find_extend_random(e, g, n=3000, k=4, seed=11)
    for seed in range(4):
        rng = random.Random(300 + seed)
        paths = random_paths(rng, n_nodes=rng.choice([3, 6, 12]), n_paths=rng.choice([10, 40]), max_len=rng.choice([8, 20]))
        img = image_of(gb.build_bwt(gb.bidirectional_sequences(paths + [[2, 4]])))
        for layout in ("auto", "runs"):
            g, e = orc.GBWT.load(img), b200.GBWT.from_bytes(img, layout=layout)
            pc.check_find_extend_random(e, g, n=4000, k=5, seed=seed)
            seqs = [p for p in pc.all_sequences(g) if len(p) >= 3]
            pats = np.array([p[j:j + 3] for p in seqs for j in range(len(p) - 2)], dtype=np.uint64)
            if len(pats):
                assert pc.states_equal(e.find_extend(pats), g.find_extend_batch(pats))
    for sigma, bits in [(3, (1, 4, 10)), (64, (1, 5)), (255, (1, 9)), (300, (1, 12))]:
        rng = random.Random(sigma)
        edges, runs, total = wide_record_index(sigma, rng, bits)
        img, _ = records_image(edges, runs, sequences=total, size=3 * total, offset=0)
        g, e = orc.GBWT.load(img), b200.GBWT.from_bytes(img)
        pats = np.array([[1, 2 + v] for v in range(sigma)] + [[1, 1], [1, 0], [1, sigma + 5]], dtype=np.uint64)
        assert pc.states_equal(e.find_extend(pats), g.find_extend_batch(pats))
