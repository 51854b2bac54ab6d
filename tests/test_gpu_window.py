"""GPU parity tests for the round-2 kernels, through the C ABI against the CPU oracle, bit-exact:

* the record-window search kernel (find_window.cu: records staged into shared memory with bulk copies, two-hop
  steps) with windows small enough that queries leave them, bodies that do not fit, record formats it defers, every
  way a pattern can fail, and 32-bit patterns;
* checkpointed path extraction (k_build_checkpoints / k_extract_checkpointed) against the chain walks and the oracle.
"""
import os
import random

import numpy as np
import pytest

import gbwt_builder as gb
import parity_checks as pc
from oracle import oracle as orc
from synth import synth

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def b200():
    import gbwt_rs_b200
    return gbwt_rs_b200


def image_of(b, bidirectional=True):
    flags = 4 | (1 if bidirectional else 0)
    return synth.gbwt_image(b["sequences"], b["size"], b["offset"], b["alphabet_size"], flags, b["starts"], b["data"])


def force_windows(monkeypatch, **knobs):
    monkeypatch.setenv("GBWT_B200_LOCALITY", "1")        # sort even tiny batches on tiny indexes
    monkeypatch.setenv("GBWT_B200_FIND_WINDOW", "2")     # windows on any index they can run on
    monkeypatch.setenv("GBWT_B200_WINDOW_STATS", "1")
    for k, v in knobs.items():
        monkeypatch.setenv("GBWT_B200_" + k, str(v))


def damaged_batches(g, pats, rng):
    """The ways a pattern can fail: a node that is no node, the other strand, a valid node in the wrong place."""
    n, k = pats.shape
    rows, cols = np.arange(n), rng.integers(0, k, n)
    bad = pats.copy()
    values = np.array([0, 1, 2**32, 2**63, g.alphabet_size(), g.alphabet_size() + 5], dtype=np.uint64)
    bad[rows, cols] = values[rng.integers(0, 6, n)]
    swapped = pats.copy()
    swapped[rows, cols] ^= np.uint64(1)
    skipped = pats.copy()
    skipped[:, k // 2:] = np.roll(skipped[:, k // 2:], 1, axis=1)
    return [pats, bad, swapped, skipped]


@pytest.mark.parametrize("threads", [256, 512, 1024])
def test_window_kernel_matches_oracle(b200, threads, monkeypatch):
    force_windows(monkeypatch, WINDOW_THREADS=threads)
    S, H, seed = 3000, 64, 42
    img = synth.bubble_chain(S, H, seed)
    g, e = orc.GBWT.load(img.array), b200.GBWT.from_bytes(img.array)
    info = e.window_info()
    assert info["can_run"] and info["default"] and info["threads"] == threads
    pats = synth.patterns(S, H, seed, n=200_000, k=32)
    out = e.find_extend(pats)
    assert pc.states_equal(out, g.find_extend_batch(pats))
    assert np.all(out["end"] > out["start"]) and np.array_equal(out["node"], pats[:, -1])
    info = e.window_info()
    assert info["queries"] >= 200_000 and info["deferred"] * 100 < info["queries"], info   # answered from shared memory
    assert pc.states_equal(e.find_extend_u32(pats.astype(np.uint32)), out)


@pytest.mark.parametrize("k", [2, 7, 16, 17, 40, 64, 100])
def test_window_kernel_pattern_lengths(b200, k, monkeypatch):
    """Patterns shorter and longer than the 32 nodes an index's windows are planned for (a batch gets margins for its own
    length), lengths that are not a whole number of 16-node segments or 32-byte sectors, 64- and 32-bit nodes."""
    force_windows(monkeypatch)
    S, H, seed = 4000, 64, 5
    img = synth.bubble_chain(S, H, seed)
    g, e = orc.GBWT.load(img.array), b200.GBWT.from_bytes(img.array)
    pats = synth.patterns(S, H, seed, n=30_000, k=k)
    rng = np.random.default_rng(k)
    if k > 1:
        rows = rng.integers(0, len(pats), 3000)
        pats[rows, rng.integers(1, k, 3000)] ^= np.uint64(1)  # a third of the damaged ones fail somewhere in the middle
    want = g.find_extend_batch(pats)
    w0 = e.window_info()
    assert pc.states_equal(e.find_extend(pats), want)
    assert pc.states_equal(e.find_extend_u32(pats.astype(np.uint32)), want)
    w1 = e.window_info()
    assert w1["queries"] - w0["queries"] == 2 * len(pats)      # both went through the windows ...
    assert w1["deferred"] - w0["deferred"] < len(pats) // 4    # ... and mostly stayed there


# ---- bidirectional searches from record windows (k_bd_window) ---------------------------------------------------------

@pytest.mark.parametrize("layout", ["auto", "runs"])
def test_bd_window_kernel_on_fixtures_and_random_graphs(b200, layout, monkeypatch):
    """Every (first, start, end) of every sequence (src/gbwt/tests.rs:365-462) with the window kernel forced: fixtures (two
    orientations of a node among a record's successors, empty records), random graphs with high outdegrees (deferred)."""
    force_windows(monkeypatch)
    for name in ("example.gbwt", "with-empty.gbwt", "translation.gbz"):
        raw = open(os.path.join(GOLDEN, name), "rb").read()
        g, e = orc.GBWT.load(raw), b200.GBWT.from_bytes(raw, layout=layout)
        assert pc.check_bd(e, g) > 0
    from test_hostsim_layout import random_paths
    from test_run_checkpoints import sparse_paths
    for seed in range(6):
        rng = random.Random(300 + seed)
        paths = sparse_paths(rng, rng.choice([6, 10, 40]), rng.choice([50, 200])) if seed % 2 else \
            random_paths(rng, n_nodes=rng.choice([3, 12, 40]), n_paths=rng.choice([10, 40]), max_len=rng.choice([8, 40, 70]))
        if not any(paths):
            paths.append([2, 4])
        img = image_of(gb.build_bwt(gb.bidirectional_sequences(paths)))
        g, e = orc.GBWT.load(img), b200.GBWT.from_bytes(img, layout=layout)
        pc.check_bd(e, g)


@pytest.mark.parametrize("model", [dict(), dict(alt_ppm=50_000, tri_mod=4)])
def test_bd_window_kernel_on_bubble_chains(b200, model, monkeypatch):
    """Random (first, start, end) on sampled subpaths of a bubble chain: paired successors never occur here, plain bubbles
    always; subpaths longer than the tile and tri-allelic sites (records the window does not decode) are deferred; damaged
    paths fail where the oracle fails."""
    force_windows(monkeypatch)
    S, H, seed = 3000, 64, 17
    img = synth.bubble_chain(S, H, seed, **model)
    g, e = orc.GBWT.load(img.array), b200.GBWT.from_bytes(img.array)
    rng = np.random.default_rng(8)
    for k in (32, 9, 40):
        n = 40_000
        pats = synth.patterns(S, H, seed, n=n, k=k, **model)
        bad = rng.integers(0, n, n // 10)
        pats[bad, rng.integers(0, k, len(bad))] ^= np.uint64(1)
        first = rng.integers(0, k, size=n).astype(np.uint64)
        start = (first * rng.random(n)).astype(np.uint64)
        end = (first + 1 + ((k - first - 1) * rng.random(n)).astype(np.uint64)).astype(np.uint64)
        offs = np.arange(n + 1, dtype=np.uint64) * k
        flat = pats.reshape(-1)
        want = g.bd_search_batch(flat, offs, first, start, end)
        assert pc.states_equal(e.bd_search(flat, offs, first, start, end), want)
        assert np.count_nonzero(want["forward"]["end"] > want["forward"]["start"]) > n // 2


@pytest.mark.parametrize("direct", ["1", "0"])
def test_window_placement_of_a_crowded_batch(b200, direct, monkeypatch):
    """A batch whose queries all start in a few windows: with the one-pass placement (every window owns twice its share of
    the batch) most of them find their window full and are finished by the general kernels; with the exact sort they all
    fit. Same results either way, for find/extend and for bidirectional searches."""
    force_windows(monkeypatch)
    monkeypatch.setenv("GBWT_B200_WINDOW_DIRECT", direct)
    S, H, seed = 6000, 64, 31
    img = synth.bubble_chain(S, H, seed)
    g, e = orc.GBWT.load(img.array), b200.GBWT.from_bytes(img.array)
    assert e.window_info()["windows"] >= 32
    n, k = 60_000, 32
    rng = np.random.default_rng(12)
    seqs = [synth.sequence(S, H, seed, i) for i in range(8)]
    pats = np.empty((n, k), dtype=np.uint64)
    for q in range(n):
        p = seqs[q % 8]
        t = int(rng.integers(0, 300)) if q % 8 % 2 == 0 else int(rng.integers(len(p) - 300 - k, len(p) - k))  # (both strands start near record 0)
        pats[q] = p[t:t + k]
    want = g.find_extend_batch(pats)
    w0 = e.window_info()
    assert pc.states_equal(e.find_extend(pats), want)
    assert pc.states_equal(e.find_extend_u32(pats.astype(np.uint32)), want)
    w1 = e.window_info()
    if direct == "1":
        assert w1["deferred"] - w0["deferred"] > n      # (most of both batches overflowed)
    first = rng.integers(0, k, size=n).astype(np.uint64)
    start = (first * rng.random(n)).astype(np.uint64)
    end = (first + 1 + ((k - first - 1) * rng.random(n)).astype(np.uint64)).astype(np.uint64)
    offs = np.arange(n + 1, dtype=np.uint64) * k
    flat = pats.reshape(-1)
    assert pc.states_equal(e.bd_search(flat, offs, first, start, end), g.bd_search_batch(flat, offs, first, start, end))


def test_window_kernel_edge_cases(b200, monkeypatch):
    force_windows(monkeypatch)
    S, H, seed = 400, 40, 17
    img = synth.bubble_chain(S, H, seed)
    g, e = orc.GBWT.load(img.array), b200.GBWT.from_bytes(img.array)
    rng = np.random.default_rng(8)
    for k in (1, 2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 64):
        pats = synth.patterns(S, H, seed, n=3000, k=k)
        for batch in damaged_batches(g, pats, rng):
            want = g.find_extend_batch(batch)
            assert pc.states_equal(e.find_extend(batch), want), k
            narrow = batch[np.all(batch < 2**32, axis=1)]
            assert pc.states_equal(e.find_extend_u32(narrow.astype(np.uint32)), g.find_extend_batch(narrow)), k
    assert e.window_info()["queries"] > 0
    pc.check_find_extend_random(e, g, n=30_000, k=7, seed=2)


@pytest.mark.parametrize("knobs", [dict(WINDOW=32, WINDOW_MARGIN=32), dict(WINDOW=64, WINDOW_MARGIN=0),
                                   dict(WINDOW_SMEM_KB=32, WINDOW=128, WINDOW_MARGIN=64)])
def test_window_kernel_defers_what_it_cannot_answer(b200, knobs, monkeypatch):
    # windows smaller than a pattern's reach, no margin at all, and too little shared memory for the 1024-haplotype
    # bodies: the deferred list and the general kernel must give the same answers
    force_windows(monkeypatch, **knobs)
    S, H, seed = 500, 1024, 3
    img = synth.bubble_chain(S, H, seed)
    g, e = orc.GBWT.load(img.array), b200.GBWT.from_bytes(img.array)
    pats = synth.patterns(S, H, seed, n=60_000, k=32)
    rng = np.random.default_rng(1)
    for batch in damaged_batches(g, pats, rng):
        assert pc.states_equal(e.find_extend(batch), g.find_extend_batch(batch))
    info = e.window_info()
    assert info["deferred"] > 0 and info["queries"] >= 4 * 60_000, info


@pytest.mark.parametrize("layout", ["auto", "runs"])
def test_window_kernel_on_fixtures_and_random_graphs(b200, layout, monkeypatch):
    # indexes the windows are not meant for (run-length bodies, outdegree > 2, empty records, tiny): forced on, same answers
    force_windows(monkeypatch)
    for name in ("example.gbwt", "with-empty.gbwt", "translation.gbz"):
        raw = open(os.path.join(GOLDEN, name), "rb").read()
        g, e = orc.GBWT.load(raw), b200.GBWT.from_bytes(raw, layout=layout)
        pc.check_find_extend_subpaths(e, g)
        pc.check_find_extend_random(e, g, n=4001, k=4, seed=11)
        pats = np.array([[2**40, 22], [0, 0], [22, 2**63]], dtype=np.uint64)
        assert pc.states_equal(e.find_extend(pats), g.find_extend_batch(pats))
        assert e.window_info()["queries"] > 0
    from test_hostsim_layout import random_paths
    for seed in range(4):
        rng = random.Random(seed)
        paths = random_paths(rng, n_nodes=rng.choice([6, 12, 40]), n_paths=rng.choice([10, 40]), max_len=rng.choice([8, 20]))
        if not any(paths):
            paths.append([2, 4])
        img = image_of(gb.build_bwt(gb.bidirectional_sequences(paths)))
        g, e = orc.GBWT.load(img), b200.GBWT.from_bytes(img, layout=layout)
        pc.check_find_extend_subpaths(e, g)
        pc.check_find_extend_random(e, g, n=5000, k=5, seed=seed)


def test_u32_patterns_without_windows(b200, monkeypatch):
    # the 32-bit entry point on the plain kernels (small batch: no sort) and on the sorted general kernel
    S, H, seed = 600, 64, 5
    img = synth.bubble_chain(S, H, seed)
    pats = synth.patterns(S, H, seed, n=50_000, k=32)
    for layout in ("auto", "runs"):
        g, e = orc.GBWT.load(img.array), b200.GBWT.from_bytes(img.array, layout=layout)
        want = g.find_extend_batch(pats)
        assert pc.states_equal(e.find_extend_u32(pats.astype(np.uint32)), want)
        monkeypatch.setenv("GBWT_B200_LOCALITY", "1")
        monkeypatch.setenv("GBWT_B200_FIND_WINDOW", "0")
        assert pc.states_equal(e.find_extend_u32(pats.astype(np.uint32)), want)
        monkeypatch.delenv("GBWT_B200_LOCALITY")
        monkeypatch.delenv("GBWT_B200_FIND_WINDOW")
    assert len(e.find_extend_u32(np.zeros((0, 32), dtype=np.uint32))) == 0


# ---- checkpointed extraction ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("window", ["0", "2"])   # one-lane walks from global memory / CTAs walking staged record windows
@pytest.mark.parametrize("shift", [6, 8])
def test_checkpointed_extraction_matches_chain_walks(b200, shift, window, monkeypatch):
    monkeypatch.setenv("GBWT_B200_CHECKPOINT_SHIFT", str(shift))
    monkeypatch.setenv("GBWT_B200_EXTRACT_WINDOW", window)
    S, H, seed = 1500, 48, 9
    img = synth.bubble_chain(S, H, seed)
    g = orc.GBWT.load(img.array)
    e = b200.GBWT.from_bytes(img.array, checkpoints=True)
    plain = b200.GBWT.from_bytes(img.array, checkpoints=False)
    info = e.checkpoint_info()
    assert info["present"] and info["interval"] == 1 << shift and info["max_segments"] >= (2 * S + 1) >> shift
    assert not plain.checkpoint_info()["present"]
    ids = np.concatenate([np.arange(2 * H, dtype=np.uint64), np.array([2 * H, 2**40, 3, 3], dtype=np.uint64)])
    o_off, o_nodes = g.extract_batch(ids)
    for engine in (e, plain):
        offsets, nodes, lengths = engine.extract(ids)
        assert np.array_equal(offsets, o_off) and np.array_equal(nodes, o_nodes)
    assert np.array_equal(e.sequence_lengths(ids), plain.sequence_lengths(ids))
    for i in (0, 1, 2 * H - 1):
        assert np.array_equal(np.array(list(e.sequence(i)), dtype=np.uint64), synth.sequence(S, H, seed, i))
    # slots shorter and longer than the sequences: truncated, and the slack comes back zeroed
    lengths = e.sequence_lengths(ids[:6])
    caps = np.array([0, 5, int(lengths[2]), int(lengths[3]) + 7, 1, 64], dtype=np.uint64)
    offs = np.zeros(7, dtype=np.uint64)
    np.cumsum(caps, out=offs[1:])
    nodes = np.full(int(offs[-1]), 77, dtype=np.uint64)
    got_len = np.zeros(6, dtype=np.uint64)
    lib = b200.library()
    rc = lib.gbwt_b200_extract(e._h, ids[:6].ctypes.data, 6, offs.ctypes.data, nodes.ctypes.data, got_len.ctypes.data)
    assert rc == 0 and np.array_equal(got_len, lengths)
    for i in range(6):
        full = synth.sequence(S, H, seed, int(ids[i]))
        a, b = int(offs[i]), int(offs[i + 1])
        n = min(b - a, len(full))
        assert np.array_equal(nodes[a:a + n], full[:n]) and np.all(nodes[a + n:b] == 0)


@pytest.mark.parametrize("window", ["0", "2"])
@pytest.mark.parametrize("layout", ["auto", "runs"])
def test_checkpointed_extraction_on_fixtures_and_random_graphs(b200, layout, window, monkeypatch):
    monkeypatch.setenv("GBWT_B200_CHECKPOINT_SHIFT", "6")
    monkeypatch.setenv("GBWT_B200_EXTRACT_WINDOW", window)
    for name in ("example.gbwt", "with-empty.gbwt", "translation.gbz"):
        raw = open(os.path.join(GOLDEN, name), "rb").read()
        g, e = orc.GBWT.load(raw), b200.GBWT.from_bytes(raw, layout=layout, checkpoints=True)
        assert e.checkpoint_info()["present"]
        pc.check_navigation(e, g)
    from test_hostsim_layout import random_paths
    for seed in range(4):
        rng = random.Random(100 + seed)
        paths = random_paths(rng, n_nodes=rng.choice([3, 12, 40]), n_paths=rng.choice([10, 40]), max_len=rng.choice([8, 200, 700]))
        if not any(paths):
            paths.append([2, 4])
        img = image_of(gb.build_bwt(gb.bidirectional_sequences(paths)))
        g, e = orc.GBWT.load(img), b200.GBWT.from_bytes(img, layout=layout, checkpoints=True)
        ids = np.arange(g.sequences() + 2, dtype=np.uint64)
        o_off, o_nodes = g.extract_batch(ids)
        offsets, nodes, lengths = e.extract(ids)
        assert np.array_equal(offsets, o_off) and np.array_equal(nodes, o_nodes)


@pytest.mark.parametrize("window", ["0", "2"])
def test_checkpoints_on_an_index_with_invalid_edge_targets(b200, window, monkeypatch):
    monkeypatch.setenv("GBWT_B200_EXTRACT_WINDOW", window)
    # a damaged index (an edge to a node beyond the alphabet): the build walk and the segment walks take their
    # bounds-checked instantiation and stop where the reference's iterator stops
    edges = [[(1, 0)], [(2, 0), (40, 0)], [(0, 0)]]
    runs = [[(0, 2)], [(1, 1), (0, 1)], [(0, 1)]]
    g0 = orc.GBWT.from_records(edges, runs, sequences=2, size=6, offset=0)
    img = synth.gbwt_image(2, 6, 0, 3, 4, g0.record_starts(), g0.bwt_data())
    g, e = orc.GBWT.load(img), b200.GBWT.from_bytes(img, checkpoints=True)
    assert e.checkpoint_info()["present"]
    ids = np.arange(3, dtype=np.uint64)
    o_off, o_nodes = g.extract_batch(ids)
    offsets, nodes, lengths = e.extract(ids)
    assert np.array_equal(offsets, o_off) and np.array_equal(nodes, o_nodes) and lengths[2] == np.uint64(2**64 - 1)


def test_window_extraction_of_many_sequences(b200, monkeypatch):
    """Enough sequences for several 512-lane CTAs per segment (the default dispatch), both strands mixed in one batch,
    tri-allelic sites (records the windows do not decode: those steps come from global memory), slots of every length."""
    monkeypatch.setenv("GBWT_B200_CHECKPOINT_SHIFT", "7")
    for model in (dict(), dict(alt_ppm=50_000, tri_mod=4)):
        S, H, seed = 700, 700, 21
        img = synth.bubble_chain(S, H, seed, **model)
        g, e = orc.GBWT.load(img.array), b200.GBWT.from_bytes(img.array, checkpoints=True)
        assert e.checkpoint_info()["present"]
        rng = np.random.default_rng(4)
        ids = rng.permutation(2 * H).astype(np.uint64)
        ids[5] = np.uint64(2 * H + 3)   # not a sequence
        o_off, o_nodes = g.extract_batch(ids)
        offsets, nodes, lengths = e.extract(ids)
        assert np.array_equal(offsets, o_off) and np.array_equal(nodes, o_nodes)
        monkeypatch.setenv("GBWT_B200_EXTRACT_WINDOW", "0")
        offsets0, nodes0, lengths0 = e.extract(ids)
        monkeypatch.delenv("GBWT_B200_EXTRACT_WINDOW")
        assert np.array_equal(nodes, nodes0) and np.array_equal(lengths, lengths0)


# ---- build once, replicate (CUDA IPC export / import) ----------------------------------------------------------------

def _ipc_child(blob, pats, ids, conn):
    import numpy as np
    import gbwt_rs_b200 as b200
    try:
        e = b200.GBWT.import_ipc(blob, device=0)
        offsets, nodes, lengths = e.extract(ids)
        conn.send(("ok", e.find_extend(pats), offsets, nodes, e.checkpoint_info(), e.window_info()["can_run"], e.serialize()))
    except Exception as exc:  # pragma: no cover
        conn.send(("error", repr(exc)))
    conn.close()


def test_exported_index_is_imported_by_another_process(b200, monkeypatch):
    # rank 0 builds, the other ranks import: the arrays are copied device to device through CUDA IPC handles
    import multiprocessing as mp
    monkeypatch.setenv("GBWT_B200_CHECKPOINT_SHIFT", "6")
    S, H, seed = 900, 32, 11
    img = synth.bubble_chain(S, H, seed)
    g, e = orc.GBWT.load(img.array), b200.GBWT.from_bytes(img.array, checkpoints=True)
    pats = synth.patterns(S, H, seed, n=30_000, k=32)
    ids = np.arange(2 * H, dtype=np.uint64)
    blob = e.export_ipc()
    ctx = mp.get_context("spawn")
    parent, child = ctx.Pipe()
    proc = ctx.Process(target=_ipc_child, args=(blob, pats, ids, child))
    proc.start()
    msg = parent.recv()
    proc.join(timeout=60)
    assert msg[0] == "ok", msg
    _, states, offsets, nodes, ckpt, can_run, image = msg
    assert pc.states_equal(states, g.find_extend_batch(pats))
    o_off, o_nodes = g.extract_batch(ids)
    assert np.array_equal(offsets, o_off) and np.array_equal(nodes, o_nodes)
    assert ckpt["present"] and ckpt["entries"] == e.checkpoint_info()["entries"] and can_run == 1
    assert image == e.serialize()
    with pytest.raises(IOError):
        b200.GBWT.import_ipc(b"not a blob" * 100)
