"""Brute-force GBWT construction from explicit paths (test support, pure Python, small inputs).

The reference ships no construction algorithm (gbwt-rs src/bwt.rs:207-210); its fixtures were
built by the C++ `gfa2gbwt`. This builder restates the GBWT definition -- the visits to a node are
sorted by their reversed prefixes, ties at the start of a sequence broken by sequence id -- and is
pinned by reproducing the BWT bytes and record starts of the reference's fixtures exactly
(tests/test_synth.py). It then serves as the truth for the closed-form bubble-chain generator.
"""
from __future__ import annotations


def varint(v: int) -> bytes:
    out = bytearray()
    while v > 0x7F:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def encode_run(sigma: int, value: int, length: int) -> bytes:
    if sigma >= 255:
        return varint(value) + varint(length - 1)
    threshold = 256 // sigma
    if length < threshold:
        return bytes([value + sigma * (length - 1)])
    return bytes([value + sigma * (threshold - 1)]) + varint(length - threshold)


def encode_record(edges, runs) -> bytes:
    """BWTBuilder::append layout (gbwt-rs src/bwt.rs:241-253)."""
    if not edges:
        return b"\x00"
    out = bytearray(varint(len(edges)))
    prev = 0
    for node, offset in edges:
        out += varint(node - prev) + varint(offset)
        prev = node
    for value, length in runs:
        out += encode_run(len(edges), value, length)
    return bytes(out)


def bidirectional_sequences(paths):
    """Sequence 2i = path i, 2i+1 = its reverse (gbwt-rs src/support.rs:310-314)."""
    seqs = []
    for p in paths:
        seqs.append(list(p))
        seqs.append([x ^ 1 for x in reversed(p)])
    return seqs


def build_records(seqs):
    """Returns (offset, alphabet_size, records) with records[r] = (edges, runs); record 0 = endmarker."""
    nodes = [x for s in seqs for x in s]
    if not nodes:
        return 0, 1, [([(0, 0)], [(0, len(seqs))])] if seqs else []
    offset = min(nodes) - 1
    alphabet_size = max(nodes) + 1
    # visits[v] = list of (key, successor); key = reversed prefix + (0, seq id)
    visits = {}
    for j, s in enumerate(seqs):
        full = [0] + list(s) + [0]  # endmarker, nodes, endmarker
        for p in range(len(full) - 1):
            v = full[p]
            succ = full[p + 1]
            if p == 0:
                key = (j,)
                visits.setdefault(0, []).append((key, succ))
            else:
                key = tuple(reversed(s[:p - 1])) + (0, j)
                visits.setdefault(v, []).append((key, succ))
    # number of visits to w from each predecessor v: for edge offsets
    into = {}
    for v, lst in visits.items():
        for _, succ in lst:
            into.setdefault(succ, {}).setdefault(v, 0)
            into[succ][v] += 1
    records = []
    for r in range(alphabet_size - offset):
        v = 0 if r == 0 else r + offset
        lst = visits.get(v)
        if not lst:
            records.append(([], []))
            continue
        lst.sort(key=lambda t: t[0])
        succs = sorted(set(s for _, s in lst))
        edges = []
        for w in succs:
            if w == 0:
                edges.append((0, 0))
            else:
                edges.append((w, sum(c for u, c in into[w].items() if u < v)))
        rank = {w: i for i, w in enumerate(succs)}
        runs = []
        for _, s in lst:
            if runs and runs[-1][0] == rank[s]:
                runs[-1][1] += 1
            else:
                runs.append([rank[s], 1])
        records.append((edges, [tuple(x) for x in runs]))
    return offset, alphabet_size, records


def build_bwt(seqs):
    """Returns dict(offset, alphabet_size, sequences, size, starts, data)."""
    offset, alphabet_size, records = build_records(seqs)
    data = bytearray()
    starts = []
    for edges, runs in records:
        starts.append(len(data))
        data += encode_record(edges, runs)
    return dict(offset=offset, alphabet_size=alphabet_size, sequences=len(seqs),
                size=sum(len(s) + 1 for s in seqs), starts=starts, data=bytes(data), records=records)
