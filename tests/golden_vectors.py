"""Known-answer values transcribed from the reference's own unit tests and doc-tests.

Every constant cites the gbwt-rs file:line that asserts it (crate `gbz` 0.5.1). These are
literal expected values (data), not code. See SURVEY.md App. B.
"""

# ---- codecs -----------------------------------------------------------------------------
# src/support.rs:1042-1045 (ByteCode doc-test)
BYTECODE_VALUES = [123, 456, 789]
BYTECODE_BYTES = bytes([123, 72 + 128, 3, 21 + 128, 6])
# src/support.rs:1186-1190 (RLE doc-test), sigma = 4
RLE4_RUNS = [(3, 12), (2, 721), (0, 34)]
RLE4_BYTES = bytes([3 + 4 * 11, 2 + 4 * 63, 17 + 128, 5, 0 + 4 * 33])
# src/support/tests.rs:452-469
RLE_ROUNDTRIP_SIGMAS = [(591, 4), (366, 254), (421, 255), (283, 14901), (330, 0)]
RLE_THRESHOLD_SIGMAS = [1, 4, 5, 128, 129, 254]
# src/support/tests.rs:471-513 (hand-assembled record)
RECORD_SIGMA = 4
RECORD_EDGES = [(0, 0), (13, 7), (22, 1), (44, 0)]

# ---- BWT examples -----------------------------------------------------------------------
# src/bwt/tests.rs:10-32 and src/bwt.rs:10-19: the GBWT example from the paper
PAPER_EDGES = [
    [(1, 0)],
    [(2, 0), (3, 0)],
    [(4, 0), (5, 0)],
    [(4, 1)],
    [(5, 1), (6, 0)],
    [(7, 0)],
    [(7, 2)],
    [(0, 0)],
]
PAPER_RUNS = [
    [(0, 3)],
    [(0, 2), (1, 1)],
    [(0, 1), (1, 1)],
    [(0, 1)],
    [(1, 1), (0, 1)],
    [(0, 2)],
    [(0, 1)],
    [(0, 3)],
]
PAPER_INVALID_NODE = 8
# src/bwt.rs:22-35 (module doc-test)
PAPER_DOC = dict(len=8, rec=2, outdegree=2, successor1=5, offset1=0, rec_len=2, lf1=(5, 0),
                 follow=((0, 2), 5, (0, 1)), total_len=17)

# src/bwt/tests.rs:35-87: bidirectional version of the example
BIDIR_EDGES = [
    [(2, 0), (15, 0)],
    [(4, 0), (6, 0)], [(0, 0)],
    [(8, 0), (10, 0)], [(3, 0)],
    [(8, 1)], [(3, 2)],
    [(10, 1), (12, 0)], [(5, 0), (7, 0)],
    [(14, 0)], [(5, 1), (9, 0)],
    [(14, 2)], [(9, 1)],
    [(0, 0)], [(11, 0), (13, 0)],
]
BIDIR_RUNS = [
    [(0, 3), (1, 3)],
    [(0, 2), (1, 1)], [(0, 3)],
    [(0, 1), (1, 1)], [(0, 2)],
    [(0, 1)], [(0, 1)],
    [(1, 1), (0, 1)], [(1, 1), (0, 1)],
    [(0, 2)], [(0, 1), (1, 1)],
    [(0, 1)], [(0, 1)],
    [(0, 3)], [(1, 1), (0, 2)],
]
BIDIR_INVALID_NODE = 16
# src/bwt/tests.rs:313-318: records 2 and 6 of the paper example emptied
EMPTY_RECORDS = (2, 6)

# ---- fixture statistics -----------------------------------------------------------------
# src/gbwt/tests.rs:44-56; src/gbwt.rs:55-58
STATS = {
    "example.gbwt": dict(len=68, sequences=12, alphabet_size=52, alphabet_offset=21),
    "with-empty.gbwt": dict(len=70, sequences=14, alphabet_size=52, alphabet_offset=21),
}


def encode_node(node_id, reverse=False):
    """support::encode_node, src/support.rs:155."""
    return 2 * node_id + (1 if reverse else 0)


def flip_node(node):
    """support::flip_node, src/support.rs:188."""
    return node ^ 1


def reverse_path(path):
    """support::reverse_path, src/support.rs:310-314."""
    return [flip_node(x) for x in reversed(path)]


def true_paths(with_empty):
    """src/gbwt/tests.rs:116-162 (original-graph ids -> GBWT nodes)."""
    F, R = False, True
    e = encode_node
    result = [
        [e(11, F), e(12, F), e(14, F), e(15, F), e(17, F)],
        [e(21, F), e(22, F), e(24, F), e(25, F)],
        [e(11, F), e(12, F), e(14, F), e(15, F), e(17, F)],
        [e(11, F), e(13, F), e(14, F), e(16, F), e(17, F)],
    ]
    if with_empty:
        result.append([])
    result.append([e(21, F), e(22, F), e(24, F), e(23, R), e(21, R)])
    result.append([e(21, F), e(22, F), e(24, F), e(25, F)])
    return result


def true_nodes():
    """src/gbwt/tests.rs:242-250."""
    out = set()
    for n in [11, 12, 13, 14, 15, 16, 17, 21, 22, 23, 24, 25]:
        out.add(encode_node(n, False))
        out.add(encode_node(n, True))
    return out


def count_occurrences(paths, subpath):
    """src/gbwt/tests.rs:252-266 (brute force over both orientations)."""
    result = 0
    rev = reverse_path(subpath)
    k = len(subpath)
    for path in paths:
        for i in range(len(path)):
            if path[i:i + k] == subpath:
                result += 1
            if i + 1 >= k and path[i + 1 - k:i + 1] == rev:
                result += 1
    return result


# src/gbwt.rs:546-548 (SequenceIter doc-test): path 3 reverse
SEQ7 = [35, 33, 29, 27, 23]
# src/gbwt.rs:60-83 (GBWT doc-test)
DOC_FIND = dict(nodes=[encode_node(12), encode_node(14), encode_node(15)], final_node=encode_node(15), len=2)
DOC_BD = dict(first=encode_node(14), back=encode_node(12), fwd=encode_node(15),
              forward_node=encode_node(15), reverse_node=encode_node(12, True), len=2)
# src/gbz/tests.rs:375-379: translation.gbz paths (GBWT nodes)
TRANSLATION_PATHS = [
    [2, 4, 6, 10, 12, 18, 22],
    [2, 4, 6, 10, 12, 18, 22],
    [2, 4, 8, 10, 12, 20, 22],
]
# src/gbz.rs:1189-1208 (StateIter doc-test on example.gbz): (from, to, len), F = forward
DOC_STATEITER = dict(node=14, len=3, successors=2,
                     predecessors=[((12, False), (15, False), 2), ((13, False), (16, False), 1)])

# src/graph/tests.rs:21-34 (example.gg; the same Graph is embedded in example.gbz) and :66-79 (translation).
GRAPH_SEQUENCES = ["G", "A", "T", "T", "A", "C", "A", "", "", "", "G", "A", "T", "T", "A"]
GRAPH_SEQUENCES_TRANSLATION = ["GA", "T", "T", "A", "CA", "G", "", "", "A", "T", "TA"]
# src/support/tests.rs:13-42
REVERSE_COMPLEMENTS = [(b"", b""), (b"C", b"G"), (b"GATTACA", b"TGTAATC"), (b"GATTACAT", b"ATGTAATC")]


def complement_table():
    """support::COMPLEMENT, src/support.rs:87-98 (restated as data for the tests)."""
    table = bytearray(b"N" * 256)
    for a, b in (("A", "T"), ("C", "G"), ("G", "C"), ("T", "A")):
        table[ord(a)] = ord(b)
        table[ord(a.lower())] = ord(b)
    return bytes(table)


def true_dna(node_sequence, path, endmarker=b"\x00"):
    """extract_sequence of src/bin/gbz-extract.rs:173-189 over a list of GBWT node identifiers."""
    table = complement_table()
    out = bytearray()
    for node in path:
        seq = node_sequence(node // 2)
        out += seq[::-1].translate(table) if node & 1 else seq
    return bytes(out) + endmarker

# src/gbz/tests.rs:237-247 (example.gbz) and :332-342 (translation.gbz): (node id, sequence) of every node.
GBZ_NODES = [(11, "G"), (12, "A"), (13, "T"), (14, "T"), (15, "A"), (16, "C"), (17, "A"),
             (21, "G"), (22, "A"), (23, "T"), (24, "T"), (25, "A")]
GBZ_NODES_TRANSLATION = [(1, "GA"), (2, "T"), (3, "T"), (4, "A"), (5, "CA"), (6, "G"), (9, "A"), (10, "T"), (11, "TA")]
