"""Builds and binds tests/hostsim (TEST-ONLY CPU run of the product's layout builder and one-lane query code).

This lets `-m "not gpu"` tests check the device layout (K0) and the scan logic of
gbwt-rs_b200/csrc/record_scan.cuh against the oracle without a GPU. The product library never links it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "hostsim", "hostsim.cpp"),
       os.path.join(ROOT, "gbwt-rs_b200", "csrc", "layout_builder.cpp"),
       os.path.join(ROOT, "gbwt-rs_b200", "csrc", "sds_loader.cpp"),
       os.path.join(ROOT, "gbwt-rs_b200", "csrc", "layout_writer.cpp")]
DEPS = SRC + [os.path.join(ROOT, "gbwt-rs_b200", "csrc", f) for f in
              ("layout.h", "layout_builder.h", "sds_loader.h", "layout_writer.h", "record_scan.cuh")] + \
       [os.path.join(ROOT, "include", "gbwt_b200.h")]
OUT = os.path.join(HERE, "hostsim", "libgbwt_hostsim.so")

STATE = np.dtype([("node", "<u8"), ("start", "<u8"), ("end", "<u8")])
BDSTATE = np.dtype([("forward", STATE), ("reverse", STATE)])
POS = np.dtype([("node", "<u8"), ("offset", "<u8")])

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        stale = (not os.path.exists(OUT)) or any(os.path.getmtime(s) > os.path.getmtime(OUT) for s in DEPS)
        if stale:
            subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                            "-o", OUT] + SRC + ["-ldl"], check=True, capture_output=True)
        L = C.CDLL(OUT)
        p, u64 = C.c_void_p, C.c_uint64
        L.hs_load.restype = p
        L.hs_load.argtypes = [p, C.c_size_t, C.c_int, C.c_char_p, C.c_size_t]
        L.hs_free.argtypes = [p]
        L.hs_serialize.restype = u64
        L.hs_serialize.argtypes = [p, C.c_int, p, u64]
        L.hs_skips.restype = p
        L.hs_skips.argtypes = [p]
        L.hs_edges_valid.argtypes = [p]
        L.hs_records.restype = u64
        L.hs_records.argtypes = [p]
        L.hs_desc_words.argtypes = [p, u64, p]
        L.hs_has_graph.argtypes = [p]
        L.hs_drop_graph_section.argtypes = [p]
        L.hs_label_count.restype = u64
        L.hs_label_count.argtypes = [p]
        L.hs_label_starts.restype = p
        L.hs_label_starts.argtypes = [p]
        L.hs_label_bytes.restype = p
        L.hs_label_bytes.argtypes = [p]
        L.hs_format_counts.argtypes = [p, p]
        L.hs_body_bytes.restype = u64
        L.hs_body_bytes.argtypes = [p]
        L.hs_checkpointed_records.restype = u64
        L.hs_checkpointed_records.argtypes = [p]
        L.hs_record_format.argtypes = [p, u64]
        L.hs_find.argtypes = [p, p, C.c_size_t, p]
        L.hs_extend.argtypes = [p, p, p, C.c_size_t, p]
        L.hs_find_extend.argtypes = [p, p, C.c_size_t, C.c_size_t, p]
        L.hs_bd_find.argtypes = [p, p, C.c_size_t, p]
        L.hs_bd_extend.argtypes = [p, p, p, C.c_size_t, C.c_int, p]
        L.hs_bd_search.argtypes = [p, p, p, p, p, p, C.c_size_t, p]
        L.hs_follow.argtypes = [p, p, C.c_size_t, C.c_int, p, p, p]
        L.hs_start.argtypes = [p, p, C.c_size_t, p]
        L.hs_forward.argtypes = [p, p, C.c_size_t, p]
        L.hs_backward.argtypes = [p, p, C.c_size_t, p]
        L.hs_sequence_lengths.argtypes = [p, p, C.c_size_t, p]
        L.hs_extract.argtypes = [p, p, C.c_size_t, p, p, p]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


class HostSim:
    """Same batched interface as gbwt_rs_b200.GBWT (see tests/parity_checks.py)."""

    name = "hostsim"

    def __init__(self, image, layout: int = 0):
        arr = np.ascontiguousarray(np.frombuffer(image, dtype=np.uint8) if not isinstance(image, np.ndarray) else image)
        err = C.create_string_buffer(256)
        self._h = lib().hs_load(_p(arr), arr.nbytes, layout, err, 256)
        if not self._h:
            raise IOError(err.value.decode())
        self._L = lib()

    def __del__(self):
        try:
            if self._h:
                self._L.hs_free(self._h)
                self._h = None
        except Exception:
            pass

    def serialize(self, gbz: bool = False) -> bytes:
        """The layout written back as a Simple-SDS GBWT (or GBZ) image (csrc/layout_writer.cpp)."""
        n = self._L.hs_serialize(self._h, int(gbz), None, 0)
        assert n > 0, "serialization failed"
        buf = np.zeros(n, dtype=np.uint8)
        assert self._L.hs_serialize(self._h, int(gbz), _p(buf), n) == n
        return buf.tobytes()

    def records(self):
        return self._L.hs_records(self._h)

    def edges_valid(self):
        return bool(self._L.hs_edges_valid(self._h))

    def skips(self):
        """(records, 4) uint32: {n0, o0, n1, o1} per record (IndexView::skips)."""
        n = self.records()
        return np.ctypeslib.as_array(C.cast(self._L.hs_skips(self._h), C.POINTER(C.c_uint32)), (n, 4)).copy()

    def desc_words(self, rec):
        out = np.zeros(8, dtype=np.uint32)
        self._L.hs_desc_words(self._h, rec, _p(out))
        return out

    def drop_graph_section(self):
        self._L.hs_drop_graph_section(self._h)

    def labels(self):
        """(starts, bytes) of the node labels the product loader parsed from a GBZ image, or None for a plain GBWT."""
        if not self._L.hs_has_graph(self._h):
            return None
        n = self._L.hs_label_count(self._h)
        starts = np.ctypeslib.as_array(C.cast(self._L.hs_label_starts(self._h), C.POINTER(C.c_uint64)), (n + 1,)).copy()
        total = int(starts[-1])
        data = np.ctypeslib.as_array(C.cast(self._L.hs_label_bytes(self._h), C.POINTER(C.c_uint8)), (total,)).copy() \
            if total else np.zeros(0, np.uint8)
        return starts, data

    def checkpointed_records(self):
        """Run bodies that carry a checkpoint table (layout.h)."""
        return int(self._L.hs_checkpointed_records(self._h))

    def format_counts(self):
        out = np.zeros(7, dtype=np.uint64)
        self._L.hs_format_counts(self._h, _p(out))
        return [int(x) for x in out]

    def record_format(self, rec):
        return self._L.hs_record_format(self._h, rec)

    def find(self, nodes):
        nodes = _u64(nodes); out = np.zeros(len(nodes), STATE)
        self._L.hs_find(self._h, _p(nodes), len(nodes), _p(out)); return out

    def extend(self, states, nodes):
        states = np.ascontiguousarray(states, STATE); nodes = _u64(nodes); out = np.zeros(len(nodes), STATE)
        self._L.hs_extend(self._h, _p(states), _p(nodes), len(nodes), _p(out)); return out

    def find_extend(self, patterns):
        patterns = _u64(patterns); n, k = patterns.shape; out = np.zeros(n, STATE)
        self._L.hs_find_extend(self._h, _p(patterns), n, k, _p(out)); return out

    def find_extend_ragged(self, nodes, offsets):
        nodes, offsets = _u64(nodes), _u64(offsets)
        out = np.zeros(len(offsets) - 1, STATE)
        for q in range(len(offsets) - 1):  # hostsim has no ragged entry point; reuse the fixed one per query
            pat = nodes[int(offsets[q]):int(offsets[q + 1])].reshape(1, -1)
            if pat.shape[1]:
                out[q] = self.find_extend(pat)[0]
        return out

    def bd_find(self, nodes):
        nodes = _u64(nodes); out = np.zeros(len(nodes), BDSTATE)
        self._L.hs_bd_find(self._h, _p(nodes), len(nodes), _p(out)); return out

    def extend_forward(self, states, nodes):
        states = np.ascontiguousarray(states, BDSTATE); nodes = _u64(nodes); out = np.zeros(len(nodes), BDSTATE)
        self._L.hs_bd_extend(self._h, _p(states), _p(nodes), len(nodes), 0, _p(out)); return out

    def extend_backward(self, states, nodes):
        states = np.ascontiguousarray(states, BDSTATE); nodes = _u64(nodes); out = np.zeros(len(nodes), BDSTATE)
        self._L.hs_bd_extend(self._h, _p(states), _p(nodes), len(nodes), 1, _p(out)); return out

    def bd_search(self, nodes, offsets, first, start, end):
        nodes, offsets, first, start, end = map(_u64, (nodes, offsets, first, start, end))
        out = np.zeros(len(first), BDSTATE)
        self._L.hs_bd_search(self._h, _p(nodes), _p(offsets), _p(first), _p(start), _p(end), len(first), _p(out))
        return out

    def follow(self, states, backward=False):
        states = np.ascontiguousarray(states, BDSTATE); n = len(states)
        counts = np.zeros(n, np.uint64)
        self._L.hs_follow(self._h, _p(states), n, int(backward), None, None, _p(counts))
        sizes = np.where(counts == np.uint64(2**64 - 1), np.uint64(0), counts)
        offsets = np.zeros(n + 1, np.uint64); np.cumsum(sizes, out=offsets[1:])
        out = np.zeros(int(offsets[-1]), BDSTATE)
        again = np.zeros(n, np.uint64)
        self._L.hs_follow(self._h, _p(states), n, int(backward), _p(offsets), _p(out), _p(again))
        assert np.array_equal(again, counts)
        return offsets, out, counts

    def start(self, ids):
        ids = _u64(ids); out = np.zeros(len(ids), POS)
        self._L.hs_start(self._h, _p(ids), len(ids), _p(out)); return out

    def forward(self, positions):
        positions = np.ascontiguousarray(positions, POS); out = np.zeros(len(positions), POS)
        self._L.hs_forward(self._h, _p(positions), len(positions), _p(out)); return out

    def backward(self, positions):
        positions = np.ascontiguousarray(positions, POS); out = np.zeros(len(positions), POS)
        if self._L.hs_backward(self._h, _p(positions), len(positions), _p(out)) != 0:
            raise AssertionError("Following sequences backward requires a bidirectional GBWT")
        return out

    def sequence_lengths(self, ids):
        ids = _u64(ids); out = np.zeros(len(ids), np.uint64)
        self._L.hs_sequence_lengths(self._h, _p(ids), len(ids), _p(out)); return out

    def extract(self, ids):
        ids = _u64(ids)
        lengths = self.sequence_lengths(ids)
        sizes = np.where(lengths == np.uint64(2**64 - 1), np.uint64(0), lengths)
        offsets = np.zeros(len(ids) + 1, np.uint64)
        np.cumsum(sizes, out=offsets[1:])
        nodes = np.zeros(int(offsets[-1]), np.uint64)
        got = np.zeros(len(ids), np.uint64)
        self._L.hs_extract(self._h, _p(ids), len(ids), _p(offsets), _p(nodes), _p(got))
        assert np.array_equal(got, lengths)
        return offsets, nodes, lengths
