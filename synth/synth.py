"""ctypes binding for the synthetic bubble-chain GBWT generator (synth/gbwt_synth.c).

Test / benchmark input only: it produces Simple-SDS GBWT images and the
query patterns of SURVEY.md 8(d). Not part of the product and not the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    path = os.path.join(_HERE, "libgbwt_synth.so")
    src = [os.path.join(_HERE, f) for f in ("gbwt_synth.c", "Makefile")]
    stale = (not os.path.exists(path)) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return path


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u64, p = C.c_uint64, C.c_void_p
        L.synth_mix64.restype = u64
        L.synth_mix64.argtypes = [u64]
        L.synth_allele.restype = C.c_uint
        L.synth_allele.argtypes = [u64, u64, u64, u64]
        L.synth_bubble_chain_gbwt.restype = p
        L.synth_bubble_chain_gbwt.argtypes = [u64, u64, u64, C.c_int, C.POINTER(u64)]
        L.synth_bubble_chain_gbwt_v.restype = p
        L.synth_bubble_chain_gbwt_v.argtypes = [u64, u64, u64, u64, u64, C.c_int, C.POINTER(u64)]
        L.synth_allele_v.restype = C.c_uint
        L.synth_allele_v.argtypes = [u64, u64, u64, u64, u64, u64]
        L.synth_sequence_v.argtypes = [u64, u64, u64, u64, u64, u64, p]
        L.synth_patterns_v.argtypes = [u64, u64, u64, u64, u64, u64, u64, u64, u64, p, C.c_int]
        L.synth_gbwt_image.restype = p
        L.synth_gbwt_image.argtypes = [u64, u64, u64, u64, u64, p, u64, p, u64, C.POINTER(u64)]
        L.synth_bwt_section.restype = p
        L.synth_bwt_section.argtypes = [p, u64, p, u64, C.POINTER(u64)]
        L.synth_gbz_image.restype = p
        L.synth_gbz_image.argtypes = [p, u64, u64, u64, p, p, C.c_int, C.POINTER(u64)]
        L.synth_free.argtypes = [p]
        L.synth_sequence.argtypes = [u64, u64, u64, u64, p]
        L.synth_patterns.argtypes = [u64, u64, u64, u64, u64, u64, u64, p, C.c_int]
        _LIB = L
    return _LIB


class Image:
    """A malloc'd Simple-SDS GBWT image owned by the C library (zero-copy numpy view)."""

    def __init__(self, ptr: int, n: int):
        self.ptr, self.nbytes = ptr, n
        self.array = np.ctypeslib.as_array((C.c_uint8 * n).from_address(ptr))

    def tobytes(self) -> bytes:
        return C.string_at(self.ptr, self.nbytes)

    def free(self):
        if self.ptr:
            self.array = None
            lib().synth_free(self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def bubble_chain(sites: int, haplotypes: int, seed: int = 42, threads: int = 0, alt_ppm: int = 0, tri_mod: int = 0) -> Image:
    """Bubble chain with `sites` sites and `haplotypes` haplotypes. Default model: N = 3*sites + 1 nodes, two equally
    likely alleles. alt_ppm > 0 selects the variant model: N = 4*sites + 1 node ids, alternative alleles with
    probability alt_ppm / 10^6, every tri_mod-th site (by hash) tri-allelic."""
    n = C.c_uint64(0)
    if alt_ppm:
        ptr = lib().synth_bubble_chain_gbwt_v(sites, haplotypes, seed, alt_ppm, tri_mod, threads, C.byref(n))
    else:
        ptr = lib().synth_bubble_chain_gbwt(sites, haplotypes, seed, threads, C.byref(n))
    if not ptr:
        raise ValueError("invalid bubble-chain parameters")
    return Image(ptr, n.value)


def gbwt_image(sequences, size, offset, alphabet_size, flags, rec_starts, data: bytes) -> bytes:
    starts = np.ascontiguousarray(rec_starts, dtype=np.uint64)
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    n = C.c_uint64(0)
    ptr = lib().synth_gbwt_image(sequences, size, offset, alphabet_size, flags, starts.ctypes.data_as(C.c_void_p),
                                 len(starts), buf.ctypes.data_as(C.c_void_p), len(buf), C.byref(n))
    out = C.string_at(ptr, n.value)
    lib().synth_free(ptr)
    return out


def node_labels(n_labels: int, seed: int = 42, max_anchor: int = 32, first_node_id: int = 1):
    """Deterministic node labels for a bubble chain: (starts[n_labels + 1], bytes). Node ids 1, 4, 7, ... are the
    anchors between sites (1..max_anchor random bases), the two nodes after each anchor are the alleles of a site
    (single bases, SNP-like). Sprinkles lower-case bases and 'N' so that reverse complements exercise the whole
    COMPLEMENT table."""
    rng = np.random.default_rng(seed)
    ids = np.arange(n_labels, dtype=np.int64) + first_node_id
    lengths = np.where(ids % 3 == 1, rng.integers(1, max_anchor + 1, n_labels), 1).astype(np.uint64)
    starts = np.zeros(n_labels + 1, dtype=np.uint64)
    np.cumsum(lengths, out=starts[1:])
    total = int(starts[-1])
    data = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, total)].copy()
    odd = rng.integers(0, 64, total)
    data[odd == 0] |= 0x20          # lower case
    data[odd == 1] = ord("N")
    return starts, data


def gbz_image(gbwt_image, label_starts, label_bytes, graph_version: int = 4, nodes: int = 0, as_image: bool = False):
    """A GBZ image around a GBWT image (bytes or Image) with the given node labels; graph_version 3 = packed
    labels (GBZ v1 files), 4 = Zstandard (GBZ v2). Returns bytes, or with as_image the malloc'd Image (no copy;
    needed above 2 GiB)."""
    if isinstance(gbwt_image, Image):
        src_ptr, src_len = gbwt_image.ptr, gbwt_image.nbytes
    else:
        keep = np.ascontiguousarray(gbwt_image, dtype=np.uint8) if isinstance(gbwt_image, np.ndarray) \
            else np.frombuffer(bytes(gbwt_image), dtype=np.uint8)
        src_ptr, src_len = keep.ctypes.data, len(keep)
    starts = np.ascontiguousarray(label_starts, dtype=np.uint64)
    data = np.ascontiguousarray(label_bytes, dtype=np.uint8)
    n = C.c_uint64(0)
    ptr = lib().synth_gbz_image(src_ptr, src_len, nodes, len(starts) - 1, starts.ctypes.data_as(C.c_void_p),
                                data.ctypes.data_as(C.c_void_p), graph_version, C.byref(n))
    if not ptr:
        raise RuntimeError("could not write the GBZ image (libzstd missing?)")
    if as_image:
        return Image(ptr, n.value)
    out = C.string_at(ptr, n.value)
    lib().synth_free(ptr)
    return out


def bwt_section(rec_starts, data: bytes) -> bytes:
    starts = np.ascontiguousarray(rec_starts, dtype=np.uint64)
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    n = C.c_uint64(0)
    ptr = lib().synth_bwt_section(starts.ctypes.data_as(C.c_void_p), len(starts), buf.ctypes.data_as(C.c_void_p),
                                  len(buf), C.byref(n))
    out = C.string_at(ptr, n.value)
    lib().synth_free(ptr)
    return out


def sequence(sites: int, haplotypes: int, seed: int, seq_id: int, alt_ppm: int = 0, tri_mod: int = 0) -> np.ndarray:
    out = np.zeros(2 * sites + 1, dtype=np.uint64)
    lib().synth_sequence_v(sites, haplotypes, seed, alt_ppm, tri_mod if alt_ppm else 0, seq_id, out.ctypes.data_as(C.c_void_p))
    return out


def patterns(sites: int, haplotypes: int, seed: int, n: int, k: int = 32, seed_q: int = 7, q0: int = 0,
             threads: int = 0, out: np.ndarray | None = None, alt_ppm: int = 0, tri_mod: int = 0) -> np.ndarray:
    if out is None:
        out = np.empty((n, k), dtype=np.uint64)
    assert out.dtype == np.uint64 and out.flags.c_contiguous and out.size == n * k
    lib().synth_patterns_v(sites, haplotypes, seed, alt_ppm, tri_mod if alt_ppm else 0, seed_q, q0, n, k,
                           out.ctypes.data_as(C.c_void_p), threads)
    return out


# ---- device-side pattern generator (benchmark input only) --------------------------------------------

_CUDA_LIB = None


def build_cuda(force: bool = False) -> str:
    path = os.path.join(_HERE, "libgbwt_synth_cuda.so")
    src = os.path.join(_HERE, "gbwt_synth_cuda.cu")
    if force or not os.path.exists(path) or os.path.getmtime(src) > os.path.getmtime(path):
        subprocess.run(["make", "-C", _HERE, "-B", "libgbwt_synth_cuda.so"], check=True, capture_output=True)
    return path


def patterns_device(sites: int, haplotypes: int, seed: int, n: int, d_out: int, k: int = 32, seed_q: int = 7,
                    q0: int = 0, stream: int = 0, alt_ppm: int = 0, tri_mod: int = 0) -> None:
    """Writes patterns q0 .. q0+n (row-major, k u64 each) to the device address `d_out`."""
    global _CUDA_LIB
    if _CUDA_LIB is None:
        L = C.CDLL(build_cuda())
        u64 = C.c_uint64
        L.synth_patterns_device.argtypes = [u64, u64, u64, u64, u64, u64, u64, C.c_void_p, C.c_void_p]
        L.synth_patterns_device_v.argtypes = [u64, u64, u64, u64, u64, u64, u64, u64, u64, C.c_void_p, C.c_void_p]
        _CUDA_LIB = L
    rc = _CUDA_LIB.synth_patterns_device_v(sites, haplotypes, seed, alt_ppm, tri_mod if alt_ppm else 0, seed_q, q0, n, k, d_out, stream)
    if rc != 0:
        raise RuntimeError(f"synth_patterns_device failed with CUDA error {rc}")
