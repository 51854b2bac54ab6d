"""CPU-only checks of the C ABI: the library loads, exports every symbol include/gbwt_b200.h declares,
and refuses to work without a GPU instead of falling back to the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gbwt_b200.h")).read()
    return sorted(set(re.findall(r"GBWT_B200_API\s+[^;(]*?\b(gbwt_b200_\w+)\s*\(", text)))


def test_header_declares_the_path():
    names = declared_symbols()
    for required in ["gbwt_b200_find", "gbwt_b200_extend", "gbwt_b200_find_extend", "gbwt_b200_bd_find",
                     "gbwt_b200_extend_forward", "gbwt_b200_extend_backward", "gbwt_b200_bd_search", "gbwt_b200_start",
                     "gbwt_b200_forward", "gbwt_b200_backward", "gbwt_b200_sequence_lengths", "gbwt_b200_extract",
                     "gbwt_b200_index_from_bytes", "gbwt_b200_index_load_file", "gbwt_b200_index_from_parts"]:
        assert required in names
    assert len(names) >= 40


def test_library_exports_every_declared_symbol():
    import gbwt_rs_b200 as gb
    lib = ctypes.CDLL(gb.LIBRARY)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} is declared in include/gbwt_b200.h but not exported"
    assert "sm_100a" in gb.version()


def test_cubin_targets_sm_100a():
    import shutil
    import subprocess
    import gbwt_rs_b200 as gb
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", gb.LIBRARY], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_a_device():
    import gbwt_rs_b200 as gb
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(gb.GBWTError) as info:
        gb.GBWT.load(os.path.join(ROOT, "tests", "golden", "example.gbwt"))
    assert info.value.code == gb.E_NO_DEVICE
    # malformed input is still reported as InvalidData before any device work
    with pytest.raises(IOError):
        gb.GBWT.from_bytes(b"\x00" * 64)


def test_shard_range_partitions():
    import gbwt_rs_b200 as gb
    for n in (0, 1, 7, 1000, 2**27 + 3):
        for world in (1, 2, 3, 8):
            blocks = [gb.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for a, b in zip(blocks, blocks[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_sass_of_the_hot_kernels():
    """The properties DESIGN.md claims for the hot kernels, read from the SASS of the built library: sector-sized
    256-bit loads in the search and walk kernels, warp shuffles + L2 prefetches in the walks, and no local-memory
    stack frame in the walk kernels (the descriptor must live in registers)."""
    import re
    import shutil
    import subprocess
    import gbwt_rs_b200 as gb
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", gb.LIBRARY], capture_output=True, text=True).stdout
    kernels = {}
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name is not None:
            kernels[name].append(line)

    def body(fragment):
        hits = [k for k in kernels if fragment in k]
        assert hits, f"no kernel matching {fragment}"
        return "\n".join(kernels[hits[0]])

    for fragment in ("k_find_extend_leanILi5", "k_find_extend_lean_mixed", "13k_find_extendILb0", "k_extract_splitILb0", "k_extract_dna_relayILb0"):
        assert "LDG.E.ENL2.256" in body(fragment), f"{fragment}: no 256-bit sector loads"
    for fragment in ("k_extractILb0", "k_extract_splitILb0", "k_extract_dna_relayILb0"):
        text = body(fragment)
        assert "SHFL" in text and "CCTL.E.PF2" in text, f"{fragment}: warp shuffles / L2 prefetches missing"
    log = open(os.path.join(os.path.dirname(gb.LIBRARY), "build_ptxas.log")).read()
    for fragment in ("k_extractILb0", "k_extract_splitILb0"):
        m = re.search(r"Function properties for \S*" + fragment + r"\S*\n\s*(\d+) bytes stack frame", log)
        assert m and int(m.group(1)) == 0, f"{fragment}: stack frame in the walk kernel"
