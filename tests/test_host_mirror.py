"""Builds the C++ mirror of the reference API (gbwt-rs_b200/host/gbwt.hpp) against libgbwt_b200.so and runs the
reference's doc-test scenario through it. On a box without a GPU the binary must refuse to run (exit 77)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_mirror", "host_mirror_test.cpp")
EXE = os.path.join(ROOT, "tests", "host_mirror", "host_mirror_test")
LIBDIR = os.path.join(ROOT, "gbwt-rs_b200")


def build():
    import gbwt_rs_b200  # noqa: F401  builds the library if needed
    deps = [SRC, os.path.join(LIBDIR, "host", "gbwt.hpp"), os.path.join(ROOT, "include", "gbwt_b200.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-o", EXE, SRC, "-L" + LIBDIR, "-lgbwt_b200",
                        "-Wl,-rpath," + LIBDIR], check=True, capture_output=True)
    return EXE


def run():
    return subprocess.run([build(), os.path.join(ROOT, "tests", "golden")], capture_output=True, text=True)


def test_host_mirror_compiles_and_refuses_without_gpu():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    proc = run()
    if has_gpu:
        assert proc.returncode == 0, proc.stderr
    else:
        assert proc.returncode == 77, (proc.returncode, proc.stderr)


@pytest.mark.gpu
def test_host_mirror_doc_example():
    proc = run()
    assert proc.returncode == 0, proc.stderr
    assert "host mirror ok" in proc.stdout
