"""Runs tests/fuzz_loader.py for a few hundred damaged images in a child process (a crash or a hang there must not
take the test session down with it)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("seed", [1, 11])
def test_damaged_images_are_rejected_or_harmless(seed):
    proc = subprocess.run([sys.executable, os.path.join(HERE, "fuzz_loader.py"), str(seed), "300"], capture_output=True, text=True,
                          timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    loaded, rejected = (int(x) for x in proc.stdout.split()[1::2])
    assert loaded + rejected == 300 and rejected > 50 and loaded > 20


@pytest.mark.gpu
def test_damaged_images_on_the_gpu():
    # the same through the C ABI on cuda:0: every kernel family on whatever still loads (tests/fuzz_gpu.py)
    proc = subprocess.run([sys.executable, os.path.join(HERE, "fuzz_gpu.py"), "3", "200"], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr[-2000:]
    loaded, rejected = (int(x) for x in proc.stdout.split()[1::2])
    assert loaded + rejected == 200 and loaded > 20
