"""examples/staged_driver.cpp on the GPU (kept last in collection order: it is a demo, the parity tests come first)."""
import os
import subprocess

import pytest

import test_cpp_shim

ROOT = test_cpp_shim.ROOT


@pytest.mark.gpu
def test_example_driver_runs():
    test_cpp_shim.test_example_driver_compiles_and_links()
    out = subprocess.check_output([os.path.join(ROOT, "examples", "staged_driver")]).decode()
    assert "bestAlignmentIndex: 0" in out and "140=120D104= at 560, deletion ref[700,820) cipos [0,0]" in out
    assert "leftSize 70 rightSize 80 mismatches 0/0" in out
