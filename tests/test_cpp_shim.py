"""The C++ shim (include/manta_b200_shim.hpp) mirrors the reference call seams: it must compile against the
library on any box, and replay the reference's known-answer vectors on the GPU box."""
import os
import subprocess

import pytest

import manta_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "shim_test")


def build():
    if not os.path.exists(manta_b200.library_path()):
        import __graft_entry__
        __graft_entry__.build()
    src = os.path.join(ROOT, "tests", "cpp", "shim_test.cpp")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < os.path.getmtime(src):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                               manta_b200.library_path(), "-Wl,-rpath," + os.path.dirname(manta_b200.library_path())])


def test_shim_compiles_and_links():
    build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_shim_known_answers():
    build()
    out = subprocess.check_output([EXE]).decode()
    assert "shim_test ok" in out


def test_shim_sv_alignment_info_vs_reference():
    """mb200::SVAlignmentInfo (host arithmetic of row a10 in the shim) against the reference class through oracle/_ref"""
    ref = os.path.join(ROOT, "oracle", "_ref", "libmanta_ref.so")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref is not built here")
    exe = os.path.join(ROOT, "tests", "cpp", "align_info_test")
    src = os.path.join(ROOT, "tests", "cpp", "align_info_test.cpp")
    hdr = os.path.join(ROOT, "include", "manta_b200_shim.hpp")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-ldl"])
    assert "align_info_test ok" in subprocess.check_output([exe, ref]).decode()


def test_example_driver_compiles_and_links():
    """examples/staged_driver.cpp (the INTEGRATION.md call sequence through the C ABI and the shim) builds against the library"""
    build()
    src = os.path.join(ROOT, "examples", "staged_driver.cpp")
    exe = os.path.join(ROOT, "examples", "staged_driver")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", exe, manta_b200.library_path(),
                           "-Wl,-rpath," + os.path.dirname(manta_b200.library_path())])
    assert os.path.exists(exe)

