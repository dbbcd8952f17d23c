"""The C++ host mirror of the reference interface (instaslice_b200/host) — compiled against the C ABI and run on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_host_mirror_selftest(tmp_path):
    pkg = os.path.join(ROOT, "instaslice_b200")
    assert os.path.exists(os.path.join(pkg, "libislhost.so")), "build() did not produce libislhost.so"
    exe = str(tmp_path / "host_mirror_selftest")
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host_mirror_selftest.cpp"),
                    "-L" + pkg, "-l:libislhost.so", "-l:libislplace.so", "-Wl,-rpath," + pkg], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "PASS" in out.stdout, out.stdout + out.stderr


def test_host_mirror_builds_and_links():
    """CPU-only: the mirror compiles and links against the C ABI library (no device call is made)."""
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "instaslice_b200", "libislhost.so"))
    assert lib is not None
