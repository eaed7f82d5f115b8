"""Builds and runs the C++ host mirror test (include/plonky2_b200.hpp over the C ABI) on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_host_mirror(glb, ctx, tmp_path):
    exe = str(tmp_path / "host_mirror_test")
    libdir = os.path.join(ROOT, "plonky2-lib_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), "-L", libdir, "-lgl_b200",
                           f"-Wl,-rpath,{libdir}", "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host_mirror_test ok" in out.stdout


def test_cpp_host_mirror_compiles(glb, tmp_path):
    """CPU-side: the header compiles and links against the C ABI (no compute call is made)."""
    obj = str(tmp_path / "host_mirror_test.o")
    subprocess.check_call(["g++", "-std=c++17", "-c", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), "-o", obj])
