"""The C++ host-side mirror of the reference's API (include/mcf_network_simplex.hpp) and the reference's xUnit solver tests
restated on it (tests/cpp/test_network_simplex.cpp, built by csrc/Makefile): NetworkSimplexTests.cs:28-245, OptimizationTests.cs:14-70."""
import os
import subprocess

import pytest

from conftest import ROOT
import mincostflow_b200 as mcf

EXE = os.path.join(ROOT, "tests", "cpp", "test_network_simplex")


def _run(dimacs_dir, *args):
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "mincostflow_b200", "csrc"), "all"])
    return subprocess.run([EXE, *args, dimacs_dir], capture_output=True, text=True, timeout=600)


def test_cpp_mirror_builds_and_fails_loudly_without_a_device(dimacs_dir):
    if mcf.device_count() > 0:
        pytest.skip("a GPU is present: covered by the gpu test")
    r = _run(dimacs_dir, "--no-device")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout


@pytest.mark.gpu
def test_reference_xunit_tests_in_cpp(dimacs_dir):
    r = _run(dimacs_dir)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout
