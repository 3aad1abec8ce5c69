"""Runs the C++ host-side KAT program (tests/cpp/test_kats.cpp): the reference's gtest suites
restated against the B200-native ControllerBase / ModelBase / CostBase classes."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "mppi_tf_b200", "_build", "test_kats")


def test_cpp_kat_program_is_built():
    assert os.path.exists(EXE), "run __graft_entry__.build()"


@pytest.mark.gpu
def test_cpp_kats():
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " 0 failures" in r.stdout
