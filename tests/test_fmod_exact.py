"""not gpu: te::fmod_exact (csrc/te_fmod.h, used by the device-side angle wrap / unwrap) is bit-identical to C fmod --
the reference's unwrap branches (geometry.hpp:31-76) depend on it."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize("flags", [[], ["-DTE_FMOD_RECIPROCAL"]], ids=["division", "reciprocal-estimate (device path)"])
def test_fmod_exact_matches_libm(tmp_path, flags):
    exe = str(tmp_path / "fmod_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off"] + flags + ["-I", os.path.join(ROOT, "target_estimation_b200", "csrc"),
                           "-o", exe, os.path.join(ROOT, "tests", "fmod_check.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "0 mismatches" in out.stdout
