"""-m gpu: the C++ host surface (TargetManager / TargetInterface / EstimatorView / IntersectionSolver / TickTargetManager of
include/target_estimation_b200/target_manager.hpp) compiled into a small program that reads like the reference's own
test/target_manager_test.cpp and linked against lib/libtarget_c.so -- what a C++ caller of the reference switches to."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_api_program(tmp_path):
    lib = os.path.join(ROOT, "target_estimation_b200", "lib")
    exe = str(tmp_path / "api_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "api_test.cpp"),
                           "-o", exe, "-L", lib, "-ltarget_c", "-lte_pool", "-Wl,-rpath," + lib])
    r = subprocess.run([exe, os.path.join(ROOT, "models")], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and r.stdout.strip().splitlines()[-1].startswith("ok:")
