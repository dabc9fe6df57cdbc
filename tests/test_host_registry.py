"""not gpu: IdRegistry (sorted flat registry of target ids in the host library) against std::map on random single and batched
operations -- tests/registry_check.cpp, compiled here and linked against lib/libtarget_c.so (no device call is made)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_id_registry_matches_std_map(tmp_path):
    lib = os.path.join(ROOT, "target_estimation_b200", "lib")
    if not os.path.exists(os.path.join(lib, "libtarget_c.so")):
        pytest.skip("libtarget_c.so not built")
    exe = str(tmp_path / "registry_check")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "registry_check.cpp"),
                           "-o", exe, "-L", lib, "-ltarget_c", "-lte_pool", "-Wl,-rpath," + lib])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().startswith("ok:"), r.stdout + r.stderr
