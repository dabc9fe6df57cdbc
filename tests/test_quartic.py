"""not gpu: te::lowest_real_root4 (csrc/te_quartic.h, the device-side root finder of the batched IntersectionSolver) against
the oracle's restatement of Eigen's PolynomialSolver + smallestRealRoot (src/intersection_solver.cpp:4-17) on the solver's
own coefficient shapes, random quartics, prescribed roots and edge cases -- same source compiled for the host."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_quartic_matches_oracle(tmp_path):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    exe = str(tmp_path / "quartic_check")
    odir = os.path.join(ROOT, "oracle")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "target_estimation_b200", "csrc"), "-o", exe,
                           os.path.join(ROOT, "tests", "quartic_check.cpp"), "-L", odir, "-lte_oracle", "-Wl,-rpath," + odir])
    out = subprocess.run([exe, "300000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-3000:]
    assert "\n0 mismatches" in out.stdout
    # the physical coefficient shapes must not need the tolerance bucket for near-multiple roots at all
    phys = [l for l in out.stdout.splitlines() if l.startswith("physical")][0]
    assert " 0 mismatches, 0 ill-conditioned" in phys, phys
