"""not gpu: the C-ABI libraries load and export every symbol include/*.h declares (no compute calls), the
product refuses to run without a CUDA device, and nothing in the product links or loads the oracle."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "target_estimation_b200", "lib")


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//.*", "", src)
    return sorted(set(re.findall(r"\b(te_[a-z0-9_]+|target_manager_[a-z0-9_]+)\s*\(", src)) - {"target_manager_c"})


@pytest.fixture(scope="module", autouse=True)
def built():
    import __graft_entry__ as g
    g.build()


@pytest.mark.parametrize("header,so", [("te_pool.h", "libte_pool.so"), ("target_manager_c.h", "libtarget_c.so")])
def test_exports(header, so):
    names = _declared(header)
    assert len(names) >= 10
    lib = C.CDLL(os.path.join(LIB, so))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_header():
    import target_estimation_b200._lib as L
    assert sorted(L.SIGNATURES) == _declared("te_pool.h")


def test_reference_abi_names_present():
    # the ten symbols of /root/reference/include/target_estimation/target_manager_c.h:28-37
    ref = ["target_manager_new", "target_manager_init", "target_manager_update_meas", "target_manager_update",
           "target_manager_get_est_pose", "target_manager_get_est_twist", "target_manager_get_est_acceleration",
           "target_manager_get_n_measurements", "target_manager_log", "target_manager_delete"]
    lib = C.CDLL(os.path.join(LIB, "libtarget_c.so"))
    for n in ref:
        assert hasattr(lib, n), n


def test_product_does_not_link_the_oracle():
    for so in ("libte_pool.so", "libtarget_c.so"):
        out = subprocess.run(["ldd", os.path.join(LIB, so)], capture_output=True, text=True).stdout
        assert "oracle" not in out
    for d in ("target_estimation_b200", "include"):
        for base, _, files in os.walk(os.path.join(ROOT, d)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                    txt = open(os.path.join(base, f)).read()
                    assert "te_oracle" not in txt and "libte_oracle" not in txt and "orc_" not in txt, os.path.join(base, f)


def test_no_cpu_fallback():
    import target_estimation_b200 as te
    if te.lib.te_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(te.TeError, match="no CUDA device"):
        te.TargetPool(te.UNIFORM_ACCELERATION)
    lib = C.CDLL(os.path.join(LIB, "libtarget_c.so"))
    lib.target_manager_new.restype = C.c_void_p
    lib.target_manager_new.argtypes = [C.c_char_p]
    lib.target_manager_init.argtypes = [C.c_void_p, C.c_uint, C.c_double, C.c_void_p, C.c_double]
    lib.target_manager_last_error.restype = C.c_char_p
    h = lib.target_manager_new(os.path.join(ROOT, "models", "model_uniform_velocity_params.yaml").encode())
    assert h   # the YAML loads on the host; the first init needs the device and must fail loudly, not compute on the CPU
    p0 = (C.c_double * 7)(0, 0, 0, 0, 0, 0, 1)
    lib.target_manager_init(h, 1, 0.004, p0, 0.0)
    assert b"no CUDA device" in lib.target_manager_last_error()
    assert lib.target_manager_new(b"/nonexistent.yaml") is None
