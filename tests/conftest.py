import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_available():
    try:
        import target_estimation_b200 as te
        return te.lib.te_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # no test may hang a GPU box: a per-test wall-clock limit (pytest-timeout; the slowest test -- 4096 x 2000 against two oracles --
    # takes under a minute).  A thread-method timeout ends the whole process, and with it any resident kernel.
    if config.pluginmanager.hasplugin("timeout"):
        for it in items:
            if "gpu" in it.keywords and it.get_closest_marker("timeout") is None:
                it.add_marker(pytest.mark.timeout(420, method="thread"))
    # -m gpu on a box without a GPU must fail loudly, not skip: only auto-skip when no -m was given
    if config.getoption("-m"):
        return
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    from tests import report
    report.dump()
