"""not gpu: pins the C++ oracle (oracle/) against (a) the committed golden vectors of the independent numpy
restatement (tests/golden/make_golden.py) and (b) the survey's sanity anchors."""
import os

import numpy as np
import pytest

from tests import orc, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODELS = {"uniform_velocity": 3, "uniform_acceleration": 2, "angular_velocities": 1, "angular_rates": 0}
DT = 1.0 / 250.0


@pytest.mark.parametrize("name", list(MODELS))
def test_oracle_matches_numpy_golden(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", "kf_golden.npz"))
    y = orc.load_yaml(os.path.join(ROOT, "models", "model_%s_params.yaml" % name))
    assert y["type"] == MODELS[name]
    meas, action, scale = g[name + "/meas"], g[name + "/action"], g[name + "/scale"]
    rec_at = list(g[name + "/rec_at"])
    n_k, n_t = action.shape
    N = y["Q"].shape[0]
    # the committed streams are the ones synth.make_streams regenerates (script + fixture stay in sync)
    m2, a2, s2 = synth.make_streams(n_t, n_k, DT, seed=20240607, accel=name in ("uniform_acceleration", "angular_rates"),
                                    angular=y["R"].shape[0] == 6)
    assert np.array_equal(m2, meas) and np.array_equal(a2, action) and np.array_equal(s2, scale)
    mgr = orc.Manager()
    ids = np.arange(n_t, dtype=np.uint32)
    for i in range(n_t):
        mgr.init_full(y["type"], i, DT, 0.0, y["Q"], y["R"], scale[i] * y["P"], meas[0, i])
    worst = 0.0
    for k in range(n_k):
        mgr.step_batch(ids, DT, meas[k], action[k])
        if k in rec_at:
            st = mgr.states(ids, N)
            j = rec_at.index(k)
            # independent implementations (LAPACK inverse, different summation order): 1e-9 bound with the H2 norm
            worst = max(worst, synth.compare_h2(st["x"], g[name + "/x"][:, j]), synth.compare_h2(st["P"], g[name + "/P"][:, j]))
    assert worst <= 1.0, worst


def _reftest(n_points=10000):
    L = orc.lib()
    meas = np.zeros((4, n_points, 7)); real = np.zeros((4, n_points, 7))
    L.orc_reftest_streams(DT, n_points, 4, orc.ptr(meas), orc.ptr(real))
    return meas, real


def test_libstdcxx_normal_anchor():
    out = np.zeros(3)
    orc.lib().orc_libstdcxx_normal(0.0, 0.01, 3, orc.ptr(out))
    assert np.allclose(out, [-0.0012196578414159691, -0.010868180442613574, 0.0068428994379655488], rtol=0, atol=1e-18)


@pytest.mark.parametrize("idx,name", [(0, "uniform_velocity"), (1, "uniform_acceleration"), (2, "angular_rates"), (3, "angular_velocities")])
def test_reference_convergence_scenarios(idx, name):
    """test/target_manager_test.cpp:148-340 re-run on the oracle with the libstdc++ noise stream (draw order
    UV -> UA -> AR -> AV) and the reference's own tolerances."""
    meas, real = _reftest()
    y = orc.load_yaml(os.path.join(ROOT, "models", "model_%s_params.yaml" % name))
    mgr = orc.Manager()
    mgr.init_full(y["type"], idx, DT, 0.0, y["Q"], y["R"], y["P"], meas[idx, 0])
    n = meas.shape[1]
    N = y["Q"].shape[0]
    pose = np.zeros((n, 7)); twist = np.zeros((n, 6))
    xs = np.zeros((1, N)); Ps = np.zeros((1, N, N))
    orc.lib().orc_run_stream(mgr.h, idx, n, DT, orc.ptr(np.ascontiguousarray(meas[idx])), None, n, orc.ptr(xs), orc.ptr(Ps),
                             orc.ptr(pose), orc.ptr(twist))
    goal = np.array([0.2, 0.3, 0.4])
    assert np.all(np.abs(pose[-1, :3] - goal) < 0.01)                       # :179-181
    assert np.all(np.abs(twist[:, :3].mean(axis=0) - goal / (n * DT)) < 0.01)   # :187-189
    if name == "angular_velocities":                                           # :335-340
        omega = np.array([3.0, 0.01, 0.1])
        assert np.all(np.abs(twist[:, 3:].mean(axis=0) - omega) < 0.05)
        assert np.all(np.abs(twist[-1, 3:] - omega) < 0.01)
    # survey anchors (numpy restatement, agreement ~1e-12)
    if name == "uniform_velocity":
        ref = [0.199805530179402, 0.299985270559087, 0.400546713092019, 0.004953635691446, 0.007470202951398, 0.010084439605856]
        assert np.allclose(xs[0], ref, rtol=1e-10, atol=1e-13)
        assert np.isclose(Ps[0, 0, 0], 1.787255412936363e-07, rtol=1e-10)
        assert np.isclose(Ps[0, 0, 3], 3.996424007207628e-08, rtol=1e-10)
    if name == "uniform_acceleration":
        assert np.allclose(xs[0, :3], [0.197372079593793, 0.301013208760892, 0.399536267889484], rtol=1e-10, atol=1e-13)
        assert np.isclose(Ps[0, 0, 0], 2.312066666930547e-06, rtol=1e-9)
        assert np.isclose(Ps[0, 0, 6], 9.883720621965695e-06, rtol=1e-9)


def test_first_step_covariance_anchor():
    y = orc.load_yaml(os.path.join(ROOT, "models", "model_uniform_velocity_params.yaml"))
    meas, _ = _reftest(10)
    mgr = orc.Manager()
    mgr.init_full(3, 0, DT, 0.0, y["Q"], y["R"], y["P"], meas[0, 0])
    mgr.update_meas(0, DT, meas[0, 0])
    assert np.isclose(mgr.state(0, 6)["P"][0, 0], 9.990010005977634e-05, rtol=1e-12)
