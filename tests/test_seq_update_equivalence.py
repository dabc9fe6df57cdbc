"""not gpu: the algebra behind the streamed angular-velocities kernel (csrc/te_av_sym.cuh av_update_seq).  With R = L L^T and
T = L^-1 the whitened measurement T y = (T C) x + e has unit, uncorrelated noise, so its components may be applied one after the
other: g = P h^T, s = h g + 1, x += g (T_i (y - x)) / s, P -= g g^T / s with h = row i of T C.  Here the independent numpy
restatement of the reference's filter (tests/golden/make_golden.py: joint update with the LAPACK inverse of S, src/kalman.cpp:135-140)
runs beside the same filter with the sequential update on the SURVEY.md 8(d) stream; they must agree far inside the 1e-9 contract."""
import os
import sys

import numpy as np

from tests import synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as g  # noqa: E402

DT = 1.0 / 250.0


class SeqFilter(g.Filter):
    def step(self, dt, meas, update):
        A = self.A(dt)
        xp = self.f(dt)
        P = A @ self.P @ A.T + self.Q
        if update:
            self.n_meas += 1
            q = meas[3:7] / np.linalg.norm(meas[3:7])
            un = g.unwrap(self.prev_rpy, g.quat_to_rpy(q))
            self.prev_rpy = un
            y = np.concatenate([meas[:3], un])
            T = np.linalg.inv(np.linalg.cholesky(self.R))     # lower triangular
            for i in range(6):
                h = np.zeros(12)
                h[:i + 1] = T[i, :i + 1]
                gv = P @ h
                s = h @ gv + 1.0
                r = T[i, :i + 1] @ (y[:i + 1] - xp[:i + 1])
                xp = xp + gv * (r / s)
                P = P - np.outer(gv, gv) / s
        self.x, self.P = xp, P
        self.t += dt


def test_sequential_whitened_update_equals_the_joint_update():
    Q, R, P0 = g.load_model("angular_velocities")
    n_t, n_k = 4, 600
    meas, action, scale = synth.make_streams(n_t, n_k, DT, seed=77, accel=False, angular=True, att_rate=0.5)
    worst = 0.0
    for i in range(n_t):
        a = g.Filter("angular_velocities", Q, R, scale[i] * P0, meas[0, i])
        b = SeqFilter("angular_velocities", Q, R, scale[i] * P0, meas[0, i])
        for k in range(n_k):
            a.step(DT, meas[k, i], action[k, i] == 2)
            b.step(DT, meas[k, i], action[k, i] == 2)
            if k % 100 == 99:
                worst = max(worst, synth.compare_h2(b.x[None], a.x[None]), synth.compare_h2(b.P[None], a.P[None]))
        assert a.n_meas == b.n_meas
    assert worst <= 0.1, worst      # (bar 1.0 = 1e-9 relative; measured ~1e-2)
