#!/usr/bin/env python3
"""Golden vectors for the oracle, produced by an INDEPENDENT dense numpy restatement of the four models
(dense A(dt), dense C = [I 0], numpy matmul, numpy.linalg.inv, simple-form
P = (I - K C) P).  It shares no code with oracle/ (C++) or the CUDA kernels; agreement is expected to
~1e-12 relative, not bit-exact (different summation order / LAPACK inverse).

    python tests/golden/make_golden.py          # rewrites tests/golden/kf_golden.npz

Inputs are the seeded streams of tests/synth.py and models/*.yaml (bit-identical to the reference's
model files, tests/test_models.py).  meas_rpy_internal_ starts at 0.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests import synth  # noqa: E402

PI = np.pi


def load_model(name):
    import yaml
    node = yaml.safe_load(open(os.path.join(ROOT, "models", "model_%s_params.yaml" % name)))
    mats = []
    for k in ("Q", "R", "P"):
        v = np.array(node[k], dtype=np.float64)
        s = int(np.sqrt(v.size))
        mats.append(v.reshape(s, s).T.copy())   # Eigen::Map column-major
    return mats


def constrain(x):
    x = np.fmod(x + PI, 2 * PI)
    if x < 0:
        x += 2 * PI
    return x - PI


def angle_conv(a):
    return np.fmod(constrain(a), 2 * PI)


def angle_diff(a, b):
    d = np.fmod(b - a + PI, 2 * PI)
    if d < 0:
        d += 2 * PI
    return d - PI


def unwrap(prev, new):
    return np.array([prev[i] - angle_diff(new[i], angle_conv(prev[i])) for i in range(3)])


def quat_to_rpy(q):
    x, y, z, w = q
    s = -2 * (x * z - w * y)
    if s > 0.9999:
        return np.array([0.0, PI / 2, 2 * np.arctan2(z, w)])
    if s < -0.9999:
        return np.array([0.0, -PI / 2, 2 * np.arctan2(z, w)])
    return np.array([np.arctan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z), np.arcsin(s),
                     np.arctan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z)])


def pose7_to_pose6(p):
    q = np.array(p[3:7]) / np.linalg.norm(p[3:7])
    return np.concatenate([p[:3], quat_to_rpy(q)])


def ear_base_inv(rpy):
    cr, sr, cp, sp = np.cos(rpy[0]), np.sin(rpy[0]), np.cos(rpy[1]), np.sin(rpy[1])
    return np.array([[1, sp * sr / cp, cr * sp / cp], [0, cr, -sr], [0, sr / cp, cr / cp]])


def jac_rpy(rpy, w, dt):
    cr, sr, cp, sp = np.cos(rpy[0]), np.sin(rpy[0]), np.cos(rpy[1]), np.sin(rpy[1])
    wy, wz = w[1], w[2]
    return np.array([[dt * (wy * cr * sp - wz * sp * sr) / cp + 1, dt * (wz * cr + wy * sr) / (cp * cp), 0],
                     [-dt * (wz * cr + wy * sr), 1, 0],
                     [dt * (wy * cr - wz * sr) / cp, dt * sp * (wz * cr + wy * sr) / (cp * cp), 1]])


class Filter:
    def __init__(self, model, Q, R, P0, p0):
        self.model = model
        self.Q, self.R = Q, R
        self.P = P0.copy()
        self.n, self.m = Q.shape[0], R.shape[0]
        self.C = np.hstack([np.eye(self.m), np.zeros((self.m, self.n - self.m))])
        self.x = np.zeros(self.n)
        if model in ("uniform_velocity", "uniform_acceleration"):
            self.x[:3] = p0[:3]
        else:
            self.x[:6] = pose7_to_pose6(p0)
        self.prev_rpy = np.zeros(3)
        self.t = 0.0
        self.n_meas = 0

    def A(self, dt):
        n = self.n
        A = np.eye(n)
        if self.model == "uniform_velocity":
            A += dt * np.eye(n, k=3)
        elif self.model == "uniform_acceleration":
            A += dt * np.eye(n, k=3) + 0.5 * dt * dt * np.eye(n, k=6)
        elif self.model == "angular_rates":
            A += dt * np.eye(n, k=6) + 0.5 * dt * dt * np.eye(n, k=12)
        else:
            rpy, w = self.x[3:6], self.x[9:12]
            A[0:3, 6:9] = dt * np.eye(3)
            A[3:6, 3:6] = jac_rpy(rpy, w, dt)
            A[3:6, 9:12] = dt * ear_base_inv(rpy)
        return A

    def f(self, dt):
        if self.model != "angular_velocities":
            return self.A(dt) @ self.x
        x = self.x.copy()
        x[0:3] += dt * self.x[6:9]
        x[3:6] += (dt * ear_base_inv(self.x[3:6])) @ self.x[9:12]
        return x

    def step(self, dt, meas, update):
        A = self.A(dt)
        xp = self.f(dt)
        P = A @ self.P @ A.T + self.Q
        if update:
            self.n_meas += 1
            if self.m == 3:
                y = meas[:3].copy()
            else:
                q = meas[3:7] / np.linalg.norm(meas[3:7])
                un = unwrap(self.prev_rpy, quat_to_rpy(q))
                self.prev_rpy = un
                y = np.concatenate([meas[:3], un])
            S = self.C @ P @ self.C.T + self.R
            K = P @ self.C.T @ np.linalg.inv(S)
            xp = xp + K @ (y - self.C @ xp)
            P = (np.eye(self.n) - K @ self.C) @ P
        self.x, self.P = xp, P
        self.t += dt


def main():
    dt = 1.0 / 250.0
    out = {}
    for name in ("uniform_velocity", "uniform_acceleration", "angular_velocities", "angular_rates"):
        Q, R, P0 = load_model(name)
        n_t, n_k = 3, 300
        meas, action, scale = synth.make_streams(n_t, n_k, dt, seed=20240607, accel=name in ("uniform_acceleration", "angular_rates"),
                                                 angular=R.shape[0] == 6)
        rec_at = [0, 1, 9, 99, 299]
        xs = np.zeros((n_t, len(rec_at), Q.shape[0]))
        Ps = np.zeros((n_t, len(rec_at), Q.shape[0], Q.shape[0]))
        for i in range(n_t):
            f = Filter(name, Q, R, scale[i] * P0, meas[0, i])
            for k in range(n_k):
                f.step(dt, meas[k, i], action[k, i] == 2)
                if k in rec_at:
                    xs[i, rec_at.index(k)] = f.x
                    Ps[i, rec_at.index(k)] = f.P
        out[name + "/x"] = xs
        out[name + "/P"] = Ps
        out[name + "/meas"] = meas
        out[name + "/action"] = action
        out[name + "/scale"] = scale
        out[name + "/rec_at"] = np.array(rec_at)
    np.savez_compressed(os.path.join(HERE, "kf_golden.npz"), **out)
    print("wrote kf_golden.npz", {k: v.shape for k, v in out.items() if k.endswith("/x")})


if __name__ == "__main__":
    main()
