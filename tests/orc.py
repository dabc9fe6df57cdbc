"""ctypes access to the CPU oracle (oracle/libte_oracle.so) -- TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ORACLE_DIR, "libte_oracle.so")


def build():
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("te_oracle.cpp", "oracle_c.cpp", "te_oracle.hpp", "Makefile")]
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return SO


_lib = None
_d, _i, _u, _p, _ll = C.c_double, C.c_int, C.c_uint, C.c_void_p, C.c_longlong


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.orc_manager_new.restype = _p; L.orc_manager_new.argtypes = [C.c_char_p]
        L.orc_manager_delete.argtypes = [_p]
        L.orc_load_yaml.restype = _i; L.orc_load_yaml.argtypes = [C.c_char_p, _p, _p, _p, _p, _p, _p, _p]
        L.orc_init_default.restype = _i; L.orc_init_default.argtypes = [_p, _u, _d, _p, _d]
        L.orc_init_full.argtypes = [_p, _i, _u, _d, _d, _p, _i, _p, _i, _p, _p, _p, _p]
        L.orc_update_meas.restype = _i; L.orc_update_meas.argtypes = [_p, _u, _d, _p]
        L.orc_update.restype = _i; L.orc_update.argtypes = [_p, _u, _d]
        L.orc_update_all.argtypes = [_p, _d]
        L.orc_erase.restype = _i; L.orc_erase.argtypes = [_p, _u]
        for f in ("orc_get_est_pose", "orc_get_est_twist", "orc_get_est_acceleration"):
            getattr(L, f).restype = _i; getattr(L, f).argtypes = [_p, _u, _p]
        L.orc_get_n_measurements.restype = _ll; L.orc_get_n_measurements.argtypes = [_p, _u]
        L.orc_num_targets.restype = _i; L.orc_num_targets.argtypes = [_p]
        L.orc_get_ids.restype = _i; L.orc_get_ids.argtypes = [_p, _p, _i]
        L.orc_get_state.restype = _i; L.orc_get_state.argtypes = [_p, _u, _p, _p, _p, _p, _p]
        for f in ("orc_get_pose_at", "orc_get_twist_at", "orc_get_acc_at"):
            getattr(L, f).restype = _i; getattr(L, f).argtypes = [_p, _u, _d, _p]
        L.orc_get_measured_pose.restype = _i; L.orc_get_measured_pose.argtypes = [_p, _u, _p]
        L.orc_get_pose_internal.restype = _i; L.orc_get_pose_internal.argtypes = [_p, _u, _p]
        L.orc_get_period_estimate.restype = _d; L.orc_get_period_estimate.argtypes = [_p, _u]
        L.orc_step_batch.argtypes = [_p, _i, _p, _d, _p, _p]
        L.orc_step_batch_mt.argtypes = [_p, _i, _i, _p, _d, _p, _p]
        L.orc_step_ticks_mt.argtypes = [_p, _i, _i, _p, _i, _d, _p, _p]
        L.orc_init_batch_mt.argtypes = [_p, _i, _i, _i, _p, _d, _p, _p, _i, _p, _i, _p, _p, _p]
        L.orc_get_states_mt.argtypes = [_p, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p]
        L.orc_run_stream.argtypes = [_p, _u, _i, _d, _p, _p, _i, _p, _p, _p, _p]
        L.orc_reftest_streams.argtypes = [_d, _i, _i, _p, _p]
        L.orc_libstdcxx_normal.argtypes = [_d, _d, _i, _p]
        L.orc_quat_to_rpy.argtypes = [_p, _p]; L.orc_rpy_to_quat.argtypes = [_p, _p]
        L.orc_quat_to_rot.argtypes = [_p, _p]; L.orc_rot_to_quat.argtypes = [_p, _p]; L.orc_rot_to_rpy.argtypes = [_p, _p]
        L.orc_unwrap3.argtypes = [_p, _p, _p]
        L.orc_constrain_angle.restype = _d; L.orc_constrain_angle.argtypes = [_d]
        L.orc_angle_diff.restype = _d; L.orc_angle_diff.argtypes = [_d, _d]
        L.orc_wrap_min_max.restype = _d; L.orc_wrap_min_max.argtypes = [_d, _d, _d]
        L.orc_qtran.argtypes = [_d, _p, _p]
        L.orc_to_sec.restype = _d; L.orc_to_sec.argtypes = [_u, _u]
        L.orc_inverse.argtypes = [_p, _i, _p]
        L.orc_mavg_new.restype = _p; L.orc_mavg_new.argtypes = [_u]
        L.orc_mavg_delete.argtypes = [_p]
        L.orc_mavg_update.restype = _d; L.orc_mavg_update.argtypes = [_p, _d]
        L.orc_mavg_variance.restype = _d; L.orc_mavg_variance.argtypes = [_p]
        L.orc_avg_new.restype = _p; L.orc_avg_new.argtypes = [_u]
        L.orc_avg_delete.argtypes = [_p]
        L.orc_avg_update.restype = _d; L.orc_avg_update.argtypes = [_p, _d]
        L.orc_get_id.restype = _i; L.orc_get_id.argtypes = [C.c_char_p, _p]
        L.orc_poly_roots.restype = _i; L.orc_poly_roots.argtypes = [_p, _i, _p, _p]
        L.orc_lowest_real_root.restype = _d; L.orc_lowest_real_root.argtypes = [_p, _i]
        L.orc_isolver_new.restype = _p; L.orc_isolver_new.argtypes = [_p, _u]
        L.orc_isolver_delete.argtypes = [_p]
        L.orc_isolver_time.restype = _d; L.orc_isolver_time.argtypes = [_p, _u, _d, _p, _d]
        L.orc_isolver_pose.restype = _i; L.orc_isolver_pose.argtypes = [_p, _u, _d, _d, _d, _p, _d, _p]
        L.orc_tick_new.restype = _p; L.orc_tick_new.argtypes = [_i, _p, _i, _p, _i, _p]
        L.orc_tick_set_expiration.argtypes = [_p, _d]
        L.orc_tick_set_token.argtypes = [_p, C.c_char_p]
        L.orc_tick_callback.argtypes = [_p, _i, C.c_char_p, _p, _p]
        L.orc_tick_callback_ids.argtypes = [_p, _i, _p, _p, _p]
        L.orc_tick_update.restype = _i; L.orc_tick_update.argtypes = [_p, _d, _u, _u, _p, _i]
        L.orc_tick_time.restype = _d; L.orc_tick_time.argtypes = [_p]
        L.orc_tick_mailboxes.restype = _i; L.orc_tick_mailboxes.argtypes = [_p]
        L.orc_bench_steps.restype = _d
        L.orc_bench_steps.argtypes = [_i, _p, _i, _p, _i, _p, _i, _i, _i, _d, _p, _d, _p]
        L.orc_hardware_threads.restype = _i
    return _lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def colmajor(M):
    """row-major numpy matrix -> the flat column-major list Mat::MapColMajor expects"""
    return np.ascontiguousarray(np.asarray(M, dtype=np.float64).T).reshape(-1)


class Manager:
    """oracle::TargetManager through the orc_* C surface."""

    def __init__(self, yaml_file=None):
        self.L = lib()
        self.h = self.L.orc_manager_new(yaml_file.encode() if yaml_file else None)
        if not self.h:
            raise RuntimeError("oracle manager: cannot load %r" % yaml_file)

    def close(self):
        if self.h:
            self.L.orc_manager_delete(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def init_default(self, id_, dt0, p0, t0):
        p0 = np.ascontiguousarray(p0, dtype=np.float64)
        return self.L.orc_init_default(self.h, id_, dt0, ptr(p0), t0)

    def init_full(self, type_, id_, dt0, t0, Q, R, P0, p0, v0=None, a0=None):
        Qc, Rc, Pc = colmajor(Q), colmajor(R), colmajor(P0)
        p0 = np.ascontiguousarray(p0, dtype=np.float64)
        v0 = np.ascontiguousarray(v0, dtype=np.float64) if v0 is not None else None
        a0 = np.ascontiguousarray(a0, dtype=np.float64) if a0 is not None else None
        n = int(np.sqrt(Qc.size)); m = int(np.sqrt(Rc.size))
        self.L.orc_init_full(self.h, type_, id_, dt0, t0, ptr(Qc), n, ptr(Rc), m, ptr(Pc), ptr(p0), ptr(v0), ptr(a0))

    def update_meas(self, id_, dt, meas):
        meas = np.ascontiguousarray(meas, dtype=np.float64)
        return bool(self.L.orc_update_meas(self.h, id_, dt, ptr(meas)))

    def update(self, id_, dt):
        return bool(self.L.orc_update(self.h, id_, dt))

    def update_all(self, dt):
        self.L.orc_update_all(self.h, dt)

    def erase(self, id_):
        return bool(self.L.orc_erase(self.h, id_))

    def step_batch(self, ids, dt, meas, action):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        meas = np.ascontiguousarray(meas, dtype=np.float64)
        action = np.ascontiguousarray(action, dtype=np.uint8)
        self.L.orc_step_batch(self.h, ids.size, ptr(ids), dt, ptr(meas), ptr(action))

    def ids(self):
        n = self.L.orc_num_targets(self.h)
        out = np.zeros(max(n, 1), dtype=np.uint32)
        self.L.orc_get_ids(self.h, ptr(out), n)
        return out[:n]

    def state(self, id_, n):
        x = np.zeros(n); P = np.zeros((n, n)); t = C.c_double(); nm = C.c_longlong(); prev = np.zeros(3)
        r = self.L.orc_get_state(self.h, id_, ptr(x), ptr(P), C.byref(t), C.byref(nm), ptr(prev))
        if r == 0:
            return None
        return {"x": x, "P": P, "t": t.value, "n_meas": nm.value, "prev_rpy": prev}

    def states(self, ids, n):
        xs = np.zeros((len(ids), n)); Ps = np.zeros((len(ids), n, n)); ts = np.zeros(len(ids))
        nms = np.zeros(len(ids), dtype=np.int64); prevs = np.zeros((len(ids), 3))
        for k, i in enumerate(ids):
            s = self.state(int(i), n)
            xs[k], Ps[k], ts[k], nms[k], prevs[k] = s["x"], s["P"], s["t"], s["n_meas"], s["prev_rpy"]
        return {"x": xs, "P": Ps, "t": ts, "n_meas": nms, "prev_rpy": prevs}

    def _get(self, fn, id_, k, *a):
        out = np.zeros(k)
        ok = fn(self.h, id_, *a, ptr(out))
        return bool(ok), out

    def pose(self, id_): return self._get(self.L.orc_get_est_pose, id_, 7)
    def twist(self, id_): return self._get(self.L.orc_get_est_twist, id_, 6)
    def acc(self, id_): return self._get(self.L.orc_get_est_acceleration, id_, 6)
    def pose_at(self, id_, t1): return self._get(self.L.orc_get_pose_at, id_, 7, t1)
    def twist_at(self, id_, t1): return self._get(self.L.orc_get_twist_at, id_, 6, t1)
    def acc_at(self, id_, t1): return self._get(self.L.orc_get_acc_at, id_, 6, t1)
    def measured_pose(self, id_): return self._get(self.L.orc_get_measured_pose, id_, 7)
    def pose_internal(self, id_): return self._get(self.L.orc_get_pose_internal, id_, 6)
    def n_measurements(self, id_): return int(self.L.orc_get_n_measurements(self.h, id_))


class ShardedManager:
    """T oracle TargetManagers, id k in manager k % T, stepped by T host threads (orc_*_mt): the same per-target arithmetic as
    Manager, fast enough for parity runs at bench scale.  One (Q, R, P0) for all targets, optional per-target P0 scale."""

    def __init__(self, threads=None):
        self.L = lib()
        self.T = int(threads or min(32, max(1, self.L.orc_hardware_threads())))
        self.hs = (C.c_void_p * self.T)(*[self.L.orc_manager_new(None) for _ in range(self.T)])

    def close(self):
        if getattr(self, "hs", None) is not None:
            for h in self.hs:
                self.L.orc_manager_delete(h)
            self.hs = None

    def __del__(self):
        self.close()

    def init_batch(self, type_, ids, dt0, Q, R, P0, p0, scale=None, t0=None):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        Qc, Rc, Pc = colmajor(Q), colmajor(R), colmajor(P0)
        p0 = np.ascontiguousarray(p0, dtype=np.float64)
        scale = np.ascontiguousarray(scale, dtype=np.float64) if scale is not None else None
        t0 = np.ascontiguousarray(t0, dtype=np.float64) if t0 is not None else None
        n = int(np.sqrt(Qc.size)); m = int(np.sqrt(Rc.size))
        self.L.orc_init_batch_mt(self.hs, self.T, type_, ids.size, ptr(ids), dt0, ptr(t0), ptr(Qc), n, ptr(Rc), m, ptr(Pc), ptr(scale), ptr(p0))

    def step_batch(self, ids, dt, meas, action):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        meas = np.ascontiguousarray(meas, dtype=np.float64)
        action = np.ascontiguousarray(action, dtype=np.uint8)
        self.L.orc_step_batch_mt(self.hs, self.T, ids.size, ptr(ids), dt, ptr(meas), ptr(action))

    def step_ticks(self, ids, dt, meas, action):
        """meas [n_ticks][n][7], action [n_ticks][n]: all ticks in one call (no Python between ticks)"""
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        meas = np.ascontiguousarray(meas, dtype=np.float64)
        action = np.ascontiguousarray(action, dtype=np.uint8)
        self.L.orc_step_ticks_mt(self.hs, self.T, ids.size, ptr(ids), action.shape[0], dt, ptr(meas), ptr(action))

    def states(self, ids, n):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        k = ids.size
        xs = np.zeros((k, n)); Ps = np.zeros((k, n, n)); ts = np.zeros(k); nms = np.zeros(k, dtype=np.int64); prevs = np.zeros((k, 3))
        found = np.zeros(k, dtype=np.uint8)
        self.L.orc_get_states_mt(self.hs, self.T, k, ptr(ids), n, ptr(xs), ptr(Ps), ptr(ts), ptr(nms), ptr(prevs), ptr(found))
        return {"x": xs, "P": Ps, "t": ts, "n_meas": nms, "prev_rpy": prevs, "found": found}


def load_yaml(path):
    L = lib()
    Q = np.zeros(18 * 18); R = np.zeros(36); P = np.zeros(18 * 18)
    n, m, t = C.c_int(), C.c_int(), C.c_int()
    f = C.c_double()
    ok = L.orc_load_yaml(path.encode(), ptr(Q), ptr(R), ptr(P), C.byref(n), C.byref(m), C.byref(t), C.byref(f))
    if not ok:
        return None
    n, m = n.value, m.value
    # flat column-major -> row-major numpy
    return {"type": t.value, "frequency": f.value, "Q": Q[: n * n].reshape(n, n).T.copy(), "R": R[: m * m].reshape(m, m).T.copy(),
            "P": P[: n * n].reshape(n, n).T.copy()}
