#!/usr/bin/env python3
"""Secondary measurements for BASELINE.json configs[2..3] (not the bench.py headline): prints one JSON object.

  C3  1 Mi angular-rates targets with per-tick add/erase churn: every tick 1 % of the live ids fall silent (expire
      8 ticks later by the reference's predicate), as many fresh ids appear; tick = dense step + stamp + expiry
      compaction + append.
  C4  1 Mi targets + one IntersectionSolver query per target per tick (device resident), for uniform-velocity (reference
      semantics: every query returns -1) and uniform-acceleration (the quartic path is exercised).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import target_estimation_b200 as te  # noqa: E402
from tests.synth import rpy_to_quat  # noqa: E402

DT = 1.0 / 250.0


def make_pool(model, n, stream, v0=None):
    mtype, _, Q, R, P0 = te.load_model(model)
    pool = te.TargetPool(mtype, stream=stream.cuda_stream)
    pool.register_class(Q, R, P0)
    pool.reserve(int(n * 1.3))
    rng = np.random.default_rng(1)
    p0 = np.zeros((n, 7)); p0[:, :3] = rng.uniform(-5, 5, (n, 3))
    p0[:, 3:] = rpy_to_quat(rng.uniform(-0.4, 0.4, (n, 3)))
    pool.add(np.arange(n, dtype=np.uint32), p0, p0_scale=rng.uniform(0.5, 2.0, n), v0=v0)
    return pool, p0


def c3_churn(n=1 << 20, ticks=48, fused=False):
    stream = torch.cuda.Stream()
    pool, p0 = make_pool("angular_rates", n, stream)
    rng = np.random.default_rng(2)
    timeout = 8 * DT
    # ---- host model of the churn, computed BEFORE the timed loop: per tick the action mask in slot order, the ids that
    # ---- must expire, and the fresh ids.  The timed loop then contains library calls only.
    ids = pool.ids().astype(np.int64)
    silent_at = np.full(ids.size, 1 << 30, dtype=np.int64)
    last_tick = np.full(ids.size, -1, dtype=np.int64)      # last tick with a measurement (-1: never stamped -> never expires)
    next_id = int(ids.max()) + 1
    sched = []

    def to_sec(tk):   # toSec of the synthetic clock: two roundings, like the device
        ns_ = 1000 * 10 ** 9 + tk * 4000000
        return float(ns_ // 10 ** 9) + 1e-9 * float(ns_ % 10 ** 9)

    for k in range(ticks):
        speaking = np.nonzero(silent_at > k)[0]
        quit_ = rng.choice(speaking, size=max(1, speaking.size // 100), replace=False)
        silent_at[quit_] = k
        speaks = silent_at > k
        act = np.where(speaks, 2, 1).astype(np.uint8)
        last_tick[speaks] = k
        # the reference's predicate in FP64: last > 0 && (now - last) >= timeout
        now_ = to_sec(k)
        cand = np.nonzero(~speaks & (last_tick >= 0))[0]
        expired = np.zeros(ids.size, dtype=bool)
        if cand.size:
            uniq = {int(t_): to_sec(int(t_)) for t_ in np.unique(last_tick[cand])}
            lasts = np.array([uniq[int(t_)] for t_ in last_tick[cand]])
            expired[cand] = (now_ - lasts) >= timeout
        exp_ids = ids[expired]
        ids, silent_at, last_tick = ids[~expired], silent_at[~expired], last_tick[~expired]
        fresh = np.arange(next_id, next_id + quit_.size, dtype=np.int64)
        next_id += quit_.size
        pf = np.zeros((fresh.size, 7)); pf[:, :3] = rng.uniform(-5, 5, (fresh.size, 3)); pf[:, 6] = 1.0
        sched.append((torch.from_numpy(act).cuda(), exp_ids.astype(np.uint32), fresh.astype(np.uint32), pf, act.size))
        ids = np.concatenate([ids, fresh]); silent_at = np.concatenate([silent_at, np.full(fresh.size, 1 << 30, dtype=np.int64)])
        last_tick = np.concatenate([last_tick, np.full(fresh.size, -1, dtype=np.int64)])
    n_max = max(s_[4] for s_ in sched) + 1
    meas = torch.zeros((n_max, 7), dtype=torch.float64, device="cuda"); meas[:, 6] = 1.0
    meas[:, :3] = torch.rand((n_max, 3), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    n_erased = n_added = steps = 0
    parts = {"step+stamp+expire (fused)" if fused else "step+stamp": 0.0, "expire": 0.0, "add": 0.0}
    t0 = time.perf_counter()
    for k, (d_act, exp_ids, fresh, pf, live) in enumerate(sched):
        ns = 1000 * 10 ** 9 + k * 4000000
        sec, nsec = ns // 10 ** 9, ns % 10 ** 9
        ta = time.perf_counter()
        if fused:   # one call: stamps + expiry flags + scan, then the step writes the survivors to their compacted slots
            erased = pool.step_dense_expire(DT, meas, 7, d_act, te.ACT_UPDATE, (sec, nsec), (sec, nsec), timeout)
            tb = tc = time.perf_counter()
        else:
            pool.step_dense(DT, meas, 7, d_act)
            pool.stamp_dense(sec, nsec, d_act)
            pool.sync()
            tb = time.perf_counter()
            erased = pool.expire(sec, nsec, timeout)
            tc = time.perf_counter()
        pool.add(fresh, pf, t0=np.full(fresh.size, k * DT))
        td = time.perf_counter()
        parts["step+stamp+expire (fused)" if fused else "step+stamp"] += tb - ta; parts["expire"] += tc - tb; parts["add"] += td - tc
        assert np.array_equal(erased, exp_ids), k          # erase decisions bit-exact against the host model
        steps += live; n_erased += erased.size; n_added += fresh.size
    pool.sync()
    dt_wall = time.perf_counter() - t0
    assert np.array_equal(pool.ids().astype(np.int64), ids)
    out = {"targets": n, "ticks": ticks, "erased": int(n_erased), "added": int(n_added), "ms_per_tick": 1e3 * dt_wall / ticks,
           "target_steps_per_s": steps / dt_wall, "ms_per_tick_parts": {k_: 1e3 * v / ticks for k_, v in parts.items()},
           "note": ("library calls only in the timed loop (te_pool_step_dense_expire: stamps, expiry flags, scan, then ONE pass in which "
                    "the step kernel writes the survivors to their compacted slots; append of the fresh ids from host arrays)"
                    if fused else
                    "library calls only in the timed loop (dense step + stamp, expiry = flags + scan + stable gather of the whole pool "
                    "into the second buffer, append of the fresh ids from host arrays)") +
                   "; every tick's erase list and the final id set are checked against a host model of the churn"}
    pool.close()
    return out


def c3_mailbox(n=1 << 20, ticks=48, model="angular_rates", device_records=False, prefetch=False):
    """C3 through the node loop itself: per tick ONE /tf message from pinned host memory (a record = id, stamp, pose7 for every
    speaking id, 68 B each) -> te_pool_mailbox_ingest, then te_pool_mailbox_tick (first-sight init of the fresh ids, sticky
    update / predict, expiry, one stable rebuild + one step launch).  Same churn as c3_churn: 1 % of the speaking ids fall
    silent per tick and expire 8 ticks later, as many fresh ids appear.  Erase lists and the final id set are checked against a
    host model of the reference's predicate."""
    stream = torch.cuda.Stream()
    mtype, _, Q, R, P0 = te.load_model(model)
    pool = te.TargetPool(mtype, stream=stream.cuda_stream)
    pool.register_class(Q, R, P0)
    pool.reserve(int(n * 1.3))
    rng = np.random.default_rng(2)
    timeout = 8 * DT

    def clock(tk):
        ns_ = 1000 * 10 ** 9 + tk * 4000000
        return ns_ // 10 ** 9, ns_ % 10 ** 9

    def to_sec(tk):
        s_, n_ = clock(tk)
        return float(s_) + 1e-9 * float(n_)

    ids = np.arange(n, dtype=np.int64)
    silent_at = np.full(n, 1 << 30, dtype=np.int64)
    last_tick = np.full(n, -1, dtype=np.int64)
    next_id = n
    sched = []
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    n_max = int(n * 1.05)
    poses = np.zeros((n_max, 7)); poses[:, :3] = rng.uniform(-5, 5, (n_max, 3)); poses[:, 3:] = rpy_to_quat(rng.uniform(-0.4, 0.4, (n_max, 3)))
    poses = pin(poses)
    for k in range(ticks + 1):   # tick 0 populates the pool (untimed)
        if k > 0:
            old = np.nonzero(silent_at > k)[0]
            quit_ = rng.choice(old, size=max(1, old.size // 100), replace=False)
            silent_at[quit_] = k
            fresh = np.arange(next_id, next_id + quit_.size, dtype=np.int64); next_id += quit_.size
            ids = np.concatenate([ids, fresh]); silent_at = np.concatenate([silent_at, np.full(fresh.size, 1 << 30, dtype=np.int64)])
            last_tick = np.concatenate([last_tick, np.full(fresh.size, -1, dtype=np.int64)])
        speaks = silent_at > k
        last_tick[speaks] = k
        now_ = to_sec(k)
        cand = np.nonzero(~speaks & (last_tick >= 0))[0]
        expired = np.zeros(ids.size, dtype=bool)
        if cand.size:
            uniq = {int(t_): to_sec(int(t_)) for t_ in np.unique(last_tick[cand])}
            lasts = np.array([uniq[int(t_)] for t_ in last_tick[cand]])
            expired[cand] = (now_ - lasts) >= timeout
        rec_ids = ids[speaks].astype(np.uint32)
        sec, nsec = clock(k)
        sched.append((pin(rec_ids), pin(np.full(rec_ids.size, sec, dtype=np.uint32)), pin(np.full(rec_ids.size, nsec, dtype=np.uint32)),
                      ids[expired].astype(np.uint32), int(ids.size)))
        ids, silent_at, last_tick = ids[~expired], silent_at[~expired], last_tick[~expired]
    d_sched, d_poses = None, None
    if device_records:   # the messages already sit in device memory (another CUDA stage / a receive buffer): te_pool_mailbox_ingest_dev
        d_poses = torch.from_numpy(poses).cuda()
        d_sched = [tuple(torch.from_numpy(a.view(np.int32)).cuda() for a in s_[:3]) for s_ in sched]
        torch.cuda.synchronize()
    r_ids, r_sec, r_nsec, exp_ids, live = sched[0]
    pool.mailbox_ingest(r_ids, r_sec, r_nsec, poses[:r_ids.size])
    erased, added = pool.mailbox_tick(DT, 0.0, clock(0), timeout)
    assert added == n and erased.size == 0
    pool.sync()
    n_erased = n_added = steps = records = 0
    parts = {"ingest": 0.0, "tick": 0.0}
    per_tick = []
    if prefetch:   # message 1 is on its way before the clock starts, as message k + 1 is while tick k runs
        pool.mailbox_prefetch(sched[1][0], sched[1][1], sched[1][2], poses[:sched[1][0].size])
        pool.sync()
    t0 = time.perf_counter()
    for k in range(1, ticks + 1):
        r_ids, r_sec, r_nsec, exp_ids, live = sched[k]
        ta = time.perf_counter()
        if device_records:
            pool.mailbox_ingest_dev(r_ids.size, d_sched[k][0], d_sched[k][1], d_sched[k][2], d_poses)
        elif prefetch:
            pool.mailbox_ingest_prefetched()           # message k takes effect (its copy ran under tick k - 1)
            if k < ticks:                              # message k + 1 starts its way to the device: runs under tick k
                nx = sched[k + 1]
                pool.mailbox_prefetch(nx[0], nx[1], nx[2], poses[:nx[0].size])
        else:
            pool.mailbox_ingest(r_ids, r_sec, r_nsec, poses[:r_ids.size])
        tb = time.perf_counter()
        erased, added = pool.mailbox_tick(DT, k * DT, clock(k), timeout)
        tc = time.perf_counter()
        parts["ingest"] += tb - ta; parts["tick"] += tc - tb
        per_tick.append(tc - ta)
        assert np.array_equal(erased, exp_ids), k
        steps += live - erased.size; n_erased += erased.size; n_added += added; records += r_ids.size
    dt_wall = time.perf_counter() - t0
    assert np.array_equal(pool.ids().astype(np.int64), ids)
    out = {"model": model, "targets": n, "ticks": ticks, "erased": int(n_erased), "added": int(n_added), "records_per_tick": records / ticks,
           "h2d_bytes_per_tick": 0 if device_records else 68 * records / ticks, "records_in": "device memory" if device_records else ("pinned host memory, next message prefetched under the tick" if prefetch else "pinned host memory"),
           "ms_per_tick": 1e3 * dt_wall / ticks, "target_steps_per_s": steps / dt_wall,
           "ms_per_tick_median": 1e3 * float(np.median(per_tick)), "ms_per_tick_worst": 1e3 * float(np.max(per_tick)),   # (library calls only)
           "ms_per_tick_parts": {k_: 1e3 * v / ticks for k_, v in parts.items()},
           "note": "the node loop through te_pool_mailbox_ingest + te_pool_mailbox_tick: records come from pinned HOST memory every tick (ids, "
                   "stamps, poses), mailboxes, first-sight init, sticky update / predict and expiry run on the device; every tick's erase "
                   "list and the final id set are checked against a host model of the churn"}
    pool.close()
    return out


def c3_tick_manager(n=1 << 20, ticks=40, model="angular_rates", publish=False):
    """the same churn through the REFERENCE-FACING library (libtarget_c.so): target_tick_manager_callback_ids + target_tick_manager_update
    (TickTargetManager: host registry of ids + the device mailboxes), optionally with the per-tick gather of every filtered pose
    (what the node broadcasts).  Messages come from ordinary host arrays."""
    from target_estimation_b200.manager import TickManagerC
    mgr = TickManagerC(os.path.join(ROOT, "models", "model_%s_params.yaml" % model), device=torch.cuda.current_device())
    timeout = 8 * DT
    mgr.set_expiration(timeout); mgr.set_publish(publish)
    rng = np.random.default_rng(2)

    def clock(tk):
        ns_ = 1000 * 10 ** 9 + tk * 4000000
        return ns_ // 10 ** 9, ns_ % 10 ** 9
    ids = np.arange(n, dtype=np.int64); silent_at = np.full(n, 1 << 30, dtype=np.int64); next_id = n
    poses = np.zeros((int(n * 1.1), 7)); poses[:, :3] = rng.uniform(-5, 5, (poses.shape[0], 3)); poses[:, 6] = 1.0
    msgs = []
    for k in range(ticks + 1):
        if k > 0:
            old = np.nonzero(silent_at > k)[0]
            quit_ = rng.choice(old, size=max(1, old.size // 100), replace=False)
            silent_at[quit_] = k
            fresh = np.arange(next_id, next_id + quit_.size, dtype=np.int64); next_id += quit_.size
            ids = np.concatenate([ids, fresh]); silent_at = np.concatenate([silent_at, np.full(fresh.size, 1 << 30, dtype=np.int64)])
        sec, nsec = clock(k)
        r = ids[silent_at > k].astype(np.uint32)
        msgs.append((r, np.full(r.size, sec, dtype=np.uint32), np.full(r.size, nsec, dtype=np.uint32)))
    r, s_, ns_ = msgs[0]
    mgr.callback_ids(r, s_, ns_, poses[:r.size]); mgr.tick(DT, *clock(0))
    parts = {"callback": [], "tick": []}
    n_erased = 0
    t0 = time.perf_counter()
    for k in range(1, ticks + 1):
        r, s_, ns_ = msgs[k]
        ta = time.perf_counter()
        mgr.callback_ids(r, s_, ns_, poses[:r.size])
        tb = time.perf_counter()
        n_erased += mgr.tick(DT, *clock(k), cap=1 << 16).size
        tc = time.perf_counter()
        parts["callback"].append(tb - ta); parts["tick"].append(tc - tb)
    dt_wall = time.perf_counter() - t0
    n_live = mgr.ids().size
    out = {"model": model, "targets": n, "ticks": ticks, "erased": int(n_erased), "live_at_end": int(n_live), "publish": bool(publish),
           "ms_per_tick": 1e3 * dt_wall / ticks, "ms_per_tick_parts": {k_: 1e3 * float(np.mean(v)) for k_, v in parts.items()},
           "ms_per_tick_parts_median": {k_: 1e3 * float(np.median(v)) for k_, v in parts.items()},
           "ms_per_tick_parts_max": {k_: 1e3 * float(np.max(v)) for k_, v in parts.items()},
           "note": "libtarget_c.so: target_tick_manager_callback_ids + target_tick_manager_update, pageable host arrays; mean / median / "
                   "worst tick (the host side of a freshly booted box is noisy: page faults of first-touched memory)"}
    mgr.close()
    return out


def c4_intersect(model, n=1 << 20, ticks=20):
    stream = torch.cuda.Stream()
    v0 = np.zeros((n, 6)); v0[:, 0] = 1.0                       # every target flies along +x at 1 m/s ...
    pool, p0 = make_pool(model, n, stream, v0=v0)
    solver = te.IntersectionSolver(pool, n_streams=n, filters_length=250)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    base = torch.from_numpy(p0).cuda()
    origin = base[:, :3].clone()
    origin[:, 0] += 2.0                                         # ... towards a sphere 2 m ahead (the target starts outside it)
    origin += 0.1 * (torch.rand((n, 3), dtype=torch.float64, device="cuda", generator=g) - 0.5)
    origin = origin.contiguous()
    radius = (0.3 + 0.4 * torch.rand((n,), dtype=torch.float64, device="cuda", generator=g)).contiguous()
    delta = torch.empty(n, dtype=torch.float64, device="cuda")
    pose = torch.empty((n, 7), dtype=torch.float64, device="cuda")
    conv = torch.empty(n, dtype=torch.uint8, device="cuda")
    meas = base.clone(); meas[:, 0] += DT; meas[:, 2] -= 0.0001   # consistent with the motion, slight downward pull -> non-zero acceleration estimate
    for _ in range(3):
        pool.step_dense(DT, meas, 7, None, te.ACT_UPDATE)
        solver.query_dense(origin, radius, 0.05, 0.1, None, delta, pose, conv)
    pool.sync()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    tq = 0.0
    e0.record(stream)
    for _ in range(ticks):
        pool.step_dense(DT, meas, 7, None, te.ACT_UPDATE)
    e1.record(stream)
    for _ in range(ticks):
        solver.query_dense(origin, radius, 0.05, 0.1, None, delta, pose, conv)
    e2.record(stream)
    pool.sync(); torch.cuda.synchronize()
    ms_step, ms_q = e0.elapsed_time(e1) / ticks, e1.elapsed_time(e2) / ticks
    found = int((delta >= 0).sum().item())
    out = {"model": model, "targets": n, "queries_per_tick": n, "ms_per_tick_step": ms_step, "ms_per_tick_queries": ms_q,
           "queries_per_s": n / (ms_q * 1e-3), "interceptions_found": found, "converged": int(conv.sum().item())}
    solver.close(); pool.close()
    return out


def replay(model="uniform_acceleration", n=4 << 20, T=16, launches=10):
    """temporal blocking: T buffered ticks per launch, state crosses HBM once per launch"""
    stream = torch.cuda.Stream()
    pool, p0 = make_pool(model, n, stream)
    base = torch.from_numpy(p0).cuda()
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    meas = base.unsqueeze(0).repeat(T, 1, 1).contiguous()
    meas[:, :, :3] += 0.01 * torch.randn((T, n, 3), dtype=torch.float64, device="cuda", generator=g)
    act = torch.where(torch.rand((T, n), device="cuda", generator=g) < 0.05, 1, 2).to(torch.uint8).contiguous()
    for _ in range(2):
        pool.step_dense_ticks(T, DT, meas, 7, act)
    pool.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(launches):
        pool.step_dense_ticks(T, DT, meas, 7, act)
    e1.record(stream)
    pool.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out = {"model": model, "targets": n, "ticks_per_launch": T, "launches": launches, "ms_per_launch": ms / launches,
           "target_steps_per_s": n * T * launches / (ms * 1e-3),
           "note": "te_pool_step_dense_ticks: same arithmetic and results as T single ticks; HBM traffic per target-step = "
                   "(state in + out) / T + measurement"}
    pool.close()
    return out


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "benchline":
        # bench.py's secondary figures, one process for all of them: BASELINE configs[2] (1 Mi ANGULAR-RATES targets through the node
        # loop with churn), the same loop for the benchmarked motion model, configs[3] (1 Mi UV + 1 Mi UA interception queries).
        # Each leg reports its own failure instead of taking the others with it.
        res = {}

        def leg(key, fn):
            try:
                res[key] = fn()
            except Exception as e:   # noqa: BLE001
                res[key] = {"error": ("%s: %s" % (type(e).__name__, e))[:300]}
            print(json.dumps({key: res[key]}), file=sys.stderr, flush=True)
        # (te_pool_mailbox_prefetch / te_pool_mailbox_ingest_prefetched: the copy of message k + 1 runs under tick k; the figure of
        #  the synchronous te_pool_mailbox_ingest rides along as *_sync)
        leg("c3", lambda: c3_mailbox(ticks=24, model="angular_rates", prefetch=True))
        leg("c3_sync", lambda: c3_mailbox(ticks=24, model="angular_rates"))
        if sys.argv[2] != "angular_rates":
            leg("node_loop", lambda: c3_mailbox(ticks=24, model=sys.argv[2], prefetch=True))
            leg("node_loop_sync", lambda: c3_mailbox(ticks=24, model=sys.argv[2]))
        leg("c4_uniform_velocity", lambda: c4_intersect("uniform_velocity"))
        leg("c4_uniform_acceleration", lambda: c4_intersect("uniform_acceleration"))
        print(json.dumps(res))
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[1] == "mailbox1":   # one model, short: bench.py's secondary node-loop figure
        print(json.dumps(c3_mailbox(ticks=24, model=sys.argv[2])))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "tickmgr":
        print(json.dumps({"c3_tick_manager_angular_rates": c3_tick_manager(), "c3_tick_manager_angular_rates_publish": c3_tick_manager(publish=True)}))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "mailbox":
        res = {}
        for m_ in ("angular_rates", "uniform_acceleration", "angular_velocities", "uniform_velocity"):
            res["c3_mailbox_" + m_] = c3_mailbox(model=m_)
            res["c3_mailbox_" + m_ + "_device_records"] = c3_mailbox(model=m_, device_records=True)
        print(json.dumps(res))
        sys.exit(0)
    res = {"c3_churn_angular_rates": c3_churn(), "c3_churn_angular_rates_fused": c3_churn(fused=True), "c4_intersect_uniform_velocity": c4_intersect("uniform_velocity"),
           "c4_intersect_uniform_acceleration": c4_intersect("uniform_acceleration"),
           "replay16_uniform_acceleration_4Mi": replay()}
    print(json.dumps(res))
