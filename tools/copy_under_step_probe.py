#!/usr/bin/env python3
"""Probe: how long does a 71 MB pinned host-to-device copy take while a step kernel saturates HBM?
(the node loop's prefetched /tf message travels under the tick: tools/bench_configs.py c3_mailbox)
  python tools/copy_under_step_probe.py [model]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import target_estimation_b200 as te

model = sys.argv[1] if len(sys.argv) > 1 else "angular_rates"
n = 1 << 20
mtype, _, Q, R, P0 = te.load_model(model)
s_step = torch.cuda.Stream()
s_copy = torch.cuda.Stream()
pool = te.TargetPool(mtype, stream=s_step.cuda_stream)
pool.register_class(Q, R, P0)
rng = np.random.default_rng(0)
p0 = np.zeros((n, 7)); p0[:, :3] = rng.uniform(-5, 5, (n, 3)); p0[:, 6] = 1
pool.add(np.arange(n, dtype=np.uint32), p0)
meas = torch.from_numpy(p0).cuda()
host = torch.empty(71 << 20, dtype=torch.uint8).pin_memory()
dev = torch.empty(71 << 20, dtype=torch.uint8, device="cuda")
res = {"model": model, "copy_mb": 71}


def run(name, with_step, with_copy, reps=10, steps_per_rep=3):
    tc, ts = [], []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if with_step:
            with torch.cuda.stream(s_step):
                k0.record(s_step)
            for _ in range(steps_per_rep):
                pool.step_dense(0.004, meas, 7, None, te.ACT_UPDATE)
            k1.record(s_step)
        if with_copy:
            with torch.cuda.stream(s_copy):
                e0.record(s_copy)
                dev.copy_(host, non_blocking=True)
                e1.record(s_copy)
        torch.cuda.synchronize()
        if with_copy:
            tc.append(e0.elapsed_time(e1))
        if with_step:
            ts.append(k0.elapsed_time(k1) / steps_per_rep)
    if tc:
        res[name + "_copy_ms"] = round(float(np.median(tc)), 3)
    if ts:
        res[name + "_step_ms"] = round(float(np.median(ts)), 3)


run("alone", False, True)
run("step_alone", True, False)
run("together", True, True)
print(json.dumps(res))
