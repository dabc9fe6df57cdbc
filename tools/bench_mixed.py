#!/usr/bin/env python3
"""BASELINE.json configs[4]: mixed models sharded by id across the GPUs of one box + all-gather of estimates.

    python tools/bench_mixed.py [--targets-per-gpu 8388608] [--ticks 32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/bench_mixed.py ...

owner(id) = id mod G; model(id) by (id div 16) mod 10: 0-3 uniform velocity, 4-7 uniform acceleration, 8 angular
velocities, 9 angular rates (SURVEY.md 8(d) C5 says "id mod 10", which is not independent of id mod G for even G -- rank 0 of 2
would own every AV and no AR target; taking the residue of id div 16 keeps the 40/40/10/10 mix on every rank for G | 16).  One pool per model per rank, four launches per tick on one stream, no collective on
the hot path; afterwards every rank contributes its [pose7 | twist6] records to an NCCL all-gather.  Prints one JSON line.
"""
import argparse
import json
import os
import sys


import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import target_estimation_b200 as te  # noqa: E402
from tests.synth import rpy_to_quat  # noqa: E402

DT = 1.0 / 250.0
MIX = [("uniform_velocity", (0, 1, 2, 3)), ("uniform_acceleration", (4, 5, 6, 7)), ("angular_velocities", (8,)), ("angular_rates", (9,))]


def run_c5(rank, world, local, stream, n, ticks, warmup, barrier=None, seed=100):
    """one rank's share of C5: four pools (40/40/10/10 by (id div 16) mod 10) of the n ids this rank owns, `ticks` timed ticks
    (max over ranks), then -- world > 1 -- the NCCL all-gather of every rank's [pose7 | twist6] records.  Returns the result
    dict on rank 0, None elsewhere.  Every rank must call it (collectives)."""
    if barrier is None:
        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
    ids_all = (np.arange(n, dtype=np.int64) * world + rank)           # the ids this rank owns
    rng = np.random.default_rng(seed + rank)
    pools, inputs, alg_bytes = [], [], 0.0
    for name, residues in MIX:
        ids = ids_all[np.isin((ids_all // 16) % 10, residues)].astype(np.uint32)
        mtype, _, Q, R, P0 = te.load_model(name)
        pool = te.TargetPool(mtype, device=local, stream=stream.cuda_stream)
        pool.register_class(Q, R, P0)
        pool.reserve(ids.size)
        k = ids.size
        p0 = np.zeros((k, 7)); p0[:, :3] = rng.uniform(-5, 5, (k, 3)); p0[:, 3:] = rpy_to_quat(rng.uniform(-0.4, 0.4, (k, 3)))
        pool.add(ids, p0, p0_scale=rng.uniform(0.5, 2.0, k))
        base = torch.from_numpy(p0).cuda()
        g = torch.Generator(device="cuda"); g.manual_seed(7 + rank)
        sets = []
        for j in range(2):
            m = base.clone(); m[:, :3] += 0.01 * torch.randn((k, 3), dtype=torch.float64, device="cuda", generator=g)
            a = torch.where(torch.rand((k,), device="cuda", generator=g) < 0.05, 1, 2).to(torch.uint8)
            sets.append((m.contiguous(), a.contiguous()))
        pools.append(pool); inputs.append(sets)
        alg_bytes += k * te.bytes_per_step(mtype)                      # (update-step figure; 5 % predict-only steps are a few bytes less)

    def tick(t):
        for pool, sets in zip(pools, inputs):
            m, a = sets[t % 2]
            pool.step_dense(DT, m, 7, a)

    for t in range(warmup):
        tick(t)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(ticks):
        tick(t)
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    # all-gather of the estimate records (off the hot path)
    counts = [len(p) for p in pools]
    rec = torch.empty((sum(counts), 13), dtype=torch.float64, device="cuda")
    ag_ms = rec_ms = None
    with torch.cuda.stream(stream):
        def gather_records():
            off = 0
            for p, c in zip(pools, counts):
                p.estimates_dev(rec[off:off + c])
                off += c
        gather_records()
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record(stream)
        for _ in range(3):
            gather_records()
        r1.record(stream)
        barrier()
        rec_ms = r0.elapsed_time(r1) / 3
        if world > 1:
            out = torch.empty((world * rec.shape[0], 13), dtype=torch.float64, device="cuda")
            dist.all_gather_into_tensor(out, rec)
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(3):
                dist.all_gather_into_tensor(out, rec)
            a1.record(stream)
            barrier()
            t_ag = torch.tensor([a0.elapsed_time(a1) / 3], dtype=torch.float64, device="cuda")
            dist.all_reduce(t_ag, op=dist.ReduceOp.MAX)
            ag_ms = float(t_ag.item())
            del out
    torch.cuda.synchronize()
    res = None
    if rank == 0:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        nbytes = int(rec.numel() * 8)
        res = {"config": "C5 mixed models (40% UV / 40% UA / 10% AV / 10% AR by (id div 16) mod 10), owner = id mod G", "n_gpus": world,
               "targets_per_gpu": n, "targets_total": n * world, "ticks": ticks, "ms_per_tick": ms / ticks,
               "target_steps_per_s": n * world * ticks / (ms * 1e-3), "per_gpu_counts": dict(zip([m for m, _ in MIX], counts)),
               "alg_gbs_per_gpu": alg_bytes / (ms / ticks * 1e-3) / 1e9, "frac_of_hbm_peak": alg_bytes / (ms / ticks * 1e-3) / 1e9 / peak,
               "device_bytes_per_gpu": int(sum(p.device_bytes() for p in pools)),
               "estimate_records_ms": rec_ms,
               "allgather": None if ag_ms is None else {"ms": ag_ms, "bytes_per_rank": nbytes, "records": "pose7|twist6 per target, every rank receives all",
                                                        "bus_gbs": (world - 1) * nbytes / (ag_ms * 1e-3) / 1e9, "backend": "NCCL all_gather over NVLink"}}
    for p in pools:
        p.close()
    del rec, inputs
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--targets-per-gpu", type=int, default=8 << 20)
    ap.add_argument("--ticks", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)      # NCCL's version banner goes to fd 1: keep stdout for the JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier(); torch.cuda.synchronize()
        sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    stream = torch.cuda.Stream()
    res = run_c5(rank, world, local, stream, args.targets_per_gpu, args.ticks, args.warmup)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
