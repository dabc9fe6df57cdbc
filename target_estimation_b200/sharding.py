"""Host-side multi-GPU plumbing: targets shard by id, `owner(id) = id mod G`, one process per GPU,
no collective on the hot path; the only exchange is the optional gather of per-target estimate records
([id | pose7 | twist6]) to the publishing rank(s).  Works with any torch.distributed backend (NCCL over NVLink on
the GPU box, gloo in the CPU tests)."""
import numpy as np


def owner(ids, world):
    """rank that owns each target id"""
    return np.asarray(ids, dtype=np.uint32) % np.uint32(world)


def route(ids, world, *payloads):
    """Split one tick's batch by owner.  Returns a list (per rank) of (ids_r, payload_r...) with the original
    relative order kept inside each shard (ascending ids stay ascending)."""
    ids = np.asarray(ids, dtype=np.uint32)
    own = owner(ids, world)
    out = []
    for r in range(world):
        sel = np.nonzero(own == r)[0]
        out.append((ids[sel],) + tuple(None if p is None else np.asarray(p)[sel] for p in payloads))
    return out


def merge_sorted(shards):
    """G ascending id arrays (+ optional row payloads) -> one globally ascending array: the order of the
    reference's std::map iteration (target_manager.hpp:36)."""
    ids = np.concatenate([s[0] for s in shards]) if shards else np.zeros(0, dtype=np.uint32)
    order = np.argsort(ids, kind="stable")
    rest = []
    for k in range(1, len(shards[0]) if shards else 0):
        rest.append(np.concatenate([s[k] for s in shards])[order])
    return (ids[order],) + tuple(rest)


def all_gather_records(local_ids, local_records, dist, device=None):
    """All-gather of variable-length per-target records.  local_ids: torch int64 [n_r]; local_records: torch
    float64 [n_r, W] (W = 13 for pose7|twist6).  Counts are exchanged first (one all_gather of one int), shards are
    padded to the maximum count for a single all_gather_into_tensor-style exchange, and the padding is dropped on
    arrival.  Returns (ids [N], records [N, W]) in globally ascending id order on every rank."""
    import torch
    world = dist.get_world_size()
    n_local = torch.tensor([local_ids.shape[0]], dtype=torch.int64, device=local_ids.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    n_max = max(counts) if counts else 0
    W = local_records.shape[1]
    pad_ids = torch.zeros(n_max, dtype=torch.int64, device=local_ids.device)
    pad_rec = torch.zeros((n_max, W), dtype=local_records.dtype, device=local_records.device)
    pad_ids[: local_ids.shape[0]] = local_ids
    pad_rec[: local_ids.shape[0]] = local_records
    g_ids = [torch.zeros_like(pad_ids) for _ in range(world)]
    g_rec = [torch.zeros_like(pad_rec) for _ in range(world)]
    dist.all_gather(g_ids, pad_ids)
    dist.all_gather(g_rec, pad_rec)
    ids = torch.cat([g_ids[r][: counts[r]] for r in range(world)])
    rec = torch.cat([g_rec[r][: counts[r]] for r in range(world)])
    order = torch.argsort(ids, stable=True)
    return ids[order], rec[order]
