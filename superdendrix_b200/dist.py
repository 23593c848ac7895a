"""Multi-GPU sharding of the significance path: one process per GPU (torchrun),
``torch.distributed`` for the plumbing.

The path has no exchange step inside any kernel: nulls, subset
ranges and (profile, module) jobs are independent units, so ranks take disjoint
ranges and only reduce

  * exceedance counts            all_reduce(SUM) of an int64 vector (one per profile),
  * scan results                 all_gather of each rank's top-n records + a local merge,
                                 all_reduce(SUM) of the number of subsets scored,
  * per-job statistics           all_gather of the per-rank slices.

Random streams are keyed on GLOBAL null / permutation indices, so the reduced
result is identical for 1, 2, 4 or 8 ranks (the reference's own sharding by
``-pre`` has the same property of disjoint index ranges, src/generate_null_matrices.py:29,82).
"""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------------------
# pure partitioning / merging helpers (no torch)
# ---------------------------------------------------------------------------
def shard_range(n: int, rank: int, world: int):
    """Contiguous balanced split of range(n): sizes differ by at most one."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def scan_feature_ranges(n_features: int, k: int, world: int):
    """Split the scan by smallest feature index g so every rank owns about the same number of subsets.
    g owns 1 + (F-1-g) + C(F-1-g, 2) subsets (sizes 1..3, truncated at k)."""
    F = int(n_features)
    g = np.arange(F, dtype=np.float64)
    r = F - 1 - g
    work = np.ones(F)
    if k >= 2:
        work += r
    if k >= 3:
        work += r * (r - 1) / 2
    cum = np.concatenate([[0.0], np.cumsum(work)])
    cuts = [int(np.searchsorted(cum, cum[-1] * i / world, side="left")) for i in range(world + 1)]
    cuts[0], cuts[-1] = 0, F
    cuts = np.maximum.accumulate(np.clip(cuts, 0, F))
    return [(int(cuts[i]), int(cuts[i + 1])) for i in range(world)]


def merge_hits(parts, topn: int):
    """Merge per-rank scan results (dicts with W, features, coverage) into the global top-n.
    Order: W descending, then enumeration order (size-major, lexicographic) — the order of a single-GPU scan."""
    W = np.concatenate([p["W"] for p in parts])
    F = np.concatenate([p["features"] for p in parts])
    C = np.concatenate([p["coverage"] for p in parts])
    valid = F[:, 0] >= 0
    W, F, C = W[valid], F[valid], C[valid]
    size = (F >= 0).sum(axis=1)
    key = np.where(F >= 0, F, 0)
    order = np.lexsort((key[:, 2], key[:, 1], key[:, 0], size, -W))
    order = order[:topn]
    out_W = np.full(topn, -np.inf)
    out_F = np.full((topn, 3), -1, dtype=np.int32)
    out_C = np.zeros(topn, dtype=np.int32)
    out_W[:len(order)], out_F[:len(order)], out_C[:len(order)] = W[order], F[order], C[order]
    return {"W": out_W, "features": out_F, "coverage": out_C}


# ---------------------------------------------------------------------------
# collectives (torch.distributed; NCCL on GPUs, gloo in the CPU tests)
# ---------------------------------------------------------------------------
def _dist():
    import torch
    import torch.distributed as dist
    return torch, dist


def _world(group=None):
    _, dist = _dist()
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def _device(group=None):
    torch, dist = _dist()
    if dist.is_initialized() and dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def allreduce_sum_int64(values, group=None) -> np.ndarray:
    torch, dist = _dist()
    v = np.ascontiguousarray(values, dtype=np.int64)
    if _world(group)[1] == 1:
        return v.copy()
    t = torch.from_numpy(v.copy()).to(_device(group))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def allgather_array(a, group=None):
    """All ranks contribute an array of identical shape and dtype (float64 / int64 / int32)."""
    torch, dist = _dist()
    a = np.ascontiguousarray(a)
    world = _world(group)[1]
    if world == 1:
        return [a.copy()]
    t = torch.from_numpy(a.copy()).to(_device(group))
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return [o.cpu().numpy() for o in outs]


# ---------------------------------------------------------------------------
# sharded operations.  `engine` is a superdendrix_b200.engine.Engine (or any object with the same methods).
# ---------------------------------------------------------------------------
def shard_chains(n, chain_len, rank, world, pair_group=1):
    """Null range of a rank as whole chains (so no rank has to replay the prefix of a chain another rank owns) — whole
    blocks of pair_group chains when the chains of a block share their schedule of row pairs (one warp runs a block)."""
    unit = int(chain_len) * max(int(pair_group), 1)
    n_units = -(-int(n) // unit)
    u_lo, u_hi = shard_range(n_units, rank, world)
    return min(u_lo * unit, n), min(u_hi * unit, n)


def build_pool(engine, seed, n_trades=-1, pool_size=None, group=None):
    """The burn-in pool built once across the ranks: rank r burns its slice of the pool states, the slices are exchanged
    with ONE all-gather over NVLink directly on the pool's device memory (include/sdx.h: sdx_pool_begin / build /
    commit), every rank commits the same pool.  Falls back to a local build for a single rank, an engine without the
    part-wise entry points (the oracle stand-in of the CPU tests) or a pool that does not split evenly."""
    rank, world = _world(group)
    if world == 1 or not hasattr(engine, "pool_begin") or pool_size is None or pool_size % world:
        return False
    torch, dist = _dist()
    if dist.get_backend(group) != "nccl" or engine.pool_valid(seed, n_trades):        # the same on every rank: same sequence of calls
        return False
    ptr, state_bytes = engine.pool_begin(seed, n_trades)
    per = pool_size // world
    engine.pool_build(rank * per, per)

    class _Pool:                      # the library's device buffer as a torch tensor (no copy)
        __cuda_array_interface__ = {"shape": (pool_size * state_bytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    whole = torch.as_tensor(_Pool(), device=_device(group))
    mine = whole[rank * per * state_bytes:(rank + 1) * per * state_bytes]
    dist.all_gather_into_tensor(whole, mine, group=group)        # in place: every rank's slice lands at its offset
    torch.cuda.current_stream().synchronize()
    engine.pool_commit()
    return True


def null_pvalue(engine, seed, n_nulls, modules, null0=0, n_trades=-1, stat=0, z_obs=None, chain_len=None, burn=None,
                pool_size=None, pair_group=None, group=None, shared_pool=True):
    """Permutation p-values over n_nulls nulls sharded by null index (the reference shards one `-p` over processes
    by `-pre`, src/generate_null_matrices.py:29,82).  Every null is chain_len ordinary rounds away from a burnt-in pool
    state (Engine.set_burn_in), so it is a draw from the stationary distribution — what the reference gets from its
    single long chain — and there are enough independent units to fill every GPU.  Streams and pool are keyed on global
    indices and ranks take whole blocks of chains: the counts do not depend on the number of ranks.
    Returns dict(n_ge, p, z_obs, n_nulls)."""
    from . import defaults
    chain_len = defaults.CHAIN_LEN if chain_len is None else chain_len
    burn = defaults.BURN_ROUNDS if burn is None else burn
    pool_size = defaults.POOL_SIZE if pool_size is None else pool_size
    pair_group = defaults.PAIR_GROUP if pair_group is None else pair_group
    rank, world = _world(group)
    lo, hi = shard_chains(n_nulls, chain_len, rank, world, pair_group)
    engine.set_burn_in(burn, pool_size)
    if hasattr(engine, "set_pair_group"):
        engine.set_pair_group(pair_group)
    if shared_pool and burn > 0 and stat == 0:
        build_pool(engine, seed, n_trades, pool_size, group)
    if z_obs is None:
        z_obs = engine.null_pvalue(seed, 0, 0, modules, n_trades=n_trades, stat=stat)["z_obs"]     # same on every rank
    res = engine.null_pvalue(seed, null0 + lo, hi - lo, modules, n_trades=n_trades, chain_len=chain_len, stat=stat, z_obs=z_obs)
    n_ge = allreduce_sum_int64(res["n_ge"], group)
    return {"n_ge": n_ge, "p": n_ge / float(max(n_nulls, 1)), "z_obs": np.asarray(z_obs), "n_nulls": int(n_nulls)}


def scan(engine, profile=0, k=3, topn=8, group=None):
    """Exhaustive scan sharded by smallest feature index; returns the global top-n and the total count."""
    rank, world = _world(group)
    lo, hi = scan_feature_ranges(engine.F, k, world)[rank]
    part = engine.scan(profile, k, topn, g_lo=lo, g_hi=hi)
    parts = [{"W": w, "features": f, "coverage": c} for w, f, c in zip(
        allgather_array(part["W"], group), allgather_array(part["features"], group), allgather_array(part["coverage"], group))]
    out = merge_hits(parts, topn)
    out["n_scored"] = int(allreduce_sum_int64([part["n_scored"]], group)[0])
    return out


def _gather_jobs(local, n_jobs, world, group):
    """local: this rank's slice (rows lo..hi of an n_jobs-row array); returns the full array on every rank."""
    width = max(shard_range(n_jobs, r, world)[1] - shard_range(n_jobs, r, world)[0] for r in range(world))
    pad = np.zeros((width,) + local.shape[1:], dtype=local.dtype)
    pad[:len(local)] = local
    parts = allgather_array(pad, group)
    return np.concatenate([parts[r][:shard_range(n_jobs, r, world)[1] - shard_range(n_jobs, r, world)[0]] for r in range(world)])


def condperm_batch(engine, modules, n_perm, seed, perm_id0=0, profile_of_job=None, group=None):
    """Conditional permutation counts for many (profile, module) jobs, sharded by job."""
    rank, world = _world(group)
    modules = np.atleast_2d(np.asarray(modules, dtype=np.int32))
    n = len(modules)
    pj = np.zeros(n, dtype=np.int32) if profile_of_job is None else np.asarray(profile_of_job, dtype=np.int32)
    lo, hi = shard_range(n, rank, world)
    if hi > lo:
        counts, real = engine.condperm_batch(modules[lo:hi], n_perm, seed, perm_id0, pj[lo:hi])
    else:
        counts, real = np.zeros((0, modules.shape[1]), dtype=np.int64), np.zeros((0, modules.shape[1]))
    return _gather_jobs(counts, n, world, group), _gather_jobs(real, n, world, group)


def ranksum(engine, modules, profile_of_job=None, group=None):
    """Rank-sum z / p for many (profile, module) jobs, sharded by job."""
    rank, world = _world(group)
    modules = np.atleast_2d(np.asarray(modules, dtype=np.int32))
    n = len(modules)
    pj = np.zeros(n, dtype=np.int32) if profile_of_job is None else np.asarray(profile_of_job, dtype=np.int32)
    lo, hi = shard_range(n, rank, world)
    if hi > lo:
        r = engine.ranksum(modules[lo:hi], pj[lo:hi])
    else:
        r = {"z": np.zeros(0), "p": np.zeros(0), "twice_rank_sum": np.zeros(0, dtype=np.int64), "n1": np.zeros(0, dtype=np.int32)}
    return {key: _gather_jobs(np.asarray(r[key]), n, world, group) for key in ("z", "p", "twice_rank_sum", "n1")}
