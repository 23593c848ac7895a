"""world_size-2 gloo test of the multi-GPU path's host logic (superdendrix_b200.dist): disjoint null / subset /
job ranges per rank and the reductions give exactly the single-process result.  The compute under the sharding is
the oracle (tests/_oracle_engine.py); on the GPU box the same functions run over Engine + NCCL."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as tdist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _inputs():
    rng = np.random.default_rng(42)
    dense = (rng.random((60, 24)) < 0.2).astype(np.uint8)
    dense[:, 0] = rng.random(60) < 0.5
    w = np.clip(2 * rng.standard_normal((2, 60)), -20, 20)
    modules = np.array([[0, 3, 5], [1, 2, -1]], dtype=np.int32)
    return dense, w, modules


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    from _oracle_engine import OracleEngine
    from superdendrix_b200 import dist
    dense, w, modules = _inputs()
    eng = OracleEngine(dense, w)
    pv = dist.null_pvalue(eng, seed=7, n_nulls=37, modules=modules, chain_len=4, burn=3, pool_size=5, pair_group=1)
    pg = dist.null_pvalue(eng, seed=7, n_nulls=150, modules=modules, chain_len=1, burn=2, pool_size=64, pair_group=32)      # blocks of 32 chains per rank
    sc = dist.scan(eng, profile=0, k=3, topn=6)
    jobs = np.array([[0, 3, 5], [1, 2, -1], [4, -1, -1], [0, 1, 2], [6, 7, -1]], dtype=np.int32)
    pj = np.array([0, 1, 1, 0, 1], dtype=np.int32)
    cp = dist.condperm_batch(eng, jobs, 50, seed=3, perm_id0=8, profile_of_job=pj)
    rs = dist.ranksum(eng, jobs, pj)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), n_ge=pv["n_ge"], p=pv["p"], g_ge=pg["n_ge"], sW=sc["W"], sF=sc["features"], sC=sc["coverage"],
             sN=sc["n_scored"], cp=cp[0], cpW=cp[1], z=rs["z"], s2=rs["twice_rank_sum"])
    tdist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_reduce_to_the_single_process_result(tmp_path):
    sys.path.insert(0, HERE)
    from _oracle_engine import OracleEngine
    from superdendrix_b200 import dist
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    dense, w, modules = _inputs()
    eng = OracleEngine(dense, w)
    one = dist.null_pvalue(eng, seed=7, n_nulls=37, modules=modules, chain_len=4, burn=3, pool_size=5, pair_group=1)      # world = 1 path
    one_g = dist.null_pvalue(eng, seed=7, n_nulls=150, modules=modules, chain_len=1, burn=2, pool_size=64, pair_group=32)
    assert dist.shard_chains(150, 1, 0, 2, 32) == (0, 96) and dist.shard_chains(150, 1, 1, 2, 32) == (96, 150)
    sc1 = eng.scan(0, 3, 6)
    jobs = np.array([[0, 3, 5], [1, 2, -1], [4, -1, -1], [0, 1, 2], [6, 7, -1]], dtype=np.int32)
    pj = np.array([0, 1, 1, 0, 1], dtype=np.int32)
    cp1 = eng.condperm_batch(jobs, 50, 3, 8, pj)
    rs1 = eng.ranksum(jobs, pj)
    for r in range(2):
        z = np.load(tmp_path / ("rank%d.npz" % r))
        assert np.array_equal(z["n_ge"], one["n_ge"]) and np.allclose(z["p"], one["p"])
        assert np.array_equal(z["g_ge"], one_g["n_ge"])
        assert np.array_equal(z["sF"], sc1["features"]) and np.array_equal(z["sW"], sc1["W"]) and np.array_equal(z["sC"], sc1["coverage"])
        assert int(z["sN"]) == sc1["n_scored"] == 24 + 24 * 23 // 2 + 24 * 23 * 22 // 6
        assert np.array_equal(z["cp"], cp1[0]) and np.array_equal(z["cpW"], cp1[1])
        assert np.array_equal(z["z"], rs1["z"]) and np.array_equal(z["s2"], rs1["twice_rank_sum"])
