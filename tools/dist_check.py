"""torchrun check of superdendrix_b200.dist over NCCL: sharded results equal the single-rank results.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py
"""
import os, sys
import numpy as np
import torch
import torch.distributed as tdist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superdendrix_b200 import bitpack, defaults, dist, synth
from superdendrix_b200.engine import Engine

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = tdist.get_rank(), tdist.get_world_size()
dense, _, _ = synth.feature_matrix(770, 400)
w = synth.weights(770, 3)
mods = np.array([[0, 1, 2], [3, 4, -1], [5, 6, 7]], dtype=np.int32)
with Engine(local) as eng:
    eng.upload_dense(dense)
    eng.upload_weights(w)
    pv = dist.null_pvalue(eng, 1764, 20001, mods)
    one = eng.null_pvalue(1764, 0, 20001, mods, chain_len=defaults.CHAIN_LEN)["n_ge"]   # dist.null_pvalue has set the same burn-in pool (built in slices + all-gather) and pair groups
    assert np.array_equal(pv["n_ge"], one), (pv["n_ge"], one)
    sc = dist.scan(eng, 1, 3, 8)
    whole = eng.scan(1, 3, 8)
    assert np.array_equal(sc["features"], whole["features"]) and np.array_equal(sc["W"], whole["W"]) and sc["n_scored"] == whole["n_scored"]
    jobs = np.tile(mods, (5, 1))
    pj = np.arange(15, dtype=np.int32) % 3
    cp = dist.condperm_batch(eng, jobs, 500, 7, 0, pj)
    assert np.array_equal(cp[0], eng.condperm_batch(jobs, 500, 7, 0, pj)[0])
    rs = dist.ranksum(eng, jobs, pj)
    assert np.array_equal(rs["twice_rank_sum"], eng.ranksum(jobs, pj)["twice_rank_sum"])
if rank == 0:
    print("dist check ok on %d ranks: n_ge %s, scan optimum %s" % (world, pv["n_ge"].tolist(), sc["features"][0].tolist()))
tdist.destroy_process_group()
