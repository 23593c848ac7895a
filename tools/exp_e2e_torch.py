import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from superdendrix_b200 import bitpack, synth
from superdendrix_b200.engine import Engine
dense, _, _ = synth.feature_matrix(770, 400)
w = synth.weights(770, 1)
module = synth.pick_module(dense, 3)[None, :].astype(np.int32)
rows = bitpack.pack_rows(dense.T)
torch.cuda.set_device(0)
eng = Engine(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
eng.set_burn_in(6, 1480)
flush = torch.empty(256*1024*1024, dtype=torch.uint8, device="cuda")
for it in range(4):
    flush.zero_()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(stream)
    t0 = time.perf_counter(); eng.upload_matrix(rows, 770); t1 = time.perf_counter(); eng.upload_weights(w); t2 = time.perf_counter()
    eng.null_pvalue(1764, (10000 + it) * 100000, 100000, module, chain_len=16); t3 = time.perf_counter()
    b.record(stream); torch.cuda.synchronize()
    print("upload %.1f weights %.1f nulls %.1f ms; events %.1f ms; lib %s" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), a.elapsed_time(b), eng.timing()))
