"""Debug: burnt-in nulls of the default and the work-queue kernel against the oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle
from superdendrix_b200 import bitpack, synth
from superdendrix_b200.engine import Engine
dense, _, _ = synth.feature_matrix(770, 400)
S, F = dense.shape
obs = bitpack.pack_rows(dense if F >= S else dense.T)
R = obs.shape[0]
burn, pool, cl, n = (int(x) for x in sys.argv[1:5]) if len(sys.argv) > 4 else (3, 7, 4, 20)
want, _ = oracle.curveball_nulls(obs, 5, 2, n, 5 * R, cl, burn=burn, pool_size=pool)
with Engine(0) as eng:
    eng.upload_dense(dense)
    eng.set_burn_in(burn, pool)
    got, _ = eng.curveball(5, 2, n, -1, cl, want_applied=False)
bad = [i for i in range(n) if not np.array_equal(got[i], want[i])]
print("persist=%s burn=%d pool=%d chain=%d: %d of %d nulls differ %s" % (os.environ.get("SDX_TRADE_PERSIST"), burn, pool, cl, len(bad), n, bad[:20]))
