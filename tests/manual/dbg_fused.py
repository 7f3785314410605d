import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import halo2_gpu_specific_b200 as h2
from halo2_gpu_specific_b200 import _fr, _lib
from halo2_gpu_specific_b200.arithmetic import Srs
from concurrent.futures import ThreadPoolExecutor
k = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = 1 << k
dom = h2.EvaluationDomain(5, k)
g = Srs.synthetic(n, 0, 1); gl = Srs.synthetic(n, n, 1)
params = h2.Params(k, g, gl)
tbl = np.stack([_fr.to_mont(v) for v in range(1 << 16)])
rng = np.random.default_rng(1)
cols = _lib.pinned_empty((8, n, 4))
for i in range(8):
    cols[i] = tbl[rng.integers(0, 1 << 16, size=n)]
cols0 = np.array(cols)
for i in range(8):
    try:
        p = params.commit_lagrange_batch(cols[i:i+1], 16, ifft=(dom.omega_inv, dom.ifft_divisor))
        print("single ok", i)
    except Exception as e:
        print("single FAIL", i, e)
cols[:] = cols0
def one(i):
    _lib.set_device(0)
    try:
        params.commit_lagrange_batch(cols[i:i+1], 16, ifft=(dom.omega_inv, dom.ifft_divisor)); return (i, "ok")
    except Exception as e:
        return (i, str(e))
with ThreadPoolExecutor(3) as ex:
    print(list(ex.map(one, range(8))))
cols[:] = cols0
try:
    params.commit_lagrange_batch(cols, 16, ifft=(dom.omega_inv, dom.ifft_divisor)); print("batch ok")
except Exception as e:
    print("batch FAIL", e)
