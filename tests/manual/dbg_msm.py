import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import halo2_gpu_specific_b200 as h2
from halo2_gpu_specific_b200 import _lib
from halo2_gpu_specific_b200.arithmetic import Srs
from oracle import cref, bn254 as o
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
scalars = cref.random_fr_mont(n, 0xB2000003 + n)
ks = cref.from_mont(0, cref.random_fr_mont(n, 0x51 + n))
bases = cref.g1_mul_gen(ks)
want = cref.jac_to_affine(cref.best_multiexp(scalars, bases, 8))[0]
srs = Srs.register(bases)
for rep in range(4):
    got = h2.best_multiexp(scalars, srs)
    print(n, rep, np.array_equal(got[:8], want), {k: round(v, 3) for k, v in _lib.last_msm_phases().items()})
srs2 = Srs.register(bases).precompute()
for rep in range(2):
    got = h2.best_multiexp(scalars, srs2)
    print('pre', n, rep, np.array_equal(got[:8], want))
