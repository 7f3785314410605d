"""The reference's in-process multi-GPU model (one process, N_GPU devices, arithmetic.rs:413-440) on >= 2 GPUs.
Skipped on a single-GPU box; run with `gpurun --gpus 2`."""
import numpy as np
import pytest

import halo2_gpu_specific_b200 as h2
from halo2_gpu_specific_b200.arithmetic import MultiGpuSrs, Srs
from oracle import bn254 as o
from oracle import cref

pytestmark = pytest.mark.gpu


def _need2(gpu):
    if gpu.lib().b2_device_count() < 2:
        pytest.skip("needs 2 GPUs")


def test_in_process_range_sharded_msm(gpu):
    _need2(gpu)
    n = 20001
    scalars = cref.random_fr_mont(n, 0x71)
    bases = cref.g1_mul_gen(cref.from_mont(0, cref.random_fr_mont(n, 0x72)))
    want = cref.jac_to_affine(cref.best_multiexp(scalars, bases, 8))[0]
    multi = MultiGpuSrs(bases)
    assert len(multi.shards) == gpu.lib().b2_device_count()
    got = h2.gpu_multiexp_bound(scalars, multi, 254)
    assert np.array_equal(got[:8], want)
    small = cref.random_fr_small_mont(n, 0x73, 16)
    got = h2.gpu_multiexp_bound(small, multi, 16)
    assert np.array_equal(got[:8], cref.jac_to_affine(cref.best_multiexp(small, bases, 8))[0])
    multi.free()


def test_per_thread_device_selection_and_wrong_device_error(gpu):
    _need2(gpu)
    n = 3000
    scalars = cref.random_fr_mont(n, 0x74)
    bases = cref.g1_mul_gen(cref.from_mont(0, cref.random_fr_mont(n, 0x75)))
    want = cref.jac_to_affine(cref.best_multiexp(scalars, bases, 8))[0]
    gpu.set_device(1)
    srs1 = Srs.register(bases)
    assert np.array_equal(h2.best_multiexp(scalars, srs1)[:8], want)
    k = 12
    dom = h2.EvaluationDomain(5, k)
    x = cref.random_fr_mont(1 << k, 0x76)
    a = x.copy()
    dom.lagrange_to_coeff(a)            # NTT on device 1
    assert np.array_equal(a, cref.ifft(x, dom.omega_inv, dom.ifft_divisor, k, 8))
    gpu.set_device(0)
    with pytest.raises(gpu.B2Error):    # SRS lives on device 1
        h2.best_multiexp(scalars, srs1)
    gpu.set_device(1)
    srs1.free()
    gpu.set_device(0)
