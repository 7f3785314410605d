"""Witness file (halo2_proofs/src/helpers.rs:919-1015): commitments straight from the file equal the per-column commits."""
import ctypes

import numpy as np
import pytest

from oracle import cref

import halo2_gpu_specific_b200 as h2
from halo2_gpu_specific_b200 import _lib, helpers
from halo2_gpu_specific_b200.arithmetic import Srs
from halo2_gpu_specific_b200.evaluation import DeviceBuffer

pytestmark = pytest.mark.gpu


def test_commit_witness_file_matches_per_column_commits(gpu, tmp_path):
    k, cols = 12, 7
    n = 1 << k
    g = Srs.synthetic(n, 0, 0xB2000071)
    gl = Srs.synthetic(n, n, 0xB2000071)
    params = h2.Params(k, g, gl)
    advice = [cref.random_fr_small_mont(n, 0x700 + i, 16) if i % 2 else cref.random_fr_mont(n, 0x700 + i) for i in range(cols)]
    advice[3][::2] = 0
    path = str(tmp_path / "witness.bin")
    helpers.store_witness(path, advice, k)
    n_cols = ctypes.c_uint32()
    _lib.check(_lib.lib().b2_witness_file_columns(path.encode(), ctypes.byref(n_cols)))
    assert n_cols.value == cols
    want = np.stack([params.commit_lagrange(a) for a in advice])
    got = helpers.commit_witness_file(params, path)
    assert np.array_equal(got, want)
    # a sub-range, kept resident: the device copy is the file's bytes
    keep = DeviceBuffer(3 * n)
    got2 = helpers.commit_witness_file(params, path, first=2, count=3, d_keep=keep.ptr)
    assert np.array_equal(got2, want[2:5])
    assert np.array_equal(keep.download(3 * n).reshape(3, n, 4), np.stack(advice[2:5]))
    keep.free()
    with pytest.raises(_lib.B2Error):
        helpers.commit_witness_file(params, path, first=5, count=4)
