"""The C-ABI library loads, exports every symbol include/b2pcs.h declares, and (without a GPU)
refuses to compute instead of falling back to a CPU path."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b2pcs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from halo2_gpu_specific_b200 import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in b2pcs.h but not exported"
    assert sorted(_lib.SYMBOLS) == names
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    exported = set(re.findall(r"\sT\s+(b2_\w+)", out))
    assert set(names) <= exported


def test_header_is_plain_c():
    """no torch / C++ types in the signatures: the header compiles as C"""
    subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "b2pcs.h")])


def test_library_is_sm100a_native():
    from halo2_gpu_specific_b200 import _lib
    out = subprocess.check_output(["cuobjdump", "-lelf", _lib.LIB_PATH]).decode()
    assert "sm_100a" in out


def test_no_cpu_fallback_without_gpu():
    from halo2_gpu_specific_b200 import _lib
    import halo2_gpu_specific_b200 as h2
    L = _lib.lib()
    assert L.b2_version() >= 100
    if L.b2_device_count() > 0:
        pytest.skip("a GPU is present")
    a = np.zeros((4, 4), dtype=np.uint64)
    om = np.zeros(4, dtype=np.uint64)
    with pytest.raises(_lib.B2Error):
        h2.best_fft(a, om, 2)
    with pytest.raises(_lib.B2Error):
        h2.best_multiexp(a, np.zeros((4, 8), dtype=np.uint64))
    # raw ABI: error code, not a crash
    assert L.b2_best_fft(a.ctypes.data_as(ctypes.c_void_p), om.ctypes.data_as(ctypes.c_void_p), 2) != 0
    assert len(L.b2_last_error()) > 0


def test_argument_errors_mirror_reference_asserts():
    import halo2_gpu_specific_b200 as h2
    from halo2_gpu_specific_b200 import _lib
    a = np.zeros((5, 4), dtype=np.uint64)
    with pytest.raises(_lib.B2Error):           # assert_eq!(n, 1 << log_n)  arithmetic.rs:569
        h2.best_fft(a, np.zeros(4, dtype=np.uint64), 2)
    with pytest.raises(_lib.B2Error):           # assert_eq!(coeffs.len(), bases.len())  arithmetic.rs:466
        h2.best_multiexp(np.zeros((4, 4), dtype=np.uint64), np.zeros((3, 8), dtype=np.uint64))
    # empty input / max_bits == 0 -> identity without touching a device (arithmetic.rs:346, 443)
    ident = h2.best_multiexp_gpu_cond(np.zeros((0, 4), dtype=np.uint64), np.zeros((0, 8), dtype=np.uint64))
    assert ident[8:].sum() == 0
    ident = h2.gpu_multiexp_single_gpu_with_bound(np.ones((2, 4), dtype=np.uint64), np.ones((2, 8), dtype=np.uint64), 0)
    assert ident[8:].sum() == 0


def test_domain_constants_match_oracle():
    """EvaluationDomain::new (poly/domain.rs:44-149): host-side constants, no GPU needed"""
    import halo2_gpu_specific_b200 as h2
    from oracle import bn254 as o
    for j, k in ((1, 3), (5, 6), (4, 10), (9, 12)):
        d, r = h2.EvaluationDomain(j, k), o.EvaluationDomain(j, k)
        assert d.extended_k == r.extended_k and d.n == r.n
        enc = lambda v: o.fr_encode([v])[0]  # noqa: E731
        for name in ("omega", "omega_inv", "extended_omega", "extended_omega_inv", "g_coset", "g_coset_inv",
                     "ifft_divisor", "extended_ifft_divisor", "barycentric_weight"):
            assert np.array_equal(getattr(d, name), enc(getattr(r, name))), name
        assert np.array_equal(d.t_evaluations, o.fr_encode(r.t_evaluations))


def test_rust_ffi_file_covers_the_abi():
    """integration/b2pcs.rs (generated from include/b2pcs.h by tools/gen_rust_ffi.py) is up to date and declares every
    symbol the library exports exactly once, with the arity the ctypes prototypes use"""
    import sys
    from halo2_gpu_specific_b200 import _lib
    root = ROOT
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_rust_ffi.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    text = open(os.path.join(root, "integration", "b2pcs.rs")).read()
    decl = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (b2_\w+)\((.*?)\) -> ", text)}
    L = _lib.lib()
    for name in _lib.SYMBOLS:
        assert text.count(f"pub fn {name}(") == 1, name
        argtypes = getattr(getattr(L, name), "argtypes", None)
        if argtypes is not None:
            n_rust = 0 if not decl[name].strip() else decl[name].count(",") + 1
            assert n_rust == len(argtypes), (name, decl[name], len(argtypes))
    assert "pub struct B2NttDesc" in text and "pub struct B2QuotientArgs" in text and "B2_MAX_BITS_AUTO: u32 = 0xFFFFFFFF" in text
