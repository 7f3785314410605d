"""Mirror of EvaluationDomain (halo2_proofs/src/poly/domain.rs) over the C ABI."""
from __future__ import annotations

import numpy as np

from . import _fr
from ._lib import B2_ERR_ARG, B2Error, as_fr, check, lib, ptr, require_gpu
from .arithmetic import gpu_ifft


class EvaluationDomain:
    """poly/domain.rs:24-149.  `zeta` is a parameter because which primitive cube root the
    pinned pairing crate exports as Fr::ZETA cannot be read from the reference tree; proofs
    do not depend on the choice (the coset is entered and left consistently)."""

    def __init__(self, j: int, k: int, zeta: int = _fr.ZETA):
        R = _fr.R_MOD
        self.quotient_poly_degree = j - 1
        self.k = k
        self.n = 1 << k
        extended_k = k
        while (1 << extended_k) < self.n * self.quotient_poly_degree:
            extended_k += 1
        if extended_k > _fr.S:
            raise B2Error(B2_ERR_ARG, "extended_k exceeds Fr::S")
        self.extended_k = extended_k
        ext_omega = _fr.ROOT_OF_UNITY
        for _ in range(extended_k, _fr.S):
            ext_omega = ext_omega * ext_omega % R
        omega = ext_omega
        for _ in range(k, extended_k):
            omega = omega * omega % R
        self._omega, self._omega_inv = omega, _fr.inv(omega)
        self._ext_omega, self._ext_omega_inv = ext_omega, _fr.inv(ext_omega)
        self._zeta, self._zeta_sq = zeta % R, zeta * zeta % R
        orig, step = pow(zeta, self.n, R), pow(ext_omega, self.n, R)
        t_ev, cur = [], orig
        while True:
            t_ev.append(cur)
            cur = cur * step % R
            if cur == orig:
                break
        assert len(t_ev) == 1 << (extended_k - k)
        self._t_evaluations = [_fr.inv((t - 1) % R) for t in t_ev]
        # Montgomery-form constants, as the reference stores them
        self.omega = _fr.to_mont(omega)
        self.omega_inv = _fr.to_mont(self._omega_inv)
        self.extended_omega = _fr.to_mont(ext_omega)
        self.extended_omega_inv = _fr.to_mont(self._ext_omega_inv)
        self.g_coset = _fr.to_mont(self._zeta)
        self.g_coset_inv = _fr.to_mont(self._zeta_sq)
        self.ifft_divisor = _fr.to_mont(_fr.inv(self.n % R))
        self.extended_ifft_divisor = _fr.to_mont(_fr.inv((1 << extended_k) % R))
        self.t_evaluations = np.stack([_fr.to_mont(t) for t in self._t_evaluations])
        self.barycentric_weight = _fr.to_mont(_fr.inv(self.n % R))

    def extended_len(self) -> int:
        return 1 << self.extended_k

    # ---- transforms --------------------------------------------------------------------
    def lagrange_to_coeff(self, a: np.ndarray) -> np.ndarray:
        """:233-243 (consumes `a`: transformed in place and returned)"""
        if a.size != 4 << self.k:
            raise B2Error(B2_ERR_ARG, "assert_eq!(a.values.len(), 1 << self.k)")
        gpu_ifft(a, self.omega_inv, self.k, self.ifft_divisor)
        return a

    lagrange_to_coeff_st = lagrange_to_coeff  # :249-266 (cuda: gpu_ifft)

    def lagrange_to_coeff_batch(self, cols: np.ndarray) -> np.ndarray:
        """`cols`: (columns, n, 4), transformed in place -- the prover's par_iter over
        columns (plonk/prover.rs:643-646) as one batched call."""
        from ._lib import NttDesc
        import ctypes
        if cols.ndim != 3 or cols.shape[1] != self.n or cols.shape[2] != 4:
            raise B2Error(B2_ERR_ARG, f"expected (columns, {self.n}, 4)")
        require_gpu()
        d = NttDesc()
        d.log_n, d.location = self.k, 0
        d.omega, d.divisor = self.omega_inv.ctypes.data, self.ifft_divisor.ctypes.data
        d.n_in = d.n_out = d.in_stride = d.out_stride = self.n
        d.columns = cols.shape[0]
        d.in_ = d.out = cols.ctypes.data
        check(lib().b2_ntt_exec(ctypes.byref(d)))
        return cols

    def coeff_to_extended(self, a: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """:270-287; accepts (n,4) or a batch (columns, n, 4); returns (.., 2^extended_k, 4).
        `out` may be a preallocated (e.g. pinned) buffer of that shape."""
        arr = as_fr(a) if a.ndim <= 2 else np.ascontiguousarray(a)
        batch = arr.ndim == 3
        cols = arr.shape[0] if batch else 1
        if (arr.shape[1] if batch else arr.shape[0]) != self.n:
            raise B2Error(B2_ERR_ARG, "assert_eq!(a.values.len(), 1 << self.k)")
        require_gpu()
        if out is None:
            out = np.empty((cols, self.extended_len(), 4), dtype=np.uint64)
        elif out.size != cols * self.extended_len() * 4 or not out.flags.c_contiguous:
            raise B2Error(B2_ERR_ARG, "out has the wrong size")
        out = out.reshape(cols, self.extended_len(), 4)
        check(lib().b2_coeff_to_extended(ptr(arr), ptr(out), cols, self.k, self.extended_k, ptr(self.g_coset),
                                         ptr(self.g_coset_inv), ptr(self.extended_omega)))
        return out if batch else out[0]

    def extended_to_coeff(self, a: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """:328-350; returns n * quotient_poly_degree coefficients (`out`: optional preallocated buffer)"""
        arr = as_fr(a)
        if arr.shape[0] != self.extended_len():
            raise B2Error(B2_ERR_ARG, "assert_eq!(a.values.len(), self.extended_len())")
        require_gpu()
        n_out = self.n * self.quotient_poly_degree
        if out is None:
            out = np.empty((n_out, 4), dtype=np.uint64)
        elif out.size != n_out * 4 or not out.flags.c_contiguous:
            raise B2Error(B2_ERR_ARG, "out has the wrong size")
        out = out.reshape(n_out, 4)
        check(lib().b2_extended_to_coeff(ptr(arr), ptr(out), n_out, self.extended_k, ptr(self.g_coset),
                                         ptr(self.g_coset_inv), ptr(self.extended_omega_inv),
                                         ptr(self.extended_ifft_divisor)))
        return out

    def divide_by_vanishing_poly(self, a: np.ndarray) -> np.ndarray:
        """:354-373, in place"""
        if a.size != 4 << self.extended_k:
            raise B2Error(B2_ERR_ARG, "assert_eq!(a.values.len(), self.extended_len())")
        require_gpu()
        check(lib().b2_divide_by_vanishing_poly(ptr(a), self.extended_k, ptr(self.t_evaluations),
                                                self.t_evaluations.shape[0]))
        return a
