"""Mirror of the hot-path entry points of halo2_proofs/src/arithmetic.rs over the C ABI.

Arrays are numpy uint64 in the reference's memory layout: scalars (n,4) Montgomery Fr,
affine bases (n,8), results (12,) Jacobian normalised to Z = 1.  `bases` may also be a
resident `Srs` (or a slice of one), which skips the per-call upload the reference does in
gpu_multiexp_single_gpu_with_bound (arithmetic.rs:349-360).
"""
from __future__ import annotations

import ctypes
from typing import Union

import numpy as np

from . import _fr
from ._lib import B2_ERR_ARG, B2Error, as_fr, as_fr1, check, lib, ptr, require_gpu


class Srs:
    """Bases resident in HBM (b2_srs_register).  Slicing gives a view (offset, length)."""

    def __init__(self, handle: int, n: int, offset: int = 0, owner: "Srs | None" = None):
        self.handle, self.n, self.offset, self._owner = handle, n, offset, owner

    @classmethod
    def register(cls, bases: np.ndarray) -> "Srs":
        require_gpu()
        b = np.ascontiguousarray(np.asarray(bases, dtype=np.uint64))
        if b.ndim != 2 or b.shape[1] not in (8, 9):
            raise B2Error(B2_ERR_ARG, f"bases must be (n,8) [or (n,9) with a trailing flag word], got {b.shape}")
        h = ctypes.c_uint64()
        check(lib().b2_srs_register(ptr(b), b.shape[0], b.shape[1] * 8, ctypes.byref(h)))
        return cls(h.value, b.shape[0])

    @classmethod
    def register_compressed(cls, data: bytes, n: int, sign_bit: int = 7) -> "Srs":
        """n 32-byte compressed points as they sit in a params file (poly/commitment.rs:262-273): decompressed on
        the device straight into the resident SRS"""
        require_gpu()
        buf = np.frombuffer(data, dtype=np.uint8, count=32 * n)
        h = ctypes.c_uint64()
        check(lib().b2_srs_register_compressed(ctypes.c_void_p(buf.ctypes.data), n, sign_bit, ctypes.byref(h)))
        return cls(h.value, n)

    def read_compressed(self, sign_bit: int = 7) -> bytes:
        out = np.empty(32 * self.n, dtype=np.uint8)
        check(lib().b2_srs_read_compressed(self.handle, self.offset, self.n, sign_bit, ctypes.c_void_p(out.ctypes.data)))
        return out.tobytes()

    @classmethod
    def synthetic(cls, n: int, first_index: int = 0, seed: int = 0xB2000003) -> "Srs":
        require_gpu()
        h = ctypes.c_uint64()
        check(lib().b2_srs_synthetic(n, first_index, seed, ctypes.byref(h)))
        return cls(h.value, n)

    @classmethod
    def from_scalars_dev(cls, d_scalars: int, n: int) -> "Srs":
        """bases[i] = [k_i] G for n Montgomery-form scalars at device pointer d_scalars (b2_srs_from_scalars_dev)"""
        require_gpu()
        h = ctypes.c_uint64()
        check(lib().b2_srs_from_scalars_dev(ctypes.c_void_p(d_scalars), n, ctypes.byref(h)))
        return cls(h.value, n)

    def precompute(self, window_bits: int = 0) -> "Srs":
        """build the window table (b2_srs_precompute): one shared bucket set, wider windows"""
        check(lib().b2_srs_precompute(self.handle, int(window_bits)))
        return self

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, s: slice) -> "Srs":
        lo, hi, step = s.indices(self.n)
        if step != 1:
            raise B2Error(B2_ERR_ARG, "Srs slices must be contiguous")
        return Srs(self.handle, max(hi - lo, 0), self.offset + lo, owner=self._owner or self)

    def read(self, offset: int = 0, count: int | None = None) -> np.ndarray:
        count = self.n - offset if count is None else count
        out = np.empty((count, 8), dtype=np.uint64)
        check(lib().b2_srs_read(self.handle, self.offset + offset, count, ptr(out)))
        return out

    def free(self) -> None:
        if self._owner is None and self.handle:
            lib().b2_srs_free(self.handle)
            self.handle = 0


Bases = Union[np.ndarray, Srs]


def _identity() -> np.ndarray:
    out = np.zeros(12, dtype=np.uint64)
    q_one = [0xd35d438dc58f0d9d, 0x0a78eb28f5c70b3d, 0x666ea36f7879462c, 0x0e0a77c19a07df2f]
    out[4:8] = q_one
    return out


class MsmFuture:
    """A multiexp in flight (b2_msm_async): `result()` blocks until the point is there.  Keeps the scalar array alive."""

    def __init__(self, ticket: int, out: np.ndarray, keep):
        self._ticket, self._out, self._keep, self._done = ticket, out, keep, False

    def result(self) -> np.ndarray:
        if not self._done:
            self._done = True
            check(lib().b2_msm_wait(self._ticket))
            self._keep = None
        return self._out


def gpu_multiexp_async(coeffs, srs: "Srs", max_bits: int = _fr.NUM_BITS) -> MsmFuture:
    """gpu_multiexp_single_gpu_with_bound (arithmetic.rs:334-367) without waiting for the result: the copy of the
    scalars, the MSM and the read-back are enqueued on a free lane and a future is returned, so ONE caller thread can
    keep the copy of column c + 1 under the kernels of column c (the reference overlaps them with several rayon
    workers).  `coeffs` must not be modified until `.result()` returns; pinned memory keeps this call non-blocking."""
    c = as_fr(coeffs)
    if c.shape[0] != len(srs):
        raise B2Error(B2_ERR_ARG, f"coeffs ({c.shape[0]}) and bases ({len(srs)}) differ in length")
    require_gpu()
    out = np.zeros(12, dtype=np.uint64)
    t = ctypes.c_uint64()
    check(lib().b2_msm_async(srs.handle, srs.offset, ptr(c), c.shape[0], int(max_bits), ptr(out), ctypes.byref(t)))
    return MsmFuture(t.value, out, c)


def gpu_multiexp_single_gpu_with_bound(coeffs, bases: Bases, max_bits: int) -> np.ndarray:
    """arithmetic.rs:334-367.  max_bits == 0 -> identity (:346)."""
    c = as_fr(coeffs)
    n = c.shape[0]
    nb = len(bases) if isinstance(bases, Srs) else np.asarray(bases).shape[0]
    if n != nb:
        raise B2Error(B2_ERR_ARG, f"coeffs ({n}) and bases ({nb}) differ in length")  # assert_eq! :466
    if max_bits == 0 or n == 0:
        return _identity()
    require_gpu()
    out = np.zeros(12, dtype=np.uint64)
    if isinstance(bases, Srs):
        check(lib().b2_msm(bases.handle, bases.offset, ptr(c), n, int(max_bits), ptr(out)))
    else:
        srs = Srs.register(bases)
        try:
            check(lib().b2_msm(srs.handle, 0, ptr(c), n, int(max_bits), ptr(out)))
        finally:
            srs.free()
    return out


def gpu_multiexp_single_gpu(coeffs, bases: Bases) -> np.ndarray:
    """arithmetic.rs:309-311"""
    return gpu_multiexp_single_gpu_with_bound(coeffs, bases, 254)


class MultiGpuSrs:
    """The reference's in-process multi-GPU model (N_GPU devices behind one pool, plonk/prover.rs:56-74):
    the bases are split by point range, part_len = ceil(n / n_gpu) (arithmetic.rs:426), and shard g lives
    on device g.  `devices` defaults to every visible device."""

    def __init__(self, bases: np.ndarray, devices=None, precompute: bool = True):
        require_gpu()
        b = np.ascontiguousarray(np.asarray(bases, dtype=np.uint64))
        self.devices = list(range(lib().b2_device_count())) if devices is None else list(devices)
        self.n = b.shape[0]
        ng = len(self.devices)
        self.part_len = (self.n + ng - 1) // ng if self.n else 0
        self.shards = []
        prev = lib().b2_get_device()
        try:
            for g, dev in enumerate(self.devices):
                lo = min(g * self.part_len, self.n)
                hi = min(lo + self.part_len, self.n)
                if hi == lo:
                    break
                check(lib().b2_set_device(dev))
                srs = Srs.register(b[lo:hi])
                if precompute:
                    srs.precompute()
                self.shards.append((dev, lo, hi, srs))
        finally:
            lib().b2_set_device(prev)

    def __len__(self) -> int:
        return self.n

    def free(self) -> None:
        for _, _, _, srs in self.shards:
            srs.free()
        self.shards = []


def gpu_multiexp_bound(coeffs, bases, max_bits: int) -> np.ndarray:
    """arithmetic.rs:413-440: split by point range over the GPUs, one partial per device, sum of the
    partials.  With a MultiGpuSrs the split runs inside this process (one host thread per device, as the
    reference's par_chunks does); with a single-device Srs / array it is the single-GPU call, and the
    one-process-per-GPU variant lives in parallel.sharded_msm."""
    if not isinstance(bases, MultiGpuSrs):
        return gpu_multiexp_single_gpu_with_bound(coeffs, bases, max_bits)
    c = as_fr(coeffs)
    if c.shape[0] != len(bases):
        raise B2Error(B2_ERR_ARG, f"coeffs ({c.shape[0]}) and bases ({len(bases)}) differ in length")
    if max_bits == 0 or c.shape[0] == 0:
        return _identity()
    from concurrent.futures import ThreadPoolExecutor

    def part(shard):
        dev, lo, hi, srs = shard
        check(lib().b2_set_device(dev))       # per-thread device selection
        return gpu_multiexp_single_gpu_with_bound(c[lo:hi], srs, max_bits)

    with ThreadPoolExecutor(len(bases.shards)) as ex:
        partials = list(ex.map(part, bases.shards))
    return g1_sum(np.stack(partials))


def gpu_multiexp(coeffs, bases: Bases) -> np.ndarray:
    """arithmetic.rs:370-372"""
    return gpu_multiexp_bound(coeffs, bases, _fr.NUM_BITS)


def best_multiexp(coeffs, bases: Bases) -> np.ndarray:
    """arithmetic.rs:465-492 (panics on length mismatch -> B2Error)."""
    return gpu_multiexp_bound(coeffs, bases, _fr.NUM_BITS)


def small_multiexp(coeffs, bases: Bases) -> np.ndarray:
    """arithmetic.rs:112-132.  The reference keeps this double-and-add loop on the CPU for a handful of
    points; the same point comes out of the engine's MSM, so the mirror routes it there."""
    return gpu_multiexp_bound(coeffs, bases, _fr.NUM_BITS)


def best_multiexp_gpu_cond(coeffs, bases: Bases) -> np.ndarray:
    """arithmetic.rs:442-458: empty -> identity; otherwise the GPU path.  (The reference
    keeps n <= 2^14 on the CPU; this engine has no CPU path, the result is the same point.)"""
    if as_fr(coeffs).shape[0] == 0:
        return _identity()
    return gpu_multiexp(coeffs, bases)


def gpu_multiexp_bound_and_fft(coeffs: np.ndarray, bases: Srs, max_bits: int, omega, divisor, log_n: int) -> np.ndarray:
    """arithmetic.rs:375-410: MSM of `coeffs` against bases and in-place iFFT of the same
    vector with one upload.  `coeffs` (a writable (2^log_n,4) array) is overwritten."""
    if not isinstance(bases, Srs):
        raise B2Error(B2_ERR_ARG, "gpu_multiexp_bound_and_fft needs a resident Srs")
    if not (isinstance(coeffs, np.ndarray) and coeffs.dtype == np.uint64 and coeffs.flags.c_contiguous
            and coeffs.flags.writeable):
        raise B2Error(B2_ERR_ARG, "coeffs must be a writable C-contiguous uint64 array")
    if coeffs.size != 4 << log_n or len(bases) < (1 << log_n):
        raise B2Error(B2_ERR_ARG, "length mismatch")
    require_gpu()
    out = np.zeros(12, dtype=np.uint64)
    om, dv = as_fr1(omega), as_fr1(divisor)
    if bases.offset != 0:
        raise B2Error(B2_ERR_ARG, "commit+ifft uses the SRS from its start")
    check(lib().b2_msm_and_ifft(bases.handle, ptr(coeffs), int(max_bits), ptr(om), ptr(dv), int(log_n), ptr(out)))
    return out


def _inplace(a) -> np.ndarray:
    if not (isinstance(a, np.ndarray) and a.dtype == np.uint64 and a.flags.c_contiguous and a.flags.writeable):
        raise B2Error(B2_ERR_ARG, "a must be a writable C-contiguous uint64 array (it is transformed in place)")
    return a


def best_fft(a: np.ndarray, omega, log_n: int) -> None:
    """arithmetic.rs:546-554: in-place, natural order in and out.  assert_eq!(n, 1 << log_n) (:569)."""
    a = _inplace(a)
    if a.size != 4 << log_n:
        raise B2Error(B2_ERR_ARG, f"len {a.size // 4} != 1 << {log_n}")
    require_gpu()
    om = as_fr1(omega)
    check(lib().b2_best_fft(ptr(a), ptr(om), int(log_n)))


gpu_fft = best_fft  # arithmetic.rs:495-512


def gpu_ifft(a: np.ndarray, omega, log_n: int, divisor) -> None:
    """arithmetic.rs:515-534"""
    a = _inplace(a)
    if a.size != 4 << log_n:
        raise B2Error(B2_ERR_ARG, f"len {a.size // 4} != 1 << {log_n}")
    require_gpu()
    om, dv = as_fr1(omega), as_fr1(divisor)
    check(lib().b2_gpu_ifft(ptr(a), ptr(om), int(log_n), ptr(dv)))


def g1_sum(points) -> np.ndarray:
    """Sum of Jacobian points: the host-side reduce of gpu_multiexp_bound (arithmetic.rs:428-435)."""
    p = np.ascontiguousarray(np.asarray(points, dtype=np.uint64).reshape(-1, 12))
    require_gpu()
    out = np.zeros(12, dtype=np.uint64)
    check(lib().b2_g1_sum(ptr(p), p.shape[0], ptr(out)))
    return out


def eval_polynomial(poly: np.ndarray, point) -> np.ndarray:
    """arithmetic.rs:714-735: poly (n, 4) coefficients, point (4,) -> (4,) Montgomery"""
    require_gpu()
    p = as_fr(poly)
    pt = as_fr1(point)
    out = np.empty(4, dtype=np.uint64)
    check(lib().b2_eval_polynomial(ptr(p), p.shape[0], ptr(pt), ptr(out)))
    return out


def kate_division(a: np.ndarray, b) -> np.ndarray:
    """arithmetic.rs:752-773: (a(X) - a(b)) / (X - b), n - 1 coefficients"""
    require_gpu()
    p = as_fr(a)
    if p.shape[0] < 2:
        raise B2Error(B2_ERR_ARG, "kate_division needs at least two coefficients")
    out = np.empty((p.shape[0] - 1, 4), dtype=np.uint64)
    check(lib().b2_kate_division(ptr(p), p.shape[0], ptr(as_fr1(b)), ptr(out)))
    return out


def poly_combine(polys, v) -> np.ndarray:
    """poly/multiopen/gwc/prover.rs:47-56: fold `batch = batch * v + poly` over the polynomials opened at one point.
    polys: sequence of (n, 4) coefficient arrays (same n); v: (4,) Montgomery -> (n, 4)"""
    require_gpu()
    cols = [as_fr(p) for p in polys]
    if not cols or any(c.shape != cols[0].shape for c in cols):
        raise B2Error(B2_ERR_ARG, "poly_combine needs at least one polynomial and equal lengths")
    import ctypes
    arr = (ctypes.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    out = np.empty_like(cols[0])
    check(lib().b2_poly_combine(arr, len(cols), cols[0].shape[0], ptr(as_fr1(v)), ptr(out)))
    return out
