"""Mirror of Params (halo2_proofs/src/poly/commitment.rs:23-29, 129-222) over the C ABI."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _fr
from ._lib import B2_ERR_ARG, B2Error, as_fr, as_fr1, check, lib, ptr, require_gpu
from .arithmetic import Srs, best_multiexp_gpu_cond, gpu_multiexp_bound_and_fft, gpu_multiexp_single_gpu_with_bound


class Params:
    """KZG prover parameters with both bases resident in HBM.

    g / g_lagrange: (n,8) uint64 affine arrays (as Params::read would produce), or Srs
    objects that are already resident."""

    def __init__(self, k: int, g, g_lagrange, precompute: bool = True, additional_data: bytes = b""):
        self.additional_data = bytes(additional_data)
        self.k = k
        self.n = 1 << k
        self.g = g if isinstance(g, Srs) else Srs.register(g)
        self.g_lagrange = g_lagrange if isinstance(g_lagrange, Srs) else Srs.register(g_lagrange)
        if len(self.g) != self.n or len(self.g_lagrange) != self.n:
            raise B2Error(B2_ERR_ARG, "g and g_lagrange must hold 2^k points")
        if precompute:
            # window tables of both bases (254/c + 1 copies each in HBM): fewer point additions
            # per commit and no per-window reduction; built once, like the upload
            self.g.precompute()
            self.g_lagrange.precompute()

    @classmethod
    def unsafe_setup(cls, k: int, s: int, precompute: bool = True) -> "Params":
        """Params::unsafe_setup (:56-124) on the device, for a caller-chosen s (the reference draws it from OsRng,
        :61, and forgets it): g[i] = [s^i] G (:63-83), g_lagrange[i] = [(s^n - 1)/n * w^i / (s - w^i)] G (:85-112).
        The scalar side is a prefix product, a batch inversion and three element-wise passes on resident vectors;
        the point side one fixed-base multiplication per point.  2 * 2^k points without touching the host --
        SURVEY 8d config 3's "synthetic SRS built on the GPU by the engine itself"."""
        import ctypes
        from .evaluation import DeviceBuffer
        if k > _fr.S:
            raise B2Error(B2_ERR_ARG, "assert!(k <= Fr::S)")                  # :60
        require_gpu()
        R = _fr.R_MOD
        n = 1 << k
        s %= R
        L = lib()
        vp = ctypes.c_void_p
        one = _fr.to_mont(1)
        const, pw, den = DeviceBuffer(n), DeviceBuffer(n), DeviceBuffer(n)
        try:
            def fill(value: int) -> None:
                const.upload(np.tile(_fr.to_mont(value), (n, 1)))

            def powers(base: int, out: DeviceBuffer) -> None:               # out[i] = base^i
                fill(base)
                check(L.b2_prefix_scan_dev(0, vp(const.ptr), n, ptr(one), None, vp(out.ptr), n, None))

            powers(s, pw)
            g = Srs.from_scalars_dev(pw.ptr, n)
            root = _fr.ROOT_OF_UNITY
            for _ in range(k, _fr.S):
                root = root * root % R
            powers(root, pw)                                                 # w^i
            fill(s)
            check(L.b2_fr_vec_dev(2, vp(const.ptr), vp(pw.ptr), n, vp(den.ptr), None))       # s - w^i
            check(L.b2_batch_invert_dev(vp(den.ptr), n, None))
            check(L.b2_fr_vec_dev(0, vp(pw.ptr), vp(den.ptr), n, vp(den.ptr), None))         # w^i / (s - w^i)
            fill((pow(s, n, R) - 1) * _fr.inv(n % R) % R)
            check(L.b2_fr_vec_dev(0, vp(den.ptr), vp(const.ptr), n, vp(den.ptr), None))      # * (s^n - 1) / n
            check(L.b2_synchronize())
            g_lagrange = Srs.from_scalars_dev(den.ptr, n)
        finally:
            const.free(); pw.free(); den.free()
        return cls(k, g, g_lagrange, precompute=precompute)

    def commit(self, poly) -> np.ndarray:
        """:129-133"""
        p = as_fr(poly)
        size = p.shape[0]
        if len(self.g) < size:
            raise B2Error(B2_ERR_ARG, "assert!(self.g.len() >= size)")
        return best_multiexp_gpu_cond(p, self.g[0:size])

    def commit_lagrange(self, poly) -> np.ndarray:
        """:138-142"""
        p = as_fr(poly)
        size = p.shape[0]
        if len(self.g) < size:
            raise B2Error(B2_ERR_ARG, "assert!(self.g.len() >= size)")
        return best_multiexp_gpu_cond(p, self.g_lagrange[0:size])

    def commit_lagrange_with_bound(self, poly, max_bits: int) -> np.ndarray:
        """:199-222.  The reference first drops zero scalars on the CPU (:204-212); zero
        scalars produce no bucket entries here, so the same point comes out without that
        pass.  max_bits is a contract: a larger scalar raises (B2_ERR_BOUND)."""
        p = as_fr(poly)
        return gpu_multiexp_single_gpu_with_bound(p, self.g_lagrange[0:p.shape[0]], max_bits)

    def commit_lagrange_and_ifft(self, poly: np.ndarray, omega_inv, ifft_divisor):
        """:144-170 (cuda flavour): returns (coefficients, commitment); `poly` is consumed
        (overwritten with its coefficient form), as the reference moves the Vec."""
        c = gpu_multiexp_bound_and_fft(poly, self.g_lagrange, _fr.NUM_BITS, omega_inv, ifft_divisor, self.k)
        return poly, c

    def commit_lagrange_batch(self, cols: np.ndarray, max_bits: int = _fr.NUM_BITS, ifft=None):
        """The prover's per-column commits (plonk/prover.rs:293-299, 470-501, 561-593) as one
        call: cols (columns, n, 4) -> (columns, 12).  ifft=(omega_inv, divisor) also replaces
        every column by its coefficient form (commit_lagrange_and_ifft per column)."""
        if cols.ndim != 3 or cols.shape[2] != 4 or cols.shape[1] > self.n:
            raise B2Error(B2_ERR_ARG, "expected (columns, n, 4)")
        if not (cols.flags.c_contiguous and cols.dtype == np.uint64):
            raise B2Error(B2_ERR_ARG, "cols must be C-contiguous uint64")
        require_gpu()
        out = np.zeros((cols.shape[0], 12), dtype=np.uint64)
        if ifft is None:
            check(lib().b2_commit_batch(self.g_lagrange.handle, ptr(cols), cols.shape[0], cols.shape[1],
                                        int(max_bits), 0, None, None, self.k, ptr(out)))
        else:
            om, dv = as_fr1(ifft[0]), as_fr1(ifft[1])
            check(lib().b2_commit_batch(self.g_lagrange.handle, ptr(cols), cols.shape[0], cols.shape[1],
                                        int(max_bits), 1, ptr(om), ptr(dv), self.k, ptr(out)))
        return out

    def commit_batch(self, cols: np.ndarray) -> np.ndarray:
        """`commit` (:129-133) for several coefficient-form polynomials of one length (the h(X) pieces,
        vanishing/prover.rs:86-96, and the multiopen witness polynomials, gwc/prover.rs:162): (columns, m, 4),
        m <= n, against g[0..m] -> (columns, 12) normalised."""
        if cols.ndim != 3 or cols.shape[2] != 4 or cols.shape[1] > self.n or cols.shape[1] == 0:
            raise B2Error(B2_ERR_ARG, "expected (columns, m <= n, 4)")
        cols = np.ascontiguousarray(cols, dtype=np.uint64)
        require_gpu()
        out = np.zeros((cols.shape[0], 12), dtype=np.uint64)
        check(lib().b2_commit_batch(self.g.handle, ptr(cols), cols.shape[0], cols.shape[1], _fr.NUM_BITS, 0, None, None,
                                    self.k, ptr(out)))
        return out

    def write(self, writer, sign_bit: int = 7) -> None:
        """:241-253: k || g || g_lagrange (32-byte compressed points) || len || additional_data.  The points are
        compressed on the device from the resident SRS."""
        writer.write(int(self.k).to_bytes(4, "little"))
        writer.write(self.g.read_compressed(sign_bit))
        writer.write(self.g_lagrange.read_compressed(sign_bit))
        writer.write(len(self.additional_data).to_bytes(4, "little"))
        writer.write(self.additional_data)

    @classmethod
    def read(cls, reader, sign_bit: int = 7, precompute: bool = True) -> "Params":
        """:256-294.  The compressed points go to the device as read; the Fq square roots the reference computes
        with `parallelize` on the CPU (:262-273) run there.  Raises B2Error(B2_ERR_ARG) on an invalid point."""
        head = reader.read(4)
        if len(head) != 4:
            raise B2Error(B2_ERR_ARG, "params: truncated header")
        k = int.from_bytes(head, "little")
        if k > _fr.S:
            raise B2Error(B2_ERR_ARG, f"params: k = {k} exceeds Fr::S")
        n = 1 << k
        parts = []
        for _ in range(2):
            b = reader.read(32 * n)
            if len(b) != 32 * n:
                raise B2Error(B2_ERR_ARG, "params: truncated point section")
            parts.append(b)
        ln = reader.read(4)
        if len(ln) != 4:
            raise B2Error(B2_ERR_ARG, "params: truncated additional-data length")
        ln = int.from_bytes(ln, "little")
        extra = reader.read(ln)
        if len(extra) != ln:
            raise B2Error(B2_ERR_ARG, "params: truncated additional data")
        g = Srs.register_compressed(parts[0], n, sign_bit)
        try:
            gl = Srs.register_compressed(parts[1], n, sign_bit)
        except B2Error:
            g.free()
            raise
        return cls(k, g, gl, precompute=precompute, additional_data=extra)

    def free(self) -> None:
        self.g.free()
        self.g_lagrange.free()


class ParamsVerifier:
    """poly/commitment.rs:33-40, 384-389: only the piece on the hot path, the small MSM over the public
    inputs (`commit_lagrange` against the first public_inputs_size Lagrange bases)."""

    def __init__(self, k: int, g_lagrange):
        self.k = k
        self.n = 1 << k
        self.g_lagrange = g_lagrange if isinstance(g_lagrange, Srs) else Srs.register(g_lagrange)

    def commit_lagrange(self, scalars) -> np.ndarray:
        """:384-389"""
        p = as_fr(scalars)
        if p.shape[0] > len(self.g_lagrange):
            raise B2Error(B2_ERR_ARG, "more scalars than Lagrange bases")
        return best_multiexp_gpu_cond(p, self.g_lagrange[0:p.shape[0]])

    def free(self) -> None:
        self.g_lagrange.free()
