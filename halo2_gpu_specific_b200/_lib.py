"""ctypes binding of libb2pcs.so (include/b2pcs.h).  Fails loudly; no fallback."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2PCS_LIB") or os.path.join(_HERE, "libb2pcs.so")  # env: kernel-variant experiments

B2_OK, B2_ERR_ARG, B2_ERR_CUDA, B2_ERR_OOM, B2_ERR_BOUND, B2_ERR_HANDLE = 0, -1, -2, -3, -4, -5


class B2Error(RuntimeError):
    """Raised where the reference panics (.unwrap()/.expect()/assert_eq!)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"b2pcs error {code}: {msg}")
        self.code = code


class NttDesc(ctypes.Structure):
    _fields_ = [
        ("log_n", ctypes.c_uint32),
        ("location", ctypes.c_uint32),
        ("omega", ctypes.c_void_p),
        ("divisor", ctypes.c_void_p),
        ("coset_in", ctypes.c_void_p),
        ("coset_out", ctypes.c_void_p),
        ("n_in", ctypes.c_uint64),
        ("n_out", ctypes.c_uint64),
        ("columns", ctypes.c_uint64),
        ("in_", ctypes.c_void_p),
        ("in_stride", ctypes.c_uint64),
        ("out", ctypes.c_void_p),
        ("out_stride", ctypes.c_uint64),
        ("stream", ctypes.c_void_p),
        ("coset_gen", ctypes.c_void_p),
    ]


# every symbol include/b2pcs.h declares (checked by tests/test_abi.py without a GPU)
SYMBOLS = [
    "b2_version", "b2_last_error", "b2_device_count", "b2_set_device", "b2_get_device", "b2_synchronize",
    "b2_launch_count", "b2_stream_create", "b2_stream_synchronize", "b2_stream_destroy", "b2_srs_register", "b2_srs_synthetic", "b2_srs_from_scalars_dev", "b2_memcpy_d2d", "b2_vanishing_random_poly_dev", "b2_fr_max_bits_dev", "b2_srs_precompute", "b2_srs_len", "b2_srs_read", "b2_srs_free",
    "b2_msm", "b2_msm_dev", "b2_best_multiexp", "b2_g1_sum", "b2_g1_normalize", "b2_g1_sum_dev", "b2_g1_sum_groups_dev", "b2_ntt_exec", "b2_best_fft", "b2_gpu_ifft",
    "b2_coeff_to_extended", "b2_extended_to_coeff", "b2_divide_by_vanishing_poly", "b2_msm_and_ifft",
    "b2_commit_batch", "b2_commit_batch_resident", "b2_host_alloc", "b2_host_free", "b2_host_register", "b2_host_unregister", "b2_dev_alloc", "b2_dev_free", "b2_memcpy_h2d",
    "b2_memcpy_d2h", "b2_field_vec", "b2_imad_probe", "b2_shoup_probe", "b2_mul_probe", "b2_pipe_probe", "b2_msm_async", "b2_msm_wait", "b2_logup_multiplicity_dev", "b2_eval_polynomials_dev", "b2_dfma_probe", "b2_mixed_probe", "b2_affine_batch_probe", "b2_last_timing", "b2_last_msm_phases", "b2_msm_config",
    "b2_quotient_program_create", "b2_quotient_program_free", "b2_quotient_program_info", "b2_quotient_program_slot_classes", "b2_quotient_program_dump", "b2_quotient_eval",
    "b2_g1_decompress", "b2_g1_compress", "b2_srs_register_compressed", "b2_srs_read_compressed",
    "b2_eval_polynomial", "b2_eval_polynomial_dev", "b2_kate_division", "b2_kate_division_dev", "b2_poly_combine", "b2_poly_combine_dev", "b2_witness_file_columns", "b2_commit_witness_file",
    "b2_batch_invert", "b2_batch_invert_dev", "b2_prefix_scan", "b2_prefix_scan_dev", "b2_fr_vec_dev",
]

_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2Error(B2_ERR_CUDA, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                                       f"g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.b2_last_error.restype = ctypes.c_char_p
        L.b2_launch_count.restype = ctypes.c_uint64
        L.b2_launch_count.argtypes = [ctypes.c_int]
        vp, sz, u32, u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint64
        L.b2_stream_create.argtypes = [ctypes.POINTER(ctypes.c_void_p)]
        L.b2_stream_synchronize.argtypes = [ctypes.c_void_p]
        L.b2_stream_destroy.argtypes = [ctypes.c_void_p]
        L.b2_srs_register.argtypes = [vp, sz, sz, ctypes.POINTER(u64)]
        L.b2_srs_synthetic.argtypes = [sz, u64, u64, ctypes.POINTER(u64)]
        L.b2_srs_from_scalars_dev.argtypes = [vp, sz, ctypes.POINTER(u64)]
        L.b2_memcpy_d2d.argtypes = [vp, vp, sz]
        L.b2_vanishing_random_poly_dev.argtypes = [vp, vp, u32, sz, vp, vp]
        L.b2_fr_max_bits_dev.argtypes = [vp, sz, ctypes.POINTER(u32)]
        L.b2_srs_precompute.argtypes = [u64, u32]
        L.b2_g1_normalize.argtypes = [vp, sz]
        L.b2_srs_len.argtypes = [u64, ctypes.POINTER(sz)]
        L.b2_srs_read.argtypes = [u64, sz, sz, vp]
        L.b2_srs_free.argtypes = [u64]
        L.b2_msm.argtypes = [u64, sz, vp, sz, u32, vp]
        L.b2_msm_dev.argtypes = [u64, sz, vp, sz, u32, vp, vp]
        L.b2_msm_async.argtypes = [u64, sz, vp, sz, u32, vp, ctypes.POINTER(u64)]
        L.b2_msm_wait.argtypes = [u64]
        L.b2_logup_multiplicity_dev.argtypes = [vp, u32, vp, u64, u64, vp, ctypes.POINTER(u64)]
        L.b2_best_multiexp.argtypes = [vp, vp, sz, vp]
        L.b2_g1_sum.argtypes = [vp, sz, vp]
        L.b2_g1_sum_dev.argtypes = [vp, sz, vp, vp]
        L.b2_g1_sum_groups_dev.argtypes = [vp, sz, sz, vp, vp]
        L.b2_ntt_exec.argtypes = [ctypes.POINTER(NttDesc)]
        L.b2_best_fft.argtypes = [vp, vp, u32]
        L.b2_gpu_ifft.argtypes = [vp, vp, u32, vp]
        L.b2_coeff_to_extended.argtypes = [vp, vp, u64, u32, u32, vp, vp, vp]
        L.b2_extended_to_coeff.argtypes = [vp, vp, u64, u32, vp, vp, vp, vp]
        L.b2_divide_by_vanishing_poly.argtypes = [vp, u32, vp, u32]
        L.b2_msm_and_ifft.argtypes = [u64, vp, u32, vp, vp, u32, vp]
        L.b2_commit_batch.argtypes = [u64, vp, u64, sz, u32, ctypes.c_int, vp, vp, u32, vp]
        L.b2_commit_batch_resident.argtypes = [u64, vp, ctypes.c_int, vp, u64, sz, u32, ctypes.c_int, vp, vp, u32, vp]
        L.b2_host_alloc.argtypes = [sz, ctypes.POINTER(vp)]
        L.b2_host_free.argtypes = [vp]
        L.b2_host_register.argtypes = [vp, sz]
        L.b2_host_unregister.argtypes = [vp]
        L.b2_dev_alloc.argtypes = [sz, ctypes.POINTER(vp)]
        L.b2_dev_free.argtypes = [vp]
        L.b2_memcpy_h2d.argtypes = [vp, vp, sz]
        L.b2_memcpy_d2h.argtypes = [vp, vp, sz]
        L.b2_field_vec.argtypes = [ctypes.c_int, ctypes.c_int, vp, vp, sz, vp]
        L.b2_imad_probe.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        L.b2_shoup_probe.argtypes = [ctypes.POINTER(ctypes.c_double)]
        L.b2_mul_probe.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
        L.b2_dfma_probe.argtypes = [ctypes.POINTER(ctypes.c_double)]
        L.b2_pipe_probe.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
        L.b2_mixed_probe.argtypes = [u32, u32, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
        L.b2_affine_batch_probe.argtypes = [sz, u32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                            vp, vp, vp, sz]
        L.b2_last_timing.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        L.b2_last_msm_phases.argtypes = [ctypes.POINTER(ctypes.c_double)]
        L.b2_msm_config.argtypes = [u64, sz, u32, ctypes.POINTER(u32), ctypes.POINTER(u32), ctypes.POINTER(u32)]
        L.b2_quotient_program_create.argtypes = [vp, ctypes.POINTER(u64)]
        L.b2_quotient_program_free.argtypes = [u64]
        L.b2_quotient_program_info.argtypes = [u64] + [ctypes.POINTER(u32)] * 4
        L.b2_quotient_program_slot_classes.argtypes = [u64] + [ctypes.POINTER(u32)] * 2
        L.b2_quotient_eval.argtypes = [u64, vp]
        L.b2_g1_decompress.argtypes = [vp, sz, u32, vp]
        L.b2_g1_compress.argtypes = [vp, sz, u32, vp]
        L.b2_srs_register_compressed.argtypes = [vp, sz, u32, ctypes.POINTER(u64)]
        L.b2_srs_read_compressed.argtypes = [u64, sz, sz, u32, vp]
        L.b2_eval_polynomial.argtypes = [vp, u64, vp, vp]
        L.b2_eval_polynomial_dev.argtypes = [vp, u64, u64, u64, vp, vp]
        L.b2_eval_polynomials_dev.argtypes = [ctypes.POINTER(vp), u64, u64, vp, vp]
        L.b2_kate_division.argtypes = [vp, u64, vp, vp]
        L.b2_kate_division_dev.argtypes = [vp, u64, vp, vp, vp]
        L.b2_poly_combine_dev.argtypes = [ctypes.POINTER(vp), ctypes.c_uint32, u64, vp, vp, vp]
        L.b2_poly_combine.argtypes = [ctypes.POINTER(vp), ctypes.c_uint32, u64, vp, vp]
        L.b2_witness_file_columns.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint32)]
        L.b2_commit_witness_file.argtypes = [ctypes.c_uint64, ctypes.c_char_p, ctypes.c_uint32, u64, u64, ctypes.c_uint32, vp, vp]
        L.b2_batch_invert.argtypes = [vp, sz]
        L.b2_batch_invert_dev.argtypes = [vp, sz, vp]
        L.b2_prefix_scan.argtypes = [ctypes.c_int, vp, sz, vp, vp, sz]
        L.b2_prefix_scan_dev.argtypes = [ctypes.c_int, vp, sz, vp, vp, vp, sz, vp]
        L.b2_fr_vec_dev.argtypes = [ctypes.c_int, vp, vp, sz, vp, vp]
        L.b2_quotient_program_dump.argtypes = [u64, vp, sz, ctypes.POINTER(u32), vp, sz, ctypes.POINTER(u32)]
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise B2Error(rc, lib().b2_last_error().decode("utf-8", "replace"))


def require_gpu() -> None:
    """The product path must fail loudly when there is no CUDA device."""
    if lib().b2_device_count() < 1:
        raise B2Error(B2_ERR_CUDA, "no CUDA device visible; this engine has no CPU fallback")


def ptr(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


def as_fr(a, copy: bool = False) -> np.ndarray:
    """(n,4) uint64, C-contiguous"""
    arr = np.asarray(a, dtype=np.uint64)
    if arr.ndim == 1:
        arr = arr.reshape(-1, 4)
    if arr.shape[-1] != 4:
        raise B2Error(B2_ERR_ARG, f"expected (...,4) uint64 limbs, got {arr.shape}")
    arr = np.ascontiguousarray(arr)
    return arr.copy() if copy else arr


def as_fr1(a) -> np.ndarray:
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.uint64).reshape(4))
    return arr


def bind_host_to_gpu(device_index: int) -> Optional[list]:
    """Restrict this process to the CPUs NVML reports as local to the GPU, so that page-locked buffers allocated
    afterwards land on the GPU's NUMA node (first-touch) and its host copies do not cross the socket interconnect.
    Meant for one-process-per-GPU runs with N > 1, where N ranks otherwise pull their columns out of whichever
    node the scheduler happened to place them on.  Returns the CPU list, or None when NVML / the cpuset does not
    allow it (nothing is changed then).  B2_NUMA_BIND=0 disables it."""
    if os.environ.get("B2_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if visible:
            ids = [v for v in visible.split(",") if v.strip() != ""]
            if device_index < len(ids) and ids[device_index].strip().isdigit():
                phys = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def pinned_empty(shape, dtype=np.uint64) -> np.ndarray:
    """numpy array backed by page-locked host memory (b2_host_alloc); freed with the array."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = ctypes.c_void_p()
    check(lib().b2_host_alloc(max(n, 1), ctypes.byref(p)))
    buf = (ctypes.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.ctypes.data] = p.value  # keep: freed explicitly with pinned_free
    return arr


_PINNED = {}


def pinned_free(arr: np.ndarray) -> None:
    p = _PINNED.pop(arr.ctypes.data, None)
    if p is not None:
        lib().b2_host_free(ctypes.c_void_p(p))


def set_device(dev: int) -> None:
    check(lib().b2_set_device(int(dev)))


def last_timing():
    k, t = ctypes.c_double(), ctypes.c_double()
    check(lib().b2_last_timing(ctypes.byref(k), ctypes.byref(t)))
    return k.value, t.value


def last_msm_phases():
    a = (ctypes.c_double * 8)()
    check(lib().b2_last_msm_phases(a))
    names = ["digits", "scan", "scatter", "accumulate", "fixup", "reduce", "final", "total"]
    return dict(zip(names, list(a)))
