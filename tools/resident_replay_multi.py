#!/usr/bin/env python
"""Device-resident create_proof schedule replay over N GPUs of one node (zkWasm-scale k = 22
on 8 x B200).   torchrun --nproc-per-node N tools/resident_replay_multi.py [--k 22] [--reps 2] [--out file.json]

Same schedule and synthetic shape as tools/resident_replay.py (halo2_proofs/src/plonk/prover.rs:206-850), one process
per GPU.  What is sharded and what is exchanged:

  phases 1-3  witness columns are dealt to ranks in contiguous blocks (parallel.column_range); every rank pulls ITS
              columns through ITS PCIe link, commits them there (b2_commit_batch_resident).          no collective
  exchange A  all-gather of the advice columns in Lagrange form (the z construction reads them)      NCCL, A * 2^k * 32 B
  phases 5-6  the P + S + H z columns are dealt the same way: built, committed and inverse-transformed on their rank
              (the 32-byte hand-over of z[last] between permutation sets is not modelled across ranks)  no collective
  phase  8    every rank inverse-transforms its own advice / m columns                                no collective
  exchange B  all-gather of the coefficient forms of every witness-dependent polynomial              NCCL, 97 * 2^k * 32 B
  phase  9    evaluate_h: rows of the extended domain split coset-major (parallel.quotient_tasks); ranks that share a
              coset split its transforms and swap the halves point-to-point; all-gather of the h slices;
              extended_to_coeff on every rank                                  NCCL, 2^(k+2) * 32 B (+ P2P 97/2 * 2^k * 32 B)
  phase 10    the D pieces of h are committed on D different ranks                                    no collective
  phase 11    evaluations: every rank evaluates its block of each polynomial group                    no collective
  phase 12    multiopen: the fold poly_batch = poly_batch * v + poly is split by COEFFICIENT range (it is element-wise),
              the slices are all-gathered, and point set s is divided and committed on rank s          NCCL, R * 2^k * 32 B

The run prints a digest of h(X) coefficients and of the folded multiopen polynomials; it is the same for every N
(`profiles/`), which is the parity check of the sharded schedule against the single-GPU one.

Commitments / evaluations (96 / 32 bytes each) would be gathered on the transcript's rank; they are left where they
are produced (the bytes are negligible next to the exchanges above).  Every phase is bracketed by a barrier +
synchronize and timed by wall clock on rank 0 after the barrier, i.e. the maximum over ranks.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import halo2_gpu_specific_b200 as h2  # noqa: E402
from halo2_gpu_specific_b200 import _fr, _lib, parallel  # noqa: E402
from halo2_gpu_specific_b200 import evaluation as E  # noqa: E402
from halo2_gpu_specific_b200._lib import NttDesc  # noqa: E402
from halo2_gpu_specific_b200.arithmetic import Srs  # noqa: E402
from halo2_gpu_specific_b200.grand_product import ExprCompiler  # noqa: E402
import quotient_bench as qb  # noqa: E402

R = _fr.R_MOD
vp = ctypes.c_void_p


class TBuf:
    """Fr elements in a torch tensor (so that NCCL can move them); the library sees the raw pointer"""

    def __init__(self, elems: int):
        self.t = torch.empty((max(1, elems), 4), dtype=torch.int64, device="cuda")
        self.ptr, self.elems = self.t.data_ptr(), elems

    def rows(self, lo: int, cnt: int) -> torch.Tensor:
        return self.t[lo:lo + cnt]

    def upload(self, a: np.ndarray, offset_elems: int = 0):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        _lib.check(_lib.lib().b2_memcpy_h2d(vp(self.ptr + offset_elems * 32), vp(a.ctypes.data), a.nbytes))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=22)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--out", default="")
    ap.add_argument("--j", type=int, default=5, help="EvaluationDomain degree parameter (5: four cosets; 3: two cosets, "
                                                     "which lets 4 ranks exercise the shared-coset path)")
    ap.add_argument("--no-split-transforms", action="store_true",
                    help="ranks that share a coset each transform all of its polynomials (no point-to-point exchange)")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(v, d)) for v, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    _lib.require_gpu()
    _lib.set_device(local)
    numa_cpus = _lib.bind_host_to_gpu(local) if world > 1 else None      # pinned columns on the GPU's NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    k = a.k
    n = 1 << k
    dom = h2.EvaluationDomain(a.j, k)
    nc = 1 << (dom.extended_k - k)
    sh = dict(A=64, I=1, F=32, L=8, S=12, H=4, P=8, D=dom.quotient_poly_degree, R=3, perm_cols=24)
    lookups = (2, 2, 2, 2, 1, 1, 1, 1)
    n_z = sh["P"] + sh["S"] + sh["H"]
    for cnt in (sh["A"], sh["L"], n_z, n, nc * n):
        if cnt % world:
            raise SystemExit(f"world size {world} must divide {cnt}")

    def block(count):
        return parallel.column_range(count, world, rank)

    def sync():
        L.b2_synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    t0 = time.time()
    g = Srs.synthetic(n, 0, 0xB2000003)
    gl = Srs.synthetic(n, n, 0xB2000003)
    params = h2.Params(k, g, gl)
    ev, lk, shf, n_sets = qb.synthetic_evaluator(A=sh["A"], F=sh["F"], I=sh["I"], lookups=lookups, shuffles=sh["H"],
                                                 perm_cols=sh["perm_cols"])
    prog = ev.program(n_sets, lk, shf)
    assert n_sets == sh["P"]

    # ---- resident coefficient buffer, same slot layout as the single-GPU tool; the z groups are contiguous
    slots, cur = {}, 0
    for name, cnt in (("fixed", sh["F"]), ("sigma", sh["perm_cols"]), ("advice", sh["A"]), ("instance", sh["I"]),
                      ("perm_z", sh["P"]), ("lookup_z", sh["S"]), ("shuffle_z", sh["H"]), ("lookup_m", sh["L"])):
        slots[name] = (cur, cnt)
        cur += cnt
    slots["z_all"] = (slots["perm_z"][0], n_z)
    n_polys, n_key = cur, sh["F"] + sh["perm_cols"]
    coef = TBuf(n_polys * n)
    cos = TBuf((n_polys + 3) * n)                 # coset evaluations; also the staging area of the exchanges
    key_cos = E.DeviceBuffer((n_key + 3) * nc * n)
    sigma_lagrange = E.DeviceBuffer(sh["perm_cols"] * n)
    work = E.DeviceBuffer(max(sh["P"], 1) * n)
    hext = E.DeviceBuffer(dom.extended_len())
    hcoef = E.DeviceBuffer(n * dom.quotient_poly_degree)
    hloc = TBuf(nc * n // world)
    hfull = TBuf(nc * n)
    batch = TBuf(sh["R"] * n)                     # folded polynomials of the R point sets
    d96 = E.DeviceBuffer(8)

    def slot_ptr(name, i=0):
        return coef.ptr + (slots[name][0] + i) * n * 32

    rng = np.random.default_rng(3)                # same data on every rank
    pool = 6
    big = _lib.pinned_empty((pool, n, 4))
    big[:] = rng.integers(0, 2**64, size=(pool, n, 4), dtype=np.uint64)
    big[:, :, 3] &= np.uint64((1 << 60) - 1)
    small_tbl = np.stack([_fr.to_mont(v) for v in range(1 << 16)])
    spool = max(4, min(32, sh["A"] // world))       # a multiple of 4: entry i repeats entry i % 4
    small = _lib.pinned_empty((spool, n, 4))
    for i in range(spool):
        small[i] = small_tbl[rng.integers(0, 1 << 16, size=n)] if i < 4 else small[i % 4]
    small[:, ::3] = 0
    lag = [big[i % pool] for i in range(3)]

    # keygen-time residency (untimed)
    for i in range(sh["F"]):
        coef.upload(big[i % pool], slots["fixed"][0] * n + i * n)
    for i in range(sh["perm_cols"]):
        coef.upload(big[(i + 1) % pool], slots["sigma"][0] * n + i * n)
        sigma_lagrange.upload(big[(i + 2) % pool], i * n)
    for c in range(nc):
        g_c = dom._zeta * pow(dom._ext_omega, c, R) % R
        E.coeff_to_coset_dev(dom, coef.ptr, n_key, g_c, key_cos.ptr + c * (n_key + 3) * n * 32)
        for i in range(3):
            key_cos.upload(lag[i], (c * (n_key + 3) + n_key + i) * n)
    setup_s = time.time() - t0

    def commit_resident(src, name, lo, hi, bits, ifft):
        out = np.zeros((max(1, hi - lo), 12), dtype=np.uint64)
        done, plen = 0, src.shape[0]
        while done < hi - lo:
            at = (lo + done) % plen              # column j always carries pool entry j % plen: same data for every N
            run = min(plen - at, hi - lo - done)
            cols = src[at: at + run]
            _lib.check(L.b2_commit_batch_resident(params.g_lagrange.handle, vp(cols.ctypes.data), 0,
                                                  vp(slot_ptr(name, lo + done)), run, n, bits, 1 if ifft else 0,
                                                  vp(dom.omega_inv.ctypes.data), vp(dom.ifft_divisor.ctypes.data), k,
                                                  vp(out[done:].ctypes.data)))
            done += run
        return out

    def commit_on_device(name, lo, hi, bits, ifft):
        out = np.zeros((max(1, hi - lo), 12), dtype=np.uint64)
        if hi > lo:
            _lib.check(L.b2_commit_batch_resident(params.g_lagrange.handle, None, 1, vp(slot_ptr(name, lo)), hi - lo, n, bits,
                                                  1 if ifft else 0, vp(dom.omega_inv.ctypes.data),
                                                  vp(dom.ifft_divisor.ctypes.data), k, vp(out.ctypes.data)))
        return out

    def intt_on_device(name, lo, hi):
        if hi <= lo:
            return
        d = NttDesc()
        d.log_n, d.location = k, 1
        d.omega, d.divisor = dom.omega_inv.ctypes.data, dom.ifft_divisor.ctypes.data
        d.n_in = d.n_out = d.in_stride = d.out_stride = n
        d.columns = hi - lo
        d.in_ = d.out = slot_ptr(name, lo)
        _lib.check(L.b2_ntt_exec(ctypes.byref(d)))

    def gather_group(name):
        """all-gather of a slot group whose columns were produced in per-rank blocks (out of place, then copied)"""
        if world == 1:
            return
        first, cnt = slots[name]
        lo, hi = block(cnt)
        mine = coef.rows((first + lo) * n, (hi - lo) * n)
        stage = cos.rows(0, cnt * n)
        dist.all_gather_into_tensor(stage.view(-1), mine.reshape(-1))
        coef.rows(first * n, cnt * n).copy_(stage)

    def bcast_group(name, src=0):
        if world > 1:
            first, cnt = slots[name]
            dist.broadcast(coef.rows(first * n, cnt * n), src)

    # z construction (permutation shape), as in the single-GPU tool
    ch = [5, 7, 0] + [11 + j for j in range(3)]
    c1 = ExprCompiler()
    acc = None
    for j in range(3):
        t = c1.emit(("Mul", ("Challenge", 0), ("Aux", j, 0)))
        t = c1.emit(("AddChallenge", c1.emit(("Add", t, ("Advice", j, 0))), "Gamma"))
        acc = t if acc is None else c1.emit(("Mul", acc, t))
    zp1 = E.QuotientProgram(c1.rotations, c1.constants, c1.calcs, acc, 0, 3, 0, 3, len(ch))
    c2 = ExprCompiler()
    acc = ("Aux", 0, 0)
    for j in range(3):
        t = c2.emit(("Mul", ("CosetX",), ("Challenge", 3 + j)))
        t = c2.emit(("AddChallenge", c2.emit(("Add", t, ("Advice", j, 0))), "Gamma"))
        acc = c2.emit(("Mul", acc, t))
    zp2 = E.QuotientProgram(c2.rotations, c2.constants, c2.calcs, acc, 0, 3, 0, 1, len(ch))

    def build_z(lo, hi):
        adv0 = slots["advice"][0]
        for s0 in range(lo, hi, sh["P"]):
            grp = range(s0, min(hi, s0 + sh["P"]))
            for s in grp:
                cols = [coef.ptr + (adv0 + (3 * s + j) % sh["A"]) * n * 32 for j in range(3)]
                sig = [sigma_lagrange.ptr + ((3 * s + j) % sh["perm_cols"]) * n * 32 for j in range(3)]
                zp1.eval(k, 1, [], cols, [], sig, ch, work.ptr + (s - s0) * n * 32)
            _lib.check(L.b2_batch_invert_dev(vp(work.ptr), len(grp) * n, None))
            for s in grp:
                cols = [coef.ptr + (adv0 + (3 * s + j) % sh["A"]) * n * 32 for j in range(3)]
                w = work.ptr + (s - s0) * n * 32
                zp2.eval(k, 1, [], cols, [], [w], ch, w, x0=1, x_step=dom._omega)
                dst = slot_ptr("z_all", s)
                if s % 3 == 0:     # chains of three columns: block boundaries of 1, 2, 4 and 8 ranks all fall on them, so
                                   # the polynomials (and the digest below) do not depend on the number of ranks
                    _lib.check(L.b2_prefix_scan_dev(0, vp(w), n, None, None, vp(dst), n, None))
                else:
                    _lib.check(L.b2_prefix_scan_dev(0, vp(w), n, None, vp(slot_ptr("z_all", s - 1) + (n - 6) * 32), vp(dst), n, None))

    # evaluate_h pointer tables (key cosets resident per coset; witness cosets in `cos`)
    def col_ptrs(c):
        key_base = key_cos.ptr + c * (n_key + 3) * n * 32

        def p(name, i):
            if name in ("fixed", "sigma"):
                return key_base + (slots[name][0] + i) * n * 32
            return cos.ptr + (slots[name][0] + i) * n * 32
        fixed = [p("fixed", i) for i in range(sh["F"])]
        advice = [p("advice", i) for i in range(sh["A"])]
        inst = [p("instance", i) for i in range(sh["I"])]
        aux = [key_base + (n_key + i) * n * 32 for i in range(3)]
        aux += [p("sigma", i) for i in range(sh["perm_cols"])] + [p("perm_z", i) for i in range(sh["P"])]
        zi = 0
        for li, sets in enumerate(lookups):
            aux += [p("lookup_z", zi + i) for i in range(sets)] + [p("lookup_m", li)]
            zi += sets
        aux += [p("shuffle_z", i) for i in range(sh["H"])]
        return fixed, advice, inst, aux

    tables = [col_ptrs(c) for c in range(nc)]
    challenges = [(i + 2) * 0x123456789ABCDEF % R for i in range(prog.n_challenges)]
    tasks = parallel.quotient_tasks(nc, n, world, rank)

    # ranks sharing a coset (world > number of cosets): each transforms a share of the witness polynomials and the
    # group swaps the shares over NVLink (point-to-point), instead of every rank transforming all of them
    n_wit = n_polys - n_key
    grp_first, shares = parallel.coset_transform_shares(nc, world, rank, n_wit)
    share = len(shares)

    def coset_transform(c):
        g_c = dom._zeta * pow(dom._ext_omega, c, R) % R
        if share == 1 or a.no_split_transforms:
            E.coeff_to_coset_dev(dom, coef.ptr + n_key * n * 32, n_wit, g_c, cos.ptr + n_key * n * 32)
            return
        bounds = [(plo, phi) for _, plo, phi in shares]
        lo, hi = bounds[rank - grp_first]
        E.coeff_to_coset_dev(dom, coef.ptr + (n_key + lo) * n * 32, hi - lo, g_c, cos.ptr + (n_key + lo) * n * 32)
        L.b2_synchronize()
        ops = []
        for j, (plo, phi) in enumerate(bounds):
            peer = grp_first + j
            if peer == rank:
                continue
            ops.append(dist.P2POp(dist.isend, cos.rows((n_key + lo) * n, (hi - lo) * n), peer))
            ops.append(dist.P2POp(dist.irecv, cos.rows((n_key + plo) * n, (phi - plo) * n), peer))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        torch.cuda.synchronize()

    def evaluate_h():
        loaded, written = None, 0
        for c, begin, count in tasks:
            if c != loaded:
                coset_transform(c)
                loaded = c
            fx, adv, ins, aux = tables[c]
            prog.eval(k, 1, fx, adv, ins, aux, challenges, hloc.ptr, x0=pow(dom._ext_omega, c, R), x_step=dom._omega,
                      scale=dom.t_evaluations[c:c + 1], out_stride=1, out_offset=written, row_begin=begin, row_count=count)
            written += count
        L.b2_synchronize()
        if world > 1:
            dist.all_gather_into_tensor(hfull.t.view(-1), hloc.t.view(-1))
            torch.cuda.synchronize()
            src = hfull.ptr
        else:
            src = hloc.ptr
        E.interleave_cosets_dev(dom, src, hext.ptr)

    def h_to_coeff():
        z = np.concatenate([dom.g_coset_inv, dom.g_coset])
        d = NttDesc()
        d.log_n, d.location = dom.extended_k, 1
        d.omega, d.divisor = dom.extended_omega_inv.ctypes.data, dom.extended_ifft_divisor.ctypes.data
        d.coset_out = z.ctypes.data
        d.n_in = d.in_stride = dom.extended_len()
        d.n_out = d.out_stride = n * dom.quotient_poly_degree
        d.columns, d.in_, d.out = 1, hext.ptr, hcoef.ptr
        _lib.check(L.b2_ntt_exec(ctypes.byref(d)))

    def msm_dev_to_host(d_scalars, count):
        out = np.zeros((1, 12), dtype=np.uint64)
        _lib.check(L.b2_msm_dev(params.g.handle, 0, vp(d_scalars), count, 254, vp(d96.ptr), None))
        L.b2_synchronize()
        _lib.check(L.b2_memcpy_d2h(vp(out.ctypes.data), vp(d96.ptr), 96))
        _lib.check(L.b2_g1_normalize(vp(out.ctypes.data), 1))
        return out

    def commit_h_pieces():
        return [msm_dev_to_host(hcoef.ptr + i * n * 32, n) for i in range(sh["D"]) if i % world == rank]

    x_int = 0x1F2E3D4C5B6A79881F2E3D4C5B6A7988 % R
    x_pt, x_next = _fr.to_mont(x_int), _fr.to_mont(x_int * dom._omega % R)
    x_last = _fr.to_mont(x_int * pow(dom._omega, n - 6, R) % R)
    v_ch = _fr.to_mont(0x0123456789ABCDEF0FEDCBA987654321 % R)
    n_adv_next = sh["A"] // 4
    groups_x = ["advice", "instance", "fixed", "sigma", "perm_z", "lookup_z", "lookup_m", "shuffle_z"]

    def eval_block(base_ptr, count, pt):
        lo, hi = block(count) if count >= world else ((0, count) if rank == 0 else (0, 0))
        if hi > lo:
            out = np.empty((hi - lo, 4), dtype=np.uint64)
            _lib.check(L.b2_eval_polynomial_dev(vp(base_ptr + lo * n * 32), hi - lo, n, n, vp(pt.ctypes.data), vp(out.ctypes.data)))
        return hi - lo

    def evaluate_all():
        cnt = 0
        for pt, names in ((x_pt, groups_x), (x_next, ("perm_z", "lookup_z", "shuffle_z"))):
            for gname in names:
                cnt += eval_block(slot_ptr(gname), slots[gname][1], pt)
        cnt += eval_block(slot_ptr("advice"), n_adv_next, x_next)
        cnt += eval_block(slot_ptr("perm_z"), sh["P"] - 1, x_last)
        cnt += eval_block(hcoef.ptr, sh["D"], x_pt)
        return cnt

    def open_sets():
        at_x = [slot_ptr(gn, i) for gn in groups_x for i in range(slots[gn][1])] + [hcoef.ptr + i * n * 32 for i in range(sh["D"])]
        at_next = [slot_ptr("advice", i) for i in range(n_adv_next)] + \
                  [slot_ptr(gn, i) for gn in ("perm_z", "lookup_z", "shuffle_z") for i in range(slots[gn][1])]
        at_last = [slot_ptr("perm_z", i) for i in range(sh["P"] - 1)]
        return [(x_pt, at_x), (x_next, at_next), (x_last, at_last)]

    def multiopen():
        sets = open_sets()
        lo, hi = block(n)                                   # this rank's coefficient range of every fold
        for s, (pt, ptrs) in enumerate(sets):
            arr = (vp * len(ptrs))(*[p + lo * 32 for p in ptrs])
            dst = (hloc.ptr if world > 1 else batch.ptr + s * n * 32)
            _lib.check(L.b2_poly_combine_dev(arr, len(ptrs), hi - lo, vp(v_ch.ctypes.data), vp(dst), None))
            if world > 1:
                L.b2_synchronize()
                dist.all_gather_into_tensor(batch.rows(s * n, n).view(-1), hloc.rows(0, hi - lo).reshape(-1))
                torch.cuda.synchronize()         # hloc is reused by the next fold
        outs = []
        for s, (pt, ptrs) in enumerate(sets):
            if s % world != rank:
                continue
            wit = cos.ptr
            _lib.check(L.b2_kate_division_dev(vp(batch.ptr + s * n * 32), n, vp(pt.ctypes.data), vp(wit), None))
            outs.append(msm_dev_to_host(wit, n - 1))
        return sum(len(p) for _, p in sets)

    def phase(fn):
        sync()
        t = time.perf_counter()
        out = fn()
        sync()
        return time.perf_counter() - t, out

    a_lo, a_hi = block(sh["A"])
    l_lo, l_hi = block(sh["L"])
    z_lo, z_hi = block(n_z)
    results = []
    n_evals = n_opened = 0
    for rep in range(a.reps + 1):
        ph = {}
        ph["1_instance_commit_ifft"], _ = phase(lambda: commit_resident(big, "instance", 0, sh["I"], 254, True) if rank == 0 else None)
        ph["2_advice_commit"], _ = phase(lambda: commit_resident(small, "advice", a_lo, a_hi, 16, False))
        ph["3_lookup_m_commit"], _ = phase(lambda: commit_resident(small, "lookup_m", l_lo, l_hi, 16, False))
        ph["4_allgather_advice_lagrange"], _ = phase(lambda: gather_group("advice"))
        ph["5_z_construct_on_device"], _ = phase(lambda: build_z(z_lo, z_hi))
        ph["6_z_commit_and_ifft"], _ = phase(lambda: commit_on_device("z_all", z_lo, z_hi, 254, True))
        ph["7_vanishing_commit"], _ = phase(lambda: params.commit(big[0]) if rank == 0 else None)
        ph["8_advice_m_ifft_on_device"], _ = phase(lambda: (intt_on_device("advice", a_lo, a_hi),
                                                             intt_on_device("lookup_m", l_lo, l_hi)))
        ph["8b_allgather_coefficient_forms"], _ = phase(lambda: (gather_group("advice"), gather_group("z_all"),
                                                                  gather_group("lookup_m"), bcast_group("instance")))
        ph["9_evaluate_h"], _ = phase(evaluate_h)
        ph["10_h_to_coeff"], _ = phase(h_to_coeff)
        ph["10_h_commits"], _ = phase(commit_h_pieces)
        ph["11_evaluations_on_device"], n_evals = phase(evaluate_all)
        ph["12_multiopen_on_device"], n_opened = phase(multiopen)
        ph["total"] = sum(ph.values())
        results.append(ph)
    best = min(results[1:], key=lambda p: p["total"])
    import hashlib
    sync()
    hbytes = np.empty((4096, 4), dtype=np.uint64)
    _lib.check(L.b2_memcpy_d2h(vp(hbytes.ctypes.data), vp(hcoef.ptr + (n - 2048) * 32), hbytes.nbytes))
    bbytes = np.empty((sh["R"], 64, 4), dtype=np.uint64)
    for s_ in range(sh["R"]):
        _lib.check(L.b2_memcpy_d2h(vp(bbytes[s_].ctypes.data), vp(batch.ptr + (s_ * n + n // 2) * 32), bbytes[s_].nbytes))
    digest = hashlib.sha256(hbytes.tobytes() + bbytes.tobytes()).hexdigest()[:16]
    if rank == 0:
        doc = {"workload": "create_proof schedule replay, zkWasm-scale shape, device-resident, one process per GPU",
               "k": k, "shape": sh, "n_gpus": world, "phases_s": best, "wall_s": best["total"], "setup_s": setup_s,
               "exchanges_GiB": {"advice_lagrange": sh["A"] * n * 32 / 2**30 if world > 1 else 0,
                                 "coefficient_forms": (sh["A"] + n_z + sh["L"] + sh["I"]) * n * 32 / 2**30 if world > 1 else 0,
                                 "h": nc * n * 32 / 2**30 if world > 1 else 0,
                                 "multiopen_folds": sh["R"] * n * 32 / 2**30 if world > 1 else 0},
               "h2d_GiB_per_rank": ((a_hi - a_lo) + (l_hi - l_lo) + sh["I"] + 1) * n * 32 / 2**30,
               "digest_h_and_folds": digest, "host_cpus_rank0": (len(numa_cpus) if numa_cpus else None),
               "evaluations_rank0": n_evals, "polynomials_folded": n_opened, "tasks_rank0": tasks,
               "timing": "per phase: barrier + synchronize on both sides, wall clock on rank 0 (= max over ranks)",
               "program": prog.info(),
               "excluded": "CPU-side protocol logic (witness synthesis, the logup multiplicities, transcript hashing); "
                           "the z[last] hand-over between permutation sets on different ranks"}
        print(json.dumps(doc), flush=True)
        if a.out:
            json.dump(doc, open(a.out, "w"), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
