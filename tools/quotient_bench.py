"""Times the fused quotient-evaluation kernel (b2_quotient_eval) on a synthetic zkWasm-scale shape
(A = 64 advice, F = 32 fixed, I = 1, 8 lookups with 12 input sets, 4 shuffles,
8 permutation sets over 24 columns, degree 5), columns resident in HBM.

    python tools/quotient_bench.py --k 20 [--gates 96] [--reps 5] [--json out.json]

Reports rows/s, field multiplications/s against the measured integer peak (b2_imad_probe) and the
column bytes read per second against the HBM peak."""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_gpu_specific_b200 as h2  # noqa: E402
from halo2_gpu_specific_b200 import _fr, _lib  # noqa: E402
from halo2_gpu_specific_b200 import evaluation as E  # noqa: E402


def synthetic_evaluator(A=64, F=32, I=1, gates=96, lookups=(2, 2, 2, 2, 1, 1, 1, 1), shuffles=4, perm_cols=24,
                        degree=5, seed=1, long_lived=0):
    """Calculation lists in the reference's form (what Evaluator::new would hand over), no CSE needed.
    long_lived: that many of the first gates' products are read again by extra gates at the END of the gate list -- a
    circuit whose gates share sub-expressions far apart (the y-fold keeps each such value alive across the whole
    program: what the engine's global slot class is for)."""
    rng = random.Random(seed)
    rotations = [0, 1, -1, 2]
    constants = [0, 1, 2, 7]
    calcs = []

    def emit(c):
        calcs.append(c)
        return ("Intermediate", len(calcs) - 1)

    def col(kind, n):
        return emit(("Store", (kind, rng.randrange(n), rng.randrange(len(rotations)))))

    value_parts = []
    for g in range(gates):   # q * (a*b - c) and q * (a*b*c + 7*d - e): degree 3 / 4
        q = col("Fixed", F)
        a, b, c = col("Advice", A), col("Advice", A), col("Advice", A)
        if g % 2 == 0:
            t = emit(("Sub", emit(("Mul", a, b)), c))
        else:
            d, e = col("Advice", A), col("Advice", A)
            t = emit(("Mul", emit(("Mul", a, b)), c))
            t = emit(("Sub", emit(("Add", t, emit(("Mul", d, ("Constant", 3))))), e))
        value_parts.append(emit(("Mul", q, t)))
    for g in range(min(long_lived, gates)):   # q_g * t_g * (advice + fixed): reads gate g's value again, at the end
        value_parts.append(emit(("Mul", value_parts[g], emit(("Add", col("Advice", A), col("Fixed", F))))))
    lookup_results = []
    for sets in lookups:
        def compressed(width):
            lc = col("Advice", A)
            for _ in range(width - 1):
                lc = emit(("LcTheta", lc, col("Advice", A)))
            return lc
        table = ("AddChallenge", emit(("LcTheta", col("Fixed", F), col("Fixed", F))), "Beta")
        prods, sums = [], []
        for _ in range(sets):
            ins = [emit(("AddChallenge", compressed(2), "Beta")) for _ in range(2)]
            prods.append(("Store", emit(("Mul", ins[0], ins[1]))))
            sums.append(("Store", emit(("Add", ins[1], ins[0]))))
        lookup_results.append((table, prods, sums))
    shuffle_results = []
    for _ in range(shuffles):
        a, b = col("Advice", A), col("Advice", A)
        shuffle_results.append((("AddChallenge", a, "Beta"), ("AddChallenge", b, "Beta")))
    perm = [("Advice", i) for i in range(perm_cols)]
    ev = h2.Evaluator(rotations, constants, calcs, value_parts, lookup_results, shuffle_results, F, A, I, perm, degree, 5)
    return ev, list(lookups), shuffles, (perm_cols + degree - 3) // (degree - 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--gates", type=int, default=96)
    ap.add_argument("--long-lived", type=int, default=0,
                    help="extra gates at the end of the gate list that read the first gates' values again (A/B of the "
                         "global slot class: run with B2_Q_HYBRID=0 and without)")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--json", default=None)
    ap.add_argument("--pipeline", action="store_true",
                    help="time the whole evaluate_h path in coset mode from device-resident coefficient forms: "
                         "per coset one batched size-2^k transform of every polynomial + the fused kernel, then "
                         "extended_to_coeff of h back to the host")
    args = ap.parse_args()
    _lib.require_gpu()
    _lib.set_device(0)
    ev, lookups, shuffles, n_sets = synthetic_evaluator(gates=args.gates, long_lived=args.long_lived)
    prog = ev.program(n_sets, lookups, shuffles)
    info = prog.info()
    ext_k = args.k + 2
    rows = 1 << ext_k
    ncols = prog.n_fixed + prog.n_advice + prog.n_instance + prog.n_aux
    if args.pipeline:
        return pipeline(args, prog, info, ncols)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    # synthetic resident columns: 8 distinct random columns, the rest device copies of them
    rng = np.random.default_rng(5)
    base = rng.integers(0, 1 << 62, size=(8, rows, 4), dtype=np.uint64)   # < 2^254: valid field elements (< r)
    base[:, :, 3] &= np.uint64((1 << 60) - 1)
    buf = E.DeviceBuffer(ncols * rows)
    t0 = time.time()
    for c in range(ncols):
        buf.upload(base[c % 8], c * rows)
    upload_s = time.time() - t0
    out = E.DeviceBuffer(rows)
    ptrs = [buf.ptr + c * rows * 32 for c in range(ncols)]
    nf, na, ni = prog.n_fixed, prog.n_advice, prog.n_instance
    challenges = [(i + 2) * 0x123456789ABCDEF % _fr.R_MOD for i in range(prog.n_challenges)]
    dom = h2.EvaluationDomain(5, args.k)
    times = []
    for _ in range(args.reps + 2):
        prog.eval(ext_k, 4, ptrs[:nf], ptrs[nf:nf + na], ptrs[nf + na:nf + na + ni], ptrs[nf + na + ni:], challenges,
                  out.ptr, x0=1, x_step=dom._ext_omega, scale=dom.t_evaluations)
        times.append(_lib.last_timing()[0])
    times = times[2:]
    ms = float(np.median(times))
    ab = None
    if args.long_lived and os.environ.get("B2_Q_HYBRID") is None:
        # the same circuit lowered with every slot in shared memory, on the same resident columns
        os.environ["B2_Q_HYBRID"] = "0"
        plain = synthetic_evaluator(gates=args.gates, long_lived=args.long_lived)[0].program(n_sets, lookups, shuffles)
        del os.environ["B2_Q_HYBRID"]
        out2 = E.DeviceBuffer(rows)
        t2 = []
        for _ in range(args.reps + 2):
            plain.eval(ext_k, 4, ptrs[:nf], ptrs[nf:nf + na], ptrs[nf + na:nf + na + ni], ptrs[nf + na + ni:], challenges,
                       out2.ptr, x0=1, x_step=dom._ext_omega, scale=dom.t_evaluations)
            t2.append(_lib.last_timing()[0])
        same = bool(np.array_equal(out.download(), out2.download()))
        ab = {"all_slots_shared": {"program": plain.info(), "kernel_ms": float(np.median(t2[2:]))},
              "results_equal": same, "speedup_of_the_global_slot_class": float(np.median(t2[2:])) / ms}
        out2.free()
    wide, modmul = ctypes.c_double(), ctypes.c_double()
    _lib.check(_lib.lib().b2_imad_probe(ctypes.byref(wide), ctypes.byref(modmul)))
    instr, _, _ = prog.dump()
    col_reads = sum(1 for ins in instr for w in (ins[2:] if ins[0] not in (3, 4) else ins[2:3]) if (w >> 28) == 2)
    muls = info["n_mul"] + 2   # + coset point + vanishing scale
    res = {
        "k": args.k, "extended_k": ext_k, "rows": rows, "columns_resident": ncols,
        "resident_GiB": ncols * rows * 32 / 2**30, "program": info, "column_reads_per_row": col_reads,
        "kernel_ms": ms, "rows_per_s": rows / (ms * 1e-3),
        "modmul_per_s": muls * rows / (ms * 1e-3), "modmul_peak_per_s": modmul.value,
        "int_frac": muls * rows / (ms * 1e-3) / modmul.value,
        "column_GBps": col_reads * 32 * rows / (ms * 1e-3) / 1e9,
        "upload_s": upload_s,
    }
    if ab is not None:
        res["ab_global_slot_class"] = ab
    print(json.dumps(res))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(res, f, indent=1)
    buf.free()
    out.free()


def pipeline(args, prog, info, ncols):
    k, n = args.k, 1 << args.k
    dom = h2.EvaluationDomain(5, k)
    nc = 1 << (dom.extended_k - k)
    R = _fr.R_MOD
    rng = np.random.default_rng(7)
    base = rng.integers(0, 1 << 62, size=(8, n, 4), dtype=np.uint64)
    base[:, :, 3] &= np.uint64((1 << 60) - 1)
    coef = E.DeviceBuffer(ncols * n)
    for c in range(ncols):
        coef.upload(base[c % 8], c * n)
    cos = E.DeviceBuffer(ncols * n)
    out = E.DeviceBuffer(dom.extended_len())
    ptrs = [cos.ptr + c * n * 32 for c in range(ncols)]
    nf, na, ni = prog.n_fixed, prog.n_advice, prog.n_instance
    challenges = [(i + 2) * 0x123456789ABCDEF % R for i in range(prog.n_challenges)]
    runs = []
    for rep in range(args.reps + 1):
        _lib.lib().b2_synchronize()
        t0 = time.perf_counter()
        ntt_ms = eval_ms = 0.0
        for c in range(nc):
            g_c = dom._zeta * pow(dom._ext_omega, c, R) % R
            E.coeff_to_coset_dev(dom, coef.ptr, ncols, g_c, cos.ptr)
            ntt_ms += _lib.last_timing()[0]
            prog.eval(k, 1, ptrs[:nf], ptrs[nf:nf + na], ptrs[nf + na:nf + na + ni], ptrs[nf + na + ni:], challenges,
                      out.ptr, x0=pow(dom._ext_omega, c, R), x_step=dom._omega, scale=dom.t_evaluations[c:c + 1],
                      out_stride=nc, out_offset=c)
            eval_ms += _lib.last_timing()[0]
        t1 = time.perf_counter()
        h = E.extended_to_coeff_dev(dom, out)
        t2 = time.perf_counter()
        runs.append({"wall_s": t2 - t0, "cosets_s": t1 - t0, "ntt_kernel_ms": ntt_ms, "eval_kernel_ms": eval_ms,
                     "extended_to_coeff_s": t2 - t1})
    best = min(runs[1:], key=lambda r: r["wall_s"])
    res = {"workload": "evaluate_h + h(X) coefficients, coset mode, coefficient forms resident in HBM",
           "k": k, "extended_k": dom.extended_k, "polynomials": ncols, "program": info,
           "resident_GiB": (2 * ncols * n + dom.extended_len()) * 32 / 2**30,
           "extended_layout_GiB": ncols * dom.extended_len() * 32 / 2**30, **best,
           "ntt_Melem_per_s": ncols * n * nc / (best["ntt_kernel_ms"] * 1e-3) / 1e6,
           "eval_rows_per_s": n * nc / (best["eval_kernel_ms"] * 1e-3)}
    print(json.dumps(res))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(res, f, indent=1)
    coef.free(); cos.free(); out.free()
    assert h.shape[0] == n * dom.quotient_poly_degree


if __name__ == "__main__":
    main()
