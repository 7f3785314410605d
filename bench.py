#!/usr/bin/env python
"""bench.py -- BN254 MSM Mpts/s (+ Fr NTT Melem/s) at 2^22 on 1..8 B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--logn 22]

A "step" is one pass of the commitment hot path over one synthetic column per GPU: one MSM of
2^22 uniformly random Fr scalars against that rank's resident point-range shard of the SRS
(north_star: "MSM point ranges per GPU with one partial G1 point combined per rank"), the NCCL
all-gather of one 96-byte partial per rank, and the sum of the partials -- gathered and summed
for a block of steps at a time (--gather-every; inside the timed region), as the sharded prover
gathers the points of a block of columns.  Per-GPU work is fixed ("weak" scaling): at N ranks
every step is one MSM of N * 2^22 points.

  value   whole-job points/s with the scalars already resident in HBM; device time (CUDA events
          on the launching stream, max over ranks).  The steps are independent MSMs (one column
          each), issued on --value-streams caller streams in turn (default 2) so that the digit
          sort of one runs under the bucket reduction of the previous; the first stream forks the
          others after the start event and joins them before the end event
  e2e     same metric through the public host API (b2_msm_async on a pinned host column against
          the resident Srs, one caller thread, up to --e2e-depth columns in flight): H2D of the
          scalars and D2H of the point are inside the timed region; the blocking call is reported
          next to it
  parity_check  the N-rank result against the closed form [sum s_i h_i mod r] G, after the timed region
  strong_scaling  one MSM of fixed total size 2^18 .. 2^26 split over the ranks
  sharded_create_proof  N > 1 only: the real prover divided over the ranks (zkWasm-shaped circuit, k = 18), every
          rank's proof bytes compared with each other and with a single-GPU proof
  roofline  dominant kernel (msm_accumulate): integer-pipe roofline, measured with CUDA events
          inside this run; peak = this run's own carry-chained IMAD.WIDE probe
  ntt     N == 1 only: 64 columns of a k=22 forward NTT, device resident: Melem/s, HBM GB/s vs
          the measured copy peak, and the same integer roofline
  quotient  N == 1 only: evaluate_h + h(X) of a synthetic zkWasm-scale constraint system at k = logn from
          HBM-resident coefficient forms (coset mode), wall seconds, kernel times and the fused kernel's
          integer roofline; its own cpu_baseline (C restatement of the reference's row loop)
  create_proof  N == 1 only: BASELINE config 4, the benches/plonk.rs circuit at k = 18, whole create_proof through
          the prover mirror (plonk.create_proof), wall seconds and per-phase split
  create_proof_k22  N == 1 only: "k = 22 create_proof wall-time" as a real proof of a zkWasm-shaped circuit (64 advice, 32
          fixed, 8 lookups, 4 shuffles, 24 permutation columns) on one GPU
  cpu_baseline / --impl reference: the C restatement of the reference's rayon path
          (oracle/cpu_ref.c = arithmetic.rs:20-108, 465-492) on the box's host cores

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "BN254 G1 MSM throughput at 2^22 points per GPU"
UNIT = "Mpts/s"


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
# exports under profiles/ (key: kernel, log2 n, window tables); None for configurations that were not captured
NCU_TRAFFIC = {("msm_accumulate", 22, True): 7.513166e9 + 83.072e6}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def random_montgomery_scalars(n: int, seed: int, pinned=None) -> np.ndarray:
    """uniform 252-bit values used directly as Montgomery residues (< r, so every one is a valid
    encoding of a uniformly distributed field element).  No oracle involved."""
    rng = np.random.default_rng(seed)
    out = pinned if pinned is not None else np.empty((n, 4), dtype=np.uint64)
    out[:] = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    out[:, 3] &= np.uint64((1 << 60) - 1)
    return out


def workload_desc(logn: int, world: int, scaling: str) -> str:
    """config.workload, the SAME string on the engine arm and on the reference arm"""
    if scaling == "strong":
        return (f"BN254 G1 MSM of 2^{logn} uniformly random Fr scalars in total, split by point range ceil(n/G) over {world} "
                f"GPU(s) (gpu_multiexp_bound, arithmetic.rs:413-440)")
    return (f"BN254 G1 MSM, 2^{logn} uniformly random Fr scalars per GPU x {world} GPU(s) = one MSM of {world} x 2^{logn} "
            f"points split by point range (gpu_multiexp_bound, arithmetic.rs:413-440)")


# ---- closed-form parity check (no oracle involved): MSM(s, [h_i] G) == [sum s_i h_i mod r] G ----------------------
R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
Q_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47


def _splitmix(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def synthetic_multipliers(n: int, first: int, seed: int) -> np.ndarray:
    """h_i of b2_srs_synthetic (bases[i] = [h_i] G), restated on the host"""
    with np.errstate(over="ignore"):
        idx = np.arange(first, first + n, dtype=np.uint64)
        h = _splitmix(np.uint64(seed) ^ _splitmix(idx))
    h[h == 0] = 1
    return h


def dot_u256_u64(a: np.ndarray, h: np.ndarray) -> int:
    """sum_i a_i * h_i as a Python int (a: (n,4) u64 limbs): 16-bit pieces so every partial dot fits in uint64"""
    total = 0
    h16 = [((h >> np.uint64(16 * b)) & np.uint64(0xFFFF)) for b in range(4)]
    for limb in range(4):
        col = a[:, limb]
        for x in range(4):
            piece = (col >> np.uint64(16 * x)) & np.uint64(0xFFFF)
            if not piece.any():
                continue
            for b in range(4):
                total += int(np.dot(piece, h16[b])) << (64 * limb + 16 * x + 16 * b)
    return total


def g1_mul_gen_host(t: int):
    """[t] (1, 2) on y^2 = x^3 + 3 over Fq: affine double-and-add with Python ints; None = identity"""
    def add(P, Q):
        if P is None:
            return Q
        if Q is None:
            return P
        (x1, y1), (x2, y2) = P, Q
        if x1 == x2:
            if (y1 + y2) % Q_MOD == 0:
                return None
            lam = 3 * x1 * x1 * pow(2 * y1, -1, Q_MOD) % Q_MOD
        else:
            lam = (y2 - y1) * pow(x2 - x1, -1, Q_MOD) % Q_MOD
        x3 = (lam * lam - x1 - x2) % Q_MOD
        return x3, (lam * (x1 - x3) - y1) % Q_MOD
    acc, base = None, (1, 2)
    t %= R_MOD
    while t:
        if t & 1:
            acc = add(acc, base)
        base = add(base, base)
        t >>= 1
    return acc


def decode_normalized_point(jac12: np.ndarray):
    """(12,) u64 Jacobian with Z = R (Montgomery one) or Z = 0 -> canonical affine ints / None"""
    limbs = [int(v) for v in np.asarray(jac12, dtype=np.uint64).reshape(12)]
    val = lambda w: sum(x << (64 * i) for i, x in enumerate(w))  # noqa: E731
    if val(limbs[8:]) == 0:
        return None
    rinv = pow(1 << 256, -1, Q_MOD)
    return val(limbs[0:4]) * rinv % Q_MOD, val(limbs[4:8]) * rinv % Q_MOD


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's CPU path (C restatement; the Rust crate cannot be built
    offline) on all host cores, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cref
    cores = os.cpu_count() or 1
    world = max(1, args.gpus)
    n_full = (1 << args.logn) * (world if args.scaling == "weak" else 1)   # the engine arm's whole job at this N
    # calibrate on 2^16, then pick the largest power-of-two sample that keeps the run bounded
    cal = min(1 << 16, n_full)
    ks = np.zeros((n_full, 4), dtype=np.uint64)
    ks[:, 0] = np.arange(1, n_full + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    t0 = time.time()
    bases_cal = cref.g1_mul_gen(ks[:cal], cores)
    gen_rate = cal / (time.time() - t0)
    sc = cref.random_fr_mont(cal, 0xB2000003)
    t0 = time.time()
    cref.best_multiexp(sc, bases_cal, cores)
    rate = cal / (time.time() - t0)
    budget_s = 150.0
    total_steps = args.steps + args.warmup
    sample = n_full
    while sample > cal and (sample / gen_rate + total_steps * sample / rate) > budget_s:
        sample >>= 1
    # bases [k_i] G with 64-bit k_i (cheap to generate; MSM cost does not depend on the point values)
    bases = cref.g1_mul_gen(ks[:sample], cores)
    scalars = cref.random_fr_mont(sample, 0xB2000003)
    for _ in range(args.warmup):
        cref.best_multiexp(scalars, bases, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cref.best_multiexp(scalars, bases, cores)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt / 1e6
    sample_desc = (f"best_multiexp over {sample} of the job's {n_full} points per step ({cores} threads, chunk = n/T); "
                   f"a rate, so comparable with the engine arm's whole-job rate")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "u32x8 (256-bit Montgomery integers)", "data": "synthetic",
        "config": {"workload": workload_desc(args.logn, world, args.scaling), "sample": sample_desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C restatement of halo2_proofs/src/arithmetic.rs (oracle/cpu_ref.c), not the Rust binary: no rustc/cargo "
                "in the image and the reference's arithmetic crate is an un-vendored git dependency",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_engine(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    import halo2_gpu_specific_b200 as h2
    from halo2_gpu_specific_b200 import _lib
    from halo2_gpu_specific_b200.arithmetic import Srs
    from halo2_gpu_specific_b200 import parallel

    _lib.require_gpu()  # no CPU fallback: fail loudly
    torch.cuda.set_device(local)
    _lib.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = _lib.bind_host_to_gpu(local) if world > 1 else None   # pinned buffers on the GPU's NUMA node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    seed = 0xB2000003
    if args.scaling == "strong":       # ONE MSM of 2^logn points, part_len = ceil(n / G) (arithmetic.rs:426)
        n_total = 1 << args.logn
        first, hi = parallel.shard_range(n_total, world, rank)
        n = hi - first
    else:                              # 2^logn points per GPU
        n = 1 << args.logn
        n_total = n * world
        first = rank * n

    # --- inputs: this rank's point range of the SRS (resident) and its slice of the scalar vector
    srs = Srs.synthetic(n, first_index=first, seed=seed)
    if not args.no_precompute:
        srs.precompute()   # window table of this rank's shard (one-off, like the SRS upload itself)
    h_scalars = _lib.pinned_empty((n, 4))
    random_montgomery_scalars(n, seed + rank, pinned=h_scalars)
    d_scalars = torch.from_numpy(h_scalars.view(np.int64)).to(dev)
    # a real (non-default) stream: the library launches on it, torch's events and NCCL order on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = ctypes.c_void_p(stream.cuda_stream)
    assert stream.cuda_stream != 0

    # The steps are independent MSMs (the prover commits column after column), so the device-resident loop issues them on
    # --value-streams caller streams in turn (default 2): each stream's calls keep to a lane of their own, and the digit
    # sort of MSM i + 1 (bound by L2 atomics) runs under the bucket reduction of MSM i (a latency chain that leaves the
    # SMs idle).  --value-streams 1 is the plain single-stream loop.
    # With N > 1 the partials of --gather-every consecutive steps (default: all steps of the loop) form a block that is
    # all-gathered in ONE collective and summed in one launch -- what the sharded prover does with the points of a block
    # of columns (prover_sharded.ShardedCommits._gather): a NCCL kernel per step has to find room on SMs that the next
    # MSM's accumulate kernel fills, and holds the overlap back.  --gather-every 1 is the per-step gather.
    lanes_streams = [stream] + [torch.cuda.Stream(device=dev) for _ in range(max(0, args.value_streams - 1))]
    G = max(1, args.gather_every if args.gather_every > 0 else max(args.steps, args.warmup, 1))
    blocks = [torch.zeros((G, 12), dtype=torch.int64, device=dev) for _ in range(2)]     # double-buffered
    g_all = torch.zeros((world, G, 12), dtype=torch.int64, device=dev)
    g_sums = torch.zeros((G, 12), dtype=torch.int64, device=dev)
    st = {"block": 0, "fill": 0, "step": 0, "last": None}
    flushed = [None, None]                          # event: block b has been gathered (its rows may be overwritten)

    def flush_gather():
        """all-gather + sums of the rows filled since the last flush, on `stream`, after every lane stream's MSMs"""
        rows = st["fill"]
        if rows == 0:
            return
        b = st["block"]
        for s_ in lanes_streams[1:]:
            stream.wait_stream(s_)
        if world > 1:
            dist.all_gather_into_tensor(g_all.view(-1), blocks[b].view(-1))             # current stream == `stream`
            _lib.check(L.b2_g1_sum_groups_dev(ctypes.c_void_p(g_all.data_ptr()), world, G,
                                              ctypes.c_void_p(g_sums.data_ptr()), sp))
            st["last"] = g_sums[rows - 1]
        else:
            st["last"] = blocks[b][rows - 1]
        flushed[b] = torch.cuda.Event()
        flushed[b].record(stream)
        st["block"], st["fill"] = 1 - b, 0

    def step_resident():
        if st["fill"] == G:
            flush_gather()
        b, row = st["block"], st["fill"]
        s_ = lanes_streams[st["step"] % len(lanes_streams)]
        st["step"] += 1
        st["fill"] += 1
        if row == 0 and flushed[b] is not None:
            for t_ in lanes_streams[1:]:
                t_.wait_event(flushed[b])           # the block's previous contents have been gathered
        _lib.check(L.b2_msm_dev(srs.handle, 0, ctypes.c_void_p(d_scalars.data_ptr()), n, 254,
                                ctypes.c_void_p(blocks[b][row].data_ptr()), ctypes.c_void_p(s_.cuda_stream)))

    def result_tensor():
        flush_gather()
        return st["last"]

    def fork_streams():                            # everything after this point on `stream` precedes the other streams' work
        for s_ in lanes_streams[1:]:
            s_.wait_stream(stream)

    def join_streams():                            # `stream` has seen the end of every step, gathers and sums included
        flush_gather()

    def step_e2e():
        return parallel.sharded_msm(h_scalars, srs, 254)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # --- IMAD peak probe (integer roofline denominator), before the timed region
    macs, muls = ctypes.c_double(), ctypes.c_double()
    _lib.check(L.b2_imad_probe(ctypes.byref(macs), ctypes.byref(muls)))
    raw_wide = ctypes.c_double()                     # independent of the field code: bare IMAD.WIDE.U32 chains, no carries
    _lib.check(L.b2_pipe_probe(0, ctypes.byref(raw_wide)))

    for _ in range(args.warmup):
        step_resident()
    flush_gather()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    L.b2_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc_ms = []
    phase_sum = {}
    barrier()
    ev0.record(stream)
    fork_streams()
    for _ in range(args.steps):
        step_resident()
    join_streams()
    ev1.record(stream)
    barrier()
    launches = int(L.b2_launch_count(0))
    dev_ms = ev0.elapsed_time(ev1)
    # per-phase times of the last step (CUDA events recorded by the library on the same stream)
    ph = _lib.last_msm_phases()
    c_w = (ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32())
    L.b2_msm_config(srs.handle, n, 254, ctypes.byref(c_w[0]), ctypes.byref(c_w[1]), ctypes.byref(c_w[2]))
    c_bits, windows, bucket_sets = c_w[0].value, c_w[1].value, c_w[2].value
    # accumulate-kernel duration: measure it over several steps through the phase events
    for _ in range(min(args.steps, 5)):
        step_resident()
        flush_gather()
        torch.cuda.synchronize()
        p = _lib.last_msm_phases()
        acc_ms.append(p["accumulate"])
        for k_, v_ in p.items():
            phase_sum[k_] = phase_sum.get(k_, 0.0) + v_
    nph = max(len(acc_ms), 1)
    phases_avg = {k_: v_ / nph for k_, v_ in phase_sum.items()}

    # --- e2e through the public host API: ONE caller thread, every step copies its pinned host column to the device
    # (H2D) and reads its point back (D2H); the asynchronous form of the call (gpu_multiexp_async -> b2_msm_async) lets
    # the copies of the next steps run under the kernels of step i, up to --e2e-depth steps in flight
    def run_e2e_pipelined(steps):
        from collections import deque
        res, inflight = None, deque()
        for _ in range(steps):
            if len(inflight) == args.e2e_depth:
                res = inflight.popleft().result()
            inflight.append(parallel.sharded_msm_async(h_scalars, srs, 254))
        while inflight:
            res = inflight.popleft().result()
        return res

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    run_e2e_pipelined(3)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res_e2e = step_e2e()
    barrier()
    e2e_sync_s = time.perf_counter() - t0
    barrier()
    t0 = time.perf_counter()
    res_pipe = run_e2e_pipelined(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    assert np.array_equal(res_pipe, res_e2e), "pipelined and blocking host-API results differ"
    # the same API driven the way the reference's prover drives it: several host threads (rayon
    # workers, plonk/prover.rs:293) committing columns at once; each call still copies its own
    # scalars in and its point out, the library overlaps them across its lanes
    from concurrent.futures import ThreadPoolExecutor
    n_callers = 3

    def caller(steps):
        _lib.set_device(local)
        for _ in range(steps):
            h2.gpu_multiexp_single_gpu_with_bound(h_scalars, srs, 254)

    with ThreadPoolExecutor(n_callers) as ex:      # untimed: every lane allocates its workspace on first use
        list(ex.map(caller, [max(1, args.warmup // 2)] * n_callers))
    torch.cuda.synchronize()
    barrier()
    with ThreadPoolExecutor(n_callers) as ex:
        t0 = time.perf_counter()
        list(ex.map(caller, [args.steps] * n_callers))
        torch.cuda.synchronize()
        conc_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # consistency: the host-API result equals the device-resident result (same inputs)
    step_resident()
    res_t = result_tensor()
    torch.cuda.synchronize()
    res_dev = np.ascontiguousarray(res_t.cpu().numpy().view(np.uint64))
    _lib.check(L.b2_g1_normalize(_lib.ptr(res_dev), 1))
    assert np.array_equal(res_dev, res_e2e), "device-resident and host-API results differ"

    # parity of the N-rank result against the closed form, computed on the host without the engine or the oracle:
    # bases are [h_i] G with known 64-bit h_i, so MSM(s, bases) = [sum s_i h_i mod r] G; the scalars are Montgomery
    # residues a_i = s_i * 2^256, so sum s_i h_i = 2^-256 * sum a_i h_i
    t0 = time.perf_counter()
    part = dot_u256_u64(h_scalars, synthetic_multipliers(n, first, seed))
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, part)
    else:
        parts = [part]
    parity = None
    if rank == 0:
        t = sum(parts) * pow(1 << 256, -1, R_MOD) % R_MOD
        want = g1_mul_gen_host(t)
        got = decode_normalized_point(res_dev)
        parity = {"n_ranks": world, "ok": bool(got == want), "points": n_total,
                  "check": "engine result (device-resident path == host-API path, asserted) against "
                           "[sum_i s_i h_i mod r] G computed on the host with Python integers",
                  "host_s": time.perf_counter() - t0}
        if not parity["ok"]:
            print(f"PARITY FAILURE: got {got}, want {want}", file=sys.stderr, flush=True)

    t = torch.tensor([dev_ms, e2e_s * 1e3, conc_s * 1e3, e2e_sync_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, conc_ms, e2e_sync_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])

    strong = None
    if not args.no_strong:
        strong = strong_scaling_sweep(args, torch, dist, dev, stream, _lib, rank, world)

    ntt = None
    if world == 1 and not args.no_ntt:
        ntt = bench_ntt(args, torch, dev, _lib, h2, float(muls.value))
    quotient = None
    if world == 1 and not args.no_quotient:
        quotient = bench_quotient(args, _lib, h2, float(muls.value))

    proof = None
    if world == 1 and not args.no_proof:
        proof = bench_create_proof(args, _lib, h2)

    proof22 = None
    if world == 1 and not args.no_proof22:
        proof22 = bench_create_proof_zkwasm(args, _lib, h2)

    sharded_proof = None
    if world > 1 and not args.no_proof:
        sharded_proof = bench_sharded_proof(args, torch, dist, _lib, h2, rank, world)

    if rank == 0:
        hbm_peak, peak_src = _peaks()
        total_pts = n_total * args.steps
        value = total_pts / (dev_ms * 1e-3) / 1e6
        acc = statistics.mean(acc_ms)
        mac_per_launch = 128.0 * 10.0 * n * windows          # 10 mul-equivalents per mixed add
        achieved = mac_per_launch / (acc * 1e-3) / 1e12
        peak = float(macs.value) / 1e12
        sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
        sm_hz = float((clocks or {}).get("sm_max_mhz") or 1965.0) * 1e6
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u32x8 (256-bit Montgomery integers)", "data": "synthetic",
            "config": {
                "workload": workload_desc(args.logn, world, args.scaling),
                "detail": "each rank's point-range shard of the SRS is resident; one 96-byte partial per rank "
                          "all-gathered over NCCL and summed",
                "points_per_gpu": n, "points_total": n_total, "window_bits": c_bits, "windows": windows,
                "bucket_sets": bucket_sets,
                "srs_window_table": not args.no_precompute, "parallelism": f"range-shard x{world}",
                "cache": "inputs larger than L2: scalars 128 MiB + bases 256 MiB + sort buffers 512 MiB per step",
                "timing": "CUDA events on the launching stream, max over ranks",
                "value_streams": args.value_streams,
                "gather": ("one all-gather + one grouped sum per block of "
                           f"{args.gather_every if args.gather_every > 0 else args.steps} steps, inside the timed region (the "
                           "sharded prover gathers the points of a block of columns the same way)") if world > 1 else None,
                "value_loop": (f"the steps are independent MSMs issued on {args.value_streams} caller streams in turn (the sort of "
                               "MSM i + 1 runs under the bucket reduction of MSM i); the events bracket all of them on the "
                               "first stream, which forks the others after the start event and joins them before the end "
                               "event" if args.value_streams > 1 else "one caller stream, the MSMs run back to back"),
            },
            "e2e": {"value": total_pts / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": 96,
                    "api": "halo2_gpu_specific_b200.parallel.sharded_msm_async (pinned host column -> b2_msm_async -> "
                           f"96 B), one caller thread, up to {args.e2e_depth} steps in flight (one lane each): the H2D copies "
                           "of the next steps overlap the kernels of step i; every step still copies its own 2^logn x 32 B "
                           "in and reads its point out",
                    "blocking_call": {"value": total_pts / (e2e_sync_ms * 1e-3) / 1e6, "unit": UNIT,
                                      "ms_per_step": e2e_sync_ms / args.steps,
                                      "api": "parallel.sharded_msm -> b2_msm (copy, then compute, then read-back)"}},
            "e2e_concurrent": {"value": n_total * args.steps * n_callers / (conc_ms * 1e-3) / 1e6, "unit": UNIT,
                               "host_threads": n_callers,
                               "note": "same per-call copies; 3 concurrent callers per GPU (local partial MSMs only, "
                                       "no cross-rank combine), as the reference's rayon workers would call it"},
            "gpu_launches": launches,
            "parity_check": parity,
            "host_binding": ({"cpus_rank0": len(numa_cpus), "note": "each rank bound to the CPUs NVML reports local to its GPU "
                              "before allocating pinned buffers (B2_NUMA_BIND=0 disables)"} if numa_cpus else None),
            "roofline": {
                "kernel": "msm_accumulate_kernel", "bound": "int", "achieved": achieved, "peak": peak, "unit": "TMAC/s",
                "frac": achieved / peak,
                "frac_vs": {
                    "montgomery_product_probe": achieved / peak,
                    "raw_imad_wide_probe": achieved / (raw_wide.value / 1e12),
                    "nominal_64_per_clk_per_sm": achieved / (sm_hz * 64 * sms / 1e12),
                    "note": "raw probe = b2_pipe_probe(0) in this run: 8 independent IMAD.WIDE.U32 chains per thread, "
                            f"{raw_wide.value / 1e12:.2f} T/s = {raw_wide.value / (sm_hz * sms):.1f} "
                            "per clock per SM (the 32 x 32 + 64 form runs at half the 64 / clk / SM of plain IMAD; "
                            "SASS and ncu pipe counters of the probe: profiles/r2_ncu_summary.md)"},
                "traffic": NCU_TRAFFIC.get(("msm_accumulate", args.logn, not args.no_precompute)),
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one launch, from the committed "
                                "ncu --set full export profiles/r2_ncu_msm_accumulate.raw.csv (7.513 GB read + 0.083 GB "
                                "written at 2^22 with window tables); algorithmic gather bytes 3.7e9: DRAM is ~13 % busy",
                "model": f"128 MACs x 10 mul-equivalents x n x W = {mac_per_launch:.3e} 32x32->64 MACs per launch "
                         f"(SURVEY 8d), duration {acc:.3f} ms (CUDA events, mean of {len(acc_ms)}); the kernel issues "
                         "9.44 product-equivalents per mixed add (the two products of Y3 share one reduction), so "
                         "the multiplier pipe itself is busy 0.944x this fraction",
                "peak_source": "b2_imad_probe in this run: carry-chained IMAD.WIDE Montgomery products, "
                               f"{muls.value / 1e9:.1f} G modmul/s x 128",
                "share_of_step": acc / phases_avg.get("total", acc),
                "share_note": "accumulate kernel time over the time of ONE MSM run alone (library phase events; the ncu "
                              "launch list serialises the same way).  In the timed loop the MSMs of different caller streams "
                              f"overlap at their heads and tails, so a step takes {dev_ms / args.steps:.3f} ms there and the "
                              f"kernel is {acc / (dev_ms / args.steps):.3f} of it",
            },
            "msm_phases_ms": phases_avg,
            "clocks": clocks,
            "hbm_peak": {"gbs": hbm_peak, "source": peak_src},
        }
        if strong:
            line["strong_scaling"] = strong
        if ntt:
            line["ntt"] = ntt
        if quotient:
            line["quotient"] = quotient
        if proof:
            line["create_proof"] = proof
        if proof22:
            line["create_proof_k22"] = proof22
        if sharded_proof is not None:
            line["sharded_create_proof"] = sharded_proof
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args, srs, h_scalars, res_dev)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def strong_scaling_sweep(args, torch, dist, dev, stream, _lib, rank, world):
    """BASELINE config 3: ONE MSM of fixed total size n in {2^18 .. 2^26}, points split ceil(n/G) over the G ranks
    (arithmetic.rs:413-440), device resident, CUDA events on the launching stream, max over ranks.  Every result is
    checked against the closed form on the host.  The same command at G = 1, 2, 4, 8 gives the strong-scaling curve."""
    from halo2_gpu_specific_b200.arithmetic import Srs
    from halo2_gpu_specific_b200 import parallel
    L = _lib.lib()
    sp = ctypes.c_void_p(stream.cuda_stream)
    seed = 0xB2000003
    out = {}
    for logn in [int(x) for x in args.strong_logn.split(",") if x]:
        n_total = 1 << logn
        first, hi = parallel.shard_range(n_total, world, rank)
        n = hi - first
        srs = Srs.synthetic(n, first_index=first, seed=seed)
        if not args.no_precompute:
            srs.precompute()
        h_sc = random_montgomery_scalars(n, seed + 77 * logn + rank)
        d_sc = torch.from_numpy(h_sc.view(np.int64)).to(dev)
        d_partial = torch.zeros(12, dtype=torch.int64, device=dev)
        d_gather = torch.zeros(12 * world, dtype=torch.int64, device=dev)
        d_result = torch.zeros(12, dtype=torch.int64, device=dev)

        def step():
            _lib.check(L.b2_msm_dev(srs.handle, 0, ctypes.c_void_p(d_sc.data_ptr()), n, 254,
                                    ctypes.c_void_p(d_partial.data_ptr()), sp))
            if world > 1:
                dist.all_gather_into_tensor(d_gather, d_partial)
                _lib.check(L.b2_g1_sum_dev(ctypes.c_void_p(d_gather.data_ptr()), world,
                                           ctypes.c_void_p(d_result.data_ptr()), sp))

        reps = 10 if logn <= 22 else (5 if logn <= 24 else 3)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res = np.ascontiguousarray((d_result if world > 1 else d_partial).cpu().numpy().view(np.uint64))
        _lib.check(L.b2_g1_normalize(_lib.ptr(res), 1))
        part = dot_u256_u64(h_sc, synthetic_multipliers(n, first, seed))
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            parts = [None] * world
            dist.all_gather_object(parts, part)
        else:
            parts = [part]
        ms = float(t[0])
        ok = None
        if rank == 0:
            want = g1_mul_gen_host(sum(parts) * pow(1 << 256, -1, R_MOD) % R_MOD)
            ok = bool(decode_normalized_point(res) == want)
        out[f"2^{logn}"] = {"ms": ms, "mpts_per_s": n_total / (ms * 1e-3) / 1e6, "points_per_gpu": n, "parity_ok": ok}
        srs.free()
        del d_sc
    return {"scaling": "strong", "n_gpus": world, "sizes": out,
            "note": "one MSM of the stated TOTAL size split by point range over the ranks; device-resident scalars and "
                    "window tables, per-rank partial all-gathered (96 B) and summed; max over ranks"}


def _ntt_multiplier_occupancy(L, elems_per_s, k, passes, montgomery_per_s):
    """time the multiplier pipe would need at the probed product rates / measured time: k/2 Shoup products per element
    (butterflies) + (passes - 1) Montgomery products per element (inter-pass twiddles)"""
    shoup = ctypes.c_double()
    if L.b2_shoup_probe(ctypes.byref(shoup)) != 0 or shoup.value <= 0:
        return None
    need_s_per_elem = (k / 2.0) / shoup.value + (passes - 1) / montgomery_per_s
    return {"frac": need_s_per_elem * elems_per_s, "shoup_products_per_s": shoup.value,
            "model": "(k/2 butterfly products at the Shoup probe rate + (passes-1) twiddle products at the Montgomery probe "
                     "rate) per element / measured time per element"}


def bench_ntt(args, torch, dev, _lib, h2, modmuls_per_s):
    """64 columns, k = logn forward NTT, device resident (config 2 of BASELINE.json)"""
    from halo2_gpu_specific_b200._lib import NttDesc
    L = _lib.lib()
    k = args.logn
    n = 1 << k
    cols = args.ntt_cols
    dom = h2.EvaluationDomain(5, k)
    g = torch.Generator(device=dev)
    g.manual_seed(0xB2000002)
    # 252-bit Montgomery residues generated on the device
    x = torch.randint(-2**63, 2**63 - 1, (cols, n, 4), dtype=torch.int64, device=dev, generator=g)
    x[:, :, 3] &= (1 << 60) - 1
    stream = torch.cuda.current_stream()
    assert stream.cuda_stream != 0
    torch.cuda.synchronize()
    d = NttDesc()
    d.log_n, d.location = k, 1
    d.omega = dom.omega.ctypes.data
    d.n_in = d.n_out = d.in_stride = d.out_stride = n
    d.columns = cols
    d.in_ = d.out = x.data_ptr()
    d.stream = stream.cuda_stream
    steps = max(3, min(args.steps, 10))
    for _ in range(3):
        _lib.check(L.b2_ntt_exec(ctypes.byref(d)))
    torch.cuda.synchronize()
    L.b2_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        _lib.check(L.b2_ntt_exec(ctypes.byref(d)))
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = int(L.b2_launch_count(0)) // steps
    elems = cols * n
    hbm_peak, peak_src = _peaks()
    passes = launches
    gbs = 64.0 * elems / (ms * 1e-3) / 1e9
    macs = 64.0 * k * elems
    # e2e: host pinned column batch through lagrange_to_coeff_batch (H2D + iNTT + D2H), 8 columns
    ecols = 8
    hx = _lib.pinned_empty((ecols, n, 4))
    random_montgomery_scalars(ecols * n, 7, pinned=hx.reshape(-1, 4))
    dom.lagrange_to_coeff_batch(hx)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        dom.lagrange_to_coeff_batch(hx)
    e2e_ms = (time.perf_counter() - t0) / reps * 1e3
    _lib.pinned_free(hx)
    return {
        "metric": f"Fr NTT throughput at 2^{k}, {cols} columns, device resident", "value": elems / (ms * 1e-3) / 1e6,
        "unit": "Melem/s", "ms_per_batch": ms, "passes": passes,
        "roofline_hbm": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                         "model": "64 B per element per transform (one read + one write)", "peak_source": peak_src},
        "roofline_int": {"bound": "int", "achieved": macs / (ms * 1e-3) / 1e12, "peak": modmuls_per_s * 128 / 1e12,
                         "unit": "TMAC/s", "frac": (macs / (ms * 1e-3)) / (modmuls_per_s * 128),
                         "model": "64 * log2(n) MACs per element (SURVEY 8d); the butterflies multiply by precomputed "
                                  "twiddles with Shoup's method (92 wide MACs + 23 narrow products instead of 128 + "
                                  "8 per product), so the multiplier pipe does ~0.81x the work this model charges"},
        "multiplier_occupancy": _ntt_multiplier_occupancy(L, elems / (ms * 1e-3), k, passes, modmuls_per_s),
        "e2e": {"value": ecols * n / (e2e_ms * 1e-3) / 1e6, "unit": "Melem/s", "columns": ecols,
                "h2d_bytes": ecols * n * 32, "d2h_bytes": ecols * n * 32,
                "api": "EvaluationDomain.lagrange_to_coeff_batch (pinned host columns, iNTT)"},
    }


def bench_quotient(args, _lib, h2, modmuls_per_s):
    """evaluate_h + h(X) for a synthetic zkWasm-scale constraint system at k = logn (the
    "k = 22 create_proof" phase that follows the commitments): coefficient forms resident in HBM, the extended
    domain walked coset by coset (one batched size-2^k transform of all polynomials + ONE fused kernel per coset),
    then divide_by_vanishing_poly (folded into the kernel's store) and extended_to_coeff back to the host."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import quotient_bench as qb
    from halo2_gpu_specific_b200 import _fr
    from halo2_gpu_specific_b200 import evaluation as E
    k = args.logn
    n = 1 << k
    ev, lookups, shuffles, n_sets = qb.synthetic_evaluator()
    prog = ev.program(n_sets, lookups, shuffles)
    info = prog.info()
    ncols = prog.n_fixed + prog.n_advice + prog.n_instance + prog.n_aux
    dom = h2.EvaluationDomain(5, k)
    nc = 1 << (dom.extended_k - k)
    R = _fr.R_MOD
    base = np.empty((4, n, 4), dtype=np.uint64)
    random_montgomery_scalars(4 * n, 11, pinned=base.reshape(-1, 4))
    coef = E.DeviceBuffer(ncols * n)
    for c in range(ncols):
        coef.upload(base[c % 4], c * n)
    cos = E.DeviceBuffer(ncols * n)
    out = E.DeviceBuffer(dom.extended_len())
    ptrs = [cos.ptr + c * n * 32 for c in range(ncols)]
    nf, na, ni = prog.n_fixed, prog.n_advice, prog.n_instance
    challenges = [(i + 2) * 0x123456789ABCDEF % R for i in range(prog.n_challenges)]
    L = _lib.lib()
    runs = []
    for rep in range(3):
        L.b2_synchronize()
        L.b2_launch_count(1)
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
        h = E.extended_to_coeff_dev(dom, out)
        runs.append((time.perf_counter() - t0, ntt_ms, eval_ms, int(L.b2_launch_count(0))))
    wall, ntt_ms, eval_ms, launches = min(runs[1:])
    coef.free(); cos.free(); out.free()
    rows = n * nc
    muls = info["n_mul"] + 2                       # + the coset point and the vanishing scale
    res = {
        "metric": f"evaluate_h + h(X) coefficients at k={k} (extended_k={dom.extended_k}), synthetic zkWasm-scale "
                  f"constraint system: {ncols} polynomials, {info['n_instr']} field instructions per row",
        "value": wall, "unit": "s", "higher_is_better": False, "rows": rows, "gpu_launches": launches,
        "ntt_kernel_ms": ntt_ms, "eval_kernel_ms": eval_ms, "d2h_bytes": int(h.nbytes),
        "resident_GiB": (2 * ncols * n + rows) * 32 / 2**30, "program": info,
        "eval_rows_per_s": rows / (eval_ms * 1e-3),
        "roofline_int": {"kernel": "quotient_eval_kernel", "bound": "int", "unit": "TMAC/s",
                         "achieved": muls * 128.0 * rows / (eval_ms * 1e-3) / 1e12, "peak": modmuls_per_s * 128 / 1e12,
                         "frac": muls * rows / (eval_ms * 1e-3) / modmuls_per_s,
                         "model": f"{muls} field multiplications per row x 128 MACs (SURVEY 8d accounting); "
                                  f"additions and operand decode are not counted.  The host fuses a * b +- c * d into one "
                                  f"instruction with ONE Montgomery reduction (192 + 8 MACs for two products): the pipe "
                                  f"does less than the model counts, so the fraction can approach 1 while the pipe "
                                  f"itself stays at its ~82 % (B2_Q_NO_FUSE=1 runs the plain form)"},
    }
    if not args.no_cpu:
        # CPU: the C restatement of the reference's row loop (Calculation::evaluate, plonk/evaluation.rs:846-1001) on a
        # bounded sample of rows of the SAME program, all host cores
        from oracle import cref
        f = ev.flat_h_program(n_sets, lookups, shuffles)
        cores = os.cpu_count() or 1
        lr = 12
        crow = 1 << lr
        cols_h = [np.ascontiguousarray(base[c % 4][:crow]) for c in range(ncols)]
        enc = lambda v: np.stack([_fr.to_mont(x) for x in v])  # noqa: E731
        t0 = time.perf_counter()
        reps = 0
        while reps < 2 or (time.perf_counter() - t0 < 5.0 and reps < 50):
            cref.quotient_eval(f["rotations"], enc(f["constants"]), f["calcs"], f["result"], cols_h[:nf],
                               cols_h[nf:nf + na], cols_h[nf + na:nf + na + ni], cols_h[nf + na + ni:], enc(challenges),
                               lr, 1, x0=_fr.to_mont(1), step=dom.omega, threads=cores)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        res["cpu_baseline"] = {"value": crow / dt, "unit": "rows/s", "cores": cores, "kind": "port",
                               "sample": f"2^{lr} rows of the same program per repetition, {reps} repetitions "
                                         f"(oracle/cpu_ref.c ref_quotient_eval; expression kernel only, no transforms)"}
    return res


def bench_create_proof(args, _lib, h2):
    """BASELINE config 4: the benches/plonk.rs circuit, full create_proof (GWC) through the prover mirror
    (halo2_gpu_specific_b200.plonk) at k = --proof-k, device-resident engine with the proving key kept resident
    between proofs.  Wall clock around the whole call: host transcript, RNG and bookkeeping included, the advice
    columns start in pinned host memory and cross PCIe once.  The SRS is a real KZG one built on the device
    (Params.unsafe_setup); the same code path is checked against the oracle's verifier, pairing included, in
    tests/test_gpu_prover.py (k = 8 ... 20).  `host_api` repeats the measurement with the engine that copies the
    operands of every call in and out (how the reference's cuda build drives its GPU).  Never fatal for the main line."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import plonk_bench_circuit as bc
        from halo2_gpu_specific_b200 import plonk as HP
        k = args.proof_k
        cs = HP.ConstraintSystem(**bc.constraint_system_args())
        fixed, advice, mapping = bc.build(k)
        params = h2.Params.unsafe_setup(k, 0x2B200B200B200B200B200B200B200B2001)
        adv = _lib.pinned_empty(advice.shape)
        try:
            pk = HP.keygen(params, cs, fixed, mapping)
            L = _lib.lib()
            out = {}
            for kind in ("resident", "host_api"):
                eng = HP.ResidentEngine(params, pk.vk.domain) if kind == "resident" else HP.Engine(params, pk.vk.domain)
                times, phases, launches, nbytes = [], {}, 0, 0
                for it in range(1 + args.proof_reps):
                    adv[:] = advice
                    tm = {}
                    l0 = L.b2_launch_count(0)
                    t0 = time.perf_counter()
                    proof = HP.create_proof(params, pk, adv, [], HP.SeededRng(it), timings=tm, engine=eng)
                    dt = time.perf_counter() - t0
                    if it:                                    # first call warms plans, tables, key cosets, workspaces
                        times.append(dt)
                        launches = L.b2_launch_count(0) - l0
                        nbytes = len(proof)
                        for name, v in tm.items():
                            phases.setdefault(name, []).append(v)
                eng.free()
                out[kind] = {"value": statistics.median(times), "unit": "s", "all_s": times,
                             "phases_s": {a: statistics.median(b) for a, b in phases.items()},
                             "gpu_launches": launches, "proof_bytes": nbytes}
            cpu = None
            if not args.no_cpu:
                cpu = _cpu_proof_schedule(params, pk, advice, k)
            res = out["resident"]
            if cpu:
                res["cpu_baseline"] = cpu
            res.update({
                "metric": f"create_proof wall time, benches/plonk.rs circuit at k={k} (3 advice, 4 fixed, 1 permutation "
                          f"set, degree 5), GWC multiopen",
                "higher_is_better": False, "reps": args.proof_reps, "h2d_bytes": int(advice.nbytes),
                "srs": "KZG SRS built on the device (Params.unsafe_setup)",
                "api": "halo2_gpu_specific_b200.plonk.create_proof (pinned host advice columns -> proof bytes), "
                       "ResidentEngine",
                "host_api_engine": out["host_api"],
            })
            return res
        finally:
            _lib.pinned_free(adv)
            params.free()
    except Exception as e:                                # noqa: BLE001 -- reported, never fatal for the MSM line
        import traceback
        return {"error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc()[-600:]}


def _cpu_proof_schedule(params, pk, advice, k):
    """cpu_baseline of the create_proof section: the MSM / NTT / quotient-row calls that the SAME proof makes on the
    reference's CPU path (plonk/prover.rs without the cuda feature), issued on the C restatement (oracle/cpu_ref.c) with
    all host cores: 3 advice + 1 z commitments against g_lagrange, 1 random + 4 h-piece + 2 opening commitments against
    g, 4 inverse transforms of size 2^k, 4 coset extensions (advice and z; the proving key's cosets exist since keygen),
    the row loop of evaluate_h over 2^(k+2) rows with this circuit's program, extended_to_coeff.  The field arithmetic
    of the evaluation phase and of the multiopen folds is not included: a lower bound for the CPU prover."""
    from halo2_gpu_specific_b200 import _fr
    from oracle import cref
    cores = os.cpu_count() or 1
    dom = pk.vk.domain
    n = 1 << k
    g, gl = params.g.read(), params.g_lagrange.read()
    rnd = cref.random_fr_mont(n, 0xB2000092)
    enc = lambda v: np.stack([_fr.to_mont(x) for x in v])  # noqa: E731
    f = pk.ev.flat_h_program(1, [], 0)
    fixed_ext = [cref.coeff_to_extended(p, k, dom.extended_k, dom.g_coset, dom.g_coset_inv, dom.extended_omega, cores)
                 for p in pk.fixed_polys]                                  # keygen-time data, untimed
    sigma_ext = [cref.coeff_to_extended(p, k, dom.extended_k, dom.g_coset, dom.g_coset_inv, dom.extended_omega, cores)
                 for p in pk.sigma_polys]
    challenges = [(i + 2) * 0x123456789ABCDEF % _fr.R_MOD for i in range(f["n_challenges"])]
    t0 = time.perf_counter()
    cols = [np.ascontiguousarray(advice[i]) for i in range(advice.shape[0])] + [rnd]
    for c in cols:
        cref.best_multiexp(c, gl, cores)
    for _ in range(7):
        cref.best_multiexp(rnd, g, cores)
    t_msm = time.perf_counter() - t0
    t1 = time.perf_counter()
    coeffs = [cref.ifft(c, dom.omega_inv, dom.ifft_divisor, k, cores) for c in cols]
    ext = [cref.coeff_to_extended(c, k, dom.extended_k, dom.g_coset, dom.g_coset_inv, dom.extended_omega, cores)
           for c in coeffs]
    t_ntt = time.perf_counter() - t1
    t2 = time.perf_counter()
    aux = [pk.l0, pk.l_last, pk.l_active_row] + sigma_ext + [ext[3]]
    h = cref.quotient_eval(f["rotations"], enc(f["constants"]), f["calcs"], f["result"], fixed_ext, ext[:3], [], aux,
                           enc(challenges), dom.extended_k, 1 << (dom.extended_k - k), x0=_fr.to_mont(1),
                           step=dom.extended_omega, threads=cores)
    t_rows = time.perf_counter() - t2
    t3 = time.perf_counter()
    cref.extended_to_coeff(h, dom.extended_k, dom.g_coset, dom.g_coset_inv, dom.extended_omega_inv,
                           dom.extended_ifft_divisor, cores)
    t_ntt += time.perf_counter() - t3
    total = time.perf_counter() - t0
    return {"value": total, "unit": "s", "cores": cores, "kind": "port",
            "sample": "the 11 MSMs, 4 iNTTs, 4 coset extensions, the evaluate_h row loop and extended_to_coeff of ONE "
                      "proof of the same circuit on oracle/cpu_ref.c (a lower bound: evaluation-phase and multiopen "
                      "field arithmetic, transcript and witness handling not included)",
            "split_s": {"msm": t_msm, "ntt": t_ntt, "evaluate_h_rows": t_rows}}


def bench_create_proof_zkwasm(args, _lib, h2):
    """BASELINE's third headline number, "k = 22 create_proof wall-time", as a real proof on one GPU: the zkWasm-shaped
    synthetic circuit of tools/zkwasm_shape_circuit.py (64 advice, 32 fixed, 1 instance, 8 lookups / 12 input sets,
    4 shuffles, 24 permutation columns, degree 5; a stand-in gate set, zkWasm's own circuit is not available) through
    plonk.create_proof with the device-resident engine.  Wall clock around the call; the 8 GiB of advice start in pinned
    host memory; SRS built on the device; proving key resident (second and later proofs).  The same script with the
    oracle's verifier attached is tests/manual/prove_zkwasm_shape.py.  Never fatal for the main line."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import zkwasm_shape_circuit as zk
        from halo2_gpu_specific_b200 import plonk as HP
        k = args.proof22_k
        params = h2.Params.unsafe_setup(k, 0x2B200B200B200B200B200B200B200B2001)
        adv = None
        try:
            cs = HP.ConstraintSystem(**zk.constraint_system_args(extra_gates=300))
            dom = h2.EvaluationDomain(cs.degree(), k)
            fixed, advice, public, mapping = zk.build(k, HP.Engine(params, dom).to_mont, seed=k)
            pk = HP.keygen(params, cs, fixed, mapping)
            adv = _lib.pinned_empty(advice.shape)
            adv[:] = advice
            del advice, fixed
            eng = HP.ResidentEngine(params, pk.vk.domain, profile=True)
            L = _lib.lib()
            runs = []
            for it in range(1 + args.proof22_reps):
                tm = {}
                eng.op_times.clear()
                l0 = L.b2_launch_count(0)
                t0 = time.perf_counter()
                proof = HP.create_proof(params, pk, adv, [public], HP.SeededRng(it), timings=tm, engine=eng,
                                        use_gwc=not args.shplonk)
                runs.append({"wall_s": time.perf_counter() - t0, "phases_s": tm, "launches": int(L.b2_launch_count(0) - l0),
                             "ops": {a: [round(b[0], 5), b[1]] for a, b in sorted(eng.op_times.items())},
                             "bytes": len(proof)})
            eng.free()
            best = min(runs[1:], key=lambda r: r["wall_s"])
            return {"metric": f"create_proof wall time, zkWasm-shaped circuit at k={k} "
                              f"({'SHPLONK' if args.shplonk else 'GWC'}), one GPU",
                    "value": best["wall_s"], "unit": "s", "higher_is_better": False, "all_s": [r["wall_s"] for r in runs[1:]],
                    "first_call_s": runs[0]["wall_s"], "phases_s": best["phases_s"], "engine_ops_s_calls": best["ops"],
                    "gpu_launches": best["launches"], "proof_bytes": best["bytes"], "h2d_bytes": int(adv.nbytes),
                    "shape": {"A": zk.A, "F": zk.F, "I": zk.I, "lookup_sets": list(zk.LOOKUP_SETS), "shuffles": zk.SHUFFLES,
                              "perm_cols": zk.PERM_COLS, "degree": 5, "extra_gates": 300},
                    "h_program": pk.ev.program(8, list(zk.LOOKUP_SETS), zk.SHUFFLES).info(),
                    "api": "halo2_gpu_specific_b200.plonk.create_proof, ResidentEngine"}
        finally:
            if adv is not None:
                _lib.pinned_free(adv)
            params.free()
    except Exception as e:                                # noqa: BLE001
        import traceback
        return {"error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc()[-600:]}


def bench_sharded_proof(args, torch, dist, _lib, h2, rank, world):
    """N > 1: the REAL prover divided over the ranks (prover_sharded: advice upload, lookups, z columns, advice transforms
    and evaluate_h rows shared out; NCCL exchanges) on the zkWasm-shaped circuit at k = --sharded-proof-k, next to the
    same proof on one GPU (rank 0, plain ResidentEngine).  The driver-visible parity of the multi-GPU prover: every
    rank's proof bytes are compared with each other and with the single-GPU proof.  Never fatal for the main line; every
    rank must call it (collectives inside)."""
    k = args.sharded_proof_k
    failed = None
    out = None
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import zkwasm_shape_circuit as zk
        from halo2_gpu_specific_b200 import plonk as HP
        from halo2_gpu_specific_b200 import prover_sharded as PS
        params = h2.Params.unsafe_setup(k, 0x2B200B200B200B200B200B200B200B2001)
        adv = None
        try:
            cs = HP.ConstraintSystem(**zk.constraint_system_args(extra_gates=64))
            dom = h2.EvaluationDomain(cs.degree(), k)
            fixed, advice, public, mapping = zk.build(k, HP.Engine(params, dom).to_mont, seed=k)
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                pk = HP.keygen(params, cs, fixed, mapping)
            adv = _lib.pinned_empty(advice.shape)
            adv[:] = advice
            del advice, fixed
            setup_ok = True
        except Exception as e:                            # noqa: BLE001
            setup_ok, failed = False, {"error": f"setup on rank {rank}: {type(e).__name__}: {e}"}
        # a rank whose setup failed must not leave the others waiting in a collective: agree first
        flag = torch.tensor([1 if setup_ok else 0], dtype=torch.int64, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if adv is not None:
                _lib.pinned_free(adv)
            params.free()
            return failed or {"error": "setup failed on another rank"}
        try:
            eng = PS.ShardedResidentEngineQ(params, pk.vk.domain)
            PS.create_proof(params, pk, adv, [public], None, engine=eng)     # warm-up: synchronized rng + digest check
            best, phases, proof = None, None, None
            for _ in range(2):
                dist.barrier()
                tm = {}
                t0 = time.perf_counter()
                proof = HP.create_proof(params, pk, adv, [public], HP.SeededRng(1), engine=eng, timings=tm)
                d = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
                dist.all_reduce(d, op=dist.ReduceOp.MAX)
                if best is None or float(d.item()) < best:
                    best, phases = float(d.item()), tm
            by_range = getattr(eng, "range_commits", 0) // 3          # warm-up + two timed proofs
            eng.free()
            alone_s, same, alone_err = None, True, None
            if rank == 0:
                try:
                    plain = HP.ResidentEngine(params, pk.vk.domain)
                    HP.create_proof(params, pk, adv, [public], HP.SeededRng(0), engine=plain)
                    for _ in range(2):
                        t0 = time.perf_counter()
                        alone = HP.create_proof(params, pk, adv, [public], HP.SeededRng(1), engine=plain)
                        d = time.perf_counter() - t0
                        alone_s = d if alone_s is None else min(alone_s, d)
                    plain.free()
                    same = alone == proof
                except Exception as e:                    # noqa: BLE001  (the other ranks wait in the gather below)
                    same, alone_err = False, f"{type(e).__name__}: {e}"
            gathered = [None] * world
            dist.all_gather_object(gathered, proof)
            same = bool(same and all(g == proof for g in gathered))
            out = {"metric": f"create_proof wall time, zkWasm-shaped circuit at k={k} (GWC), divided over {world} GPUs",
                   "value": best, "unit": "s", "higher_is_better": False, "n_ranks": world, "single_gpu_s": alone_s,
                   "bytes_equal_on_all_ranks_and_to_single_gpu": same, "phases_s": phases, "proof_bytes": len(proof),
                   "columns_committed_by_point_range": by_range,
                   "api": "halo2_gpu_specific_b200.prover_sharded (ShardedResidentEngineQ), one process per GPU, NCCL"}
            if alone_err:
                out["single_gpu_error"] = alone_err
        finally:
            if adv is not None:
                _lib.pinned_free(adv)
            params.free()
    except Exception as e:                                # noqa: BLE001
        import traceback
        failed = {"error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc()[-600:]}
    return out if failed is None else failed


def cpu_baseline(args, srs=None, scalars=None, engine_point=None):
    """The same workload on the host cores (C restatement of the rayon path, oracle/cpu_ref.c): the full MSM of the
    timed step (the engine's own synthetic bases read back from HBM, the same scalars), a few repetitions -- the
    workload of `--impl reference` at N = 1 -- and the result compared with the engine's."""
    from oracle import cref
    cores = os.cpu_count() or 1
    if srs is not None:
        bases = srs.read()
        sample = bases.shape[0]
        scalars = np.ascontiguousarray(scalars[:sample])
        src = "the engine's synthetic bases read back from HBM and the timed step's scalars"
    else:
        sample = 1 << min(args.logn, 18)
        ks = np.zeros((sample, 4), dtype=np.uint64)
        ks[:, 0] = np.arange(1, sample + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        bases = cref.g1_mul_gen(ks, cores)
        scalars = cref.random_fr_mont(sample, 0xB2000003)
        src = "bases [k_i] G generated by the oracle"
    cref.best_multiexp(scalars[:4096], bases[:4096], cores)
    t0 = time.perf_counter()
    reps = 0
    res = None
    while reps < 2 or (time.perf_counter() - t0 < 10.0 and reps < 20):
        res = cref.best_multiexp(scalars, bases, cores)
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    match = None
    if engine_point is not None:
        match = bool(np.array_equal(cref.jac_to_affine(res)[0], np.asarray(engine_point).reshape(12)[:8]))
    k = min(args.logn, 20)
    x = cref.random_fr_mont(1 << k, 0xB2000002)
    from halo2_gpu_specific_b200 import _fr
    om = _fr.to_mont(pow(_fr.ROOT_OF_UNITY, 1 << (28 - k), _fr.R_MOD))
    t1 = time.perf_counter()
    cref.best_fft(x, om, k, cores)
    fft_dt = time.perf_counter() - t1
    return {"value": sample / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"best_multiexp on all {sample} points of the timed step ({src}), {reps} repetitions, "
                      f"{cores} threads, chunk = n/T (oracle/cpu_ref.c, restatement of arithmetic.rs:20-108,465-492)",
            "matches_engine_result": match,
            "ntt": {"value": (1 << k) / fft_dt / 1e6, "unit": "Melem/s", "sample": f"best_fft_cpu k={k}, one column"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--logn", type=int, default=22)
    ap.add_argument("--ntt-cols", type=int, default=64)
    ap.add_argument("--no-ntt", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-quotient", action="store_true")
    ap.add_argument("--no-proof", action="store_true")
    ap.add_argument("--proof-k", type=int, default=18)
    ap.add_argument("--proof-reps", type=int, default=3)
    ap.add_argument("--no-proof22", action="store_true", help="skip the zkWasm-shaped k=22 real proof (about 40 s of setup)")
    ap.add_argument("--proof22-k", type=int, default=22)
    ap.add_argument("--proof22-reps", type=int, default=2)
    ap.add_argument("--value-streams", type=int, default=2, choices=[1, 2, 3],
                    help="caller streams the device-resident loop issues its MSMs on in turn (1 = one stream, serial)")
    ap.add_argument("--gather-every", type=int, default=0,
                    help="N > 1: steps whose partials are all-gathered and summed together (0 = all steps of the loop, 1 = "
                         "a collective per step)")
    ap.add_argument("--e2e-depth", type=int, default=3, choices=[1, 2, 3],
                    help="b2_msm_async tickets one caller thread keeps in flight in the e2e loop (<= B2_LANES)")
    ap.add_argument("--sharded-proof-k", type=int, default=18,
                    help="N > 1: size of the zkWasm-shaped circuit proved by all ranks together (sharded_create_proof)")
    ap.add_argument("--shplonk", action="store_true", help="create_proof_with_shplonk in the proof sections")
    ap.add_argument("--no-precompute", action="store_true", help="plain bases: one bucket set per window + Horner")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 2^logn points per GPU (default); strong: one MSM of 2^logn points split over the GPUs")
    ap.add_argument("--no-strong", action="store_true", help="skip the fixed-total-size MSM sweep (strong_scaling key)")
    ap.add_argument("--strong-logn", default="18,20,22,24,26")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "engine":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
