#!/usr/bin/env python
"""N-GPU check of prover_sharded.ShardedResidentEngine (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \\
        tools/sharded_proof_check.py --k 14

Every rank proves the benches/plonk.rs circuit with its commitments divided over the ranks; rank 0 also proves it alone
and the bytes must agree.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=14)
    ap.add_argument("--circuit", default="bench", choices=["bench", "zkwasm"],
                    help="bench: benches/plonk.rs; zkwasm: tools/zkwasm_shape_circuit.py (64 advice, lookups, shuffles)")
    ap.add_argument("--split-quotient", action="store_true",
                    help="also divide evaluate_h by cosets (ShardedResidentEngineQ)")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--range-shard", default="auto", choices=["auto", "on", "off"],
                    help="few-column blocks (instance, random polynomial, h pieces, multiopen witnesses) divided by point "
                         "range: by the cost model / always / never (ShardedCommits.RANGE_SHARD)")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    import halo2_gpu_specific_b200 as h2
    from halo2_gpu_specific_b200 import _lib
    from halo2_gpu_specific_b200 import plonk as HP
    from halo2_gpu_specific_b200.prover_sharded import ShardedResidentEngine, ShardedResidentEngineQ
    import plonk_bench_circuit as bc
    import zkwasm_shape_circuit as zk
    _lib.require_gpu()
    _lib.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if a.circuit == "zkwasm":
        # every rank builds the witness and the proving key for itself: refuse to start when the host cannot hold
        # `world` copies (a box driven out of memory is a lost box)
        import psutil
        per_rank = (zk.A + zk.F) * (1 << a.k) * 32 * 3.5 + zk.PERM_COLS * (1 << a.k) * 32 * 4
        avail = psutil.virtual_memory().available
        if per_rank * world > 0.8 * avail:
            if rank == 0:
                print(json.dumps({"error": f"host memory: {world} ranks x {per_rank / 2**30:.0f} GiB needed, "
                                           f"{avail / 2**30:.0f} GiB available; use a smaller --k or fewer ranks"}))
            sys.exit(3)
    params = h2.Params.unsafe_setup(a.k, 0x2B200B200B200B200B200B200B200B2001)
    if a.circuit == "bench":
        cs = HP.ConstraintSystem(**bc.constraint_system_args())
        fixed, advice, mapping = bc.build(a.k)
        public = []
    else:
        cs = HP.ConstraintSystem(**zk.constraint_system_args(extra_gates=300))
        dom = h2.EvaluationDomain(cs.degree(), a.k)
        fixed, advice, pub, mapping = zk.build(a.k, HP.Engine(params, dom).to_mont, seed=a.k)
        public = [pub]
    pk = HP.keygen(params, cs, fixed, mapping)
    pinned = _lib.pinned_empty(advice.shape)          # the witness waits in page-locked host memory, as in bench.py
    pinned[:] = advice
    del advice, fixed
    advice = pinned
    eng = (ShardedResidentEngineQ if a.split_quotient else ShardedResidentEngine)(params, pk.vk.domain)
    eng.RANGE_SHARD = {"auto": None, "on": True, "off": False}[a.range_shard]
    from halo2_gpu_specific_b200 import prover_sharded as PS
    # warm-up through the public multi-rank entry: rank 0's OS seed broadcast, BLAKE2b stream, bytes compared across ranks
    PS.create_proof(params, pk, advice, public, None, engine=eng)
    dt, phases = None, None
    for rep in range(a.reps):
        work = advice
        if world > 1:
            dist.barrier()
        tm = {}
        t0 = time.perf_counter()
        proof = HP.create_proof(params, pk, work, public, HP.SeededRng(1), engine=eng, timings=tm)
        d = time.perf_counter() - t0
        if world > 1:                                    # a proof is done when the slowest rank is
            t = torch.tensor([d], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            d = float(t.item())
        if dt is None or d < dt:
            dt, phases = d, tm
    range_commits = getattr(eng, "range_commits", 0)
    eng.free()
    ok = True
    alone_s = None
    if rank == 0:
        plain = HP.ResidentEngine(params, pk.vk.domain)
        HP.create_proof(params, pk, advice, public, HP.SeededRng(0), engine=plain)
        alone_phases = None
        for rep in range(a.reps):
            work = advice
            tm = {}
            t0 = time.perf_counter()
            alone = HP.create_proof(params, pk, work, public, HP.SeededRng(1), engine=plain, timings=tm)
            d = time.perf_counter() - t0
            if alone_s is None or d < alone_s:
                alone_s, alone_phases = d, tm
        plain.free()
        ok = alone == proof
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, proof)
        ok = ok and all(g == proof for g in gathered)
    if rank == 0:
        print(json.dumps({"check": f"sharded create_proof, {a.circuit} circuit, quotient split: {a.split_quotient}", "k": a.k,
                          "n_gpus": world, "range_shard": a.range_shard,
                          "columns_committed_by_point_range": range_commits // (a.reps + 1),
                          "bytes_equal_on_all_ranks_and_to_single_gpu": bool(ok), "sharded_s": dt, "single_gpu_s": alone_s,
                          "sharded_phases_s": phases, "single_gpu_phases_s": alone_phases, "reps": a.reps,
                          "rng_sync_check": "warm-up proof through prover_sharded.create_proof (broadcast seed, "
                                            "digest all-gather) passed",
                          "backend": "nccl" if world > 1 else "none"}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    _lib.pinned_free(pinned)
    params.free()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
