#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/multi_gpu_quotient_check.py [--k 20]
evaluate_h + h(X) with the extended domain's rows split over N GPUs (parallel.sharded_evaluate_h): parity of every
rank's result against a single-GPU run on rank 0, and device-side timing of the sharded phase (max over ranks)."""
import argparse
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
from halo2_gpu_specific_b200 import _lib, parallel  # noqa: E402
import quotient_bench as qb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=18)
    ap.add_argument("--gates", type=int, default=96)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    _lib.require_gpu()
    _lib.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    k, n = a.k, 1 << a.k
    ev, lookups, shuffles, n_sets = qb.synthetic_evaluator(gates=a.gates)
    dom = h2.EvaluationDomain(5, k)
    ext = dom.extended_len()
    rng = np.random.default_rng(9)      # same data on every rank
    base = rng.integers(0, 1 << 62, size=(5, n, 4), dtype=np.uint64)
    base[:, :, 3] &= np.uint64((1 << 60) - 1)
    lag = rng.integers(0, 1 << 62, size=(3, ext, 4), dtype=np.uint64)
    lag[:, :, 3] &= np.uint64((1 << 60) - 1)
    pick = lambda i: base[i % 5]  # noqa: E731
    fixed = [pick(i) for i in range(ev.num_fixed)]
    advice = [pick(i + 1) for i in range(ev.num_advice)]
    inst = [pick(i + 2) for i in range(ev.num_instance)]
    sigma = [pick(i + 3) for i in range(len(ev.permutation_columns))]
    perms = [pick(i + 4) for i in range(n_sets)]
    lks = [{"z": [pick(i + j) for j in range(s)], "m": pick(i + 2)} for i, s in enumerate(lookups)]
    shf = [pick(i) for i in range(shuffles)]
    args = (dom, fixed, advice, inst, lag[0], lag[1], lag[2], sigma, 3, 5, 7, 11, lks, shf, perms)
    from halo2_gpu_specific_b200.evaluation import ResidentPolys
    aux = list(sigma) + list(perms)
    for lk in lks:
        aux += list(lk["z"]) + [lk["m"]]
    aux += list(shf)
    resident = ResidentPolys(dom, [fixed, advice, inst, aux], [lag[0], lag[1], lag[2]])   # inputs resident in HBM
    times = []
    for rep in range(3):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        got = parallel.sharded_evaluate_h(ev, *args, resident=resident)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        times.append(time.perf_counter() - t0)
    t = torch.tensor([min(times[1:])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = True
    if rank == 0:
        want = ev.evaluate_h(*args, to_coeff=True)
        ok = bool(np.array_equal(got, want))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        # every rank's copy must equal rank 0's
        mine = torch.from_numpy(got.view(np.int64)).cuda()
        ref = mine.clone()
        dist.broadcast(ref, 0)
        flag &= int(torch.equal(mine, ref))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        doc = {"workload": "sharded evaluate_h + h(X): coefficient forms resident on every rank, rows split coset-major, "
                           "NCCL all-gather of h, extended_to_coeff + D2H of the coefficients on every rank", "k": k, "n_gpus": world,
               "parity_vs_single_gpu": bool(flag.item()), "wall_s": float(t.item()),
               "tasks_rank0": parallel.quotient_tasks(ext // n, n, world, 0)}
        print(json.dumps(doc), flush=True)
        if a.out:
            json.dump(doc, open(a.out, "w"), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
