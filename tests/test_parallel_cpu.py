"""N>1 host logic on CPU: world_size-2 gloo run of the range-sharded MSM (partition, gather order,
combine) with the oracle injected as the per-rank MSM, compared with the unsharded oracle result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, outdir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from halo2_gpu_specific_b200 import parallel
    from oracle import cref
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    scalars = cref.random_fr_mont(n, 0x51)
    ks = np.zeros((n, 4), dtype=np.uint64)
    ks[:, 0] = np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    bases = cref.g1_mul_gen(ks, 2)
    lo, hi = parallel.shard_range(n, world, rank)
    res = parallel.sharded_msm(scalars[lo:hi], bases[lo:hi], 254,
                               local_msm=lambda s, b, bits: cref.best_multiexp(s, b, 2),
                               combine=cref.jac_sum)
    gathered = parallel.all_gather_partials(np.full(12, rank, dtype=np.uint64))
    assert [int(g[0]) for g in gathered] == list(range(world))
    np.save(os.path.join(outdir, f"r{rank}.npy"), cref.jac_to_affine(res)[0])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1000, 1001])
def test_sharded_msm_world2_gloo(tmp_path, n):
    from oracle import cref
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    scalars = cref.random_fr_mont(n, 0x51)
    ks = np.zeros((n, 4), dtype=np.uint64)
    ks[:, 0] = np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    bases = cref.g1_mul_gen(ks, 2)
    want = cref.jac_to_affine(cref.best_multiexp(scalars, bases, 4))[0]
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"r{r}.npy"), want)


def test_shard_ranges_cover_and_match_reference_rule():
    from halo2_gpu_specific_b200 import parallel
    for n in (0, 1, 7, 8, 9, 1 << 22, (1 << 22) + 1):
        for w in (1, 2, 4, 8):
            parts = [parallel.shard_range(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            for a, b in zip(parts, parts[1:]):
                assert a[1] == b[0]
            part_len = (n + w - 1) // w   # arithmetic.rs:426
            assert all(hi - lo <= part_len for lo, hi in parts)
    assert parallel.column_range(64, 8, 3) == (24, 32)


def test_quotient_tasks_tile_the_extended_domain():
    from halo2_gpu_specific_b200 import parallel
    for nc, n in ((4, 1 << 10), (8, 1 << 6), (2, 1 << 5)):
        for w in (1, 2, 4, 8, 16):
            seen = []
            for r in range(w):
                tasks = parallel.quotient_tasks(nc, n, w, r)
                assert sum(t[2] for t in tasks) == nc * n // w
                for c, begin, count in tasks:
                    assert 0 <= c < nc and count > 0 and begin + count <= n
                    seen += [c * n + begin + i for i in (0, count - 1)]
                    seen.append((c * n + begin, c * n + begin + count))
            ranges = sorted(x for x in seen if isinstance(x, tuple))
            assert ranges[0][0] == 0 and ranges[-1][1] == nc * n
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0]          # contiguous, disjoint, coset-major
    assert parallel.quotient_tasks(4, 16, 2, 1) == [(2, 0, 16), (3, 0, 16)]
    assert parallel.quotient_tasks(4, 16, 8, 3) == [(1, 8, 8)]


def _gather_worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from halo2_gpu_specific_b200 import parallel
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    nc, n = 4, 32
    full = (np.arange(nc * n * 4, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)).reshape(nc * n, 4)
    mine = np.concatenate([full[c * n + b: c * n + b + cnt] for c, b, cnt in parallel.quotient_tasks(nc, n, world, rank)])
    got = parallel.all_gather_rows(mine, nc * n)
    np.save(os.path.join(outdir, f"g{rank}.npy"), got)
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_rows_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_gather_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    full = (np.arange(4 * 32 * 4, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)).reshape(128, 4)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"g{r}.npy"), full)


def test_witness_file_layout_round_trip(tmp_path):
    """helpers.rs:919-1015: u32 count, column i at 4 + i * 2^(k+5), raw in-memory field elements (host I/O only)"""
    import numpy as np
    from halo2_gpu_specific_b200 import helpers
    k, cols = 6, 5
    n = 1 << k
    rng = np.random.default_rng(5)
    advice = [rng.integers(0, 2**63, size=(n, 4), dtype=np.uint64) for _ in range(cols)]
    path = str(tmp_path / "w.bin")
    helpers.store_witness(path, advice, k)
    raw = open(path, "rb").read()
    assert len(raw) == 4 + cols * (1 << (k + 5))
    assert int.from_bytes(raw[:4], "little") == cols
    off = 4 + 3 * (1 << (k + 5))
    assert raw[off:off + 32] == advice[3][0].tobytes()
    back = helpers.fetch_witness(path, k)
    assert len(back) == cols and all(np.array_equal(a, b) for a, b in zip(advice, back))


def test_coset_transform_shares_cover_every_polynomial_once():
    """ranks sharing a coset split its transforms (tools/resident_replay_multi.py): the shares of a group are disjoint,
    cover all polynomials, and groups line up with quotient_tasks (same coset for every rank of a group)"""
    from halo2_gpu_specific_b200 import parallel
    n, nc, n_polys = 1 << 10, 4, 97
    for world in (1, 2, 4, 8, 16):
        for rank in range(world):
            first, shares = parallel.coset_transform_shares(nc, world, rank, n_polys)
            assert [s[0] for s in shares] == list(range(first, first + len(shares))) and first <= rank < first + len(shares)
            covered = []
            for _, lo, hi in shares:
                covered += list(range(lo, hi))
            assert covered == list(range(n_polys))
            cosets = {parallel.quotient_tasks(nc, n, world, r)[0][0] for r, _, _ in shares}
            if world > nc:
                assert len(cosets) == 1 and all(len(parallel.quotient_tasks(nc, n, world, r)) == 1 for r, _, _ in shares)
            else:
                assert len(shares) == 1


def _prover_worker(rank, world, port, outdir, range_shard=None):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import plonk_fixture as fxm
    from halo2_gpu_specific_b200 import plonk as HP
    from halo2_gpu_specific_b200.prover_sharded import ShardedCommits, ShardedQuotient
    from oracle import bn254 as o
    from oracle import prover as PR
    from oracle_engine import CosetQuotientDouble, OracleEngine

    class ShardedOracleEngine(ShardedQuotient, ShardedCommits, CosetQuotientDouble, OracleEngine):
        RANGE_SHARD = range_shard      # None: the cost model (column-parallel at this size); True: every full-width block

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    k = 5
    fx = fxm.build(k=k, seed=11)
    ocs = fx["cs"]
    oparams = PR.Params(k, 0x2B200B200B200B2001)
    opk = PR.keygen(oparams, ocs, fx["fixed"], fx["mapping"])
    cs = HP.ConstraintSystem.like(ocs)

    class HostParams:
        pass
    hp = HostParams()
    hp.k, hp.n = k, 1 << k
    eng = ShardedOracleEngine(oparams, opk.vk.domain, ocs)
    calls = {"msm": 0}
    inner = eng._msm

    def counting(col, bases):
        calls["msm"] += 1
        return inner(col, bases)
    eng._msm = counting
    pk = HP.keygen(hp, cs, np.stack([o.fr_encode(c) for c in fx["fixed"]]), np.array(fx["mapping"], dtype=np.int64),
                   engine=OracleEngine(oparams, opk.vk.domain, ocs), transcript_repr=opk.vk.transcript_repr)
    adv = np.ascontiguousarray(np.stack([o.fr_encode(c) for c in fx["advice"]]))
    inst = [fx["instance"][0][:4]]
    out = {}
    for gwc in (True, False):
        calls["msm"] = 0
        out[gwc] = (HP.create_proof(hp, pk, adv.copy(), inst, HP.SeededRng(3), engine=eng, use_gwc=gwc), calls["msm"])
    np.save(os.path.join(outdir, f"proof_r{rank}.npy"), np.frombuffer(out[True][0] + out[False][0], dtype=np.uint8))
    np.save(os.path.join(outdir, f"msms_r{rank}.npy"), np.array([out[True][1], out[False][1]]))
    np.save(os.path.join(outdir, f"ranges_r{rank}.npy"), np.array(getattr(eng, "partial_ranges", []), dtype=np.int64).reshape(-1, 3))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_commit_prover_matches_single_process(tmp_path, world):
    """prover_sharded.ShardedCommits + ShardedQuotient over the oracle-backed engine, gloo: every rank produces the
    single-process proof bytes (GWC and SHPLONK) while committing only its share of the columns and contributing only
    its cosets of the extended domain to h"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import plonk_fixture as fxm
    from halo2_gpu_specific_b200.plonk import SeededRng
    from oracle import prover as PR
    port = _free_port()
    mp.spawn(_prover_worker, args=(world, port, str(tmp_path), False), nprocs=world, join=True)   # columns only
    fx = fxm.build(k=5, seed=11)
    oparams = PR.Params(5, 0x2B200B200B200B2001)
    opk = PR.keygen(oparams, fx["cs"], fx["fixed"], fx["mapping"])
    inst = [fx["instance"][0][:4]]
    want = PR.create_proof(oparams, opk, fx["advice"], inst, SeededRng(3)) + \
        PR.create_proof(oparams, opk, fx["advice"], inst, SeededRng(3), use_gwc=False)
    msms = []
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"proof_r{r}.npy")).tobytes()
        assert got == want, f"rank {r}"
        msms.append(np.load(os.path.join(str(tmp_path), f"msms_r{r}.npy")))
    total = np.sum(msms, axis=0)
    # GWC: 1 instance + 9 advice + 2 m + (2 + 3 + 1) z + 1 random + 4 h + 4 openings = 27 MSMs, divided, none duplicated
    assert total[0] == 27 and max(m[0] for m in msms) <= 27 // world + 7


@pytest.mark.parametrize("world,forced", [(2, True), (3, True), (4, None)])
def test_range_sharded_commits_match_single_process(tmp_path, world, forced):
    """SURVEY 8(e) option (i) inside the prover (prover_sharded.ShardedCommits.commit_by_point_range): with the range
    division forced for every full-width block, each rank multiplies its point range [r * ceil(n / N), ...) of every
    column (arithmetic.rs:426 part_len rule), the partials are all-gathered rank-major and summed -- and the proof
    bytes (GWC and SHPLONK) are still the single-process oracle prover's on every rank.  forced = None: four ranks with
    the cost model choosing (one-column blocks by range, the others by column).  Five ranks (uneven ranges of the 32
    points) were run offline with the same result."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import plonk_fixture as fxm
    from halo2_gpu_specific_b200 import parallel
    from halo2_gpu_specific_b200.plonk import SeededRng
    from oracle import prover as PR
    mp.spawn(_prover_worker, args=(world, _free_port(), str(tmp_path), forced), nprocs=world, join=True)
    fx = fxm.build(k=5, seed=11)
    oparams = PR.Params(5, 0x2B200B200B200B2001)
    opk = PR.keygen(oparams, fx["cs"], fx["fixed"], fx["mapping"])
    inst = [fx["instance"][0][:4]]
    want = PR.create_proof(oparams, opk, fx["advice"], inst, SeededRng(3)) + \
        PR.create_proof(oparams, opk, fx["advice"], inst, SeededRng(3), use_gwc=False)
    for r in range(world):
        assert np.load(os.path.join(str(tmp_path), f"proof_r{r}.npy")).tobytes() == want, f"rank {r}"
        ranges = np.load(os.path.join(str(tmp_path), f"ranges_r{r}.npy"))
        # the instance column, the random polynomial, the h pieces and the multiopen witnesses went by point range, and
        # every call used this rank's range of the 32 points
        assert {(int(a), int(b)) for a, b, _ in ranges} == {parallel.shard_range(32, world, r)}
        if forced:
            assert len(ranges) >= 8 and sum(int(c) for _, _, c in ranges) >= 1 + 1 + 4 + 4
        else:                                # instance + random polynomial (GWC), + SHPLONK's one-column blocks
            assert len(ranges) >= 4 and all(int(c) == 1 for _, _, c in ranges)


def test_range_division_cost_model():
    """_by_point_range: one column on several ranks goes by point range at proof sizes, blocks that divide evenly stay
    column-parallel, and bounded commits (their bound check lives in the batch call) never take the range path"""
    from halo2_gpu_specific_b200 import parallel
    from halo2_gpu_specific_b200.prover_sharded import ShardedCommits

    class E(ShardedCommits):
        msm_partials = sum_partials = None

        class domain:
            n = 1 << 22

    e = E()
    saved = parallel.world
    try:
        for world, expect in ((1, {}), (2, {1: True, 2: False, 3: True, 4: False, 24: False}),
                              (8, {1: True, 4: True, 8: False, 24: False, 64: False})):
            parallel.world = lambda w=world: (0, w)
            for count, by_range in expect.items():
                assert e._by_point_range(count, 254) is by_range, (world, count)
            assert e._by_point_range(1, 16) is False and e._by_point_range(0, 254) is False
        parallel.world = lambda: (0, 8)
        E.domain.n = 1 << 10            # small circuits: the fixed part of an MSM dominates, columns stay whole ... except
        assert e._by_point_range(4, 254) is False and e._by_point_range(1, 254) is True    # ... when 7 ranks would idle
    finally:
        parallel.world = saved


def test_range_division_cost_model_follows_the_committed_sweep():
    """ShardedCommits._msm_ms, the time model behind the choice between the two divisions, against the measured
    single-GPU MSM sweep committed with the round's bench line (profiles/r2_bench_1gpu_final.json, strong_scaling):
    within 5 % at every size from 2^18 to 2^26"""
    import json
    from halo2_gpu_specific_b200.prover_sharded import ShardedCommits
    line = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_1gpu_final.json")))
    sizes = line["strong_scaling"]["sizes"]
    assert len(sizes) >= 5
    for name, rec in sizes.items():
        n = 1 << int(name.split("^")[1])
        assert abs(ShardedCommits._msm_ms(n) / rec["ms"] - 1) < 0.05, name


def _rng_worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from halo2_gpu_specific_b200 import prover_sharded as ps
    from halo2_gpu_specific_b200._lib import B2Error
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = ps.synchronized_rng()
    np.save(os.path.join(outdir, f"rng_r{rank}.npy"), np.concatenate([rng.fr_vec(5).reshape(-1), rng.u64_vec(3), rng.u16_vec(2)]))
    ps.assert_ranks_agree(b"same bytes on every rank")
    raised = False
    try:
        ps.assert_ranks_agree(b"rank %d" % rank)
    except B2Error:
        raised = True
    np.save(os.path.join(outdir, f"raised_r{rank}.npy"), np.array([raised]))
    dist.barrier()
    dist.destroy_process_group()


def test_synchronized_rng_and_divergence_check(tmp_path):
    """prover_sharded.synchronized_rng: rank 0's OS seed reaches every rank, so the BLAKE2b streams are equal;
    assert_ranks_agree raises on EVERY rank when the ranks' bytes differ (what a per-rank OsRng would cause)"""
    world = 2
    mp.spawn(_rng_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = (np.load(tmp_path / f"rng_r{r}.npy") for r in range(world))
    assert np.array_equal(a, b) and a.any()
    assert all(np.load(tmp_path / f"raised_r{r}.npy")[0] for r in range(world))


def test_blake2b_rng_is_reproducible_and_uniform_below_r():
    from halo2_gpu_specific_b200 import _fr
    from halo2_gpu_specific_b200.plonk import Blake2bRng
    a, b = Blake2bRng(b"\x07" * 32), Blake2bRng(b"\x07" * 32)
    va = a.fr_vec(64)
    assert np.array_equal(va, b.fr_vec(64)) and not np.array_equal(va, Blake2bRng(b"\x08" * 32).fr_vec(64))
    vals = [_fr.from_mont(v) for v in va]
    assert all(0 <= v < _fr.R_MOD for v in vals) and len(set(vals)) == 64
    assert max(vals).bit_length() == 254          # the top third of Fr is reachable (253-bit sampling never gets there)
    assert Blake2bRng().fr_vec(0).shape == (0, 4)
