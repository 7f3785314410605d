"""Regenerates tests/golden/proof_digest.json: sha256 of the oracle's proof for the shared fixture under a fixed RNG
(tests/test_oracle_prover.py).  The digest is produced by this repository's oracle, not by the Rust reference
(which cannot be built here); it freezes transcript order, RNG order and encodings against drift.
Run from the repository root: python tests/golden/make_proof_digest.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import plonk_fixture as fxm  # noqa: E402
from halo2_gpu_specific_b200.plonk import SeededRng  # noqa: E402
from oracle import prover as PR  # noqa: E402

S_TOXIC = 0x2B200B200B200B200B200B200B200B2001
fx = fxm.build(k=5, seed=11)
params = PR.Params(5, S_TOXIC)
pk = PR.keygen(params, fx["cs"], fx["fixed"], fx["mapping"])
proof = PR.create_proof(params, pk, fx["advice"], [fx["instance"][0][:4]], SeededRng(1))
assert PR.verify_proof(params, pk.vk, [fx["instance"][0][:4]], proof)
shplonk = PR.create_proof(params, pk, fx["advice"], [fx["instance"][0][:4]], SeededRng(1), use_gwc=False)
assert PR.verify_proof(params, pk.vk, [fx["instance"][0][:4]], shplonk, use_gwc=False)
out = {"k5_seed11_rng1_sha256": hashlib.sha256(proof).hexdigest(), "proof_bytes": len(proof),
       "k5_seed11_rng1_shplonk_sha256": hashlib.sha256(shplonk).hexdigest(), "shplonk_proof_bytes": len(shplonk)}
json.dump(out, open(os.path.join(HERE, "proof_digest.json"), "w"), indent=1)
print(out)
