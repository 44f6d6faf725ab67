"""Generates tests/golden/vectors.json from the CPU oracle.

The reference is Rust and cannot be built or imported in this image, and it ships no known-answer vectors for this
path, so these fixtures are REGRESSION vectors of the oracle (itself pinned on RFC 9861 + the reference's properties),
not outputs of the reference.  They freeze the byte-level behaviour that the GPU path is compared against, so an
accidental change to either side shows up as a diff.  Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

seed = bytes.fromhex("000102030405060708090a0b0c0d0e0f101112131415161718191a1b1c1d1e1f")
out = {
    "seed": seed.hex(),
    "xof_first_64": O.turboshake128(seed, 64).hex(),
    "xof_skip": 123456 * 168 + 40,
    "xof_at_skip": O.turboshake128(seed, 32, skip=123456 * 168 + 40).hex(),
    "cases": [],
}
for n, arity, max_val, lwe_rows, db_seed, filter_rng in [(700, 3, 64, 48, 1, 3), (700, 4, 64, 48, 2, 4), (4096, 3, 200, 32, 3, 5), (37, 4, 9, 16, 4, 6)]:
    rnd = random.Random(db_seed)
    db = {}
    while len(db) < n:
        db[rnd.randbytes(rnd.randint(16, 32))] = rnd.randbytes(rnd.randint(1, max_val))
    srv, hint, fb = O.Server.setup(seed, db, arity, rng_seed=filter_rng, lwe_rows=lwe_rows)
    q = np.frombuffer(O.turboshake128(b"q" + seed, 4 * srv.D.shape[0]), dtype="<u4")
    resp = srv.respond(O.matrix_to_bytes(q[None, :]))
    out["cases"].append(
        {
            "n": n, "arity": arity, "max_val": max_val, "lwe_rows": lwe_rows, "db_seed": db_seed, "filter_rng": filter_rng,
            "filter_params": fb.hex(), "shape": list(srv.D.shape), "bit_len": srv.filter.mat_elem_bit_len,
            "d_sha256": hashlib.sha256(srv.D.tobytes()).hexdigest(), "hint_sha256": hashlib.sha256(hint).hexdigest(),
            "query_seed_hex": "", "response_sha256": hashlib.sha256(resp).hexdigest(), "response_hex": resp.hex() if len(resp) < 2000 else "",
        }
    )
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "vectors.json"), "w"), indent=1)
print("wrote", len(out["cases"]), "cases")
