"""Server::setup(seed, db) at full size, phase by phase (run with CHPIR_TRACE=1 for the library's own trace on stderr):
host row encoding / device row fill, then the same database "updated" with A taken from the ctx cache (a_cache)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chalametpir_b200 as cp  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << log2n
rs = np.random.default_rng(2024)
keys = rs.integers(0, 256, size=(n, 32), dtype=np.uint8)
keys[:, :8] = np.arange(n, dtype="<u8").view(np.uint8).reshape(n, 8)
vals = rs.integers(0, 256, size=(n, 1024), dtype=np.uint8)
vals2 = vals[::-1].copy()  # the "updated" database: same keys, other values
seed = bytes(range(32))


def run(label, v, **kw):
    print(f"--- {label}", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    srv, hint, fb = cp.Server.setup_from_arrays(seed, keys, v, 3, filter_seed_rng=7, batch_tc=2, a_expand="host", **kw)
    wall = time.perf_counter() - t0
    t = {k: round(x, 4) for k, x in srv.setup_timing().items()}
    print(json.dumps({"run": label, "wall_s": round(wall, 3), **t}), flush=True)
    del srv
    return hint, fb


h_host, f_host = run("host encode", vals)
h_dev, f_dev = run("device row fill", vals, db_encode="device")
assert h_host == h_dev and f_host == f_dev
h_fill, _ = run("device row fill, a_cache (fills the cache)", vals, db_encode="device", a_cache=True)
assert h_fill == h_host
h_hit, f_hit = run("updated database, device row fill, a_cache hit", vals2, db_encode="device", a_cache=True)
h_hit2, f_hit2 = run("updated database, host encode, a_cache hit", vals2, a_cache=True)
assert h_hit == h_hit2 and f_hit == f_hit2
print("freed", cp.drop_a_cache(), "bytes of cached A")
h_ref, f_ref = run("updated database, device row fill, no cache", vals2, db_encode="device")
assert h_ref == h_hit and f_ref == f_hit
print("hints identical with and without the cache: ok")
