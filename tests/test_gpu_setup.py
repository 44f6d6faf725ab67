"""GPU parity for the offline path: A expansion (matrix.rs:541-558), hint GEMM (matrix.rs:1040-1059) and the complete
Server::setup (server.rs:103) against the CPU oracle, plus end-to-end value recovery through the oracle's client."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import chalametpir_b200 as cp
from oracle import oracle as O
from conftest import make_db

SEED = bytes(range(32))
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "vectors.json")


def rand_u32(rng, shape):
    return rng.integers(0, 2**32, size=shape, dtype=np.uint64).astype(np.uint32)


# ------------------------------------------------------------------ A expansion
@pytest.mark.parametrize("rows,cols", [(1, 1), (1, 41), (1, 42), (1, 43), (3, 14), (7, 1000), (64, 4099)])
def test_generate_from_seed_matches_oracle(rows, cols):
    seed = bytes(random.Random(rows * 1000 + cols).randbytes(32))
    assert np.array_equal(cp.generate_from_seed(rows, cols, seed), O.generate_from_seed(rows, cols, seed))


def test_generate_from_seed_deep_in_the_stream():
    # last rows of a 1774 x 20000 matrix: 142 MB into the XOF stream, several kernel launches with carried state
    rows, cols = 1774, 20000
    got = cp.generate_from_seed(rows, cols, SEED, row_begin=rows - 2, row_count=2)
    assert np.array_equal(got, O.generate_rows_from_seed(cols, SEED, rows - 2, 2))


# ------------------------------------------------------------------ hint GEMM
@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("m,k,n,b", [(1, 1, 1, 9), (37, 301, 53, 10), (128, 4096, 128, 9), (200, 1000, 940, 9), (1774, 2048, 100, 14), (130, 77, 17, 4)])
def test_matmul_matches_oracle(variant, m, k, n, b):
    rng = np.random.default_rng(m + k + n)
    A = rand_u32(rng, (m, k))
    B = rng.integers(0, 1 << b, size=(k, n), dtype=np.uint32)
    assert np.array_equal(cp.matmul(A, B, b, variant), O.matmul(A, B))


@pytest.mark.parametrize("variant", [0, 1])
def test_matmul_identity(variant):  # matrix.rs:1275-1317
    rng = np.random.default_rng(5)
    A = rand_u32(rng, (300, 257))
    assert np.array_equal(cp.matmul(A, np.eye(257, dtype=np.uint32), 4, variant), A)


@pytest.mark.parametrize("variant", [0, 1])
def test_matmul_wraps_instead_of_saturating(variant):
    """All-ones operands over a long K: every int32 partial sum overflows many times; the result must wrap mod 2^32."""
    m, k, n, b = 128, 50000, 64, 14
    A = np.full((m, k), 0xFFFFFFFF, dtype=np.uint32)
    B = np.full((k, n), (1 << b) - 1, dtype=np.uint32)
    want = np.uint32((0xFFFFFFFF * ((1 << b) - 1) * k) & 0xFFFFFFFF)
    got = cp.matmul(A, B, b, variant)
    assert np.all(got == want)
    rng = np.random.default_rng(6)
    A = rand_u32(rng, (m, k)) | 0x80808080
    B = rng.integers(0, 1 << b, size=(k, n), dtype=np.uint32) | 0x2080
    assert np.array_equal(cp.matmul(A, B, b, variant), O.matmul(A, B))


def test_matmul_dimension_errors():
    with pytest.raises(cp.ChalametPIRError) as e:
        cp.matmul(np.ones((2, 3), np.uint32), np.ones((4, 2), np.uint32), 9)
    assert e.value.variant == "IncompatibleDimensionForMatrixMultiplication"


def test_matmul_operand_width_is_checked_not_truncated():
    """the default call must work for full-range B (u32 SIMT kernel behind the tensor-core variant's 16-bit limit), and
    entries wider than the declared width are refused instead of being cut by the limb split."""
    rng = np.random.default_rng(77)
    A, B = rand_u32(rng, (33, 129)), rand_u32(rng, (129, 40))
    assert np.array_equal(cp.matmul(A, B), O.matmul(A, B))          # default width 32 -> exact for any operand
    B16 = rng.integers(0, 1 << 16, size=(129, 40), dtype=np.uint32)
    assert np.array_equal(cp.matmul(A, B16, 16), O.matmul(A, B16))  # widest tensor-core width
    with pytest.raises(cp.ChalametPIRError) as e:
        cp.matmul(A, B16 | 0x10000, 16)
    assert e.value.variant == "InvalidArgument"
    with pytest.raises(cp.ChalametPIRError) as e:
        cp.matmul(A, np.full((129, 40), 512, np.uint32), 9)
    assert e.value.variant == "InvalidArgument"


# ------------------------------------------------------------------ Server::setup end to end
@pytest.mark.parametrize("arity", [3, 4])
@pytest.mark.parametrize("variant", [0, 1])
def test_setup_and_respond_end_to_end(arity, variant):  # integrations/src/test_pir.rs:12-142
    db = make_db(3000, seed=arity * 7 + variant, val_len=(1, 300))
    seed = bytes(random.Random(arity).randbytes(32))
    srv, hint, fbytes = cp.Server.setup(seed, db, arity, filter_seed_rng=21, gemm_variant=variant)
    assert len(fbytes) == 68
    b = cp.find_mat_elem_bit_len(len(db))
    D, fb2 = cp.encode_kv_database(db, b, arity, filter_seed_rng=21)
    assert fb2 == fbytes
    osrv, ohint = O.Server.setup_from_matrix(seed, D, b)
    assert hint == ohint  # identical hint bytes (wire format header + 1774 x N u32)
    client = O.Client.setup(seed, hint, fbytes)
    keys = random.Random(1).sample(list(db), 10)
    answered = 0
    for key in keys:
        try:
            q = client.query(key)
        except O.OracleError as ex:
            assert ex.name == "ArithmeticOverflowAddingQueryIndicator"  # caller retries, test_pir.rs:66-70
            continue
        r = srv.respond(q)
        assert r == osrv.respond(q)
        assert client.process_response(key, r) == db[key]
        answered += 1
    assert answered >= 8
    t = srv.setup_timing()
    assert t["total_s"] > 0 and t["expand_a_s"] > 0 and t["gemm_s"] > 0


def test_setup_column_slices_interleave_to_the_full_hint():
    rng = np.random.default_rng(3)
    K, N, b, lwe = 3000, 100, 9, 200
    D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
    _, full = cp.Server.setup_from_matrix(SEED, D, b, lwe_rows=lwe)
    assert full == O.Server.setup_from_matrix(SEED, D, b, lwe_rows=lwe, want_server=False)[1]
    M = O.matrix_from_bytes(full)
    for world in (2, 8):
        bounds = [N * r // world for r in range(world + 1)]
        for r in range(world):
            _, part = cp.Server.setup_from_matrix(SEED, D, b, lwe_rows=lwe, col_begin=bounds[r], col_count=bounds[r + 1] - bounds[r])
            assert np.array_equal(O.matrix_from_bytes(part), M[:, bounds[r] : bounds[r + 1]])


def test_golden_fixtures_through_the_gpu_path():
    g = json.load(open(GOLDEN))
    seed = bytes.fromhex(g["seed"])
    for case in g["cases"]:
        rnd = random.Random(case["db_seed"])
        db = {}
        while len(db) < case["n"]:
            db[rnd.randbytes(rnd.randint(16, 32))] = rnd.randbytes(rnd.randint(1, case["max_val"]))
        srv, hint, fb = cp.Server.setup(seed, db, case["arity"], filter_seed_rng=case["filter_rng"], lwe_rows=case["lwe_rows"])
        assert fb.hex() == case["filter_params"]
        assert hashlib.sha256(hint).hexdigest() == case["hint_sha256"]
        K = case["shape"][0]
        q = np.frombuffer(O.turboshake128(b"q" + seed, 4 * K), dtype="<u4")
        assert hashlib.sha256(srv.respond(O.matrix_to_bytes(q[None, :]))).hexdigest() == case["response_sha256"]


# ------------------------------------------------------------------ host-pipelined A expansion (chpir_setup_opts.a_expand)
@pytest.mark.parametrize("K,N,b,lwe,chunk", [(3000, 100, 9, 200, 0), (1001, 37, 10, 300, 7), (4099, 64, 9, 129, 128), (517, 20, 14, 128, 1)])
def test_host_pipelined_setup_gives_identical_hint(K, N, b, lwe, chunk):
    """Same seed, same D: the hint is byte-identical whether the XOF chain is walked by a GPU warp or by a host core
    (4K is never a multiple of the 168-byte XOF rate here, chunks end mid-block, panels are ragged)."""
    rng = np.random.default_rng(K + N)
    D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
    srv_h, hint_h = cp.Server.setup_from_matrix(SEED, D, b, lwe_rows=lwe, a_expand="host", host_chunk_rows=chunk)
    srv_d, hint_d = cp.Server.setup_from_matrix(SEED, D, b, lwe_rows=lwe, a_expand="device")
    assert hint_h == hint_d
    assert hint_h == O.Server.setup_from_matrix(SEED, D, b, lwe_rows=lwe, want_server=False)[1]
    t = srv_h.setup_timing()
    assert t["xof_host_busy_s"] > 0 and srv_d.setup_timing()["xof_host_busy_s"] == 0
    q = rand_u32(rng, (1, K))
    assert srv_h.respond(O.matrix_to_bytes(q)) == srv_d.respond(O.matrix_to_bytes(q))


def test_host_pipelined_full_setup_end_to_end():
    db = make_db(2000, seed=77, val_len=(1, 200))
    seed = bytes(random.Random(9).randbytes(32))
    srv, hint, fbytes = cp.Server.setup(seed, db, 3, filter_seed_rng=5, a_expand="host")
    b = cp.find_mat_elem_bit_len(len(db))
    D, _ = cp.encode_kv_database(db, b, 3, filter_seed_rng=5)
    assert hint == O.Server.setup_from_matrix(seed, D, b, want_server=False)[1]
    client = O.Client.setup(seed, hint, fbytes)
    done = 0
    for key in list(db)[:6]:
        try:
            q = client.query(key)
        except O.OracleError:
            continue
        assert client.process_response(key, srv.respond(q)) == db[key]
        done += 1
    assert done >= 4


def test_device_and_host_expanders_agree_deep_in_the_stream():
    rows, cols = 1774, 20000
    got = cp.generate_from_seed(rows, cols, SEED, row_begin=rows - 2, row_count=2)
    assert np.array_equal(got, cp.host_generate_from_seed(rows, cols, SEED, row_begin=rows - 2, row_count=2))


def test_setup_from_arrays_equals_setup_from_dict():
    rng = np.random.default_rng(4)
    n = 1500
    keys = rng.integers(0, 256, size=(n, 24), dtype=np.uint8)
    keys[:, :4] = np.arange(n, dtype="<u4").view(np.uint8).reshape(n, 4)
    vals = rng.integers(0, 256, size=(n, 64), dtype=np.uint8)
    db = {keys[i].tobytes(): vals[i].tobytes() for i in range(n)}
    _, h1, f1 = cp.Server.setup(SEED, db, 3, filter_seed_rng=3, lwe_rows=160, a_expand="host", host_chunk_rows=16)
    _, h2, f2 = cp.Server.setup_from_arrays(SEED, keys, vals, 3, filter_seed_rng=3, lwe_rows=160)
    assert h1 == h2 and f1 == f2


def test_full_size_setup_hint_rows():
    """BASELINE.json configs[2] (2^20 entries, 3-wise): the whole Server::setup on the full shape, host-pipelined A.  The oracle
    cannot multiply 1774 x 1.18M x 940 in seconds, so the hint is checked on its first two and its LAST row (the tail of the
    8.4 GB XOF stream, walked independently by the oracle) against exact dot products on sampled columns."""
    import torch

    n, arity = 1 << 20, 3
    b = cp.find_mat_elem_bit_len(n)
    K, N = cp.db_matrix_shape(arity, n, 1024, b)
    g = torch.Generator(device="cuda").manual_seed(20)
    D = torch.randint(0, 1 << b, (K, N), dtype=torch.int32, device="cuda", generator=g)
    srv, hint = cp.Server.setup_from_device_matrix(SEED, D.data_ptr(), K, N, b, a_expand="host", batch_tc=2)
    H = O.matrix_from_bytes(hint)
    assert H.shape == (cp.LWE_DIMENSION, N)
    cols = [0, 1, 7, N // 2, N - 2, N - 1]
    Dc = D[:, cols].cpu().numpy().astype(np.uint64)
    for r0, nr in ((0, 2), (cp.LWE_DIMENSION - 1, 1)):
        a = O.generate_rows_from_seed(K, SEED, r0, nr).astype(np.uint64)
        lo, hi = a & 0xFFFF, a >> 16
        want = ((lo @ Dc) + (((hi @ Dc) & 0xFFFF) << 16)) & 0xFFFFFFFF
        assert np.array_equal(H[r0 : r0 + nr][:, cols].astype(np.uint64), want), r0
    t = srv.setup_timing()
    assert t["expand_a_s"] < 60
    srv.close()


# ------------------------------------------------------------------ row encoding + dependent fill on the GPU (db_encode = "device")
@pytest.mark.parametrize("arity", [3, 4])
@pytest.mark.parametrize("n,val_len", [(1, (5, 5)), (2, (0, 3)), (7, (1, 40)), (300, (0, 120)), (5000, (1, 64)), (20000, (900, 1024))])
def test_device_row_fill_is_byte_identical_to_host_encode(arity, n, val_len):
    """Matrix::from_kv_database (matrix.rs:687-755 / :819-894): D built in HBM wave by wave == D built by the host encoder ==
    the oracle's, for every legal element width, ragged and empty values included; filter parameters identical."""
    db = make_db(n, seed=n + arity, val_len=val_len)
    for b in sorted({4, 9, 10, 14, O.find_mat_elem_bit_len(n)}):
        Dd, fd = cp.encode_kv_database_device(db, b, arity, filter_seed_rng=n)
        Dh, fh = cp.encode_kv_database(db, b, arity, filter_seed_rng=n)
        assert fd == fh
        assert np.array_equal(Dd, Dh), (arity, n, b)
    if n <= 5000:
        Do, fo = O.from_kv_database(db, b, arity, rng_seed=n)
        assert fd == fo.to_bytes() and np.array_equal(Dd, Do)


def test_device_row_fill_with_the_multi_threaded_values_upload(monkeypatch):
    """The values reach HBM through four helper threads and page-locked bounce buffers once they are large (csrc/staged_upload.cuh);
    forced here on a 9 MB value blob whose last 4 MB chunk is ragged: D must not change."""
    db = make_db(9000, seed=77, val_len=(900, 1100))
    b = O.find_mat_elem_bit_len(len(db))
    D0, f0 = cp.encode_kv_database_device(db, b, 3, filter_seed_rng=5)
    monkeypatch.setenv("CHPIR_STAGE_MIN_BYTES", "1")
    for _ in range(2):  # the second run reuses the ctx's bounce buffers
        D1, f1 = cp.encode_kv_database_device(db, b, 3, filter_seed_rng=5)
        assert f0 == f1 and np.array_equal(D0, D1)
    Dh, fh = cp.encode_kv_database(db, b, 3, filter_seed_rng=5)
    assert fh == f0 and np.array_equal(Dh, D0)


@pytest.mark.parametrize("arity", [3, 4])
def test_setup_with_device_row_fill_end_to_end(arity):
    db = make_db(4000, seed=arity + 100, val_len=(1, 256))
    seed = bytes(random.Random(arity + 5).randbytes(32))
    srv, hint, fbytes = cp.Server.setup(seed, db, arity, filter_seed_rng=8, db_encode="device", a_expand="host")
    srv2, hint2, fbytes2 = cp.Server.setup(seed, db, arity, filter_seed_rng=8)
    assert hint == hint2 and fbytes == fbytes2
    client = O.Client.setup(seed, hint, fbytes)
    done = 0
    for key in list(db)[:8]:
        try:
            q = client.query(key)
        except O.OracleError:
            continue
        r = srv.respond(q)
        assert r == srv2.respond(q)
        assert client.process_response(key, r) == db[key]
        done += 1
    assert done >= 5
    assert srv.setup_timing()["device_encode_s"] > 0


# ------------------------------------------------------------------ A kept resident between setups (chpir_setup_opts.a_cache)
@pytest.mark.parametrize("a_expand", ["host", "device"])
def test_a_cache_reuses_a_for_a_new_database_and_gives_identical_hints(a_expand):
    """A depends on (seed, lwe_rows, K) only.  The first setup with a_cache fills the cache, a second one with ANOTHER database of
    the same K takes A from HBM (no XOF chain) and must produce exactly the hint an uncached setup produces; a different seed,
    LWE dimension or K is a miss and replaces the cached matrix."""
    cp.drop_a_cache()
    K, N, b, lwe = 2051, 70, 9, 300  # ragged: 3 panels (128 + 128 + 44 rows), 4K not a multiple of the XOF rate
    rng = np.random.default_rng(31)
    D1 = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
    D2 = rng.integers(0, 1 << b, size=(K, N + 9), dtype=np.uint32)
    want1 = O.Server.setup_from_matrix(SEED, D1, b, lwe_rows=lwe, want_server=False)[1]
    want2 = O.Server.setup_from_matrix(SEED, D2, b, lwe_rows=lwe, want_server=False)[1]
    s1, h1 = cp.Server.setup_from_matrix(SEED, D1, b, lwe_rows=lwe, a_expand=a_expand, a_cache=True)
    assert h1 == want1 and s1.setup_timing()["a_cache_hit"] == 0
    s2, h2 = cp.Server.setup_from_matrix(SEED, D2, b, lwe_rows=lwe, a_expand=a_expand, a_cache=True)
    t2 = s2.setup_timing()
    assert h2 == want2 and t2["a_cache_hit"] == 1 and t2["xof_host_busy_s"] == 0
    # the cached route also serves the other expansion mode (the cache is keyed by what A is, not by how it was made)
    other = "device" if a_expand == "host" else "host"
    s3, h3 = cp.Server.setup_from_matrix(SEED, D1, b, lwe_rows=lwe, a_expand=other, a_cache=True)
    assert h3 == want1 and s3.setup_timing()["a_cache_hit"] == 1
    # a_cache off: the cache is neither consulted nor touched
    s4, h4 = cp.Server.setup_from_matrix(SEED, D1, b, lwe_rows=lwe, a_expand=a_expand)
    assert h4 == want1 and s4.setup_timing()["a_cache_hit"] == 0
    # misses: another seed, another LWE dimension, another K
    seed2 = bytes(range(32, 64))
    s5, h5 = cp.Server.setup_from_matrix(seed2, D1, b, lwe_rows=lwe, a_expand=a_expand, a_cache=True)
    assert s5.setup_timing()["a_cache_hit"] == 0 and h5 == O.Server.setup_from_matrix(seed2, D1, b, lwe_rows=lwe, want_server=False)[1]
    s6, h6 = cp.Server.setup_from_matrix(seed2, D1, b, lwe_rows=lwe - 45, a_expand=a_expand, a_cache=True)
    assert s6.setup_timing()["a_cache_hit"] == 0 and h6 == O.Server.setup_from_matrix(seed2, D1, b, lwe_rows=lwe - 45, want_server=False)[1]
    s7, h7 = cp.Server.setup_from_matrix(seed2, D1[:-3], b, lwe_rows=lwe - 45, a_expand=a_expand, a_cache=True)
    assert s7.setup_timing()["a_cache_hit"] == 0 and h7 == O.Server.setup_from_matrix(seed2, D1[:-3], b, lwe_rows=lwe - 45, want_server=False)[1]
    s8, h8 = cp.Server.setup_from_matrix(seed2, D1[:-3], b, lwe_rows=lwe - 45, a_expand=a_expand, a_cache=True)
    assert s8.setup_timing()["a_cache_hit"] == 1 and h8 == h7
    assert cp.drop_a_cache() == (lwe - 45) * (K - 3) * 4
    assert cp.drop_a_cache() == 0
    # responses do not depend on any of this
    q = O.matrix_to_bytes(rand_u32(rng, (1, K)))
    assert s1.respond(q) == s3.respond(q) == s4.respond(q)


@pytest.mark.parametrize("db_encode", ["host", "device"])
def test_a_cache_full_setup_after_a_database_update(db_encode):
    """Server::setup(seed, db) twice with the same seed and a database of the same size but new values (the filter is rebuilt, K stays):
    the second call reuses A, and its hint still lets the client recover every queried value."""
    cp.drop_a_cache()
    seed = bytes(random.Random(19).randbytes(32))
    db1 = make_db(1800, seed=5, val_len=(8, 120))
    db2 = {k: bytes(reversed(v)) + b"!" for k, v in db1.items()}
    lwe = 256
    s1, h1, f1 = cp.Server.setup(seed, db1, 3, filter_seed_rng=3, a_expand="host", a_cache=True, lwe_rows=lwe, db_encode=db_encode)
    assert s1.setup_timing()["a_cache_hit"] == 0
    s2, h2, f2 = cp.Server.setup(seed, db2, 3, filter_seed_rng=4, a_expand="host", a_cache=True, lwe_rows=lwe, db_encode=db_encode)
    assert s2.setup_timing()["a_cache_hit"] == 1 and s2.rows_k == s1.rows_k
    s2u, h2u, f2u = cp.Server.setup(seed, db2, 3, filter_seed_rng=4, a_expand="host", lwe_rows=lwe, db_encode=db_encode)
    assert h2 == h2u and f2 == f2u
    client = O.Client.setup(seed, h2, f2, lwe_rows=lwe)
    done = 0
    for key in list(db2)[:6]:
        try:
            q = client.query(key)
        except O.OracleError:
            continue
        assert client.process_response(key, s2.respond(q)) == db2[key]
        done += 1
    assert done >= 4
    cp.drop_a_cache()
