"""GPU client (chalametpir_client::Client, client.rs:21-283): the device arithmetic of query against the
oracle's client word for word, the reference's error behaviour, and complete PIR rounds GPU client <-> GPU server."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import chalametpir_b200 as cp
from oracle import oracle as O
from conftest import make_db

SEED = bytes(range(32))


def ternary(rng, n):
    return rng.choice(np.array([0, 1, 0xFFFFFFFF], dtype=np.uint32), size=n)


@pytest.mark.parametrize("arity", [3, 4])
@pytest.mark.parametrize("a_expand", ["device", "host"])
def test_query_bytes_match_the_oracle_client_for_given_randomness(arity, a_expand):
    db = make_db(1500, seed=arity, val_len=(1, 100))
    lwe = 200
    srv, hint, fbytes = cp.Server.setup(SEED, db, arity, filter_seed_rng=2, lwe_rows=lwe, a_expand="host")
    gc = cp.Client.setup(SEED, hint, fbytes, lwe_rows=lwe, a_expand=a_expand, host_chunk_rows=33)
    oc = O.Client.setup(SEED, hint, fbytes, lwe_rows=lwe)
    rng = np.random.default_rng(arity)
    done = 0
    for key in list(db)[:6]:
        s, e = ternary(rng, lwe), ternary(rng, gc.rows_k)
        try:
            want = oc.query_with(key, s, e)
        except O.OracleError as ex:
            assert ex.name == "ArithmeticOverflowAddingQueryIndicator"
            with pytest.raises(cp.ChalametPIRError) as g:
                gc.query_with(key, s, e)
            assert g.value.variant == "ArithmeticOverflowAddingQueryIndicator"
            continue
        got = gc.query_with(key, s, e)
        assert got == want  # b = s*A + e + indicator, wire format
        r = srv.respond(got)
        assert gc.process_response(key, r) == db[key] == oc.process_response(key, r)
        done += 1
    assert done >= 3
    # full-range (non-ternary) words exercise every bit of the wrapping arithmetic
    key = list(db)[10]
    s, e = rng.integers(0, 2**32, size=lwe, dtype=np.uint64).astype(np.uint32), rng.integers(0, 2**32, size=gc.rows_k, dtype=np.uint64).astype(np.uint32)
    try:
        want = oc.query_with(key, s, e)
        assert gc.query_with(key, s, e) == want
    except O.OracleError:
        pass


@pytest.mark.parametrize("arity", [3, 4])
def test_pir_round_gpu_client_gpu_server(arity):  # integrations/src/test_pir.rs:12-142
    db = make_db(3000, seed=arity + 40, val_len=(1, 300))
    seed = bytes(random.Random(arity).randbytes(32))
    srv, hint, fbytes = cp.Server.setup(seed, db, arity, a_expand="host", db_encode="device")
    client = cp.Client.setup(seed, hint, fbytes, a_expand="host")
    assert client.info()["pub_mat_a_bytes"] == 1774 * client.rows_k * 4
    recovered = 0
    for i, key in enumerate(random.Random(2).sample(list(db), 12)):
        try:
            q = client.query(key, rng_seed=i)
        except cp.ChalametPIRError as ex:
            assert ex.variant == "ArithmeticOverflowAddingQueryIndicator"
            continue
        assert len(q) == 8 + 4 * client.rows_k
        assert client.process_response(key, srv.respond(q)) == db[key]
        recovered += 1
    assert recovered >= 9
    # a key that is not in the database decodes to garbage: the digest check fails (client.rs:255-259) or the row is not decodable
    q = client.query(b"definitely-not-a-key", rng_seed=99)
    with pytest.raises(cp.ChalametPIRError) as e:
        client.process_response(b"definitely-not-a-key", srv.respond(q))
    assert e.value.variant in ("DecodedRowNotPrependedWithDigestOfKey", "RowNotDecodable")


def test_client_error_behaviour_matches_reference():
    db = make_db(200, seed=5, val_len=(1, 30))
    srv, hint, fbytes = cp.Server.setup(SEED, db, 3, lwe_rows=64, a_expand="host")
    with pytest.raises(cp.ChalametPIRError) as e:
        cp.Client.setup(SEED, hint, fbytes[:-1], lwe_rows=64)
    assert e.value.variant == "FailedToDeserializeFilterFromBytes"  # binary_fuse_filter.rs:498-500
    with pytest.raises(cp.ChalametPIRError) as e:
        cp.Client.setup(SEED, hint[:-4], fbytes, lwe_rows=64)
    assert e.value.variant == "FailedToDeserializeMatrixFromBytes"
    with pytest.raises(cp.ChalametPIRError) as e:
        cp.Client.setup(SEED, hint, fbytes, lwe_rows=65)
    assert e.value.variant == "InvalidHintMatrix"  # client.rs:47-49
    client = cp.Client.setup(SEED, hint, fbytes, lwe_rows=64)
    key = list(db)[0]
    with pytest.raises(cp.ChalametPIRError) as e:
        client.process_response(key, srv.respond(client.query(list(db)[1], rng_seed=1)))
    assert e.value.variant == "PendingQueryDoesNotExistForKey"  # client.rs:272-275
    for seed in range(1, 50):
        try:
            q = client.query(key, rng_seed=seed)
            break
        except cp.ChalametPIRError as ex:
            assert ex.variant == "ArithmeticOverflowAddingQueryIndicator"
    with pytest.raises(cp.ChalametPIRError) as e:
        client.query(key, rng_seed=3)
    assert e.value.variant == "PendingQueryExistsForKey"  # client.rs:97-99
    r = srv.respond(q)
    with pytest.raises(cp.ChalametPIRError) as e:
        client.process_response(key, r[:-4])
    assert e.value.variant == "FailedToDeserializeMatrixFromBytes"
    bad = O.matrix_to_bytes(np.zeros((1, client.cols_n + 1), np.uint32))
    with pytest.raises(cp.ChalametPIRError) as e:
        client.process_response(key, bad)
    assert e.value.variant == "InvalidResponseVector"  # client.rs:215-217
    assert client.process_response(key, r) == db[key]
    with pytest.raises(cp.ChalametPIRError) as e:
        client.process_response(key, r)
    assert e.value.variant == "PendingQueryDoesNotExistForKey"  # removed after processing, client.rs:269


def test_full_size_pir_round_recovers_every_queried_value():
    """BASELINE.json configs[2] end to end: 2^20 real entries x 1 kB, Server::setup (device row fill, host-pipelined A), GPU client
    (A resident, 8.4 GB), queries through Server::respond and the batched path, every queried value recovered byte for byte."""
    n = 1 << 20
    rs = np.random.default_rng(7)
    keys = rs.integers(0, 256, size=(n, 32), dtype=np.uint8)
    keys[:, :8] = np.arange(n, dtype="<u8").view(np.uint8).reshape(n, 8)
    vals = rs.integers(0, 256, size=(n, 1024), dtype=np.uint8)
    srv, hint, fbytes = cp.Server.setup_from_arrays(SEED, keys, vals, 3, a_expand="host", db_encode="device", batch_tc=1, respond_coalesce=True)
    assert len(hint) == 6_670_248 and len(fbytes) == 68  # README.md:33-36
    client = cp.Client.setup(SEED, hint, fbytes, a_expand="host")
    picks = [0, 1, 12345, n // 2, n - 1] + [int(x) for x in rs.integers(0, n, size=7)]
    queries = {}
    for j, i in enumerate(picks):
        key = keys[i].tobytes()
        if key in queries:
            continue
        try:
            queries[key] = (i, client.query(key, rng_seed=j))
        except cp.ChalametPIRError as ex:
            assert ex.variant == "ArithmeticOverflowAddingQueryIndicator"
    assert len(queries) >= 8
    assert all(len(q) == 4_718_600 for _, q in queries.values())
    items = list(queries.items())
    # half through Server::respond one by one, half through the batch entry point (tensor-core route)
    half = len(items) // 2
    resps = [srv.respond(q) for _, (_, q) in items[:half]] + srv.respond_batch([q for _, (_, q) in items[half:]])
    for (key, (i, _)), r in zip(items, resps):
        assert len(r) == 3_768
        assert client.process_response(key, r) == vals[i].tobytes()
    assert client.info()["last_query_kernel_ms"] < 10
