"""GPU parity for the cluster layer (csrc/cluster.cu): the sharded server on 1..8 GPUs of one process behind the reference's
two calls, Server::setup (server.rs:103) and Server::respond (server.rs:184).  Respond runs on row blocks of D by default and on
column slices with CHPIR_CLUSTER_SHARD=cols; both cuts are held to the same bytes.

Bit-exact: the complete hint and every response of an n-GPU cluster must equal the single-GPU bytes and the CPU oracle's bytes
("1-GPU vs 2/4/8-GPU outputs must be byte-identical").  Tests parametrised over the cluster size skip the sizes
this box cannot host; the one-GPU cluster exercises the same orchestration (slice plan, gather kernels, copy paths) on any box."""
import os
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import chalametpir_b200 as cp
from oracle import oracle as O
from conftest import make_db

SEED = bytes((5 * i + 1) & 0xFF for i in range(32))
SIZES = [1, 2, 3, 4, 8]
CUTS = ["rows", "cols"]


class cut_env:
    """CHPIR_CLUSTER_SHARD for the setups inside the block (read by the library at every cluster setup)."""

    def __init__(self, cut):
        self.cut = cut

    def __enter__(self):
        os.environ["CHPIR_CLUSTER_SHARD"] = self.cut

    def __exit__(self, *a):
        os.environ.pop("CHPIR_CLUSTER_SHARD", None)


def skip_redundant(n, cut):
    if n == 1 and cut == "cols":
        pytest.skip("a one-GPU cluster has one cut")


def rand_u32(rng, shape):
    return rng.integers(0, 2**32, size=shape, dtype=np.uint64).astype(np.uint32)


def qbytes(q):
    return O.matrix_to_bytes(np.asarray(q, dtype=np.uint32).reshape(1, -1))


def need(n):
    if cp.device_count() < n:
        pytest.skip(f"needs {n} GPUs, this box has {cp.device_count()}")


def exact_respond(D, q):
    """q . D mod 2^32 without overflow (16-bit halves of q)."""
    q64, D64 = q.astype(np.uint64), D.astype(np.uint64)
    lo, hi = q64 & 0xFFFF, q64 >> 16
    return (((lo @ D64) + (((hi @ D64) & 0xFFFF) << 16)) & 0xFFFFFFFF).astype(np.uint32)


@pytest.mark.parametrize("cut", CUTS)
@pytest.mark.parametrize("n", SIZES)
def test_cluster_hint_and_responses_equal_oracle_and_single_gpu(n, cut):
    need(n)
    skip_redundant(n, cut)
    rng = np.random.default_rng(100 + n)
    K, N, b, lwe = 4099, 133, 10, 150  # ragged K (not a multiple of 4), column count not divisible by n
    D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
    osrv, ohint = O.Server.setup_from_matrix(SEED, D, b, lwe_rows=lwe)
    single, shint = cp.Server.setup_from_matrix(SEED, D, b, lwe_rows=lwe)
    assert shint == ohint
    for gather in ("nccl", "p2p"):
        if n == 1 and gather == "nccl":
            continue
        os.environ["CHPIR_CLUSTER_GATHER"] = gather
        try:
            cl = cp.Cluster(n_gpus=n)
            with cut_env(cut):
                srv, hint = cp.ClusterServer.setup_from_matrix(cl, SEED, D, b, lwe_rows=lwe, batch_tc=1)
        finally:
            os.environ.pop("CHPIR_CLUSTER_GATHER", None)
        info = srv.get_info()
        assert info["n_gpus"] == n and info["cols_n"] == N and info["rows_k"] == K
        assert info["respond_by_rows"] == (1 if (n > 1 and cut == "rows") else 0)
        assert info["gather_uses_nccl"] == (1 if (n > 1 and gather == "nccl") else 0)
        assert hint == ohint, f"{n}-GPU hint ({gather} gather) differs from the oracle"
        for _ in range(3):  # a lone caller: the GEMV route
            q = qbytes(rand_u32(rng, K))
            r = srv.respond(q)
            assert r == osrv.respond(q) == single.respond(q)
        qs = [qbytes(rand_u32(rng, K)) for _ in range(21)]  # 21 >= 6: the tensor-core route, one partial M tile
        want = [osrv.respond(q) for q in qs]
        assert srv.respond_batch(qs) == want
        assert srv.respond_batch(qs[:3]) == want[:3]  # GEMV route through the batch call
        srv.close()
        cl.close()


@pytest.mark.parametrize("cut", CUTS)
@pytest.mark.parametrize("n", SIZES)
def test_cluster_concurrent_callers_are_coalesced_and_exact(n, cut):
    need(n)
    skip_redundant(n, cut)
    rng = np.random.default_rng(200 + n)
    K, N, b = 20011, 301, 9
    D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
    cl = cp.Cluster(n_gpus=n)
    with cut_env(cut):
        srv, _ = cp.ClusterServer.setup_from_matrix(cl, SEED, D, b, skip_hint=True, batch_tc=1, respond_coalesce=True)
    T, per = 24, 6
    qs = rand_u32(rng, (T * per, K))
    want = [O.matrix_to_bytes(exact_respond(D, q).reshape(1, -1)) for q in qs]
    got = [None] * (T * per)
    errs = []

    def work(t):
        try:
            for j in range(per):
                i = t * per + j
                got[i] = srv.respond(qbytes(qs[i]))
        except Exception as ex:  # pragma: no cover
            errs.append(ex)

    ths = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    assert not errs, errs
    assert got == want
    info = srv.get_info()
    assert info["queries"] == T * per and info["batches"] <= T * per
    srv.close()
    cl.close()


@pytest.mark.parametrize("cut", CUTS)
@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("b", [4, 9, 14])
def test_cluster_device_resident_paths(n, b, cut):
    """chpir_cluster_server_respond_device: query slices resident on the ranks' GPUs exactly as the PCIe ingest leaves them, result
    gathered in rank 0's HBM; GEMV route (copy-engine all-gather and the pull kernel) and tensor-core route."""
    need(n)
    skip_redundant(n, cut)
    import torch

    rng = np.random.default_rng(300 + 10 * n + b)
    K, N = 9001, 95
    D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
    cl = cp.Cluster(n_gpus=n)
    with cut_env(cut):
        srv, _ = cp.ClusterServer.setup_from_matrix(cl, SEED, D, b, skip_hint=True, batch_tc=1)
    nq = 171 if cut == "rows" else 71  # not a multiple of the GEMV chunk (32) / more than one tensor-core tile, the last one partial
    q = rand_u32(rng, (nq, K))
    want = np.stack([exact_respond(D, row) for row in q])
    ks = srv.k_pitch
    slices = []
    for r in range(n):
        pl = srv.plan(r)
        t = torch.full((nq, ks), 0x5A5A5A5A, dtype=torch.int32, device=f"cuda:{cl.devices[r]}")  # padding words must never be read as data
        if pl["k_count"]:
            t[:, : pl["k_count"]] = torch.from_numpy(q[:, pl["k_begin"] : pl["k_begin"] + pl["k_count"]].view(np.int32).copy()).to(t.device)
        slices.append(t)
    out = torch.zeros((nq, N), dtype=torch.int32, device=f"cuda:{cl.devices[0]}")
    for d in cl.devices:
        torch.cuda.synchronize(d)
    for mode, env in ((cp.RESPOND_GEMV, None), (cp.RESPOND_GEMV, "kernel"), (cp.RESPOND_TC, None)):
        out.fill_(-1)
        torch.cuda.synchronize(cl.devices[0])
        if env:
            os.environ["CHPIR_CLUSTER_QGATHER"] = env
        try:
            ms = srv.respond_device([t.data_ptr() for t in slices], nq, out.data_ptr(), mode=mode, repeats=2)
        finally:
            os.environ.pop("CHPIR_CLUSTER_QGATHER", None)
        assert ms > 0
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want), (n, b, mode, env)
    srv.close()
    cl.close()


@pytest.mark.parametrize("db_encode", ["host", "device"])
@pytest.mark.parametrize("n", [1, 2, 3, 8])
@pytest.mark.parametrize("arity", [3, 4])
def test_cluster_setup_from_db_pir_round(n, arity, db_encode):
    """Server::setup(seed, db) on the cluster -> same hint and filter bytes as the single-GPU call and as the oracle; the oracle's client
    recovers every queried value from the cluster's responses (integrations/src/test_pir.rs:12-142).  db_encode = device: every GPU
    builds its own columns of D in its HBM from one host-side peeling (csrc/device_fill.cuh), the row blocks are cut over NVLink."""
    need(n)
    db = make_db(3000, seed=40 + arity)
    cl = cp.Cluster(n_gpus=n)
    srv, hint, fbytes = cp.ClusterServer.setup(cl, SEED, db, arity, filter_seed_rng=21, lwe_rows=200, db_encode=db_encode)
    s1, h1, f1 = cp.Server.setup(SEED, db, arity, filter_seed_rng=21, lwe_rows=200)
    assert hint == h1 and fbytes == f1
    b = O.find_mat_elem_bit_len(len(db))
    D, fo = O.from_kv_database(db, b, arity, rng_seed=21)
    assert fbytes == fo.to_bytes()
    assert hint == O.Server.setup_from_matrix(SEED, D, b, lwe_rows=200)[1]
    client = O.Client.setup(SEED, hint, fbytes, lwe_rows=200)
    done = 0
    for key in list(db)[:12]:
        try:
            q = client.query(key)
        except O.OracleError:
            continue
        assert client.process_response(key, srv.respond(q)) == db[key]
        done += 1
    assert done >= 6
    srv.close()
    cl.close()


@pytest.mark.parametrize("cut", CUTS)
@pytest.mark.parametrize("n", [1, 2, 4])
def test_cluster_tiny_matrix_with_empty_query_slices(n, cut):
    """K smaller than the slice pitch: the last ranks ingest no query words at all (and, in the row cut, hold only zero rows); N barely
    covers the ranks."""
    need(n)
    skip_redundant(n, cut)
    rng = np.random.default_rng(400 + n)
    K, N, b = 7, max(n, 5), 11
    D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
    cl = cp.Cluster(n_gpus=n)
    with cut_env(cut):
        srv, hint = cp.ClusterServer.setup_from_matrix(cl, SEED, D, b, lwe_rows=20, batch_tc=1)
    osrv, ohint = O.Server.setup_from_matrix(SEED, D, b, lwe_rows=20)
    assert hint == ohint
    qs = [qbytes(rand_u32(rng, K)) for _ in range(9)]
    assert [srv.respond(q) for q in qs[:2]] == [osrv.respond(q) for q in qs[:2]]
    assert srv.respond_batch(qs) == [osrv.respond(q) for q in qs]
    srv.close()
    cl.close()


@pytest.mark.parametrize("cut", CUTS)
@pytest.mark.parametrize("n", [1, 2, 3, 8])
def test_cluster_page_locked_queries_are_fetched_by_the_gpus(n, cut):
    """Queries in chpir_host_alloc memory (any 4-byte alignment) are moved with one call per GPU and batch (cudaMemcpyBatchAsync, or the
    pull kernel with CHPIR_CLUSTER_INGEST=pull) instead of one DMA per query and GPU; pageable queries in the same batches take the
    DMA route; CHPIR_CLUSTER_INGEST=dma is the per-query route for everything.  Same bytes always."""
    need(n)
    rng = np.random.default_rng(800 + n)
    K, N, b = 30011, 203, 9
    D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
    cl = cp.Cluster(n_gpus=n)
    with cut_env(cut):
        srv, _ = cp.ClusterServer.setup_from_matrix(cl, SEED, D, b, skip_hint=True, batch_tc=1, respond_coalesce=True)
    Q = 40
    qs = rand_u32(rng, (Q, K))
    want = [O.matrix_to_bytes(exact_respond(D, q).reshape(1, -1)) for q in qs]
    qlen, rlen = 8 + 4 * K, 8 + 4 * N
    pad = 0  # qlen = 4 (mod 16): query i starts 4 * i bytes past a 16-byte boundary, which exercises the 16-, 8- and 4-byte load paths
    assert qlen % 16 == 4
    q_pin, r_pin = cp.PinnedBuffer(Q * (qlen + pad) + 64), cp.PinnedBuffer(Q * rlen)
    ptrs = []
    for i in range(Q):
        off = i * (qlen + pad)
        q_pin.array[off: off + qlen] = np.frombuffer(qbytes(qs[i]), dtype=np.uint8)
        ptrs.append(q_pin.ptr + off)
    assert {p % 16 for p in ptrs} >= {0, 4, 8, 12}
    # a lone caller (GEMV route), then 16 concurrent native callers (tensor-core route)
    assert srv.respond_into(ptrs[1], qlen, r_pin.ptr, rlen) == rlen
    assert r_pin.array[:rlen].tobytes() == want[1]
    before = srv.get_info()["pulled_queries"]
    assert before == 1
    srv.respond_concurrent(ptrs, qlen, 3 * Q, r_pin.ptr, rlen, 16)
    got = [r_pin.array[i * rlen: (i + 1) * rlen].tobytes() for i in range(Q)]
    assert got == want
    info = srv.get_info()
    assert info["pulled_queries"] == before + 3 * Q
    # pageable callers (Python bytes) beside page-locked ones
    errs, out = [], {}

    def pageable(i):
        try:
            out[i] = srv.respond(qbytes(qs[i]))
        except Exception as ex:  # pragma: no cover
            errs.append(ex)

    ths = [threading.Thread(target=pageable, args=(i,)) for i in range(8)]
    for th in ths:
        th.start()
    srv.respond_concurrent(ptrs[8:], qlen, Q - 8, r_pin.ptr + 8 * rlen, rlen, 8)
    for th in ths:
        th.join()
    assert not errs and [out[i] for i in range(8)] == want[:8]
    assert [r_pin.array[i * rlen: (i + 1) * rlen].tobytes() for i in range(8, Q)] == want[8:]
    # the batch call and the switch: the pull kernel instead of the copy engines' batch call, then per-query DMA
    assert srv.respond_batch([qbytes(q) for q in qs[:9]]) == want[:9]
    for route in ("pull", "dma"):
        os.environ["CHPIR_CLUSTER_INGEST"] = route
        try:
            r_pin.array[:] = 0
            pulled = srv.get_info()["pulled_queries"]
            srv.respond_concurrent(ptrs, qlen, Q, r_pin.ptr, rlen, 8)
            assert srv.get_info()["pulled_queries"] == pulled + (Q if route == "pull" else 0)
            assert [r_pin.array[i * rlen: (i + 1) * rlen].tobytes() for i in range(Q)] == want
        finally:
            os.environ.pop("CHPIR_CLUSTER_INGEST", None)
    srv.close()
    cl.close()
    q_pin.close()
    r_pin.close()


@pytest.mark.parametrize("n", [1, 2])
def test_cluster_error_behaviour_matches_reference(n):
    need(n)
    rng = np.random.default_rng(500 + n)
    K, N = 100, 20
    cl = cp.Cluster(n_gpus=n)
    srv, _ = cp.ClusterServer.setup_from_matrix(cl, SEED, rng.integers(0, 512, size=(K, N), dtype=np.uint32), 9, skip_hint=True)
    good = qbytes(rand_u32(rng, K))
    cases = {
        b"": "FailedToDeserializeMatrixFromBytes",
        good[:8]: "FailedToDeserializeMatrixFromBytes",
        good[:-1]: "FailedToDeserializeMatrixFromBytes",
        bytes(8) + bytes(4): "FailedToDeserializeMatrixFromBytes",
        qbytes(rand_u32(rng, K + 1)): "IncompatibleDimensionForRowVectorTransposedMatrixMultiplication",
    }
    for bad, variant in cases.items():
        with pytest.raises(cp.ChalametPIRError) as e:
            srv.respond(bad)
        assert e.value.variant == variant
    assert srv.respond(good)
    with pytest.raises(cp.ChalametPIRError) as e:
        cp.ClusterServer.setup(cl, SEED, {}, 3)
    assert e.value.variant == "EmptyKVDatabase"
    with pytest.raises(cp.ChalametPIRError) as e:
        cp.ClusterServer.setup(cl, SEED, make_db(8), 5)
    assert e.value.variant == "UnsupportedArityForBinaryFuseFilter"
    with pytest.raises(cp.ChalametPIRError) as e:  # the cluster slices: the caller may not
        cp.ClusterServer.setup_from_matrix(cl, SEED, np.ones((8, 8), np.uint32), 9, col_begin=1, col_count=2)
    assert e.value.variant == "InvalidArgument"
    srv.close()
    cl.close()


@pytest.mark.parametrize("cut", CUTS)
@pytest.mark.parametrize("n", [1, 2, 3])
def test_cluster_save_and_load(n, cut, tmp_path):
    need(n)
    skip_redundant(n, cut)
    rng = np.random.default_rng(600 + n)
    K, N, b = 3001, 77, 10
    D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
    cl = cp.Cluster(n_gpus=n)
    with cut_env(cut):
        srv, _ = cp.ClusterServer.setup_from_matrix(cl, SEED, D, b, skip_hint=True)
    qs = [qbytes(rand_u32(rng, K)) for _ in range(8)]
    want = [srv.respond(q) for q in qs]
    prefix = str(tmp_path / "srv")
    srv.save(prefix)
    srv.close()
    back = cp.ClusterServer.load(cl, prefix, batch_tc=1)  # the files, not the environment, decide the cut
    assert back.get_info()["respond_by_rows"] == (1 if (n > 1 and cut == "rows") else 0)
    assert back.get_info()["rows_k"] == K and back.get_info()["cols_n"] == N
    assert [back.respond(q) for q in qs[:2]] == want[:2]
    assert back.respond_batch(qs) == want  # tensor-core route on planes rebuilt from the packed rows
    back.close()
    if n > 1:  # a cluster of another size must refuse the files
        cl1 = cp.Cluster(n_gpus=1)
        with pytest.raises(cp.ChalametPIRError):
            cp.ClusterServer.load(cl1, prefix)
        cl1.close()
    cl.close()


def test_two_servers_on_two_gpus_from_one_thread():
    """the ring kernel's shared-memory attribute is per device; one thread serving device 0 then device 1 must work."""
    need(2)
    rng = np.random.default_rng(700)
    K, N, b = 30000, 940, 9
    D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
    s0, _ = cp.Server.setup_from_matrix(SEED, D, b, skip_hint=True, device=0)
    s1, _ = cp.Server.setup_from_matrix(SEED, D, b, skip_hint=True, device=1)
    q = rand_u32(rng, K)
    want = O.matrix_to_bytes(exact_respond(D, q).reshape(1, -1))
    for _ in range(2):
        assert s0.respond(qbytes(q)) == want
        assert s1.respond(qbytes(q)) == want


@pytest.mark.parametrize("n", [2, 8])
def test_full_size_cluster_respond_equals_single_gpu_bytes(n):
    """BASELINE.json configs[2] shape (2^20 entries: K = 1 179 648, N = 940, b = 9): the n-GPU response bytes equal the single-GPU bytes
    and exact dot products on columns from every rank's slice."""
    need(n)
    import torch

    K, N, b = 1179648, 940, 9
    g = torch.Generator(device="cuda:0")
    g.manual_seed(5)
    Dt = torch.randint(0, 1 << b, (K, N), dtype=torch.int32, device="cuda:0", generator=g)
    single, _ = cp.Server.setup_from_device_matrix(SEED, Dt.data_ptr(), K, N, b, skip_hint=True, batch_tc=1)
    cl = cp.Cluster(n_gpus=n)
    slices = []
    for r in range(n):
        pl = cp.cluster_plan(n, r, K, N)
        slices.append(Dt[:, pl["col_begin"] : pl["col_begin"] + pl["col_count"]].contiguous().to(f"cuda:{cl.devices[r]}"))
    for d in cl.devices:
        torch.cuda.synchronize(d)
    srv, _ = cp.ClusterServer.setup_from_device_slices(cl, SEED, [s.data_ptr() for s in slices], K, N, b, skip_hint=True, batch_tc=1)
    rng = np.random.default_rng(8)
    qs = [qbytes(rand_u32(rng, K)) for _ in range(9)]
    want = [single.respond(q) for q in qs]
    assert [srv.respond(q) for q in qs[:2]] == want[:2]  # GEMV route
    assert srv.respond_batch(qs) == want                 # tensor-core route
    cols = sorted({cp.cluster_plan(n, r, K, N)["col_begin"] for r in range(n)} | {N - 1})
    Dc = Dt[:, cols].cpu().numpy().astype(np.uint32)
    q0 = np.frombuffer(qs[0], dtype="<u4")[2:]
    assert np.array_equal(np.frombuffer(want[0], dtype="<u4")[2:][cols], exact_respond(Dc, q0))
    srv.close()
    cl.close()


@pytest.mark.parametrize("mode", ["shared", "per_rank"])
@pytest.mark.parametrize("n", [2, 4, 8])
def test_cluster_walks_one_xof_chain_and_forwards_the_panels(n, mode, monkeypatch):
    """A = generate_from_seed(lwe, K, seed) for n GPUs from ONE host chain: the leader's uploader forwards every 128-row panel over
    NVLink into the other ranks' rings (csrc/host_pipe.cuh start_mirror).  Four panels through two-panel rings (slots reused on the
    leader and on the mirrors), then rings as deep as A that stay behind as the ctx cache, then a setup served from those caches:
    the hint must be the oracle's every time, as with one chain per rank (CHPIR_CLUSTER_XOF=per_rank)."""
    need(n)
    monkeypatch.setenv("CHPIR_CLUSTER_XOF", mode)
    rng = np.random.default_rng(7 * n)
    K, N, b, lwe = 3001, 97, 9, 400
    D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
    ohint = O.Server.setup_from_matrix(SEED, D, b, lwe_rows=lwe)[1]
    cl = cp.Cluster(n_gpus=n)
    try:
        for a_cache in (False, True, True):
            srv, hint = cp.ClusterServer.setup_from_matrix(cl, SEED, D, b, lwe_rows=lwe, a_cache=a_cache, host_chunk_rows=37)
            assert hint == ohint, (mode, a_cache)
            srv.close()
        assert all(x == lwe * K * 4 for x in cl.drop_a_cache())
        db = make_db(1500, seed=n, val_len=(1, 90))
        srv, hint, fbytes = cp.ClusterServer.setup(cl, SEED, db, 3, filter_seed_rng=3, lwe_rows=300)  # the chain starts before D exists
        s1, h1, f1 = cp.Server.setup(SEED, db, 3, filter_seed_rng=3, lwe_rows=300)
        assert hint == h1 and fbytes == f1
        srv.close()
        s1.close()
    finally:
        cl.close()


@pytest.mark.parametrize("n", [1, 2, 8])
def test_plain_c_program_over_the_header_runs_setup_and_respond(n):
    """tests/c/dropin_test.c: includes include/chalamet_b200.h as C11, links libchalamet_b200.so, runs Server::setup -> Server::respond
    through the cluster entry points and checks hint row 0, a response and the error variants with its own arithmetic."""
    need(n)
    import subprocess

    from conftest import ROOT

    exe = os.path.join(ROOT, "build", "dropin_test")
    assert os.path.exists(exe), "build() did not produce build/dropin_test"
    out = subprocess.run([exe, str(n)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "dropin_test ok" in out.stdout


@pytest.mark.parametrize("n", [1, 2])
def test_cpp_mirror_of_server_and_client_runs_the_reference_integration_tests(n):
    """tests/cpp/test_pir.cpp over include/chalamet_b200.hpp (Server::setup / Server::respond / Client::* with the reference's names and
    error variants): the reference's end-to-end tests for both arities (integrations/src/test_pir.rs), the error behaviour of
    setup / respond / query / process_response, and one Server shared by 16 concurrent tasks -- on n GPUs via $CHPIR_GPUS."""
    need(n)
    import subprocess

    from conftest import ROOT

    exe = os.path.join(ROOT, "build", "test_pir_cpp")
    assert os.path.exists(exe), "build() did not produce build/test_pir_cpp"
    out = subprocess.run([exe, "2", "13"], capture_output=True, text=True, timeout=600, env={**os.environ, "CHPIR_GPUS": str(n)})
    assert out.returncode == 0, out.stdout + out.stderr
    assert "test_pir_cpp ok" in out.stdout
