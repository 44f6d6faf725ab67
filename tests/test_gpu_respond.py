"""GPU parity for the online path: chpir_server_respond (C ABI) against the CPU oracle's restatement of
Server::respond (server.rs:184-190 -> matrix.rs:328-485, :973-1010).  Bit-exact: integer work."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import chalametpir_b200 as cp
from oracle import oracle as O
from conftest import make_db

SEED = bytes(range(32))


def rand_u32(rng, shape):
    return rng.integers(0, 2**32, size=shape, dtype=np.uint64).astype(np.uint32)


def qbytes(q):
    return O.matrix_to_bytes(np.asarray(q, dtype=np.uint32).reshape(1, -1))


def oracle_respond(D, b, q):
    srv, _ = O.Server.setup_from_matrix(SEED, D, b, want_hint=False)
    return srv.respond(qbytes(q))


@pytest.mark.parametrize("b", range(4, 15))
def test_respond_matches_oracle_all_bit_lengths(b):
    rng = np.random.default_rng(b)
    for K, N in [(1, 1), (5, 3), (int(rng.integers(2, 3000)), int(rng.integers(1, 300))), (4099, 941), (70001, 37)]:
        D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
        srv, hint = cp.Server.setup_from_matrix(SEED, D, b, skip_hint=True)
        assert hint is None and srv.rows_k == K and srv.cols_n == N
        for _ in range(2):
            q = rand_u32(rng, K)
            got = srv.respond(qbytes(q))
            assert got == oracle_respond(D, b, q), (b, K, N)
        srv.close()


def test_respond_packed_ones_is_sum_of_query():  # the reference's own GEMV test, matrix.rs:1319-1376
    rng = np.random.default_rng(7)
    for b in (4, 9, 10, 14):
        K, N = int(rng.integers(1, 1025)), int(rng.integers(1, 1025))
        srv, _ = cp.Server.setup_from_matrix(SEED, np.ones((K, N), np.uint32), b, skip_hint=True)
        q = rand_u32(rng, K)
        r = O.matrix_from_bytes(srv.respond(qbytes(q)))
        assert r.shape == (1, N) and np.all(r == np.uint32(int(q.astype(np.uint64).sum()) & 0xFFFFFFFF))


def test_values_are_masked_like_row_wise_compress():  # matrix.rs:121 `& mat_elem_mask`
    rng = np.random.default_rng(8)
    D = rand_u32(rng, (300, 50))
    srv, _ = cp.Server.setup_from_matrix(SEED, D, 9, skip_hint=True)
    q = rand_u32(rng, 300)
    assert srv.respond(qbytes(q)) == oracle_respond(D, 9, q) == oracle_respond(D & 0x1FF, 9, q)


def test_respond_error_behaviour_matches_reference():
    rng = np.random.default_rng(9)
    K, N = 100, 20
    srv, _ = cp.Server.setup_from_matrix(SEED, rng.integers(0, 512, size=(K, N), dtype=np.uint32), 9, skip_hint=True)
    good = qbytes(rand_u32(rng, K))
    cases = {
        b"": "FailedToDeserializeMatrixFromBytes",
        good[:8]: "FailedToDeserializeMatrixFromBytes",  # len <= 8
        good[:-1]: "FailedToDeserializeMatrixFromBytes",
        good + b"\0\0\0\0": "FailedToDeserializeMatrixFromBytes",
        bytes(8) + bytes(4): "FailedToDeserializeMatrixFromBytes",  # rows*cols == 0
        qbytes(rand_u32(rng, K + 1)): "IncompatibleDimensionForRowVectorTransposedMatrixMultiplication",
        O.matrix_to_bytes(rand_u32(rng, (2, K // 2))): "IncompatibleDimensionForRowVectorTransposedMatrixMultiplication",
    }
    for bad, variant in cases.items():
        with pytest.raises(cp.ChalametPIRError) as e:
            srv.respond(bad)
        assert e.value.variant == variant
        with pytest.raises(O.OracleError) as oe:  # the oracle agrees on the variant
            O.Server.setup_from_matrix(SEED, np.ones((K, N), np.uint32), 9, want_hint=False)[0].respond(bad)
        assert oe.value.name == variant
    assert srv.respond(good)  # still serviceable after errors


def test_setup_argument_errors():
    with pytest.raises(cp.ChalametPIRError) as e:
        cp.Server.setup_from_matrix(SEED, np.ones((4, 4), np.uint32), 3, skip_hint=True)
    assert e.value.variant == "ImpossibleEncodedDBMatrixElementBitLength"
    with pytest.raises(cp.ChalametPIRError) as e:
        cp.Server.setup_from_matrix(SEED, np.ones((4, 4), np.uint32), 15, skip_hint=True)
    assert e.value.variant == "ImpossibleEncodedDBMatrixElementBitLength"
    with pytest.raises(cp.ChalametPIRError) as e:
        cp.Server.setup(SEED, {}, 3)
    assert e.value.variant == "EmptyKVDatabase"
    with pytest.raises(cp.ChalametPIRError) as e:
        cp.Server.setup(SEED, make_db(4), 5)
    assert e.value.variant == "UnsupportedArityForBinaryFuseFilter"


def test_column_slices_concatenate_to_full_response():
    rng = np.random.default_rng(10)
    K, N, b = 5000, 940, 9
    D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
    q = rand_u32(rng, K)
    full = O.matrix_from_bytes(oracle_respond(D, b, q))[0]
    for world in (2, 3, 8):
        bounds = [N * r // world for r in range(world + 1)]
        parts = []
        for r in range(world):
            srv, _ = cp.Server.setup_from_matrix(SEED, D, b, skip_hint=True, col_begin=bounds[r], col_count=bounds[r + 1] - bounds[r])
            assert srv.col_begin == bounds[r]
            parts.append(O.matrix_from_bytes(srv.respond(qbytes(q)))[0])
        assert np.array_equal(np.concatenate(parts), full)


def test_respond_is_reentrant_across_threads():  # Arc<Server> shared across tasks, examples/server.rs:45,55,85
    rng = np.random.default_rng(11)
    K, N, b = 20000, 300, 10
    D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
    srv, _ = cp.Server.setup_from_matrix(SEED, D, b, skip_hint=True)
    qs = [rand_u32(rng, K) for _ in range(8)]
    want = [oracle_respond(D, b, q) for q in qs]
    errs = []

    def worker(i):
        try:
            for _ in range(10):
                if srv.respond(qbytes(qs[i])) != want[i]:
                    errs.append(i)
        except Exception as ex:  # pragma: no cover
            errs.append(repr(ex))

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs
    assert [r for r in srv.respond_batch([qbytes(q) for q in qs])] == want


@pytest.mark.parametrize("batch_tc", [1, 2])
def test_coalesced_respond_across_threads(batch_tc):
    """chpir_setup_opts.respond_coalesce: concurrent Server::respond calls are answered by shared launches (grid.y GEMV for small
    batches, the tensor-core limb GEMM from 6 queries up when the limb planes are resident) -- every caller still gets exactly
    the bytes the oracle's Server::respond produces, and the error behaviour of a lone call is unchanged."""
    rng = np.random.default_rng(12)
    K, N, b = 20011, 301, 10
    D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
    srv, _ = cp.Server.setup_from_matrix(SEED, D, b, skip_hint=True, batch_tc=batch_tc, respond_coalesce=True)
    nthreads, per = 24, 12
    qs = [rand_u32(rng, K) for _ in range(nthreads)]
    qs[0][:] = 0xFFFFFFFF
    want = [oracle_respond(D, b, q) for q in qs]
    assert srv.respond(qbytes(qs[3])) == want[3]  # a lone caller: batch of one
    errs = []
    barrier = threading.Barrier(nthreads)

    def worker(i):
        try:
            barrier.wait()
            for _ in range(per):
                if srv.respond(qbytes(qs[i])) != want[i]:
                    errs.append(i)
        except Exception as ex:  # pragma: no cover
            errs.append(repr(ex))

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(nthreads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs
    # the host batch entry point takes the same two routes (tensor cores from 6 queries up when the planes are resident)
    assert srv.respond_batch([qbytes(q) for q in qs]) == want
    assert srv.respond_batch([qbytes(q) for q in qs[:3]]) == want[:3]
    with pytest.raises(cp.ChalametPIRError) as e:
        srv.respond(qbytes(qs[0])[:-4])
    assert e.value.variant == "FailedToDeserializeMatrixFromBytes"
    srv.close()


def test_respond_device_pointers_and_batch():
    import torch

    rng = np.random.default_rng(12)
    K, N, b = 30011, 846, 10
    D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
    srv, _ = cp.Server.setup_from_matrix(SEED, D, b, skip_hint=True)
    nq = 3
    Q = rand_u32(rng, (nq, K))
    dq = torch.from_numpy(Q.view(np.int32)).cuda()
    dr = torch.empty((nq, N), dtype=torch.int32, device="cuda")
    srv.respond_device(dq.data_ptr(), nq, dr.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = dr.cpu().numpy().view(np.uint32)
    for i in range(nq):
        assert np.array_equal(got[i], O.matrix_from_bytes(oracle_respond(D, b, Q[i]))[0])


@pytest.mark.parametrize("arity,n_log2", [(3, 18), (3, 20), (4, 20), (3, 22)])
def test_full_size_respond_properties(arity, n_log2):
    """BASELINE.json configs 2/3/4/5 shapes (2^18, 2^20, 2^22 entries, 1 kB values; at 2^22 K*N = 4.4e9 elements exceeds the
    reference's own u32 index arithmetic): the oracle cannot stream 4.4 GB in seconds, so the
    full-size run is checked through size-independent properties: unit-vector queries read rows back, an all-ones query
    gives the column sums, a sparse random query equals the hand-computed combination, and respond is linear."""
    import torch

    n = 1 << n_log2
    b = cp.find_mat_elem_bit_len(n)
    K, N = cp.db_matrix_shape(arity, n, 1024, b)
    g = torch.Generator(device="cuda").manual_seed(arity)
    D = torch.randint(0, 1 << b, (K, N), dtype=torch.int32, device="cuda", generator=g)
    srv, _ = cp.Server.setup_from_device_matrix(SEED, D.data_ptr(), K, N, b, skip_hint=True)
    assert srv.packed_bytes == K * 16 * (-(-(-(-N // (64 // b))) // 2))
    rng = np.random.default_rng(arity)

    def respond(qv):
        return O.matrix_from_bytes(srv.respond(qbytes(qv)))[0]

    for k in (0, 1, K // 2 + 3, K - 1):
        e = np.zeros(K, np.uint32)
        e[k] = 1
        assert np.array_equal(respond(e), D[k].cpu().numpy().view(np.uint32))
    colsum = torch.zeros(N, dtype=torch.int64, device="cuda")
    for r0 in range(0, K, 1 << 20):  # chunked: an int64 copy of the 2^22 matrix would be 35 GB
        colsum += D[r0 : r0 + (1 << 20)].sum(dim=0, dtype=torch.int64)
    colsum = (colsum & 0xFFFFFFFF).cpu().numpy().astype(np.uint32)
    assert np.array_equal(respond(np.ones(K, np.uint32)), colsum)
    idx = rng.choice(K, size=2000, replace=False)
    vals = rand_u32(rng, 2000)
    qs = np.zeros(K, np.uint32)
    qs[idx] = vals
    rows = D[torch.from_numpy(idx).cuda()].cpu().numpy().view(np.uint32).astype(np.uint64)
    want = ((vals.astype(np.uint64)[:, None] * rows).sum(axis=0) & 0xFFFFFFFF).astype(np.uint32)
    assert np.array_equal(respond(qs), want)
    q1, q2 = rand_u32(rng, K), rand_u32(rng, K)
    assert np.array_equal(respond(q1 + q2), respond(q1) + respond(q2))
    srv.close()


@pytest.mark.parametrize("n_log2", [18, 20])
def test_full_size_respond_bytes_equal_oracle_on_encoded_db(n_log2):
    """BASELINE.json configs[1] / configs[2] on a REAL encoded database (2^18 / 2^20 keys of 32 bytes, 1 kB values, 3-wise filter): D is
    built by the host encoder (byte-identical to the oracle's from_kv_database, checked here in full at 2^18), the oracle keeps it in the
    reference's layout (transposed + row_wise_compress'ed, matrix.rs:98-205) and answers with the reference's loop (matrix.rs:328-485);
    the GPU server's response bytes -- streaming GEMV and tensor-core batch route -- must equal the oracle's byte for byte."""
    import ctypes as C

    from chalametpir_b200._lib import FILTER_PARAM_BYTE_LEN, lib
    from chalametpir_b200.errors import check

    n = 1 << n_log2
    rs = np.random.default_rng(n_log2)
    keys = rs.integers(0, 256, size=(n, 32), dtype=np.uint8)
    keys[:, :8] = np.arange(n, dtype="<u8").view(np.uint8).reshape(n, 8)  # distinct by construction
    vals = rs.integers(0, 256, size=(n, 1024), dtype=np.uint8)
    b = cp.find_mat_elem_bit_len(n)
    K, N = cp.db_matrix_shape(3, n, 1024, b)
    ko = np.arange(n + 1, dtype=np.uint64) * np.uint64(32)
    vo = np.arange(n + 1, dtype=np.uint64) * np.uint64(1024)
    D = np.empty((K, N), dtype=np.uint32)
    fbytes = np.empty(FILTER_PARAM_BYTE_LEN, dtype=np.uint8)
    rng_seed = C.c_uint64(11)
    check(lib.chpir_encode_kv_database(3, n, keys.ctypes.data, ko.ctypes.data, vals.ctypes.data, vo.ctypes.data, b, 100, C.byref(rng_seed), D.ctypes.data,
                                       fbytes.ctypes.data))
    if n_log2 == 18:
        db = {keys[i].tobytes(): vals[i].tobytes() for i in range(n)}
        Do, fo = O.from_kv_database(db, b, 3, rng_seed=11)
        assert fo.to_bytes() == fbytes.tobytes() and np.array_equal(Do, D)
        del db, Do
    osrv, _ = O.Server.setup_from_matrix(SEED, D, b, want_hint=False)
    srv, _ = cp.Server.setup_from_matrix(SEED, D, b, skip_hint=True, batch_tc=1)
    del D
    rng = np.random.default_rng(1 + n_log2)
    qs = [rand_u32(rng, K) for _ in range(4)] + [np.ones(K, np.uint32), np.full(K, 0xFFFFFFFF, np.uint32)]
    want = [osrv.respond(qbytes(q)) for q in qs]
    assert [srv.respond(qbytes(q)) for q in qs] == want          # streaming GEMV
    assert srv.respond_batch([qbytes(q) for q in qs]) == want    # 6 queries: the tensor-core limb GEMM
    srv.close()


def test_full_size_batched_tensor_core_respond():
    """BASELINE.json configs[3]: 2^20 entries, 4-wise filter, 64 queries per batch through the int8-limb GEMM.  Checked against
    the streaming GEMV for every query and, independently of both, through unit-vector and all-ones queries inside the batch."""
    import torch

    arity, n, nq = 4, 1 << 20, 64
    b = cp.find_mat_elem_bit_len(n)
    K, N = cp.db_matrix_shape(arity, n, 1024, b)
    g = torch.Generator(device="cuda").manual_seed(44)
    D = torch.randint(0, 1 << b, (K, N), dtype=torch.int32, device="cuda", generator=g)
    srv, _ = cp.Server.setup_from_device_matrix(SEED, D.data_ptr(), K, N, b, skip_hint=True, batch_tc=1)
    Q = torch.randint(-(2**31), 2**31, (nq, K), dtype=torch.int32, device="cuda", generator=g)
    Q[0] = 0
    Q[0, K - 1] = 1            # unit vector: the response is the last row of D
    Q[1] = 1                   # all ones: column sums
    Q[2] = -1                  # all 0xFFFFFFFF: every limb saturated
    r_tc = torch.empty((nq, N), dtype=torch.int32, device="cuda")
    r_gv = torch.empty((nq, N), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    srv.respond_device_tc(Q.data_ptr(), nq, r_tc.data_ptr(), st)
    srv.respond_device(Q.data_ptr(), nq, r_gv.data_ptr(), st)
    torch.cuda.synchronize()
    assert torch.equal(r_tc, r_gv)
    assert torch.equal(r_tc[0], D[K - 1])
    colsum = D.sum(dim=0, dtype=torch.int64)
    assert torch.equal(r_tc[1].to(torch.int64) & 0xFFFFFFFF, colsum & 0xFFFFFFFF)
    assert torch.equal(r_tc[2].to(torch.int64) & 0xFFFFFFFF, (-colsum) & 0xFFFFFFFF)
    srv.close()


@pytest.mark.parametrize("kernel", ["pair", "1sm"])
@pytest.mark.parametrize("b,K,N,nq", [(9, 4099, 941, 5), (10, 30011, 846, 130), (8, 2500, 37, 64), (14, 777, 129, 3), (4, 5000, 300, 128),
                                      (9, 3001, 600, 96), (10, 2049, 257, 40), (12, 129, 513, 33)])
def test_respond_tensor_core_batch_matches_oracle(monkeypatch, kernel, b, K, N, nq):
    """Batched respond as a limb-decomposed int8 GEMM (north_star (2)): bit-exact against the oracle's Server::respond
    for every query of the batch, and identical to the streaming GEMV path."""
    import torch

    # the CTA-pair kernel (cta_group::2; A rows in passes of 32 / 64 / 96 / 128) is the default, the one-SM kernel the kept comparison
    monkeypatch.setenv("CHPIR_GEMM_KERNEL", kernel)
    rng = np.random.default_rng(b * 1000 + nq)
    D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
    srv, _ = cp.Server.setup_from_matrix(SEED, D, b, skip_hint=True, batch_tc=1)
    Q = rand_u32(rng, (nq, K))
    Q[0] = 0xFFFFFFFF  # all limbs saturated: int32 accumulators must wrap, not clamp
    dq = torch.from_numpy(Q.view(np.int32)).cuda()
    dr = torch.full((nq, N), -1, dtype=torch.int32, device="cuda")
    dg = torch.empty((nq, N), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    srv.respond_device_tc(dq.data_ptr(), nq, dr.data_ptr(), st)
    srv.respond_device(dq.data_ptr(), nq, dg.data_ptr(), st)
    torch.cuda.synchronize()
    got = dr.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, dg.cpu().numpy().view(np.uint32))
    osrv, _ = O.Server.setup_from_matrix(SEED, D, b, want_hint=False)
    for i in sorted(set([0, 1, nq // 2, nq - 1])):
        assert np.array_equal(got[i], O.matrix_from_bytes(osrv.respond(qbytes(Q[i])))[0]), i
    srv.close()


def test_respond_tensor_core_batch_needs_planes():
    srv, _ = cp.Server.setup_from_matrix(SEED, np.ones((64, 8), np.uint32), 9, skip_hint=True)
    import torch

    dq = torch.zeros((1, 64), dtype=torch.int32, device="cuda")
    dr = torch.zeros((1, 8), dtype=torch.int32, device="cuda")
    with pytest.raises(cp.ChalametPIRError) as e:
        srv.respond_device_tc(dq.data_ptr(), 1, dr.data_ptr(), 0)
    assert e.value.variant == "InvalidArgument"


# ------------------------------------------------------------------ persisted server state (chpir_server_save / chpir_server_load)
@pytest.mark.parametrize("b,K,N", [(9, 5003, 941), (10, 20000, 300), (4, 777, 17), (14, 1024, 64), (9, 5003, 121), (10, 4098, 7)])
def test_saved_server_answers_identically_after_load(tmp_path, b, K, N):
    rng = np.random.default_rng(b + K)
    D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
    srv, _ = cp.Server.setup_from_matrix(SEED, D, b, skip_hint=True, col_begin=3 if N > 20 else 0)
    path = tmp_path / "server.chpir"
    srv.save(path)
    assert path.stat().st_size == 64 + srv.packed_bytes
    qs = [rand_u32(rng, K) for _ in range(7)]
    want = [srv.respond(qbytes(q)) for q in qs]
    assert want[0] == oracle_respond(D[:, srv.col_begin :], b, qs[0])
    for batch_tc, coalesce in ((0, False), (1, True)):
        ld = cp.Server.load(path, batch_tc=batch_tc, respond_coalesce=coalesce)
        assert (ld.rows_k, ld.cols_n, ld.col_begin, ld.mat_elem_bit_len, ld.packed_bytes) == (srv.rows_k, srv.cols_n, srv.col_begin, b, srv.packed_bytes)
        assert [ld.respond(qbytes(q)) for q in qs] == want
        assert ld.respond_batch([qbytes(q) for q in qs]) == want  # 7 queries: the tensor-core route when the planes were rebuilt
        ld.close()
    # damaged files are rejected, not served
    raw = bytearray(path.read_bytes())
    raw[64 + len(raw) // 2] ^= 1
    (tmp_path / "flipped").write_bytes(raw)
    (tmp_path / "short").write_bytes(path.read_bytes()[:-16])
    (tmp_path / "magic").write_bytes(b"X" + path.read_bytes()[1:])
    for name in ("flipped", "short", "magic"):
        with pytest.raises(cp.ChalametPIRError) as e:
            cp.Server.load(tmp_path / name)
        assert e.value.variant == "InvalidSavedServer", name
    with pytest.raises(cp.ChalametPIRError) as e:
        cp.Server.load(tmp_path / "does-not-exist")
    assert e.value.variant == "IoFailed"


@pytest.mark.parametrize("b,K,N", [(9, 5003, 118), (9, 20001, 117), (9, 1, 119), (10, 777, 13), (4, 3001, 17 * 16 - 3), (14, 4097, 63 * 4), (9, 2 * 4096, 63 * 7)])
def test_tight_rows_answer_exactly_like_padded_rows(monkeypatch, tmp_path, b, K, N):
    """Narrow slices with an odd number of u64 words per row (118 columns x 9 bit = 17 words at 8-way sharding) can be stored without
    the 16-byte row padding (csrc/common.cuh PackedLayout.tight, CHPIR_TIGHT_PITCH=1).  Same bytes out as the padded layout, fewer
    bytes streamed; odd K exercises the even-row bulk copies that run into the zeroed pad row."""
    rng = np.random.default_rng(K + N)
    D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
    qs = [rand_u32(rng, K) for _ in range(5)]
    padded, _ = cp.Server.setup_from_matrix(SEED, D, b, skip_hint=True)
    monkeypatch.setenv("CHPIR_TIGHT_PITCH", "1")  # opt-in: measured slower than padded rows on the slices it was meant for
    tight, _ = cp.Server.setup_from_matrix(SEED, D, b, skip_hint=True, batch_tc=1)
    monkeypatch.delenv("CHPIR_TIGHT_PITCH")
    words = -(-N // (64 // b))
    assert words % 2 == 1 and tight.info.row_pitch_bytes == 8 * words and padded.info.row_pitch_bytes == 8 * (words + 1)
    assert tight.packed_bytes == K * 8 * words
    want = [oracle_respond(D, b, q) for q in qs]
    assert [tight.respond(qbytes(q)) for q in qs] == want
    assert [padded.respond(qbytes(q)) for q in qs] == want
    assert tight.respond_batch([qbytes(q) for q in qs]) == want
    # the saved file records its own row layout: it loads as tight whatever the default is at load time
    tight.save(tmp_path / "tight.chpir")
    ld = cp.Server.load(tmp_path / "tight.chpir")
    assert ld.info.row_pitch_bytes == 8 * words and [ld.respond(qbytes(q)) for q in qs] == want
