"""N > 1 host logic on CPU: gloo ranks shard the matrix, answer for their share (the oracle stands in for the GPU kernels here -- these
tests are about partition / padding / gather / reduce, not arithmetic).  Both cuts of csrc/cluster.cu: row blocks of D with an exact
sum of the partial responses (what respond runs on for n > 1) and column slices (hint; respond with CHPIR_CLUSTER_SHARD=cols)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chalametpir_b200 import sharding
from oracle import oracle as O

SEED = bytes(range(32))


def test_slice_of_partitions_exactly():
    for n in (1, 7, 118, 846, 940, 941):
        for world in (1, 2, 3, 4, 8):
            got = [sharding.slice_of(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == n
            for (b0, c0), (b1, _) in zip(got, got[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in got) - min(c for _, c in got) <= 1


def test_query_slices_cover_the_query_exactly():
    for K in (1, 5, 997, 1179648, 1130496):
        for world in (1, 2, 3, 8):
            parts = [sharding.query_slice(K, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == K
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert all(p[1] - p[0] <= p[2] for p in parts) and len({p[2] for p in parts}) == 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, K, N, b, lwe, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)  # same D and queries on every rank
        D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
        Q = 3
        q = rng.integers(0, 2**32, size=(Q, K), dtype=np.uint64).astype(np.uint32)
        qt = torch.from_numpy(q.view(np.int32)).clone()
        if rank != 0:
            qt.zero_()
        dist.broadcast(qt, 0)  # the query batch reaches every rank
        q_here = qt.numpy().view(np.uint32)
        # the e2e route: every rank holds only its K/world words of each query, the slices are all-gathered and re-laid out
        k0, k1, ks = sharding.query_slice(K, rank, world)
        q_slice = torch.zeros((Q, ks), dtype=torch.int32)
        q_slice[:, : k1 - k0] = torch.from_numpy(q.view(np.int32)[:, k0:k1].copy())
        q_rows = sharding.allgather_query_slices(dist, torch, q_slice, torch.empty((world, Q, ks), dtype=torch.int32), torch.empty((Q, K), dtype=torch.int32))
        assert np.array_equal(q_rows.numpy().view(np.uint32), q)
        c0, nc = sharding.slice_of(N, rank, world)
        srv, hint = O.Server.setup_from_matrix(SEED, np.ascontiguousarray(D[:, c0 : c0 + nc]), b, lwe_rows=lwe)
        local = np.stack([O.matrix_from_bytes(srv.respond(O.matrix_to_bytes(q_here[i : i + 1])))[0] for i in range(Q)])
        full = sharding.gather_response_slices(dist, torch, torch.from_numpy(local.view(np.int32)), N, world)
        hints = [None] * world
        dist.all_gather_object(hints, hint)
        if rank == 0:
            ref_srv, ref_hint = O.Server.setup_from_matrix(SEED, D, b, lwe_rows=lwe)
            for i in range(Q):
                want = ref_srv.respond(O.matrix_to_bytes(q[i : i + 1]))
                assert sharding.response_bytes(full[i].numpy().view(np.uint32)) == want
            assert sharding.interleave_hint_slices(hints) == ref_hint
            open(os.path.join(out_dir, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N", [(2, 37), (2, 940 // 8), (3, 50)])
def test_two_rank_gloo_column_sharding(tmp_path, world, N):
    mp.spawn(_worker, args=(world, _free_port(), 997, N, 9, 16, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").read_text() == "ok"


def _pipeline_worker(rank, world, port, K, N, b, steps, out_dir):
    """The serving schedule of bench.py at N > 1 on CPU tensors: query batches are broadcast one step ahead (double-buffered) on the
    default group, every rank answers for its column slice, and the response slices are gathered asynchronously on a SECOND group so
    that a gather -- which cannot finish before the slowest rank has answered -- never sits in front of the next broadcast."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pg_out = dist.new_group(backend="gloo")
    try:
        rng = np.random.default_rng(5)
        D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
        Q = 2
        batches = rng.integers(0, 2**32, size=(steps, Q, K), dtype=np.uint64).astype(np.uint32)
        c0, nc = sharding.slice_of(N, rank, world)
        counts = sharding.slice_counts(N, world)
        pad = max(counts)
        srv, _ = O.Server.setup_from_matrix(SEED, np.ascontiguousarray(D[:, c0 : c0 + nc]), b, want_hint=False)
        q_bufs = [torch.zeros((Q, K), dtype=torch.int32) for _ in range(2)]
        send = [torch.zeros((Q, pad), dtype=torch.int32) for _ in range(2)]
        gathered = [torch.zeros((world, Q, pad), dtype=torch.int32) for _ in range(2)]
        results = []

        def load(i, buf):  # only the root holds the queries before the broadcast
            if rank == 0:
                buf.copy_(torch.from_numpy(batches[i].view(np.int32)))
            else:
                buf.zero_()

        load(0, q_bufs[0])
        bw = dist.broadcast(q_bufs[0], 0, async_op=True)
        gw = [None, None]
        for i in range(steps):
            p = i & 1
            bw.wait()
            if i + 1 < steps:
                load(i + 1, q_bufs[p ^ 1])
                bw = dist.broadcast(q_bufs[p ^ 1], 0, async_op=True)
            q_here = q_bufs[p].numpy().view(np.uint32)
            local = np.stack([O.matrix_from_bytes(srv.respond(O.matrix_to_bytes(q_here[j : j + 1])))[0] for j in range(Q)])
            if gw[p] is not None:
                gw[p][0].wait()
                results.append(sharding.unpad_gathered(torch, gathered[p], counts).clone())
            send[p].zero_()
            send[p][:, :nc] = torch.from_numpy(local.view(np.int32))
            gw[p] = (dist.all_gather_into_tensor(gathered[p].view(-1), send[p].view(-1), group=pg_out, async_op=True), i)
        for w in sorted((w for w in gw if w is not None), key=lambda t: t[1]):
            w[0].wait()
            results.append(sharding.unpad_gathered(torch, gathered[w[1] & 1], counts).clone())
        assert len(results) == steps
        if rank == 0:
            ref, _ = O.Server.setup_from_matrix(SEED, D, b, want_hint=False)
            for i in range(steps):
                for j in range(Q):
                    want = ref.respond(O.matrix_to_bytes(batches[i, j : j + 1]))
                    assert sharding.response_bytes(results[i][j].numpy().view(np.uint32)) == want, (i, j)
            open(os.path.join(out_dir, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_pipelined_serving_schedule_with_a_second_group_for_the_gathers(tmp_path, world):
    mp.spawn(_pipeline_worker, args=(world, _free_port(), 503, 118, 9, 5, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").read_text() == "ok"


def _row_cut_worker(rank, world, port, K, N, b, out_dir):
    """The row cut of csrc/cluster.cu (respond, n > 1) with the library's own plan: rank r holds rows [k0_r, k0_r + kn_r) of D at full
    width, zero rows up to the common pitch, and ONLY the words [k0_r, k0_r + kn_r) of every query (garbage behind them: those words
    face zero rows); the partial responses are added mod 2^32 -- one exact reduce, no query word is exchanged."""
    import chalametpir_b200 as cp

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(321)  # same D and queries on every rank (each rank USES only its share)
        D = rng.integers(0, 1 << b, size=(K, N), dtype=np.uint32)
        Q = 4
        q = rng.integers(0, 2**32, size=(Q, K), dtype=np.uint64).astype(np.uint32)
        pl = cp.cluster_plan(world, rank, K, N)
        k0, kn, ks = pl["k_begin"], pl["k_count"], pl["k_pitch"]
        block = np.zeros((ks, N), dtype=np.uint32)
        block[:kn] = D[k0 : k0 + kn]
        q_slice = np.full((Q, ks), 0x5A5A5A5A, dtype=np.uint32)
        q_slice[:, :kn] = q[:, k0 : k0 + kn]
        srv, _ = O.Server.setup_from_matrix(SEED, block, b, want_hint=False)
        part = np.stack([O.matrix_from_bytes(srv.respond(O.matrix_to_bytes(q_slice[i : i + 1])))[0] for i in range(Q)])
        # exact sum mod 2^32: gloo has no unsigned 32-bit sum, int64 lanes cannot overflow for <= 2^31 ranks
        t = torch.from_numpy(part.astype(np.int64))
        dist.reduce(t, 0, op=dist.ReduceOp.SUM)
        if rank == 0:
            total = (t.numpy() & 0xFFFFFFFF).astype(np.uint32)
            ref, _ = O.Server.setup_from_matrix(SEED, D, b, want_hint=False)
            for i in range(Q):
                assert sharding.response_bytes(total[i]) == ref.respond(O.matrix_to_bytes(q[i : i + 1])), i
            open(os.path.join(out_dir, "ok"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,K,N", [(2, 997, 37), (3, 4099, 118), (2, 7, 5), (4, 40, 9)])
def test_gloo_row_cut_partial_responses_sum_to_the_full_response(tmp_path, world, K, N):
    mp.spawn(_row_cut_worker, args=(world, _free_port(), K, N, 9, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").read_text() == "ok"
