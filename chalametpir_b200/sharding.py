"""Column-slice sharding of the server across ranks (one process per GPU, torch.distributed).

`M[:, n0:n1] = A . D[:, n0:n1]` and `resp[n0:n1] = q . D[:, n0:n1]` need no cross-rank arithmetic:
rank r owns columns `slice_of(N, r, world)`, and collectives are used only to move the query in and the slices out.
These helpers hold the host-side logic (partition, padding, gather layout, re-interleaving of the hint); they are
backend-agnostic, so the same code runs over NCCL on GPUs and over gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def slice_of(n_cols: int, rank: int, world: int) -> Tuple[int, int]:
    """(col_begin, col_count) of `rank`: contiguous, sizes differ by at most one, earlier ranks take the remainder."""
    base, rem = divmod(n_cols, world)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def slice_counts(n_cols: int, world: int) -> List[int]:
    return [slice_of(n_cols, r, world)[1] for r in range(world)]


def gather_response_slices(dist, torch, local: "torch.Tensor", n_cols: int, world: int, group=None) -> "torch.Tensor":
    """local: (Q, col_count) int32 slice of Q responses on this rank -> (Q, n_cols) on every rank.
    Slices are padded to the widest one so that a single all_gather moves them."""
    counts = slice_counts(n_cols, world)
    pad = max(counts)
    Q = local.shape[0]
    send = torch.zeros((Q, pad), dtype=local.dtype, device=local.device)
    send[:, : local.shape[1]] = local
    out = torch.empty((world, Q, pad), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), send.view(-1), group=group)
    return unpad_gathered(torch, out, counts)


def unpad_gathered(torch, gathered: "torch.Tensor", counts: Sequence[int]) -> "torch.Tensor":
    """(world, Q, pad) -> (Q, sum(counts)): drop the padding and concatenate the column slices in rank order."""
    return torch.cat([gathered[r, :, :c] for r, c in enumerate(counts)], dim=1)


def query_slice(K: int, rank: int, world: int) -> Tuple[int, int, int]:
    """(k0, k1, ks): rank `rank` uploads words [k0, k1) of every query over its own PCIe link; ks = ceil(K / world) is the padded
    slice length every rank contributes to the all-gather (H2D-scatter + all-gather instead of a full upload
    per rank)."""
    ks = -(-K // world)
    k0 = min(K, rank * ks)
    return k0, min(K, k0 + ks), ks


def upload_query_slices(q_words: "torch.Tensor", k0: int, k1: int, q_slice: "torch.Tensor", stream) -> None:
    """q_words: (Q, K) int32 view of the batch in page-locked host memory; q_slice: (Q, ks) device buffer of this rank.  Copies words
    [k0, k1) of every query with ONE strided DMA on `stream` (a torch ``copy_`` of the strided view would be staged through pageable
    memory, and one ``copy_`` per query costs ~25 us of host time each -- more than the GPU needs per batch at 8-way sharding)."""
    from ._lib import lib
    from .errors import check

    if k1 <= k0:
        return
    Q, K = q_words.shape
    check(lib.chpir_upload_rows(q_slice.data_ptr(), q_slice.shape[1] * 4, q_words.data_ptr() + 4 * k0, K * 4, (k1 - k0) * 4, Q, stream.cuda_stream))


def allgather_query_slices(dist, torch, q_slice: "torch.Tensor", q_all: "torch.Tensor", q_rows: "torch.Tensor", group=None) -> "torch.Tensor":
    """q_slice: (Q, ks) this rank's words of Q queries (zero padded) -> q_rows: (Q, K) whole queries on every rank.
    q_all is the (world, Q, ks) all-gather landing buffer; the rank-major result is re-laid out query-major into q_rows."""
    world, Q, ks = q_all.shape
    dist.all_gather_into_tensor(q_all.view(-1), q_slice.view(-1), group=group)
    q_rows.copy_(q_all.permute(1, 0, 2).reshape(Q, world * ks)[:, : q_rows.shape[1]])
    return q_rows


def response_bytes(row: np.ndarray) -> bytes:
    """Matrix::to_bytes of a 1 x N response (matrix.rs:947-971)."""
    row = np.ascontiguousarray(row, dtype="<u4").reshape(-1)
    return np.array([1, row.size], dtype="<u4").tobytes() + row.tobytes()


def interleave_hint_slices(slices: Sequence[bytes]) -> bytes:
    """Wire-format hint slices (each rows x col_count, in rank order) -> the wire-format rows x N hint."""
    mats = []
    for s in slices:
        h = np.frombuffer(s, dtype="<u4")
        mats.append(h[2:].reshape(int(h[0]), int(h[1])))
    full = np.concatenate(mats, axis=1)
    return np.array(full.shape, dtype="<u4").tobytes() + np.ascontiguousarray(full).tobytes()
