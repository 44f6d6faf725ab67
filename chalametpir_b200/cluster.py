"""The column-sharded server on 1..8 GPUs of one process (``chpir_cluster_*``, csrc/cluster.cu).

Same two calls as the reference -- ``Server::setup(seed, db)`` (chalametpir_server/src/server.rs:103) and
``Server::respond(&self, query)`` (server.rs:184) -- with the sharding behind the handle:

    cluster = Cluster(n_gpus=8)                      # or Cluster() -> $CHPIR_GPUS, default 1
    server, hint_bytes, filter_param_bytes = ClusterServer.setup(cluster, seed_mu, db, arity=3)
    response_bytes = server.respond(query_bytes)     # thread-safe; concurrent callers share launches

The hint, the filter parameters and every response are byte-identical to the single-GPU :class:`chalametpir_b200.Server`.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Optional, Sequence, Tuple

import numpy as np

from ._lib import FILTER_PARAM_BYTE_LEN, LWE_DIMENSION, ClusterServerInfo, ServerInfo, SetupTiming, lib
from .errors import ChalametPIRError, check
from .server import Server, _flatten, _seed_arr, db_matrix_shape, find_mat_elem_bit_len

RESPOND_GEMV = 0
RESPOND_TC = 1


def cluster_plan(n_ranks: int, rank: int, rows_k: int, cols_n: int) -> dict:
    """The slice plan (pure host arithmetic): columns of D / hint / response owned by `rank`, and the words of every query it ingests."""
    c0, nc = C.c_uint32(), C.c_uint32()
    k0, kn, ks = C.c_uint64(), C.c_uint64(), C.c_uint64()
    check(lib.chpir_cluster_plan(n_ranks, rank, rows_k, cols_n, C.byref(c0), C.byref(nc), C.byref(k0), C.byref(kn), C.byref(ks)))
    return {"col_begin": c0.value, "col_count": nc.value, "k_begin": k0.value, "k_count": kn.value, "k_pitch": ks.value}


def _name_the_process_nccl() -> None:
    """The dynamic loader keeps ONE instance per SONAME: whichever libnccl.so.2 is loaded first is the one every later user gets.
    PyTorch ships its own (newer) NCCL and fails to import if an older system copy is already resident, so before the library
    resolves NCCL (csrc/cluster.cu, dlopen at the first multi-GPU hint gather) point it at the copy a later ``import torch`` would
    load -- $CHPIR_NCCL_LIB, honoured by the library; a no-op when the variable is set or no bundled copy exists."""
    import os
    import sys

    if os.environ.get("CHPIR_NCCL_LIB"):
        return
    for base in sys.path:
        cand = os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ["CHPIR_NCCL_LIB"] = cand
            return


class Cluster:
    """``n_gpus`` GPUs of this process with peer access between all of them (replaces gpu_utils::setup_gpu, gpu_utils.rs:25-79)."""

    def __init__(self, n_gpus: int = 0, devices: Optional[Sequence[int]] = None):
        _name_the_process_nccl()
        h = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            check(lib.chpir_cluster_create(len(devices), arr, C.byref(h)))
        else:
            check(lib.chpir_cluster_create(n_gpus, None, C.byref(h)))
        self._h = h
        n = C.c_int()
        check(lib.chpir_cluster_size(self._h, C.byref(n)))
        self.n_gpus = n.value
        self.devices = []
        for r in range(self.n_gpus):
            d = C.c_int()
            check(lib.chpir_cluster_ctx(self._h, r, None, C.byref(d)))
            self.devices.append(d.value)

    def ctx(self, rank: int):
        """The chpir_ctx of rank `rank` (borrowed)."""
        h = C.c_void_p()
        check(lib.chpir_cluster_ctx(self._h, rank, C.byref(h), None))
        return h

    def drop_a_cache(self) -> list:
        """Free the LWE matrix a ``setup(..., a_cache=True)`` left resident on every GPU of the cluster; returns the bytes released per rank."""
        out = []
        for r in range(self.n_gpus):
            n = C.c_uint64()
            check(lib.chpir_ctx_drop_a_cache(self.ctx(r), C.byref(n)))
            out.append(n.value)
        return out

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.chpir_cluster_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ClusterServer:
    def __init__(self, handle: C.c_void_p, cluster: Cluster):
        self._h = handle
        self.cluster = cluster  # keeps the contexts alive for as long as the server
        self.info = self.get_info()
        self.n_gpus = self.info["n_gpus"]
        self.rows_k = self.info["rows_k"]
        self.cols_n = self.info["cols_n"]
        self.mat_elem_bit_len = self.info["mat_elem_bit_len"]
        self.k_pitch = self.info["k_pitch"]

    # ------------------------------------------------------------------ setup
    @staticmethod
    def setup(cluster: Cluster, seed_mu: bytes, db: Mapping[bytes, bytes], arity: int = 3, *, filter_seed_rng: Optional[int] = None,
              **opts) -> Tuple["ClusterServer", bytes, bytes]:
        """Server::setup::<ARITY>(seed_mu, db) -> (server, hint_bytes, filter_param_bytes)   [server.rs:103]"""
        if len(db) == 0:
            raise ChalametPIRError(5)
        if arity not in (3, 4):
            raise ChalametPIRError(14)
        keys, vals = list(db.keys()), list(db.values())
        kb, ko = _flatten(keys)
        vb, vo = _flatten(vals)
        return ClusterServer._from_db(cluster, seed_mu, len(keys), kb, ko, vb, vo, max(len(v) for v in vals), arity, filter_seed_rng, opts)

    @staticmethod
    def setup_from_arrays(cluster: Cluster, seed_mu: bytes, keys: np.ndarray, values: np.ndarray, arity: int = 3, *,
                          filter_seed_rng: Optional[int] = None, **opts) -> Tuple["ClusterServer", bytes, bytes]:
        keys = np.ascontiguousarray(keys, dtype=np.uint8)
        values = np.ascontiguousarray(values, dtype=np.uint8)
        n = keys.shape[0]
        if n == 0:
            raise ChalametPIRError(5)
        if arity not in (3, 4):
            raise ChalametPIRError(14)
        ko = np.arange(n + 1, dtype=np.uint64) * np.uint64(keys.shape[1])
        vo = np.arange(n + 1, dtype=np.uint64) * np.uint64(values.shape[1])
        return ClusterServer._from_db(cluster, seed_mu, n, keys, ko, values, vo, values.shape[1], arity, filter_seed_rng, opts)

    @staticmethod
    def _from_db(cluster, seed_mu, n, kb, ko, vb, vo, max_vlen, arity, filter_seed_rng, opts):
        seed = _seed_arr(seed_mu)
        b = find_mat_elem_bit_len(n)
        K, N = db_matrix_shape(arity, n, max_vlen, b)
        o = Server._opts(**opts)
        m = o.lwe_rows or LWE_DIMENSION
        hint = np.empty(8 + 4 * m * N, dtype=np.uint8)
        fbytes = np.empty(FILTER_PARAM_BYTE_LEN, dtype=np.uint8)
        rng = C.c_uint64(filter_seed_rng) if filter_seed_rng is not None else None
        h, hl = C.c_void_p(), C.c_size_t()
        check(lib.chpir_cluster_server_setup_from_db(
            cluster._h, arity, seed.ctypes.data, n, kb.ctypes.data, ko.ctypes.data, vb.ctypes.data, vo.ctypes.data,
            C.byref(rng) if rng is not None else None, C.byref(o), hint.ctypes.data, hint.nbytes, C.byref(hl), fbytes.ctypes.data, C.byref(h)))
        return ClusterServer(h, cluster), hint[: hl.value].tobytes(), fbytes.tobytes()

    @staticmethod
    def setup_from_matrix(cluster: Cluster, seed_mu: bytes, D: np.ndarray, mat_elem_bit_len: int, **opts) -> Tuple["ClusterServer", Optional[bytes]]:
        """The device half of setup for an already-encoded D (K x N uint32, host): every rank uploads its own columns."""
        seed = _seed_arr(seed_mu)
        D = np.ascontiguousarray(D, dtype=np.uint32)
        K, N = D.shape
        o = Server._opts(**opts)
        m = o.lwe_rows or LWE_DIMENSION
        hint = None if o.skip_hint else np.empty(8 + 4 * m * N, dtype=np.uint8)
        h, hl = C.c_void_p(), C.c_size_t()
        check(lib.chpir_cluster_server_setup(cluster._h, seed.ctypes.data, D.ctypes.data, K, N, mat_elem_bit_len, C.byref(o),
                                             hint.ctypes.data if hint is not None else None, hint.nbytes if hint is not None else 0, C.byref(hl), C.byref(h)))
        return ClusterServer(h, cluster), (hint[: hl.value].tobytes() if hint is not None else None)

    @staticmethod
    def setup_from_device_slices(cluster: Cluster, seed_mu: bytes, d_ptrs: Sequence[int], rows_k: int, cols_n: int, mat_elem_bit_len: int,
                                 **opts) -> Tuple["ClusterServer", Optional[bytes]]:
        """d_ptrs[r]: device pointer on rank r's GPU to its compact K x col_count(r) uint32 slice (see :func:`cluster_plan`)."""
        seed = _seed_arr(seed_mu)
        o = Server._opts(**opts)
        m = o.lwe_rows or LWE_DIMENSION
        hint = None if o.skip_hint else np.empty(8 + 4 * m * cols_n, dtype=np.uint8)
        ptrs = (C.c_void_p * len(d_ptrs))(*d_ptrs)
        h, hl = C.c_void_p(), C.c_size_t()
        check(lib.chpir_cluster_server_setup_device(cluster._h, seed.ctypes.data, ptrs, rows_k, cols_n, mat_elem_bit_len, C.byref(o),
                                                    hint.ctypes.data if hint is not None else None, hint.nbytes if hint is not None else 0, C.byref(hl), C.byref(h)))
        return ClusterServer(h, cluster), (hint[: hl.value].tobytes() if hint is not None else None)

    # ------------------------------------------------------------------ persisted state
    def save(self, path_prefix: str) -> None:
        check(lib.chpir_cluster_server_save(self._h, str(path_prefix).encode()))

    @staticmethod
    def load(cluster: Cluster, path_prefix: str, *, batch_tc: int = 0, respond_coalesce: bool = False) -> "ClusterServer":
        o = Server._opts(batch_tc=batch_tc, respond_coalesce=respond_coalesce)
        h = C.c_void_p()
        check(lib.chpir_cluster_server_load(cluster._h, str(path_prefix).encode(), C.byref(o), C.byref(h)))
        return ClusterServer(h, cluster)

    # ------------------------------------------------------------------ respond
    def respond(self, query: bytes) -> bytes:
        """Server::respond(&self, query) -> response bytes   [server.rs:184-190]"""
        q = np.frombuffer(query, dtype=np.uint8)
        out = np.empty(8 + 4 * self.cols_n, dtype=np.uint8)
        n = C.c_size_t()
        check(lib.chpir_cluster_server_respond(self._h, q.ctypes.data if q.size else None, q.size, out.ctypes.data, out.nbytes, C.byref(n)))
        return out[: n.value].tobytes()

    def respond_into(self, query_ptr: int, query_len: int, resp_ptr: int, resp_cap: int) -> int:
        n = C.c_size_t()
        check(lib.chpir_cluster_server_respond(self._h, query_ptr, query_len, resp_ptr, resp_cap, C.byref(n)))
        return n.value

    def respond_batch(self, queries) -> list:
        nq = len(queries)
        arrs = [np.frombuffer(q, dtype=np.uint8) for q in queries]
        ptrs = (C.c_void_p * nq)(*[a.ctypes.data for a in arrs])
        lens = (C.c_size_t * nq)(*[a.size for a in arrs])
        stride = 8 + 4 * self.cols_n
        out = np.empty(nq * stride, dtype=np.uint8)
        check(lib.chpir_cluster_server_respond_batch(self._h, ptrs, lens, nq, out.ctypes.data, stride))
        return [out[i * stride : (i + 1) * stride].tobytes() for i in range(nq)]

    def respond_concurrent(self, query_ptrs: Sequence[int], query_len: int, total_calls: int, resp_ptr: int, resp_stride: int, n_threads: int) -> float:
        """`n_threads` native threads call chpir_cluster_server_respond concurrently (call j sends query j % len(query_ptrs)); returns the
        wall seconds.  Host buffers by pointer, e.g. slices of a :class:`PinnedBuffer`."""
        n = len(query_ptrs)
        ptrs = (C.c_void_p * n)(*query_ptrs)
        lens = (C.c_size_t * n)(*([query_len] * n))
        sec = C.c_double()
        check(lib.chpir_cluster_server_respond_concurrent(self._h, ptrs, lens, n, total_calls, resp_ptr, resp_stride, n_threads, C.byref(sec)))
        return sec.value

    def shard(self, rank: int) -> Server:
        """Rank `rank`'s resident slice as a (borrowed, non-owning) single-GPU :class:`Server`."""
        sh = C.c_void_p()
        check(lib.chpir_cluster_server_shard(self._h, rank, C.byref(sh)))
        s = Server(sh, self.cluster.devices[rank])
        s._borrowed = True
        s._keepalive = self
        return s

    def respond_device(self, q_slice_ptrs: Sequence[int], nq: int, resp_ptr0: int, mode: int = RESPOND_GEMV, repeats: int = 1) -> float:
        """Device-resident respond; returns the device time (ms) of all `repeats` passes, measured on rank 0's GPU."""
        ptrs = (C.c_void_p * len(q_slice_ptrs))(*q_slice_ptrs)
        ms = C.c_float()
        check(lib.chpir_cluster_server_respond_device(self._h, ptrs, nq, resp_ptr0, mode, repeats, C.byref(ms)))
        return ms.value

    # ------------------------------------------------------------------ introspection
    def get_info(self) -> dict:
        info = ClusterServerInfo()
        check(lib.chpir_cluster_server_get_info(self._h, C.byref(info)))
        return info.as_dict()

    def plan(self, rank: int) -> dict:
        return cluster_plan(self.n_gpus, rank, self.rows_k, self.cols_n)

    def shard_info(self, rank: int) -> dict:
        sh = C.c_void_p()
        check(lib.chpir_cluster_server_shard(self._h, rank, C.byref(sh)))
        si, st = ServerInfo(), SetupTiming()
        check(lib.chpir_server_get_info(sh, C.byref(si)))
        check(lib.chpir_server_setup_timing(sh, C.byref(st)))
        r, g, e = C.c_float(), C.c_float(), C.c_float()
        check(lib.chpir_server_last_kernel_ms(sh, C.byref(r), C.byref(g), C.byref(e)))
        return {"rows_k": si.rows_k, "cols_n": si.cols_n, "col_begin": si.col_begin, "row_pitch_bytes": si.row_pitch_bytes, "packed_bytes": si.packed_bytes,
                "timing": st.as_dict(), "gemm_ms": g.value, "expand_ms": e.value}

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.chpir_cluster_server_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
