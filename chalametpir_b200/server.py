"""Host-side mirror of ``chalametpir_server::Server`` (chalametpir_server/src/server.rs:15-190) over the C ABI.

Same names, argument meaning and error behaviour as the reference:

    server, hint_bytes, filter_param_bytes = Server.setup(seed_mu, db, arity=3)     # server.rs:103
    response_bytes = server.respond(query_bytes)                                    # server.rs:184

Everything numeric happens in libchalamet_b200.so (CUDA, sm_100a).  This module only marshals buffers.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Mapping, Optional, Tuple

import numpy as np

from ._lib import FILTER_PARAM_BYTE_LEN, LWE_DIMENSION, SEED_BYTE_LEN, ServerInfo, SetupOpts, SetupTiming, lib
from .errors import ChalametPIRError, check

_ctx_lock = threading.Lock()
_ctxs: dict = {}


def device_count() -> int:
    n = C.c_int()
    check(lib.chpir_device_count(C.byref(n)))
    return n.value


def get_ctx(device: int = 0):
    """One chpir_ctx per device ordinal, created on first use (replaces gpu_utils::setup_gpu, gpu_utils.rs:25-79)."""
    with _ctx_lock:
        ctx = _ctxs.get(device)
        if ctx is None:
            h = C.c_void_p()
            check(lib.chpir_ctx_create(device, C.byref(h)))
            ctx = _ctxs[device] = h
        return ctx


def drop_a_cache(device: int = 0) -> int:
    """Free the LWE matrix a ``setup(..., a_cache=True)`` left resident on this device; returns the bytes released."""
    n = C.c_uint64()
    check(lib.chpir_ctx_drop_a_cache(get_ctx(device), C.byref(n)))
    return n.value


class PinnedBuffer:
    """Page-locked host memory from chpir_host_alloc, viewed as a numpy uint8 array (``.array``); freed on close / GC.
    Queries and responses handed to :meth:`Server.respond_into` from such buffers move at the full PCIe rate."""

    def __init__(self, nbytes: int):
        p = C.c_void_p()
        check(lib.chpir_host_alloc(nbytes, C.byref(p)))
        self.ptr = p.value
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.ptr))

    def close(self) -> None:
        if getattr(self, "ptr", None):
            self.array = None
            lib.chpir_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def find_mat_elem_bit_len(db_entry_count: int) -> int:
    """server.rs:193-218"""
    b = C.c_uint32()
    check(lib.chpir_find_mat_elem_bit_len(db_entry_count, C.byref(b)))
    return b.value


def db_matrix_shape(arity: int, db_entry_count: int, max_value_byte_len: int, mat_elem_bit_len: int) -> Tuple[int, int]:
    k, n = C.c_uint64(), C.c_uint64()
    check(lib.chpir_db_matrix_shape(arity, db_entry_count, max_value_byte_len, mat_elem_bit_len, C.byref(k), C.byref(n)))
    return k.value, n.value


def _flatten(items):
    off = np.zeros(len(items) + 1, dtype=np.uint64)
    if len(items):
        off[1:] = np.cumsum(np.fromiter((len(x) for x in items), dtype=np.uint64, count=len(items)))
    blob = np.frombuffer(b"".join(items), dtype=np.uint8)
    if blob.size == 0:
        blob = np.zeros(1, dtype=np.uint8)
    return blob, off


def _seed_arr(seed: bytes) -> np.ndarray:
    if len(seed) != SEED_BYTE_LEN:
        raise ValueError(f"seed must be {SEED_BYTE_LEN} bytes")
    return np.frombuffer(bytes(seed), dtype=np.uint8)


def encode_kv_database(db: Mapping[bytes, bytes], mat_elem_bit_len: int, arity: int = 3, max_attempt_count: int = 100,
                       filter_seed_rng: Optional[int] = None):
    """Matrix::from_kv_database::<ARITY> (matrix.rs:633) on the host -> (D [K x N] uint32, filter_param_bytes)."""
    if len(db) == 0:
        raise ChalametPIRError(5)
    keys, vals = list(db.keys()), list(db.values())
    K, N = db_matrix_shape(arity, len(keys), max(len(v) for v in vals), mat_elem_bit_len)
    kb, ko = _flatten(keys)
    vb, vo = _flatten(vals)
    D = np.empty((K, N), dtype=np.uint32)
    fbytes = np.empty(FILTER_PARAM_BYTE_LEN, dtype=np.uint8)
    rng = C.c_uint64(filter_seed_rng) if filter_seed_rng is not None else None
    check(
        lib.chpir_encode_kv_database(
            arity, len(keys), kb.ctypes.data, ko.ctypes.data, vb.ctypes.data, vo.ctypes.data, mat_elem_bit_len, max_attempt_count,
            C.byref(rng) if rng is not None else None, D.ctypes.data, fbytes.ctypes.data,
        )
    )
    return D, fbytes.tobytes()


def encode_kv_database_device(db: Mapping[bytes, bytes], mat_elem_bit_len: int, arity: int = 3, max_attempt_count: int = 100,
                              filter_seed_rng: Optional[int] = None, device: int = 0):
    """Matrix::from_kv_database::<ARITY> with the row encoding and the dependent fill on the GPU (csrc/encode_dev.cu); returns the
    downloaded (D, filter_param_bytes) so that it can be compared with :func:`encode_kv_database` byte for byte."""
    if len(db) == 0:
        raise ChalametPIRError(5)
    keys, vals = list(db.keys()), list(db.values())
    K, N = db_matrix_shape(arity, len(keys), max(len(v) for v in vals), mat_elem_bit_len)
    kb, ko = _flatten(keys)
    vb, vo = _flatten(vals)
    D = np.empty((K, N), dtype=np.uint32)
    fbytes = np.empty(FILTER_PARAM_BYTE_LEN, dtype=np.uint8)
    rng = C.c_uint64(filter_seed_rng) if filter_seed_rng is not None else None
    check(
        lib.chpir_encode_kv_database_device(
            get_ctx(device), arity, len(keys), kb.ctypes.data, ko.ctypes.data, vb.ctypes.data, vo.ctypes.data, mat_elem_bit_len, max_attempt_count,
            C.byref(rng) if rng is not None else None, D.ctypes.data, fbytes.ctypes.data,
        )
    )
    return D, fbytes.tobytes()


class Server:
    """The PIR server: bit-packed D resident in HBM on one GPU (optionally a column slice of it)."""

    def __init__(self, handle: C.c_void_p, device: int):
        self._h = handle
        self.device = device
        info = ServerInfo()
        check(lib.chpir_server_get_info(self._h, C.byref(info)))
        self.info = info
        self.rows_k = info.rows_k
        self.cols_n = info.cols_n
        self.col_begin = info.col_begin
        self.mat_elem_bit_len = info.mat_elem_bit_len
        self.packed_bytes = info.packed_bytes

    # ------------------------------------------------------------------ setup
    @staticmethod
    def _opts(lwe_rows=0, col_begin=0, col_count=0, gemm_variant=0, skip_hint=False, batch_tc=0, a_expand="auto", host_chunk_rows=0,
              db_encode="host", respond_coalesce=False, a_cache=False, hint_on_device=False) -> SetupOpts:
        """a_expand: "auto" (default: the faster walker, i.e. "host"), "host" (one host core squeezes the TurboSHAKE128 chain and the
        uploads + panel GEMMs are pipelined behind it) or "device" (the chain on one GPU warp, A never leaves the device) -- same
        bytes every way, ~10x lower setup latency on the host core.
        a_cache: keep A resident in the device context and reuse it in later setups with the same seed, LWE dimension and K (a
        database update): the XOF chain is skipped and the hint is the tensor-core GEMM alone -- same bytes."""
        mode = {"auto": 0, "host": 1, "device": 2, 0: 0, 1: 1, 2: 2}[a_expand]
        enc = {"host": 0, "device": 1, 0: 0, 1: 1}[db_encode]  # setup / setup_from_arrays only: where the rows of D are encoded and filled
        return SetupOpts(lwe_rows, col_begin, col_count, gemm_variant, 1 if skip_hint else 0, batch_tc, mode, host_chunk_rows,
                         1 if respond_coalesce else 0, enc, 1 if a_cache else 0, 1 if hint_on_device else 0)

    @staticmethod
    def setup(seed_mu: bytes, db: Mapping[bytes, bytes], arity: int = 3, *, device: int = 0, filter_seed_rng: Optional[int] = None,
              **opts) -> Tuple["Server", bytes, bytes]:
        """Server::setup::<ARITY>(seed_mu, db) -> (Server, hint_bytes, filter_param_bytes)   [server.rs:103]"""
        if len(db) == 0:
            raise ChalametPIRError(5)  # EmptyKVDatabase, server.rs:105-107
        if arity not in (3, 4):
            raise ChalametPIRError(14)
        seed = _seed_arr(seed_mu)
        keys, vals = list(db.keys()), list(db.values())
        b = find_mat_elem_bit_len(len(keys))
        K, N = db_matrix_shape(arity, len(keys), max(len(v) for v in vals), b)
        o = Server._opts(**opts)
        m = o.lwe_rows or LWE_DIMENSION
        nc = o.col_count or (N - o.col_begin)
        hint = np.empty(8 + 4 * m * nc, dtype=np.uint8)
        fbytes = np.empty(FILTER_PARAM_BYTE_LEN, dtype=np.uint8)
        kb, ko = _flatten(keys)
        vb, vo = _flatten(vals)
        rng = C.c_uint64(filter_seed_rng) if filter_seed_rng is not None else None
        h = C.c_void_p()
        hl = C.c_size_t()
        check(
            lib.chpir_server_setup_from_db(
                get_ctx(device), arity, seed.ctypes.data, len(keys), kb.ctypes.data, ko.ctypes.data, vb.ctypes.data, vo.ctypes.data,
                C.byref(rng) if rng is not None else None, C.byref(o), hint.ctypes.data, hint.nbytes, C.byref(hl), fbytes.ctypes.data, C.byref(h),
            )
        )
        return Server(h, device), hint[: hl.value].tobytes(), fbytes.tobytes()

    @staticmethod
    def setup_from_arrays(seed_mu: bytes, keys: np.ndarray, values: np.ndarray, arity: int = 3, *, device: int = 0,
                          filter_seed_rng: Optional[int] = None, **opts) -> Tuple["Server", bytes, bytes]:
        """Server::setup::<ARITY> for a database given as two 2-D uint8 arrays (n x key_len, n x value_len): the same call as
        :meth:`setup` without materialising a million-entry Python dict (keys must be distinct, as HashMap keys are)."""
        keys = np.ascontiguousarray(keys, dtype=np.uint8)
        values = np.ascontiguousarray(values, dtype=np.uint8)
        n = keys.shape[0]
        if n == 0:
            raise ChalametPIRError(5)
        if arity not in (3, 4):
            raise ChalametPIRError(14)
        seed = _seed_arr(seed_mu)
        b = find_mat_elem_bit_len(n)
        K, N = db_matrix_shape(arity, n, values.shape[1], b)
        o = Server._opts(**opts)
        m = o.lwe_rows or LWE_DIMENSION
        nc = o.col_count or (N - o.col_begin)
        hint = np.empty(8 + 4 * m * nc, dtype=np.uint8)
        fbytes = np.empty(FILTER_PARAM_BYTE_LEN, dtype=np.uint8)
        ko = np.arange(n + 1, dtype=np.uint64) * np.uint64(keys.shape[1])
        vo = np.arange(n + 1, dtype=np.uint64) * np.uint64(values.shape[1])
        rng = C.c_uint64(filter_seed_rng) if filter_seed_rng is not None else None
        h = C.c_void_p()
        hl = C.c_size_t()
        check(
            lib.chpir_server_setup_from_db(
                get_ctx(device), arity, seed.ctypes.data, n, keys.ctypes.data, ko.ctypes.data, values.ctypes.data, vo.ctypes.data,
                C.byref(rng) if rng is not None else None, C.byref(o), hint.ctypes.data, hint.nbytes, C.byref(hl), fbytes.ctypes.data, C.byref(h),
            )
        )
        return Server(h, device), hint[: hl.value].tobytes(), fbytes.tobytes()

    @staticmethod
    def setup_from_matrix(seed_mu: bytes, D: np.ndarray, mat_elem_bit_len: int, *, device: int = 0, **opts) -> Tuple["Server", Optional[bytes]]:
        """The device half of setup for an already-encoded D (K x N uint32, host): A expansion, hint GEMM, pack."""
        seed = _seed_arr(seed_mu)
        D = np.ascontiguousarray(D, dtype=np.uint32)
        K, N = D.shape
        o = Server._opts(**opts)
        m = o.lwe_rows or LWE_DIMENSION
        nc = o.col_count or (N - o.col_begin)
        hint = None if o.skip_hint else np.empty(8 + 4 * m * max(nc, 0), dtype=np.uint8)
        h = C.c_void_p()
        hl = C.c_size_t()
        check(
            lib.chpir_server_setup(
                get_ctx(device), seed.ctypes.data, D.ctypes.data, K, N, mat_elem_bit_len, C.byref(o), hint.ctypes.data if hint is not None else None,
                hint.nbytes if hint is not None else 0, C.byref(hl), C.byref(h),
            )
        )
        return Server(h, device), (hint[: hl.value].tobytes() if hint is not None else None)

    @staticmethod
    def setup_from_device_matrix(seed_mu: bytes, d_ptr: int, rows_k: int, cols_n: int, mat_elem_bit_len: int, *, device: int = 0,
                                 **opts) -> Tuple["Server", Optional[bytes]]:
        """Same, with D already in HBM (d_ptr = device pointer to K x N uint32, e.g. a torch tensor's data_ptr())."""
        seed = _seed_arr(seed_mu)
        o = Server._opts(**opts)
        m = o.lwe_rows or LWE_DIMENSION
        nc = o.col_count or (cols_n - o.col_begin)
        hint = None if o.skip_hint else np.empty(8 + 4 * m * max(nc, 0), dtype=np.uint8)
        h = C.c_void_p()
        hl = C.c_size_t()
        check(
            lib.chpir_server_setup_device(
                get_ctx(device), seed.ctypes.data, d_ptr, rows_k, cols_n, mat_elem_bit_len, C.byref(o), hint.ctypes.data if hint is not None else None,
                hint.nbytes if hint is not None else 0, C.byref(hl), C.byref(h),
            )
        )
        return Server(h, device), (hint[: hl.value].tobytes() if hint is not None else None)

    # ------------------------------------------------------------------ persisted state
    def save(self, path: str) -> None:
        """Write the resident packed column slice to ``path`` (chpir_server_save); hint and filter bytes stay with the caller."""
        check(lib.chpir_server_save(self._h, str(path).encode()))

    @staticmethod
    def load(path: str, *, device: int = 0, batch_tc: int = 0, respond_coalesce: bool = False) -> "Server":
        """A server that answers queries straight from a file written by :meth:`save`, without re-running setup."""
        o = Server._opts(batch_tc=batch_tc, respond_coalesce=respond_coalesce)
        h = C.c_void_p()
        check(lib.chpir_server_load(get_ctx(device), str(path).encode(), C.byref(o), C.byref(h)))
        return Server(h, device)

    # ------------------------------------------------------------------ respond
    def respond(self, query: bytes) -> bytes:
        """Server::respond(&self, query) -> response bytes   [server.rs:184-190]"""
        q = np.frombuffer(query, dtype=np.uint8)
        out = np.empty(8 + 4 * self.cols_n, dtype=np.uint8)
        n = C.c_size_t()
        check(lib.chpir_server_respond(self._h, q.ctypes.data if q.size else None, q.size, out.ctypes.data, out.nbytes, C.byref(n)))
        return out[: n.value].tobytes()

    def respond_into(self, query_ptr: int, query_len: int, resp_ptr: int, resp_cap: int) -> int:
        """Pointer form of respond (no Python-side copies): host buffers, e.g. pinned torch tensors."""
        n = C.c_size_t()
        check(lib.chpir_server_respond(self._h, query_ptr, query_len, resp_ptr, resp_cap, C.byref(n)))
        return n.value

    def respond_batch(self, queries) -> list:
        nq = len(queries)
        arrs = [np.frombuffer(q, dtype=np.uint8) for q in queries]
        ptrs = (C.c_void_p * nq)(*[a.ctypes.data for a in arrs])
        lens = (C.c_size_t * nq)(*[a.size for a in arrs])
        stride = 8 + 4 * self.cols_n
        out = np.empty(nq * stride, dtype=np.uint8)
        check(lib.chpir_server_respond_batch(self._h, ptrs, lens, nq, out.ctypes.data, stride))
        return [out[i * stride : (i + 1) * stride].tobytes() for i in range(nq)]

    def respond_device(self, q_ptr: int, nq: int, resp_ptr: int, stream: int = 0) -> None:
        """Device-resident respond: q_ptr -> nq x K uint32, resp_ptr -> nq x cols_n uint32, enqueued on `stream`."""
        check(lib.chpir_server_respond_device(self._h, q_ptr, nq, resp_ptr, stream or None))  # 0/None = CUDA default stream

    def respond_device_tc(self, q_ptr: int, nq: int, resp_ptr: int, stream: int = 0) -> None:
        """Batched respond on the tensor cores (limb-decomposed int8 GEMM): same arguments as respond_device."""
        check(lib.chpir_server_respond_device_tc(self._h, q_ptr, nq, resp_ptr, stream or None))

    # ------------------------------------------------------------------ introspection
    def setup_timing(self) -> dict:
        t = SetupTiming()
        check(lib.chpir_server_setup_timing(self._h, C.byref(t)))
        return t.as_dict()

    def last_kernel_ms(self) -> dict:
        r, g, e = C.c_float(), C.c_float(), C.c_float()
        check(lib.chpir_server_last_kernel_ms(self._h, C.byref(r), C.byref(g), C.byref(e)))
        return {"respond_ms": r.value, "gemm_ms": g.value, "expand_ms": e.value}

    def close(self) -> None:
        if getattr(self, "_h", None):
            if not getattr(self, "_borrowed", False):  # a cluster's shard belongs to the cluster server
                lib.chpir_server_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def generate_from_seed(rows: int, cols: int, seed: bytes, row_begin: int = 0, row_count: Optional[int] = None, device: int = 0) -> np.ndarray:
    """Matrix::generate_from_seed (matrix.rs:541-558), expanded on the GPU."""
    row_count = rows - row_begin if row_count is None else row_count
    out = np.empty((row_count, cols), dtype=np.uint32)
    s = _seed_arr(seed)
    check(lib.chpir_generate_from_seed(get_ctx(device), s.ctypes.data, rows, cols, row_begin, row_count, out.ctypes.data))
    return out


def host_generate_from_seed(rows: int, cols: int, seed: bytes, row_begin: int = 0, row_count: Optional[int] = None, impl: int = 0) -> np.ndarray:
    """Matrix::generate_from_seed (matrix.rs:541-558) on one host core: the producer of the host-pipelined setup mode
    (csrc/host_xof.cpp).  impl: 0 = fastest on this CPU, 1 = portable scalar, 2 = BMI2, 3 = AVX-512."""
    row_count = rows - row_begin if row_count is None else row_count
    out = np.empty((row_count, cols), dtype=np.uint32)
    s = _seed_arr(seed)
    check(lib.chpir_host_generate_from_seed(s.ctypes.data, rows, cols, row_begin, row_count, impl, out.ctypes.data))
    return out


def host_xof_impl() -> str:
    return lib.chpir_host_xof_impl().decode()


def matmul(a: np.ndarray, b: np.ndarray, b_elem_bit_len: int = 32, variant: int = 0, device: int = 0) -> np.ndarray:
    """&A * &B mod 2^32 (matrix.rs:1040-1059) on the GPU. variant 0 = tensor-core limb GEMM (entries of B < 2^b_elem_bit_len <= 2^16;
    wider B falls through to the SIMT kernel), 1 = SIMT u32.  Entries of B that do not fit b_elem_bit_len raise InvalidArgument."""
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    out = np.empty((a.shape[0], b.shape[1]), dtype=np.uint32)
    check(lib.chpir_matmul(get_ctx(device), a.ctypes.data, a.shape[0], a.shape[1], b.ctypes.data, b.shape[0], b.shape[1], b_elem_bit_len, variant, out.ctypes.data))
    return out
