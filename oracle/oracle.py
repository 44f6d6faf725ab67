"""ctypes binding of the CPU oracle (oracle/chalamet_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never from chalametpir_b200/.  Parity status: "parity
unpinned" (see the header of chalamet_oracle.c and DESIGN.md section "Oracle").
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libchalamet_oracle.so")

LWE_DIMENSION = 1774
SEED_BYTE_LEN = 32

ERR = {
    0: "Ok",
    1: "InvalidMatrixDimension",
    2: "IncompatibleDimensionForMatrixMultiplication",
    3: "IncompatibleDimensionForRowVectorTransposedMatrixMultiplication",
    4: "FailedToDeserializeMatrixFromBytes",
    5: "EmptyKVDatabase",
    6: "ExhaustedAllAttemptsToBuild3WiseXorFilter",
    7: "ExhaustedAllAttemptsToBuild4WiseXorFilter",
    8: "RowNotDecodable",
    9: "DecodedRowNotPrependedWithDigestOfKey",
    10: "FailedToDeserializeFilterFromBytes",
    11: "KVDatabaseSizeTooLarge",
    12: "InvalidHintMatrix",
    13: "ArithmeticOverflowAddingQueryIndicator",
    14: "UnsupportedArityForBinaryFuseFilter",
    15: "InvalidResponseVector",
    16: "ImpossibleEncodedDBMatrixElementBitLength",
    100: "AllocationFailed",
}


class OracleError(Exception):
    def __init__(self, code: int):
        self.code = code
        self.name = ERR.get(code, f"Unknown({code})")
        super().__init__(self.name)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "chalamet_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libchalamet_oracle.so"])
    return _LIB_PATH


class _Filter(C.Structure):
    _fields_ = [
        ("seed", C.c_uint8 * 32),
        ("arity", C.c_uint32),
        ("segment_length", C.c_uint32),
        ("segment_count_length", C.c_uint32),
        ("num_fingerprints", C.c_uint64),
        ("filter_size", C.c_uint64),
        ("mat_elem_bit_len", C.c_uint64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_mix.restype = C.c_uint64
        _lib.orc_mix.argtypes = [C.c_uint64, C.c_uint64]
        _lib.orc_key_hash.restype = C.c_uint64
        _lib.orc_server_packed.restype = C.c_void_p
    return _lib


def _u8(b) -> np.ndarray:
    return np.frombuffer(bytes(b), dtype=np.uint8) if not isinstance(b, np.ndarray) else b


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _chk(rc: int):
    if rc != 0:
        raise OracleError(rc)


# ---------------------------------------------------------------- TurboSHAKE128
def turboshake128(msg: bytes, out_len: int, dsep: int = 0x1F, skip: int = 0) -> bytes:
    out = np.empty(out_len, dtype=np.uint8)
    m = _u8(msg)
    lib().orc_turboshake128_at(_p(m), C.c_size_t(len(m)), C.c_uint8(dsep), C.c_uint64(skip), _p(out), C.c_size_t(out_len))
    return out.tobytes()


# ---------------------------------------------------------------- matrix ops
def generate_from_seed(rows: int, cols: int, seed: bytes) -> np.ndarray:
    out = np.empty((rows, cols), dtype=np.uint32)
    _chk(lib().orc_generate_from_seed(C.c_uint64(rows), C.c_uint64(cols), _p(_u8(seed)), _p(out)))
    return out


def generate_rows_from_seed(cols: int, seed: bytes, row0: int, nrows: int) -> np.ndarray:
    out = np.empty((nrows, cols), dtype=np.uint32)
    _chk(lib().orc_generate_rows_from_seed(C.c_uint64(cols), _p(_u8(seed)), C.c_uint64(row0), C.c_uint64(nrows), _p(out)))
    return out


def matmul(a: np.ndarray, b: np.ndarray, fast: bool = True) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    out = np.empty((a.shape[0], b.shape[1]), dtype=np.uint32)
    fn = lib().orc_matmul_fast if fast else lib().orc_matmul
    _chk(fn(_p(a), C.c_uint64(a.shape[0]), C.c_uint64(a.shape[1]), _p(b), C.c_uint64(b.shape[0]), C.c_uint64(b.shape[1]), _p(out)))
    return out


def transpose(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint32)
    out = np.empty((a.shape[1], a.shape[0]), dtype=np.uint32)
    lib().orc_transpose(_p(a), C.c_uint64(a.shape[0]), C.c_uint64(a.shape[1]), _p(out))
    return out


def compression_factor(b: int) -> int:
    return lib().orc_compression_factor(C.c_uint(b))


def row_wise_compress(a: np.ndarray, b: int) -> np.ndarray:
    cf = compression_factor(b)
    if cf == 0:
        raise OracleError(16)
    a = np.ascontiguousarray(a, dtype=np.uint32)
    out = np.empty((a.shape[0], -(-a.shape[1] // cf)), dtype=np.uint32)
    _chk(lib().orc_row_wise_compress(_p(a), C.c_uint64(a.shape[0]), C.c_uint64(a.shape[1]), C.c_uint(b), _p(out)))
    return out


def row_wise_decompress(a: np.ndarray, b: int, num_cols: int) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint32)
    out = np.empty((a.shape[0], num_cols), dtype=np.uint32)
    _chk(lib().orc_row_wise_decompress(_p(a), C.c_uint64(a.shape[0]), C.c_uint64(a.shape[1]), C.c_uint(b), C.c_uint64(num_cols), _p(out)))
    return out


def gemv_packed(q: np.ndarray, packed: np.ndarray, decompressed_num_cols: int, b: int) -> np.ndarray:
    q = np.ascontiguousarray(q, dtype=np.uint32)
    if q.ndim == 1:
        q = q[None, :]
    packed = np.ascontiguousarray(packed, dtype=np.uint32)
    out = np.empty((1, packed.shape[0]), dtype=np.uint32)
    _chk(
        lib().orc_gemv_packed(
            _p(q), C.c_uint64(q.shape[0]), C.c_uint64(q.shape[1]), _p(packed), C.c_uint64(packed.shape[0]), C.c_uint64(packed.shape[1]),
            C.c_uint64(decompressed_num_cols), C.c_uint(b), _p(out),
        )
    )
    return out


def matrix_to_bytes(m: np.ndarray) -> bytes:
    m = np.ascontiguousarray(m, dtype=np.uint32)
    return np.array(m.shape, dtype="<u4").tobytes() + m.tobytes()


def matrix_from_bytes(b: bytes) -> np.ndarray:
    rows, cols = C.c_uint32(), C.c_uint32()
    buf = _u8(b)
    _chk(lib().orc_matrix_from_bytes(_p(buf), C.c_size_t(len(buf)), C.byref(rows), C.byref(cols)))
    return np.frombuffer(bytes(b), dtype="<u4", offset=8).reshape(rows.value, cols.value).copy()


def find_mat_elem_bit_len(n: int) -> int:
    out = C.c_uint()
    _chk(lib().orc_find_mat_elem_bit_len(C.c_uint64(n), C.byref(out)))
    return out.value


# ---------------------------------------------------------------- filter / codec
@dataclass
class Filter:
    seed: bytes
    arity: int
    segment_length: int
    segment_count_length: int
    num_fingerprints: int
    filter_size: int
    mat_elem_bit_len: int

    def _c(self) -> _Filter:
        f = _Filter()
        C.memmove(f.seed, self.seed, 32)
        f.arity, f.segment_length, f.segment_count_length = self.arity, self.segment_length, self.segment_count_length
        f.num_fingerprints, f.filter_size, f.mat_elem_bit_len = self.num_fingerprints, self.filter_size, self.mat_elem_bit_len
        return f

    @staticmethod
    def _from_c(f: _Filter) -> "Filter":
        return Filter(bytes(f.seed), f.arity, f.segment_length, f.segment_count_length, f.num_fingerprints, f.filter_size, f.mat_elem_bit_len)

    def to_bytes(self) -> bytes:
        out = np.empty(68, dtype=np.uint8)
        f = self._c()
        lib().orc_filter_to_bytes(C.byref(f), _p(out))
        return out.tobytes()

    @staticmethod
    def from_bytes(b: bytes) -> "Filter":
        f = _Filter()
        buf = _u8(b)
        _chk(lib().orc_filter_from_bytes(_p(buf), C.c_size_t(len(buf)), C.byref(f)))
        return Filter._from_c(f)

    def bits_per_entry(self) -> float:
        return self.num_fingerprints * self.mat_elem_bit_len / self.filter_size


def filter_shape(arity: int, db_size: int):
    sl, sc, nf = C.c_uint32(), C.c_uint32(), C.c_uint64()
    lib().orc_filter_shape(C.c_uint(arity), C.c_uint64(db_size), C.byref(sl), C.byref(sc), C.byref(nf))
    return sl.value, sc.value, nf.value


def mix(key: int, seed: int) -> int:
    return lib().orc_mix(C.c_uint64(key), C.c_uint64(seed))


def key_hash(key: bytes, seed: bytes) -> int:
    k = _u8(key)
    return lib().orc_key_hash(_p(k), C.c_size_t(len(k)), _p(_u8(seed)))


def hash_batch(arity: int, h: int, segment_length: int, segment_count_length: int):
    out = (C.c_uint32 * 4)()
    lib().orc_hash_batch(C.c_uint(arity), C.c_uint64(h), C.c_uint32(segment_length), C.c_uint32(segment_count_length), out)
    return tuple(out[:arity])


def encode_kv_as_row(key: bytes, value: bytes, b: int, num_cols: int) -> np.ndarray:
    row = np.empty(num_cols, dtype=np.uint32)
    k, v = _u8(key), _u8(value)
    lib().orc_encode_kv_as_row(_p(k), C.c_size_t(len(k)), _p(v), C.c_size_t(len(v)), C.c_uint(b), C.c_uint64(num_cols), _p(row))
    return row


def decode_kv_from_row(row: np.ndarray, b: int) -> bytes:
    row = np.ascontiguousarray(row, dtype=np.uint32)
    out = np.zeros(max(1, (len(row) * b) // 8), dtype=np.uint8)
    n = C.c_uint64()
    _chk(lib().orc_decode_kv_from_row(_p(row), C.c_uint64(len(row)), C.c_uint(b), _p(out), C.byref(n)))
    return out[: n.value].tobytes()


def flatten(items):
    """list[bytes] -> (blob u8, offsets u64[n+1])"""
    off = np.zeros(len(items) + 1, dtype=np.uint64)
    if items:
        off[1:] = np.cumsum([len(x) for x in items], dtype=np.uint64)
    blob = np.frombuffer(b"".join(items), dtype=np.uint8) if items else np.zeros(0, dtype=np.uint8)
    if blob.size == 0:
        blob = np.zeros(1, dtype=np.uint8)
    return blob, off


def from_kv_database(db: dict, b: int, arity: int = 3, max_attempts: int = 100, rng_seed: int = 1):
    """matrix.rs:633 from_kv_database -> (D [K x N] u32, Filter)"""
    if len(db) == 0:
        raise OracleError(5)
    keys, vals = list(db.keys()), list(db.values())
    kb, ko = flatten(keys)
    vb, vo = flatten(vals)
    rows, cols = C.c_uint64(), C.c_uint64()
    lib().orc_db_matrix_shape(C.c_uint(arity), C.c_uint64(len(keys)), C.c_uint64(max(len(v) for v in vals)), C.c_uint(b), C.byref(rows), C.byref(cols))
    D = np.zeros((rows.value, cols.value), dtype=np.uint32)
    f = _Filter()
    _chk(
        lib().orc_from_kv_database(
            C.c_uint(arity), C.c_uint64(len(keys)), _p(kb), _p(ko), _p(vb), _p(vo), C.c_uint(b), C.c_uint(max_attempts), C.c_uint64(rng_seed),
            C.byref(f), _p(D),
        )
    )
    return D, Filter._from_c(f)


def recover_value(D: np.ndarray, filt: Filter, key: bytes) -> bytes:
    D = np.ascontiguousarray(D, dtype=np.uint32)
    out = np.zeros((D.shape[1] * filt.mat_elem_bit_len) // 8 + 8, dtype=np.uint8)
    n = C.c_uint64()
    k = _u8(key)
    f = filt._c()
    _chk(lib().orc_recover_value(_p(D), C.c_uint64(D.shape[1]), C.byref(f), _p(k), C.c_size_t(len(k)), _p(out), C.byref(n)))
    return out[: n.value].tobytes()


# ---------------------------------------------------------------- server / client
class Server:
    """server.rs:15-21 + :47-78 + :184-190 (CPU build of the reference)."""

    def __init__(self, handle, N: int):
        self._h, self.N = handle, N

    @staticmethod
    def setup_from_matrix(seed: bytes, D: np.ndarray, b: int, lwe_rows: int = 0, want_hint: bool = True, want_server: bool = True):
        D = np.ascontiguousarray(D, dtype=np.uint32)
        K, N = D.shape
        m = lwe_rows or LWE_DIMENSION
        hint = np.empty(8 + 4 * m * N, dtype=np.uint8) if want_hint else None
        h = C.c_void_p()
        _chk(
            lib().orc_server_setup_from_matrix(
                _p(_u8(seed)), _p(D), C.c_uint64(K), C.c_uint64(N), C.c_uint(b), C.c_uint32(lwe_rows), _p(hint) if want_hint else None,
                C.byref(h) if want_server else None,
            )
        )
        return (Server(h, N) if want_server else None), (hint.tobytes() if want_hint else None)

    @staticmethod
    def setup(seed: bytes, db: dict, arity: int = 3, rng_seed: int = 1, lwe_rows: int = 0):
        """Server::setup::<ARITY> -> (Server, hint_bytes, filter_param_bytes)"""
        if len(db) == 0:
            raise OracleError(5)
        b = find_mat_elem_bit_len(len(db))
        D, filt = from_kv_database(db, b, arity, 100, rng_seed)
        srv, hint = Server.setup_from_matrix(seed, D, b, lwe_rows)
        srv.D, srv.filter = D, filt
        return srv, hint, filt.to_bytes()

    def respond(self, query: bytes) -> bytes:
        out = np.empty(8 + 4 * self.N, dtype=np.uint8)
        q = _u8(query)
        _chk(lib().orc_server_respond(self._h, _p(q), C.c_size_t(len(q)), _p(out)))
        return out.tobytes()

    def packed(self) -> np.ndarray:
        n, pc = C.c_uint64(), C.c_uint64()
        ptr = lib().orc_server_packed(self._h, C.byref(n), C.byref(pc))
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), shape=(n.value, pc.value)).copy()

    def __del__(self):
        try:
            if self._h:
                lib().orc_server_free(self._h)
                self._h = None
        except Exception:
            pass


class Client:
    """client.rs:21-283."""

    def __init__(self, handle, K: int, N: int, b: int):
        self._h, self.K, self.N, self.b = handle, K, N, b
        self.pending = {}

    @staticmethod
    def setup(seed: bytes, hint: bytes, filter_bytes: bytes, lwe_rows: int = 0, rng_seed: int = 7):
        h = C.c_void_p()
        hb, fb = _u8(hint), _u8(filter_bytes)
        _chk(lib().orc_client_setup(_p(_u8(seed)), _p(hb), C.c_size_t(len(hb)), _p(fb), C.c_size_t(len(fb)), C.c_uint32(lwe_rows), C.c_uint64(rng_seed), C.byref(h)))
        f = Filter.from_bytes(filter_bytes)
        N = int(np.frombuffer(hint[4:8], dtype="<u4")[0])
        return Client(h, f.num_fingerprints, N, f.mat_elem_bit_len)

    def query(self, key: bytes) -> bytes:
        if key in self.pending:
            raise OracleError(-1)
        q = np.empty(8 + 4 * self.K, dtype=np.uint8)
        c = np.empty(self.N, dtype=np.uint32)
        k = _u8(key)
        _chk(lib().orc_client_query(self._h, _p(k), C.c_size_t(len(k)), _p(q), _p(c)))
        self.pending[key] = c
        return q.tobytes()

    def query_with(self, key: bytes, secret_s: np.ndarray, error_e: np.ndarray) -> bytes:
        """client.rs:95-194 with the secret vector s and the error vector e supplied (the deterministic core of query)."""
        if key in self.pending:
            raise OracleError(-1)
        q = np.empty(8 + 4 * self.K, dtype=np.uint8)
        c = np.empty(self.N, dtype=np.uint32)
        k = _u8(key)
        s = np.ascontiguousarray(secret_s, dtype=np.uint32)
        e = np.ascontiguousarray(error_e, dtype=np.uint32)
        _chk(lib().orc_client_query_with(self._h, _p(k), C.c_size_t(len(k)), _p(s), _p(e), _p(q), _p(c)))
        self.pending[key] = c
        return q.tobytes()

    def process_response(self, key: bytes, resp: bytes) -> bytes:
        c = self.pending[key]
        out = np.zeros((self.N * self.b) // 8 + 8, dtype=np.uint8)
        n = C.c_uint64()
        k, r = _u8(key), _u8(resp)
        rc = lib().orc_client_process_response(self._h, _p(k), C.c_size_t(len(k)), _p(c), _p(r), C.c_size_t(len(r)), _p(out), C.byref(n))
        del self.pending[key]
        _chk(rc)
        return out[: n.value].tobytes()

    def __del__(self):
        try:
            if self._h:
                lib().orc_client_free(self._h)
                self._h = None
        except Exception:
            pass


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> int:
    """Resize the OpenMP pool (the stand-in for rayon's one-worker-per-hardware-thread pool); returns the size in effect."""
    lib().orc_set_num_threads.restype = C.c_int
    lib().orc_set_num_threads.argtypes = [C.c_int]
    return lib().orc_set_num_threads(int(n))
