"""Host-side mirror of ``chalametpir_client::Client`` (chalametpir_client/src/client.rs:21-283) over the C ABI.

    client = Client.setup(seed_mu, hint_bytes, filter_param_bytes)      # client.rs:39
    query_bytes = client.query(key)                                     # client.rs:95   (may raise ArithmeticOverflowAddingQueryIndicator: retry)
    value = client.process_response(key, response_bytes)                # client.rs:209

The public matrix A lives in HBM and ``s*A + e`` is computed there (csrc/client.cu); this module only marshals buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from ._lib import ClientInfo, ClientOpts, lib
from .errors import check
from .server import _seed_arr, get_ctx


class Client:
    def __init__(self, handle: C.c_void_p, device: int):
        self._h = handle
        self.device = device
        info = ClientInfo()
        check(lib.chpir_client_get_info(self._h, C.byref(info)))
        self.rows_k, self.cols_n, self.lwe_rows = info.rows_k, info.cols_n, info.lwe_rows
        self.mat_elem_bit_len, self.arity = info.mat_elem_bit_len, info.arity

    @staticmethod
    def setup(seed_mu: bytes, hint_bytes: bytes, filter_param_bytes: bytes, *, device: int = 0, lwe_rows: int = 0, a_expand: str = "auto",
              host_chunk_rows: int = 0) -> "Client":
        """Client::setup(seed_mu, hint_bytes, filter_param_bytes)   [client.rs:39-57]"""
        seed = _seed_arr(seed_mu)
        hint = np.frombuffer(hint_bytes, dtype=np.uint8)
        fb = np.frombuffer(filter_param_bytes, dtype=np.uint8)
        o = ClientOpts(lwe_rows, {"auto": 0, "host": 1, "device": 2}[a_expand], host_chunk_rows)
        h = C.c_void_p()
        check(lib.chpir_client_setup(get_ctx(device), seed.ctypes.data, hint.ctypes.data if hint.size else None, hint.size,
                                     fb.ctypes.data if fb.size else None, fb.size, C.byref(o), C.byref(h)))
        return Client(h, device)

    def query(self, key: bytes, rng_seed: Optional[int] = None) -> bytes:
        """Client::query(key) -> query bytes   [client.rs:95-194]; rng_seed = None draws from OS entropy like the reference."""
        k = np.frombuffer(key, dtype=np.uint8)
        out = np.empty(8 + 4 * self.rows_k, dtype=np.uint8)
        n = C.c_size_t()
        seed = C.c_uint64(rng_seed) if rng_seed is not None else None
        check(lib.chpir_client_query(self._h, k.ctypes.data if k.size else None, k.size, C.byref(seed) if seed is not None else None,
                                     out.ctypes.data, out.nbytes, C.byref(n)))
        return out[: n.value].tobytes()

    def query_with(self, key: bytes, secret_s: np.ndarray, error_e: np.ndarray) -> bytes:
        """The deterministic core of query: b = s*A + e (+ indicator), with s and e supplied by the caller."""
        k = np.frombuffer(key, dtype=np.uint8)
        s = np.ascontiguousarray(secret_s, dtype=np.uint32)
        e = np.ascontiguousarray(error_e, dtype=np.uint32)
        if s.size != self.lwe_rows or e.size != self.rows_k:
            raise ValueError("secret_s must have lwe_rows words and error_e K words")
        out = np.empty(8 + 4 * self.rows_k, dtype=np.uint8)
        n = C.c_size_t()
        check(lib.chpir_client_query_with(self._h, k.ctypes.data if k.size else None, k.size, s.ctypes.data, e.ctypes.data, out.ctypes.data,
                                          out.nbytes, C.byref(n)))
        return out[: n.value].tobytes()

    def process_response(self, key: bytes, response_bytes: bytes) -> bytes:
        """Client::process_response(key, response_bytes) -> value   [client.rs:209-275]"""
        k = np.frombuffer(key, dtype=np.uint8)
        r = np.frombuffer(response_bytes, dtype=np.uint8)
        out = np.empty((self.cols_n * self.mat_elem_bit_len) // 8 + 8, dtype=np.uint8)
        n = C.c_size_t()
        check(lib.chpir_client_process_response(self._h, k.ctypes.data if k.size else None, k.size, r.ctypes.data if r.size else None, r.size,
                                                out.ctypes.data, out.nbytes, C.byref(n)))
        return out[: n.value].tobytes()

    def info(self) -> dict:
        i = ClientInfo()
        check(lib.chpir_client_get_info(self._h, C.byref(i)))
        return {n: getattr(i, n) for n, _ in i._fields_}

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.chpir_client_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
