"""ctypes loader for libchalamet_b200.so (the C ABI declared in include/chalamet_b200.h).

There is no fallback: if the shared library has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)
or cannot be loaded, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libchalamet_b200.so")

LWE_DIMENSION = 1774
SEED_BYTE_LEN = 32
FILTER_PARAM_BYTE_LEN = 68
SERVER_SETUP_MAX_ATTEMPT_COUNT = 100


class SetupOpts(C.Structure):
    _fields_ = [
        ("lwe_rows", C.c_uint32),
        ("col_begin", C.c_uint32),
        ("col_count", C.c_uint32),
        ("gemm_variant", C.c_uint32),
        ("skip_hint", C.c_uint32),
        ("batch_tc", C.c_uint32),
        ("a_expand", C.c_uint32),
        ("host_chunk_rows", C.c_uint32),
        ("respond_coalesce", C.c_uint32),
        ("db_encode", C.c_uint32),
        ("a_cache", C.c_uint32),
        ("hint_on_device", C.c_uint32),
    ]


A_EXPAND_AUTO = 0
A_EXPAND_HOST_PIPELINED = 1
A_EXPAND_DEVICE = 2


class SetupTiming(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("host_encode_s", "h2d_s", "pack_s", "expand_a_s", "gemm_s", "d2h_s", "total_s", "device_encode_s", "xof_host_busy_s", "a_cache_hit", "xof_host_wait_s")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class ServerInfo(C.Structure):
    _fields_ = [
        ("rows_k", C.c_uint64),
        ("cols_n", C.c_uint32),
        ("col_begin", C.c_uint32),
        ("mat_elem_bit_len", C.c_uint32),
        ("fields_per_word", C.c_uint32),
        ("row_pitch_bytes", C.c_uint64),
        ("packed_bytes", C.c_uint64),
    ]


class ClusterServerInfo(C.Structure):
    _fields_ = [
        ("n_gpus", C.c_uint32),
        ("cols_n", C.c_uint32),
        ("rows_k", C.c_uint64),
        ("mat_elem_bit_len", C.c_uint32),
        ("lwe_rows", C.c_uint32),
        ("k_pitch", C.c_uint64),
        ("packed_bytes_total", C.c_uint64),
        ("packed_bytes_max_rank", C.c_uint64),
        ("setup_total_s", C.c_double),
        ("hint_gather_s", C.c_double),
        ("gather_uses_nccl", C.c_uint32),
        ("nccl_version", C.c_uint32),
        ("batches", C.c_uint64),
        ("queries", C.c_uint64),
        ("tc_batches", C.c_uint64),
        ("pulled_queries", C.c_uint64),
        ("respond_by_rows", C.c_uint32),
        ("reserved0", C.c_uint32),
        ("reshard_s", C.c_double),
        ("ingest_wait_s", C.c_double),
        ("ingest_s", C.c_double),
        ("exec_wait_s", C.c_double),
        ("exec_s", C.c_double),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class ClientOpts(C.Structure):
    _fields_ = [("lwe_rows", C.c_uint32), ("a_expand", C.c_uint32), ("host_chunk_rows", C.c_uint32)]


class ClientInfo(C.Structure):
    _fields_ = [
        ("rows_k", C.c_uint64),
        ("cols_n", C.c_uint32),
        ("lwe_rows", C.c_uint32),
        ("mat_elem_bit_len", C.c_uint32),
        ("arity", C.c_uint32),
        ("pub_mat_a_bytes", C.c_uint64),
        ("setup_expand_s", C.c_double),
        ("last_query_kernel_ms", C.c_float),
    ]


# every symbol include/chalamet_b200.h declares
EXPORTS = [
    "chpir_strerror",
    "chpir_last_cuda_error",
    "chpir_device_count",
    "chpir_ctx_create",
    "chpir_ctx_destroy",
    "chpir_ctx_drop_a_cache",
    "chpir_host_alloc",
    "chpir_host_free",
    "chpir_upload_rows",
    "chpir_find_mat_elem_bit_len",
    "chpir_db_matrix_shape",
    "chpir_encode_kv_database",
    "chpir_encode_kv_database_device",
    "chpir_server_setup",
    "chpir_server_setup_device",
    "chpir_server_setup_from_db",
    "chpir_server_destroy",
    "chpir_server_save",
    "chpir_server_load",
    "chpir_server_setup_timing",
    "chpir_server_get_info",
    "chpir_server_hint_device",
    "chpir_cluster_create",
    "chpir_cluster_destroy",
    "chpir_cluster_size",
    "chpir_cluster_ctx",
    "chpir_cluster_plan",
    "chpir_cluster_server_setup_from_db",
    "chpir_cluster_server_setup",
    "chpir_cluster_server_setup_device",
    "chpir_cluster_server_destroy",
    "chpir_cluster_server_shard",
    "chpir_cluster_server_save",
    "chpir_cluster_server_load",
    "chpir_cluster_server_respond",
    "chpir_cluster_server_respond_batch",
    "chpir_cluster_server_respond_device",
    "chpir_cluster_server_respond_concurrent",
    "chpir_cluster_server_get_info",
    "chpir_server_respond",
    "chpir_server_respond_batch",
    "chpir_server_respond_device",
    "chpir_server_respond_device_tc",
    "chpir_generate_from_seed",
    "chpir_matmul",
    "chpir_server_last_kernel_ms",
    "chpir_host_generate_from_seed",
    "chpir_host_xof_impl",
    "chpir_client_setup",
    "chpir_client_destroy",
    "chpir_client_query",
    "chpir_client_query_with",
    "chpir_client_process_response",
    "chpir_client_get_info",
]

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
        "chalametpir_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

_vp = C.c_void_p
_szp = C.POINTER(C.c_size_t)

lib.chpir_strerror.restype = C.c_char_p
lib.chpir_strerror.argtypes = [C.c_int]
lib.chpir_last_cuda_error.restype = C.c_char_p
lib.chpir_last_cuda_error.argtypes = []
lib.chpir_device_count.argtypes = [C.POINTER(C.c_int)]
lib.chpir_ctx_create.argtypes = [C.c_int, C.POINTER(_vp)]
lib.chpir_ctx_destroy.restype = None
lib.chpir_ctx_destroy.argtypes = [_vp]
lib.chpir_ctx_drop_a_cache.restype = C.c_int
lib.chpir_ctx_drop_a_cache.argtypes = [_vp, C.POINTER(C.c_uint64)]
lib.chpir_host_alloc.argtypes = [C.c_size_t, C.POINTER(_vp)]
lib.chpir_host_free.restype = None
lib.chpir_host_free.argtypes = [_vp]
lib.chpir_upload_rows.restype = C.c_int
lib.chpir_upload_rows.argtypes = [_vp, C.c_size_t, _vp, C.c_size_t, C.c_size_t, C.c_size_t, _vp]
lib.chpir_find_mat_elem_bit_len.argtypes = [C.c_uint64, C.POINTER(C.c_uint32)]
lib.chpir_db_matrix_shape.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
lib.chpir_encode_kv_database.argtypes = [C.c_uint32, C.c_uint64, _vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), _vp, _vp]
lib.chpir_encode_kv_database_device.argtypes = [_vp, C.c_uint32, C.c_uint64, _vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), _vp, _vp]
lib.chpir_server_setup.argtypes = [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(SetupOpts), _vp, C.c_size_t, _szp, C.POINTER(_vp)]
lib.chpir_server_setup_device.argtypes = lib.chpir_server_setup.argtypes
lib.chpir_server_setup_from_db.argtypes = [
    _vp, C.c_uint32, _vp, C.c_uint64, _vp, _vp, _vp, _vp, C.POINTER(C.c_uint64), C.POINTER(SetupOpts), _vp, C.c_size_t, _szp, _vp, C.POINTER(_vp),
]
lib.chpir_server_destroy.restype = None
lib.chpir_server_destroy.argtypes = [_vp]
lib.chpir_server_save.argtypes = [_vp, C.c_char_p]
lib.chpir_server_load.argtypes = [_vp, C.c_char_p, C.POINTER(SetupOpts), C.POINTER(_vp)]
lib.chpir_server_setup_timing.argtypes = [_vp, C.POINTER(SetupTiming)]
lib.chpir_server_get_info.argtypes = [_vp, C.POINTER(ServerInfo)]
lib.chpir_server_respond.argtypes = [_vp, _vp, C.c_size_t, _vp, C.c_size_t, _szp]
lib.chpir_server_respond_batch.argtypes = [_vp, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_uint32, _vp, C.c_size_t]
lib.chpir_server_respond_device.argtypes = [_vp, _vp, C.c_uint32, _vp, _vp]
lib.chpir_server_respond_device_tc.argtypes = [_vp, _vp, C.c_uint32, _vp, _vp]
lib.chpir_generate_from_seed.argtypes = [_vp, _vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _vp]
lib.chpir_matmul.argtypes = [_vp, _vp, C.c_uint64, C.c_uint64, _vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, _vp]
lib.chpir_host_generate_from_seed.argtypes = [_vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, _vp]
lib.chpir_host_xof_impl.restype = C.c_char_p
lib.chpir_host_xof_impl.argtypes = []
lib.chpir_client_setup.argtypes = [_vp, _vp, _vp, C.c_size_t, _vp, C.c_size_t, C.POINTER(ClientOpts), C.POINTER(_vp)]
lib.chpir_client_destroy.restype = None
lib.chpir_client_destroy.argtypes = [_vp]
lib.chpir_client_query.argtypes = [_vp, _vp, C.c_size_t, C.POINTER(C.c_uint64), _vp, C.c_size_t, _szp]
lib.chpir_client_query_with.argtypes = [_vp, _vp, C.c_size_t, _vp, _vp, _vp, C.c_size_t, _szp]
lib.chpir_client_process_response.argtypes = [_vp, _vp, C.c_size_t, _vp, C.c_size_t, _vp, C.c_size_t, _szp]
lib.chpir_client_get_info.argtypes = [_vp, C.POINTER(ClientInfo)]
lib.chpir_server_last_kernel_ms.argtypes = [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
lib.chpir_server_hint_device.argtypes = [_vp, C.POINTER(_vp), C.POINTER(C.c_uint32)]
_u32p, _u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
lib.chpir_cluster_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(_vp)]
lib.chpir_cluster_destroy.restype = None
lib.chpir_cluster_destroy.argtypes = [_vp]
lib.chpir_cluster_size.argtypes = [_vp, C.POINTER(C.c_int)]
lib.chpir_cluster_ctx.argtypes = [_vp, C.c_int, C.POINTER(_vp), C.POINTER(C.c_int)]
lib.chpir_cluster_plan.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, _u32p, _u32p, _u64p, _u64p, _u64p]
lib.chpir_cluster_server_setup_from_db.argtypes = [
    _vp, C.c_uint32, _vp, C.c_uint64, _vp, _vp, _vp, _vp, C.POINTER(C.c_uint64), C.POINTER(SetupOpts), _vp, C.c_size_t, _szp, _vp, C.POINTER(_vp),
]
lib.chpir_cluster_server_setup.argtypes = [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(SetupOpts), _vp, C.c_size_t, _szp, C.POINTER(_vp)]
lib.chpir_cluster_server_setup_device.argtypes = [_vp, _vp, C.POINTER(_vp), C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(SetupOpts), _vp, C.c_size_t, _szp, C.POINTER(_vp)]
lib.chpir_cluster_server_destroy.restype = None
lib.chpir_cluster_server_destroy.argtypes = [_vp]
lib.chpir_cluster_server_shard.argtypes = [_vp, C.c_int, C.POINTER(_vp)]
lib.chpir_cluster_server_save.argtypes = [_vp, C.c_char_p]
lib.chpir_cluster_server_load.argtypes = [_vp, C.c_char_p, C.POINTER(SetupOpts), C.POINTER(_vp)]
lib.chpir_cluster_server_respond.argtypes = [_vp, _vp, C.c_size_t, _vp, C.c_size_t, _szp]
lib.chpir_cluster_server_respond_batch.argtypes = [_vp, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_uint32, _vp, C.c_size_t]
lib.chpir_cluster_server_respond_device.argtypes = [_vp, C.POINTER(_vp), C.c_uint32, _vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
lib.chpir_cluster_server_get_info.argtypes = [_vp, C.POINTER(ClusterServerInfo)]
lib.chpir_cluster_server_respond_concurrent.argtypes = [_vp, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_uint32, C.c_uint64, _vp, C.c_size_t, C.c_uint32, C.POINTER(C.c_double)]
