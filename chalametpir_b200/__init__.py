"""chalametpir_b200 -- B200-native (sm_100a) server hot path of ChalametPIR.

Drop-in for ``chalametpir_server::Server`` behind the reference's ``gpu`` feature: hint, filter-parameter and response
byte layouts are unchanged.  The compute lives in ``libchalamet_b200.so`` (hand-written CUDA, C ABI in
``include/chalamet_b200.h``); importing this package fails if that library has not been built.
"""
from ._lib import FILTER_PARAM_BYTE_LEN, LIB_PATH, LWE_DIMENSION, SEED_BYTE_LEN, SERVER_SETUP_MAX_ATTEMPT_COUNT
from .client import Client
from .cluster import RESPOND_GEMV, RESPOND_TC, Cluster, ClusterServer, cluster_plan
from .errors import ChalametPIRError
from .server import (
    PinnedBuffer,
    Server,
    db_matrix_shape,
    device_count,
    drop_a_cache,
    encode_kv_database,
    encode_kv_database_device,
    find_mat_elem_bit_len,
    generate_from_seed,
    get_ctx,
    host_generate_from_seed,
    host_xof_impl,
    matmul,
)

__all__ = [
    "Server",
    "Cluster",
    "ClusterServer",
    "cluster_plan",
    "RESPOND_GEMV",
    "RESPOND_TC",
    "Client",
    "PinnedBuffer",
    "ChalametPIRError",
    "LWE_DIMENSION",
    "SEED_BYTE_LEN",
    "FILTER_PARAM_BYTE_LEN",
    "SERVER_SETUP_MAX_ATTEMPT_COUNT",
    "LIB_PATH",
    "db_matrix_shape",
    "device_count",
    "drop_a_cache",
    "encode_kv_database",
    "encode_kv_database_device",
    "find_mat_elem_bit_len",
    "generate_from_seed",
    "get_ctx",
    "host_generate_from_seed",
    "host_xof_impl",
    "matmul",
]
