"""Workload for ncu: one launch (or a few) of every product kernel at the BASELINE configs[2] shape (2^20 entries, 3-wise, b = 9).

  ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel|vec_x_mat|split_transpose_b|pack_kernel|split_a_limbs|fill_wave' \
      -c 40 -o gpurun_out/prof_kernels python tools/profile_kernels.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chalametpir_b200 as cp

SEED = bytes(range(32))
n = 1 << 20
b = cp.find_mat_elem_bit_len(n)
K, N = cp.db_matrix_shape(3, n, 1024, b)
g = torch.Generator(device="cuda").manual_seed(1)
D = torch.randint(0, 1 << b, (K, N), dtype=torch.int32, device="cuda", generator=g)
# one 128-row panel of A (host-expanded): pack + split_transpose_b + split_a_limbs + one gemm_tc_kernel launch
srv, hint = cp.Server.setup_from_device_matrix(SEED, D.data_ptr(), K, N, b, lwe_rows=128, a_expand="host", batch_tc=1)
q = torch.randint(-(2**31), 2**31, (128, K), dtype=torch.int32, device="cuda", generator=g)
r = torch.empty((128, N), dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
srv.respond_device_tc(q.data_ptr(), 128, r.data_ptr(), st)  # batched respond: split_a_limbs + gemm_tc_kernel
srv.respond_device(q.data_ptr(), 16, r.data_ptr(), st)      # respond_ring_kernel, 16 queries
torch.cuda.synchronize()
del D, q, r
# client: vec_x_mat_kernel over a 128-row A (same kernel, fewer rows)
client = cp.Client.setup(SEED, hint, bytes(32) + np.array([3, 8192, 1163264], dtype="<u4").tobytes() + np.array([K, n, b], dtype="<u8").tobytes(),
                         lwe_rows=128, a_expand="host")
client.query(b"some key", rng_seed=1)
# device row fill on a small real database (a few thousand waves)
rs = np.random.default_rng(0)
keys = rs.integers(0, 256, size=(1 << 14, 32), dtype=np.uint8)
keys[:, :4] = np.arange(1 << 14, dtype="<u4").view(np.uint8).reshape(-1, 4)
vals = rs.integers(0, 256, size=(1 << 14, 1024), dtype=np.uint8)
cp.Server.setup_from_arrays(SEED, keys, vals, 3, lwe_rows=16, a_expand="host", db_encode="device")
print("done")
