"""Workload for ncu: what ONE rank of an 8-GPU cluster runs per batch in the row cut -- the streaming GEMV on its row block
(147 456 x 940, b = 9; 16 queries per launch) and the tensor-core limb GEMM on the same block (128 queries) -- plus the same two on
the whole 2^20 matrix (one GPU).
  ncu --set full --clock-control none --import-source on -k regex:"respond_ring_kernel|gemm_tc_kernel" -o gpurun_out/r2_rowblock python tools/rowblock_ncu.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import chalametpir_b200 as cp  # noqa: E402

b, N = 9, 940
st = torch.cuda.current_stream().cuda_stream
for K in (147456, 1179648):
    torch.manual_seed(K)
    D = torch.randint(0, 1 << b, (K, N), dtype=torch.int32, device="cuda")
    q = torch.randint(-2**31, 2**31 - 1, (128, K), dtype=torch.int32, device="cuda")
    r = torch.empty((128, N), dtype=torch.int32, device="cuda")
    srv, _ = cp.Server.setup_from_device_matrix(bytes(32), D.data_ptr(), K, N, b, skip_hint=True, batch_tc=1)
    del D
    for _ in range(2):  # the second launch of each kind is the one to read
        srv.respond_device(q.data_ptr(), 16, r.data_ptr(), st)
    for _ in range(2):
        srv.respond_device_tc(q.data_ptr(), 128, r.data_ptr(), st)
    torch.cuda.synchronize()
    srv.close()
    del q, r
    torch.cuda.empty_cache()
print("done")
