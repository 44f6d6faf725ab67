"""Workload for ncu: the respond kernel on a 118-column slice (8-way sharding), padded rows then tight rows, 16 queries per launch.
  ncu --set full --clock-control none --import-source on -k regex:respond_ring_kernel -c 4 -o gpurun_out/slice118 python tools/slice_ncu.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import chalametpir_b200 as cp  # noqa: E402

n = 1 << 20
b = cp.find_mat_elem_bit_len(n)
K, _ = cp.db_matrix_shape(3, n, 1024, b)
ncols = int(sys.argv[1]) if len(sys.argv) > 1 else 118
torch.manual_seed(1)
D = torch.randint(0, 1 << b, (K, ncols), dtype=torch.int32, device="cuda")
q = torch.randint(-2**31, 2**31 - 1, (16, K), dtype=torch.int32, device="cuda")
r = torch.empty((16, ncols), dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for tight in ("0", "1"):
    os.environ["CHPIR_TIGHT_PITCH"] = tight
    srv, _ = cp.Server.setup_from_device_matrix(bytes(32), D.data_ptr(), K, ncols, b, skip_hint=True)
    for _ in range(2):
        srv.respond_device(q.data_ptr(), 16, r.data_ptr(), st)
    torch.cuda.synchronize()
    srv.close()
print("done")
