"""Scratch: the per-rank respond kernel of an N-way column-sharded server, reproduced on ONE GPU (a rank's kernel only sees its
own column slice), for batches of 16 queries; sweeps how many queries one CTA lifetime covers and the ring geometry."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import chalametpir_b200 as cp  # noqa: E402

SEED = bytes(range(32))
Q = 16


def run(ncols, env, iters=30):
    for k in list(os.environ):
        if k.startswith("CHPIR_R") or k.startswith("CHPIR_T"):
            del os.environ[k]
    os.environ.update({k: str(v) for k, v in env.items()})
    torch.manual_seed(ncols)
    n = 1 << 20
    b = cp.find_mat_elem_bit_len(n)
    K, _ = cp.db_matrix_shape(3, n, 1024, b)
    D = torch.randint(0, 1 << b, (K, ncols), dtype=torch.int32, device="cuda")
    srv, _ = cp.Server.setup_from_device_matrix(SEED, D.data_ptr(), K, ncols, b, skip_hint=True)
    del D
    q = torch.randint(-2**31, 2**31 - 1, (Q, K), dtype=torch.int32, device="cuda")
    r = torch.empty((Q, ncols), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(5):
        srv.respond_device(q.data_ptr(), Q, r.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        srv.respond_device(q.data_ptr(), Q, r.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (iters * Q) * 1e3
    tight = K * ncols * b / 8
    print(f"ncols={ncols:4d} {str(env):70s} {us:7.2f} us/query  {srv.packed_bytes / us / 1e3:6.0f} GB/s packed  {tight / us / 1e3:6.0f} GB/s of b-bit payload", flush=True)
    out = r.clone()
    srv.close()
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "tight":  # 8-byte row pitch against 16-byte padded rows, interleaved repeats
        for nc in (118, 117):
            for rep in range(4):
                for tp in (1, 0):
                    run(nc, {"CHPIR_TIGHT_PITCH": tp}, iters=100)
        for tp in (1, 0):
            run(118, {"CHPIR_TIGHT_PITCH": tp, "CHPIR_RING_RPT": 2}, iters=100)
            run(118, {"CHPIR_TIGHT_PITCH": tp, "CHPIR_RING_R": 48}, iters=100)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "rows":  # row lanes per stage at full width and on the slices, interleaved repeats
        for rep in range(3):
            for R in (8, 12):
                run(940, {"CHPIR_RING_R": R}, iters=60)
        for nc, Rs in ((470, (16, 20, 24)), (235, (32, 40)), (118, (64, 56))):
            for rep in range(2):
                for R in Rs:
                    run(nc, {"CHPIR_RING_R": R}, iters=60)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "grid":
        for nc in (118, 235, 470, 940):
            for env in [{}, {"CHPIR_RING_GRID_MULT": 2}, {"CHPIR_RING_GRID_MULT": 3}, {"CHPIR_RING_GRID_MULT": 4}]:
                run(nc, env)
        for env in [{"CHPIR_RING_R": 12}, {"CHPIR_RING_R": 12, "CHPIR_RING_GRID_MULT": 2}, {"CHPIR_RING_RPT": 2, "CHPIR_RING_GRID_MULT": 2}]:
            run(940, env)
        sys.exit(0)
    for nc in (118, 235, 470, 940):
        ref = None
        for qpc in (1, 2, 4, 8, 16):
            got = run(nc, {"CHPIR_RING_Q_PER_CTA": qpc})
            ref = got if ref is None else ref
            assert torch.equal(got, ref)
    for env in [{"CHPIR_RING_Q_PER_CTA": 1, "CHPIR_RING_R": 32}, {"CHPIR_RING_Q_PER_CTA": 1, "CHPIR_RING_R": 48}, {"CHPIR_RING_Q_PER_CTA": 1, "CHPIR_RING_RPT": 2},
                {"CHPIR_RING_Q_PER_CTA": 1, "CHPIR_RING_R": 32, "CHPIR_RING_RPT": 2}, {"CHPIR_RING_Q_PER_CTA": 4, "CHPIR_RING_R": 32},
                {"CHPIR_RING_Q_PER_CTA": 1, "CHPIR_RING_STAGES": 3},
                {"CHPIR_RING_Q_PER_CTA": 1, "CHPIR_RING_R": 32, "CHPIR_RING_BUDGET_KB": 100, "CHPIR_RING_GRID_MULT": 2},
                {"CHPIR_RING_Q_PER_CTA": 4, "CHPIR_RING_R": 32, "CHPIR_RING_BUDGET_KB": 100, "CHPIR_RING_GRID_MULT": 2},
                {"CHPIR_RING_Q_PER_CTA": 16, "CHPIR_RING_R": 32, "CHPIR_RING_BUDGET_KB": 100, "CHPIR_RING_GRID_MULT": 2},
                {"CHPIR_RING_Q_PER_CTA": 1, "CHPIR_RING_R": 28, "CHPIR_RING_BUDGET_KB": 72, "CHPIR_RING_GRID_MULT": 3}]:
        run(118, env)
