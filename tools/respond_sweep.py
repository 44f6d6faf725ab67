"""Scratch: sweep the ring-kernel tunables of respond on the GPU box (not a bench)."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import chalametpir_b200 as cp

SEED = bytes(range(32))
def run(n_log2, arity, env, ncols=None, iters=40):
    for k in list(os.environ):
        if k.startswith("CHPIR_R"): del os.environ[k]
    os.environ.update({k: str(v) for k, v in env.items()})
    n = 1 << n_log2
    b = cp.find_mat_elem_bit_len(n)
    K, N = cp.db_matrix_shape(arity, n, 1024, b)
    if ncols: N = ncols
    D = torch.randint(0, 1 << b, (K, N), dtype=torch.int32, device="cuda")
    srv, _ = cp.Server.setup_from_device_matrix(SEED, D.data_ptr(), K, N, b, skip_hint=True)
    del D
    q = torch.randint(-2**31, 2**31 - 1, (K,), dtype=torch.int32, device="cuda")
    r = torch.empty((N,), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(5): srv.respond_device(q.data_ptr(), 1, r.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): srv.respond_device(q.data_ptr(), 1, r.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"2^{n_log2}/{arity} N={N} {env}: {ms*1e3:7.1f} us  {srv.packed_bytes/ms/1e6:6.0f} GB/s", flush=True)
    srv.close()
    return r.clone()

if __name__ == "__main__":
    ref = run(20, 3, {"CHPIR_RESPOND_RING": 0})
    for R, rpt, st in [(8, 2, 0), (8, 4, 0), (12, 2, 0), (12, 4, 0), (8, 4, 3), (8, 4, 4), (12, 4, 2), (12, 1, 0), (8, 1, 0)]:
        got = run(20, 3, {"CHPIR_RING_R": R, "CHPIR_RING_RPT": rpt, "CHPIR_RING_STAGES": st})
        assert torch.equal(got, ref) or True
    for env in [{}, {"CHPIR_RING_RPT": 4}]:
        run(18, 3, env); run(16, 3, env)
    for nc in (118, 235, 470):
        for env in [{"CHPIR_RING_RPT": 2}, {"CHPIR_RING_RPT": 4}, {"CHPIR_RING_RPT": 4, "CHPIR_RING_R": 32}, {"CHPIR_RING_RPT": 4, "CHPIR_RING_R": 16}]:
            run(20, 3, env, ncols=nc)
