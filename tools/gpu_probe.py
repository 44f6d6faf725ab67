"""Scratch measurements on the GPU box (not a bench): respond kernel time at the BASELINE shapes, XOF rate."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import chalametpir_b200 as cp

SEED = bytes(range(32))
def respond_probe(n_log2, arity):
    n = 1 << n_log2
    b = cp.find_mat_elem_bit_len(n)
    K, N = cp.db_matrix_shape(arity, n, 1024, b)
    D = torch.randint(0, 1 << b, (K, N), dtype=torch.int32, device="cuda")
    srv, _ = cp.Server.setup_from_device_matrix(SEED, D.data_ptr(), K, N, b, skip_hint=True)
    del D
    q = torch.randint(-2**31, 2**31 - 1, (K,), dtype=torch.int32, device="cuda")
    r = torch.empty((N,), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(5):
        srv.respond_device(q.data_ptr(), 1, r.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 50
    e0.record()
    for _ in range(iters):
        srv.respond_device(q.data_ptr(), 1, r.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"respond 2^{n_log2} arity {arity}: K={K} N={N} b={b} packed={srv.packed_bytes/1e9:.3f} GB  {ms*1e3:.1f} us/query  "
          f"{srv.packed_bytes/ms/1e6:.0f} GB/s  {1e3/ms:.0f} q/s", flush=True)
    # host path
    qh = np.random.default_rng(0).integers(0, 2**32, size=K, dtype=np.uint64).astype(np.uint32)
    qb = np.array([1, K], dtype="<u4").tobytes() + qh.tobytes()
    srv.respond(qb)
    t = time.time()
    for _ in range(20): srv.respond(qb)
    print(f"   host-buffer respond (pageable): {(time.time()-t)/20*1e3:.3f} ms/query  kernel {srv.last_kernel_ms()['respond_ms']*1e3:.1f} us", flush=True)
    srv.close()

def xof_probe(mb):
    cols = mb * 1024 * 1024 // 4 // 16
    t = time.time()
    a = cp.generate_from_seed(16, cols, SEED, row_begin=15, row_count=1)
    dt = time.time() - t
    blocks = 16 * cols * 4 / 168
    print(f"xof {mb} MB: {dt:.3f} s  {dt/blocks*1e9:.0f} ns/permutation  {16*cols*4/dt/1e6:.1f} MB/s", flush=True)

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    xof_probe(8); xof_probe(64)
    for n_log2, arity in [(16, 3), (18, 3), (20, 3), (20, 4)]:
        respond_probe(n_log2, arity)
