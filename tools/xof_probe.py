"""Scratch: XOF rate on the GPU box."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chalametpir_b200 as cp
SEED = bytes(range(32))
for mb in (8, 64, 256):
    cols = mb * 1024 * 1024 // 4 // 16
    t = time.time()
    a = cp.generate_from_seed(16, cols, SEED, row_begin=15, row_count=1)
    dt = time.time() - t
    blocks = 16 * cols * 4 / 168
    print(f"xof {mb} MB: {dt:.3f} s  {dt/blocks*1e9:.0f} ns/permutation  {16*cols*4/dt/1e6:.1f} MB/s", flush=True)
