"""Measured dense int8 tensor-core throughput of this GPU (cuBLASLt through torch._int_mm), the denominator for the hint GEMM's
tensor roofline when MEASURED_PEAKS.json carries no int8 figure."""
import torch

n = 8192
a = torch.randint(-100, 100, (n, n), dtype=torch.int8, device="cuda")
b = torch.randint(-100, 100, (n, n), dtype=torch.int8, device="cuda")
for _ in range(3):
    torch._int_mm(a, b)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch._int_mm(a, b)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"int8 {n}^3: best {best:.3f} ms = {2 * n**3 / best / 1e9:.1f} TOP/s")
a16, b16 = a.to(torch.bfloat16), b.to(torch.bfloat16)
for _ in range(3):
    a16 @ b16
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    a16 @ b16
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"bf16 {n}^3: best {best:.3f} ms = {2 * n**3 / best / 1e9:.1f} TFLOP/s")
