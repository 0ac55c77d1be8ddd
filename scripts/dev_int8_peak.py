"""Developer probe (GPU box): cuBLASLt int8 and bf16 GEMM throughput on this chip (context for the roofline)."""
import torch, time
dev = "cuda"
def bench(fn, flops, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return flops / ms / 1e9
for n in (4096, 8192):
    a = torch.randint(-64, 63, (n, n), dtype=torch.int8, device=dev)
    b = torch.randint(-64, 63, (n, n), dtype=torch.int8, device=dev)
    try:
        print(f"int8 _int_mm {n}^3: {bench(lambda: torch._int_mm(a, b.t()), 2*n**3):.0f} TOPS", flush=True)
    except Exception as ex:
        print("int8 failed", ex)
    x = torch.randn(n, n, dtype=torch.bfloat16, device=dev); y = torch.randn(n, n, dtype=torch.bfloat16, device=dev)
    print(f"bf16 matmul {n}^3: {bench(lambda: x @ y, 2*n**3):.0f} TFLOPS", flush=True)
