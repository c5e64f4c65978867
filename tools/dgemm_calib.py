"""cuBLAS DGEMM calibration (not on the product path): FP64 GEMM rate this box reaches, burst and sustained."""
import torch, time, json, subprocess
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"
res = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    for _ in range(3): torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}_burst_tflops"] = 2 * n**3 / best * 1e-9
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 20 if n == 8192 else 100
    e0.record()
    for _ in range(reps): torch.matmul(a, b)
    e1.record(); torch.cuda.synchronize()
    res[f"dgemm_{n}_sustained_tflops"] = 2 * n**3 * reps / e0.elapsed_time(e1) * 1e-9
print(json.dumps(res))
print(subprocess.run("nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv", shell=True, capture_output=True, text=True).stdout)
