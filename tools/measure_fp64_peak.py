"""Measure the FP64 GEMM denominator on this B200: cuBLAS Dgemm through torch.matmul(float64),
8192^3, best-of-10 burst and a 4 s sustained loop (same method as MEASURED_PEAKS.json's bf16 entry).
Writes gpurun_out/fp64_peak.json."""
import json, os, time, torch
N = 8192
a = torch.randn(N, N, dtype=torch.float64, device="cuda")
b = torch.randn(N, N, dtype=torch.float64, device="cuda")
c = torch.empty_like(a)
for _ in range(3):
    torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * N**3 / (best * 1e-3) / 1e12
t0 = time.time(); it = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(5):
        torch.matmul(a, b, out=c)
    it += 5
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sust = 2 * N**3 * it / (e0.elapsed_time(e1) * 1e-3) / 1e12
out = {"fp64_tflops": round(burst, 2), "fp64_tflops_sustained": round(sust, 2), "n": N,
       "how": "torch.matmul float64 8192^3 (cuBLAS Dgemm): best of 10 (burst), back to back 4 s (sustained)",
       "gpu": torch.cuda.get_device_name(0)}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fp64_peak.json", "w"))
