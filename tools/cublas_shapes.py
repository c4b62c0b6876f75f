"""cuBLAS Dgemm on the shapes of the trailing update (C -= A B^T, C t x t, k pivots): the vendor kernel's
throughput on the same contraction is the practical ceiling k_update is compared with."""
import torch, json, sys
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
out = []
for t in (1024, 2048, 4096, 6016, 8192, 12288):
    for k in (128, 256, 512):
        A = torch.randn(t, k, dtype=torch.float64, device=dev)
        B = torch.randn(t, k, dtype=torch.float64, device=dev)
        C = torch.randn(t, t, dtype=torch.float64, device=dev)
        for _ in range(3): C.addmm_(A, B.t(), alpha=-1.0)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4): C.addmm_(A, B.t(), alpha=-1.0)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 4)
        tf = 2.0 * t * t * k / (best * 1e-3) / 1e12
        out.append({"t": t, "k": k, "us": best * 1e3, "tflops_full_square": tf})
        print(f"t={t:6d} k={k:4d}  {best*1e3:9.1f} us  {tf:6.2f} TF/s (full square)", flush=True)
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/cublas_shapes.json", "w"), indent=1)
