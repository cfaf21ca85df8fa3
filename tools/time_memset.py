"""Write-only ceiling: torch's fill kernel and cudaMemsetAsync on buffers of the
renderer's output sizes (graph-captured, CUDA-event timed)."""
import torch
for mb in (164, 655):
    n = mb * 1000 * 1000
    x = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, fn in (("zero_", lambda: x.zero_()), ("fill_(7)", lambda: x.fill_(7))):
        for _ in range(3): fn()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s): fn()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            for _ in range(10): fn()
        g.replay(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 10)
        print(f"{name:9s} {mb} MB: {best*1e3:.1f} us  {n/best/1e6:.0f} GB/s", flush=True)
