import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
n = 20000
d = torch.empty(n * n, dtype=torch.float64, device="cuda")
hp = torch.empty(n * n, dtype=torch.float64).pin_memory()
hu = torch.empty(n * n, dtype=torch.float64)
for name, h in (("pinned", hp), ("pageable", hu)):
    for _ in range(2):
        torch.cuda.synchronize(); t = time.perf_counter(); h.copy_(d); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"D2H {name}: {8*n*n/dt*1e-9:.1f} GB/s")
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(hp, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"H2D pinned: {8*n*n/dt*1e-9:.1f} GB/s")
