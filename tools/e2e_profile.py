"""Where does the end-to-end time go: pinned allocation, H2D, transposes, D2H."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xanthos_b200 import _cuda as C
n, m = 67420, 360
def t(fn, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
t0 = time.perf_counter(); a = torch.empty((n, m), dtype=torch.float64, pin_memory=True); print('first pinned alloc 194MB: %.1f ms' % ((time.perf_counter() - t0) * 1e3))
a.normal_()
arr = a.numpy()
print('pinned alloc (cached?) %.1f ms' % t(lambda: torch.empty((n, m), dtype=torch.float64, pin_memory=True)))
print('H2D pinned 194MB %.2f ms' % t(lambda: torch.from_numpy(arr).to('cuda', non_blocking=True)))
pag = np.array(arr)
print('H2D pageable 194MB %.2f ms' % t(lambda: torch.from_numpy(pag).to('cuda', non_blocking=True)))
f = C.Field.from_host(arr)
print('from_host (H2D+transpose) %.2f ms' % t(lambda: C.Field.from_host(arr)))
print('to_device_cell_major (transpose) %.2f ms' % t(lambda: f.to_device_cell_major()))
print('to_host (transpose + pinned alloc + D2H + sync) %.2f ms' % t(lambda: f.to_host()))
dev = f.to_device_cell_major(); host = torch.empty((n, m), dtype=torch.float64, pin_memory=True)
print('D2H into existing pinned 194MB %.2f ms' % t(lambda: host.copy_(dev, non_blocking=True)))
keep = []
def th():
    keep.append(f.to_host())
print('to_host keeping results alive (fresh pinned each time) %.2f ms' % t(th, reps=6))
