"""Write-only and copy bandwidth of the box (torch ops), to put the feature kernel's store rate in context."""
import torch
x = torch.empty(1 << 30, dtype=torch.float32, device="cuda")      # 4 GiB
y = torch.empty_like(x)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ms = t(lambda: x.fill_(1.0)); print(f"fill_   write-only : {x.numel()*4/ms/1e6:8.1f} GB/s")
ms = t(lambda: x.zero_());    print(f"zero_   (memset)   : {x.numel()*4/ms/1e6:8.1f} GB/s")
ms = t(lambda: y.copy_(x));   print(f"copy_   read+write : {2*x.numel()*4/ms/1e6:8.1f} GB/s")
ms = t(lambda: x.sum());      print(f"sum     read-only  : {x.numel()*4/ms/1e6:8.1f} GB/s")
