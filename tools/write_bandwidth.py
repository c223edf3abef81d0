"""Write-only HBM bandwidth of the box (memset / fill of 2 GiB, best of 10): the roofline of a kernel that mostly WRITES, such as
the all-pairs correlation.  MEASURED_PEAKS.json only holds the copy figure (reads + writes)."""
import torch
x = torch.empty(2 << 30, dtype=torch.uint8, device='cuda')
y = torch.empty(1 << 30, dtype=torch.float16, device='cuda')
for name, fn, nbytes in (('memset 2 GiB', lambda: x.zero_(), 2 << 30), ('fill fp16 2 GiB', lambda: y.fill_(1.0), 2 << 30)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f'{name}: {nbytes / best / 1e6:.0f} GB/s written (best of 10)')
