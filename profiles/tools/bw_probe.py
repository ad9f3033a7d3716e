"""HBM probe: pure-write, pure-read and copy bandwidth (torch ops), to tell which direction bounds a kernel."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
def t(f, reps=10):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
tw = t(lambda: a.fill_(1.0))
tr = t(lambda: a.sum())
tc = t(lambda: b.copy_(a))
print(f"write {4*n/tw/1e12:.2f} TB/s   read {4*n/tr/1e12:.2f} TB/s   copy {8*n/tc/1e12:.2f} TB/s (r+w)")
