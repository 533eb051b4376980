"""PCIe yardsticks with plain torch copies (pinned buffers): each direction alone, both at once."""
import time, torch
n = 1 << 30
h_a = torch.empty(n, dtype=torch.uint8).pin_memory(); h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
def both(): h2d(); d2h()
def both_sliced(k=16):
    m = n // k
    for i in range(k):
        with torch.cuda.stream(s1): d_a[i*m:(i+1)*m].copy_(h_a[i*m:(i+1)*m], non_blocking=True)
        with torch.cuda.stream(s2): h_b[i*m:(i+1)*m].copy_(d_b[i*m:(i+1)*m], non_blocking=True)
for name, fn in (("H2D alone", h2d), ("D2H alone", d2h), ("both at once", both), ("both, 16 slices", both_sliced)):
    t = run(fn)
    per = n / t / 1e9
    print(f"{name}: {t*1e3:.1f} ms for 1 GiB per direction = {per:.1f} GB/s per direction", flush=True)
