"""HBM read-only / write-only / copy bandwidth with plain torch ops (context for the write-heavy kernels' plateaus)."""
import torch

n = 1 << 30                                  # 1 Gi floats = 4 GiB
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")
x.fill_(1.0)


def timed(fn, bytes_moved, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return bytes_moved * reps / (a.elapsed_time(b) * 1e-3) / 1e12


print(f"write only (fill_) : {timed(lambda: y.fill_(2.0), 4 * n):.2f} TB/s")
print(f"read only  (sum)   : {timed(lambda: x.sum(), 4 * n):.2f} TB/s")
print(f"copy (read + write): {timed(lambda: y.copy_(x), 8 * n):.2f} TB/s total")
