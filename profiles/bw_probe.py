"""What HBM bandwidth does a stream with K1's read:write mix reach?  (MEASURED_PEAKS.json is a 1:1 copy.)
K1 moves 8*(2N + n_theta) bytes in and 8*(N + nnz) out per system: 24 % reads, 76 % writes."""
import torch
dev = torch.device("cuda", 0)
n = 1 << 28          # 2 GiB of float64
a = torch.empty(n, dtype=torch.float64, device=dev); b = torch.empty(n, dtype=torch.float64, device=dev)
c = torch.empty(3, n // 4, dtype=torch.float64, device=dev)


def timeit(f, bytes_, reps=10):
    for _ in range(3):
        f()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return bytes_ / best / 1e6


print("copy 1:1     GB/s", round(timeit(lambda: b.copy_(a), 2 * n * 8)))
print("fill  0:1    GB/s", round(timeit(lambda: a.zero_(), n * 8)))
print("read  1:0    GB/s", round(timeit(lambda: a.sum(), n * 8)))
src = a[: n // 4].view(1, n // 4)
print("expand 1:3   GB/s", round(timeit(lambda: c.copy_(src.expand(3, n // 4)), 4 * (n // 4) * 8)))
