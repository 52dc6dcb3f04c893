"""L2 residency probe: effective bandwidth of repeated streaming passes (in-place add, and ping-pong a->b->a)
over buffers of growing size.  Above the HBM peak => the pass (reads AND writes) stays in L2."""
import torch
def bench(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
for mb in (8, 16, 24, 32, 48, 64, 80, 96, 112, 128, 160, 192, 256, 1024):
    n = mb * (1 << 20) // 4
    a = torch.zeros(n, device="cuda"); b = torch.zeros(n, device="cuda")
    t1 = bench(lambda: a.add_(1.0))
    def pp():
        torch.add(a, 1.0, out=b); torch.add(b, 1.0, out=a)
    t2 = bench(pp) / 2
    print(f"{mb:5d} MB  in-place: {2 * n * 4 / t1 / 1e12:6.2f} TB/s   ping-pong (footprint {2 * mb} MB): {2 * n * 4 / t2 / 1e12:6.2f} TB/s", flush=True)
