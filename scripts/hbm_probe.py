"""What the HBM delivers for the three traffic mixes of the level-0 layers: write-only (torch fill_),
copy (read + write, the figure MEASURED_PEAKS.json records) and read-only (torch sum), 2 GiB buffers,
CUDA events, best of 10.  Context for the write-heavy kernels (first conv: 4 B/px in, 32 B/px out).
Careful with fill_: torch's fill kernel stores one ELEMENT-vector per thread, so 1- and 2-byte dtypes
are bound by the kernel (3.3-3.9 TB/s) and only the float32 fill shows what the memory takes."""
import torch

n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device='cuda')
b = torch.empty(n, dtype=torch.bfloat16, device='cuda')
a.fill_(1.0)


def best(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return min(t)


gb = 2 * n / 1e9
print('write-only (bf16 fill_)  %7.1f GB/s' % (gb / (best(lambda: b.fill_(2.0)) * 1e-3)))
print('write-only (f32 fill_)   %7.1f GB/s' % (gb / (best(lambda: b.view(torch.float32).fill_(3.0)) * 1e-3)))
print('write-only (f32 arange)  %7.1f GB/s' % (gb / (best(lambda: torch.arange(n // 2, out=b.view(torch.float32))) * 1e-3)))
print('memset (zero_)      %7.1f GB/s' % (gb / (best(lambda: b.zero_()) * 1e-3)))
print('copy (read+write)   %7.1f GB/s' % (2 * gb / (best(lambda: b.copy_(a)) * 1e-3)))
print('read-only (f32 sum) %7.1f GB/s' % (gb / (best(lambda: a.view(torch.float32).sum()) * 1e-3)))
print('read-only (f32 max) %7.1f GB/s' % (gb / (best(lambda: a.view(torch.float32).max()) * 1e-3)))
