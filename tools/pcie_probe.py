"""Developer probe: pinned host<->device copy bandwidth, each direction alone and both at once (the floor of bench.py's e2e)."""
import torch

n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device='cuda')
d_b = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    h2d(); d2h()


for name, fn in (('h2d', h2d), ('d2h', d2h), ('both', both)):
    ms = timed(fn)
    print('%-5s %.2f ms per GiB  -> %.1f GB/s per direction' % (name, ms, n / ms / 1e6))

# pageable host memory (what a numpy caller hands to the host entry points)
import time
import numpy as np
a = np.random.default_rng(0).random(n // 8)
t = torch.from_numpy(a)
for _ in range(2):
    t0 = time.perf_counter(); d = t.cuda(); torch.cuda.synchronize(); t1 = time.perf_counter()
print('pageable h2d %.1f ms per GiB -> %.1f GB/s' % ((t1 - t0) * 1e3, n / (t1 - t0) / 1e9))
for _ in range(2):
    t0 = time.perf_counter(); b = d.cpu(); t1 = time.perf_counter()
print('pageable d2h %.1f ms per GiB -> %.1f GB/s' % ((t1 - t0) * 1e3, n / (t1 - t0) / 1e9))
dst = np.empty_like(a)
t0 = time.perf_counter(); np.copyto(dst, a); t1 = time.perf_counter()
print('host memcpy (1 thread) %.1f ms per GiB -> %.1f GB/s' % ((t1 - t0) * 1e3, n / (t1 - t0) / 1e9))
import os
print('cpus', os.cpu_count())
