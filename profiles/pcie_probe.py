"""Host<->device copy bandwidth from pinned memory on this box: H2D alone, D2H alone, both directions at once
(two streams), and per chunk size -- the bounds of bench.py's e2e leg.   python profiles/pcie_probe.py"""
import json, time, torch
dev = torch.device('cuda', 0)
out = {}
for mb in (1, 4, 16, 64):
    n = mb * 1024 * 1024
    h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d1 = torch.empty(n, dtype=torch.uint8, device=dev); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(h2d, d2h, reps=20):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps
    run(1, 1, 3)
    a, b, c = run(1, 0), run(0, 1), run(1, 1)
    out[mb] = {'h2d_GBs': n / a / 1e9, 'd2h_GBs': n / b / 1e9, 'both_each_GBs': n / c / 1e9}
    print(mb, 'MB', json.dumps(out[mb]), flush=True)
# pageable destination (what tensor.cpu().numpy() does) vs pinned
n = 48 * 1000 * 1000
d = torch.empty(n, dtype=torch.uint8, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter(); x = d.cpu(); t1 = time.perf_counter()
print('48 MB d2h pageable: %.2f ms' % ((t1 - t0) * 1e3))
t0 = time.perf_counter(); h = torch.empty(n, dtype=torch.uint8, pin_memory=True); t1 = time.perf_counter()
print('pinned alloc 48 MB first: %.2f ms' % ((t1 - t0) * 1e3))
del h
t0 = time.perf_counter(); h = torch.empty(n, dtype=torch.uint8, pin_memory=True); t1 = time.perf_counter()
print('pinned alloc 48 MB cached: %.2f ms' % ((t1 - t0) * 1e3))
