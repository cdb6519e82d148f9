"""Host->device copy bandwidth of one 373 MB pinned arena: one cudaMemcpyAsync vs 2 / 4 concurrent chunks on separate streams."""
import torch, time
n = 373_294_784
host = torch.empty(n, dtype=torch.uint8).pin_memory()
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
for parts in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    step = (n // parts + 255) // 256 * 256
    def go():
        for i, s in enumerate(streams):
            a, b = i * step, min(n, (i + 1) * step)
            with torch.cuda.stream(s):
                dev[a:b].copy_(host[a:b], non_blocking=True)
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        go()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print(parts, "chunks:", round(n / dt / 1e9, 2), "GB/s", round(dt * 1e3, 3), "ms")
