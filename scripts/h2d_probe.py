"""Pinned H2D bandwidth at the bench's pack sizes (development aid)."""
import torch, time
dev = torch.device("cuda", 0)
for nbytes, name in ((36864000, "f32 pack"), (9216000, "u8 pack")):
    hs = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(5)]
    d = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    st = torch.cuda.Stream(dev)
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(st):
            for k in range(50):
                d[k & 1].copy_(hs[k % 5], non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 50
    print(f"{name}: {nbytes/dt/1e9:.1f} GB/s ({dt*1e6:.0f} us per copy)")
