"""Write-only HBM bandwidth on this GPU (context for the cdist roofline, whose traffic is all writes):
torch fill_ and cudaMemsetAsync over a 16 GiB buffer, CUDA events, best and median of 10."""
import json
import torch

dev = torch.device("cuda", 0)
n = 16 * (1 << 30)
buf = torch.empty(n, dtype=torch.uint8, device=dev)
f32 = buf.view(torch.float32)
res = {}
for name, fn in (("fill_f32", lambda: f32.fill_(1.5)), ("zero_u8", lambda: buf.zero_()),
                 ("copy_8g", lambda: buf[: n // 2].copy_(buf[n // 2:]))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    nbytes = n if name != "copy_8g" else n  # copy: n/2 read + n/2 written
    res[name] = {"best_gbs": nbytes / ts[0] / 1e6, "median_gbs": nbytes / ts[len(ts) // 2] / 1e6}
print(json.dumps(res))
