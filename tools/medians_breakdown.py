"""Where the time of CudaEngine.cluster_medians goes (10M x 32 x 8 fp32): CUDA-event time of every stage."""
import sys

import torch

sys.path.insert(0, ".")
import heat_b200 as hb  # noqa: E402
from heat_b200.engine import _DT, _ptr, _stream  # noqa: E402
from heat_b200._lib import check  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
n, d, k = 10_000_000, 32, 8
cent = 4 * torch.randn(k, d, device=dev, generator=g)
lab = torch.randint(0, k, (n,), device=dev, generator=g)
x = cent[lab] + torch.randn(n, d, device=dev, generator=g)
eng = hb.engine.get_engine(dev)
lib, h, st, dt = eng.lib, eng.h, _stream(dev), _DT[x.dtype]


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


keep = prefix = remaining = hist = None
for rep in range(2):
    del keep, prefix, remaining, hist
    marks = [("start", ev())]
    keep = torch.empty(n, dtype=torch.uint8, device=dev)
    check(lib.hk_row_keep(h, _ptr(x), n, d, d, dt, _ptr(keep), st), "keep")
    marks.append(("row_keep", ev()))
    prefix = torch.zeros((2, k, d), dtype=torch.int64, device=dev)
    remaining = torch.zeros((2, k, d), dtype=torch.int64, device=dev)
    hist = torch.empty((2, k, d, 256), dtype=torch.int64, device=dev)
    marks.append(("alloc", ev()))
    for p in range(4):
        hist.zero_()
        marks.append((f"zero{p}", ev()))
        check(lib.hk_select_hist(h, _ptr(x), n, d, d, dt, _ptr(lab), _ptr(keep), k, _ptr(prefix), p, _ptr(hist), st), "hist")
        marks.append((f"hist{p}", ev()))
        if p == 0:
            hist[1].copy_(hist[0])
            counts = hist[0, :, 0, :].sum(dim=1)
            remaining[0] = ((counts - 1).clamp(min=0) // 2).view(k, 1)
            remaining[1] = (counts // 2).view(k, 1)
            marks.append(("torch_counts", ev()))
        check(lib.hk_select_step(h, _ptr(hist), _ptr(remaining), _ptr(prefix), k, d, st), "step")
        marks.append((f"step{p}", ev()))
    torch.cuda.synchronize()
    if rep == 1:
        for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
            print(f"{name:14s} {a.elapsed_time(b):8.3f} ms")
        print("total", marks[0][1].elapsed_time(marks[-1][1]))
