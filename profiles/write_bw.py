"""Reference points for the roofline: pure-write (fill) and copy bandwidth of this GPU, CUDA events, best of 10."""
import json
import torch

dev = torch.device("cuda:0")
n = 1 << 29  # 2 GiB of float32
x = torch.empty(n, dtype=torch.float32, device=dev)
y = torch.empty(n, dtype=torch.float32, device=dev)


def best(fn, reps=10):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


for _ in range(3):
    x.fill_(1.0); y.copy_(x)
fill_ms = best(lambda: x.fill_(2.0))
copy_ms = best(lambda: y.copy_(x))
print(json.dumps({"fill_write_GBs": 4 * n / fill_ms / 1e6, "copy_read_plus_write_GBs": 8 * n / copy_ms / 1e6,
                  "fill_ms": fill_ms, "copy_ms": copy_ms}))
