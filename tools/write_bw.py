#!/usr/bin/env python
"""Pure-write and copy bandwidth of this GPU (context for K1's roofline: K1 writes 98 % of its bytes)."""
import json
import torch

dev = "cuda:0"
n = 512 * 1024 * 1024            # 2 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)


def t(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts)


fill = t(lambda: a.fill_(1.0))
zero = t(lambda: a.zero_())
copy = t(lambda: b.copy_(a))
read = t(lambda: a.sum())
print(json.dumps({"fill_GBps": 4 * n / fill / 1e6, "memset_GBps": 4 * n / zero / 1e6, "copy_GBps_read_plus_write": 8 * n / copy / 1e6,
                  "read_GBps(sum)": 4 * n / read / 1e6}))
