#!/usr/bin/env python
"""Summarise an ncu launch list of the bench step (tools/_ncu_r2.sh launches): keeps the LAST `--steps` steady-state steps of the
capture (a step starts at a launch of the fused warp+variance kernel's first producer, pack of conv_0_0 ... simply: at every
`image_to_rows8_kernel` / first kernel after the previous step's optimiser), writes them as CSV and prints a markdown table of
kernels by total time.  Usage: python tools/launch_summary.py gpurun_out/r02_bench_launches_all.csv profiles/r02_bench_launches.csv"""
import csv
import sys
from collections import defaultdict

src, dst = sys.argv[1], sys.argv[2]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
marker = sys.argv[4] if len(sys.argv) > 4 else "image_to_rows8_kernel"
lines = open(src, newline="").read().splitlines()
head = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
rows = list(csv.reader(lines[head:]))
hdr, rows = rows[0], [r for r in rows[1:] if len(r) == len(rows[0])]
k_name, k_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
starts = [i for i, r in enumerate(rows) if marker in r[k_name]]
if len(starts) < steps + 1:
    sys.exit(f"only {len(starts)} step markers ({marker}) in {src}")
per_step = starts[-1] - starts[-2]
first, last = starts[-steps - 1], starts[-1]      # complete steps only: the capture may end inside the last one
keep = rows[first:last]
with open(dst, "w", newline="") as f:
    w = csv.writer(f, quoting=csv.QUOTE_ALL)
    w.writerow(hdr)
    w.writerows(keep)
tot = defaultdict(lambda: [0, 0.0])
for r in keep:
    name = r[k_name].split("(")[0].replace("(anonymous namespace)", "<unnamed>")[:84]
    tot[name][0] += 1
    tot[name][1] += float(r[k_val].replace(",", "")) / 1e6
total = sum(v[1] for v in tot.values())
print(f"{len(keep)} launches = {steps} steps of {per_step}; {total:.2f} ms under ncu (cold cache, serialised)\n")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"| `{name}` | {n} | {ms:.2f} | {100 * ms / total:.1f} % |")
lib = sum(ms for name, (n, ms) in tot.items() if any(t in name for t in ("cudnn", "cutlass", "xmma", "convolve", "implicit_gemm", "gemm")))
print(f"\nlibrary GEMM / convolution kernels: {lib:.2f} ms ({100 * lib / total:.1f} %)")
