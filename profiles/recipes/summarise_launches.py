#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum) -> per-kernel totals and shares, next to the CUDA-event times of a
bench.py line (kernel_ms_per_step) so the SHARES can be compared (absolute times under ncu are cold-cache and serialised).

    python profiles/recipes/summarise_launches.py gpurun_out/r02h_launches.csv profiles/r02_bench_c3_final.json r02
"""
import collections
import csv
import json
import re
import sys

src, bench_json, tag = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
tot = collections.OrderedDict()
cnt = collections.Counter()
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]]
    m = re.search(r"(?:<unnamed>::)?([A-Za-z_0-9]+)(?:<[^(]*>)?\(", name)
    short = m.group(1) if m else name[:40]
    if "cub" in name or "DeviceRadixSort" in name or "DeviceScan" in name:
        short = "cub (radix sort / scan)"
    elif "at::" in name or "elementwise" in name or "vectorized" in name:
        short = "torch elementwise / fill / copy"
    v = float(r[ix["Metric Value"]])
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    tot[short] = tot.get(short, 0.0) + us
    cnt[short] += 1
total = sum(tot.values())
line = json.loads(open(bench_json).read().strip().splitlines()[-1])
km = line.get("kernel_ms_per_step", {})
group = {"preprocess_kernel": "preprocess", "tree_kernel": "build_tree", "blend_kernel": "blend",
         "accumulate_kernel": "accumulate", "compose_kernel": "compose_image"}
ms_keys = ("ms_count_kernel", "ms_scan_partial_kernel", "ms_scan_blocks_kernel", "ms_scan_apply_kernel", "ms_scatter_kernel")
colour_total = sum(km.get(k, 0.0) for k in ("preprocess", "depth_sort", "build_tree", "multisplit", "blend", "accumulate",
                                             "compose_image"))
out = [f"# ncu launch list, C3, 3 cameras — `{tag}`", "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none` (recipe `ncu_colour.sh`), every launch of the process.",
       "Shares of the per-camera colour pipeline under ncu (cold cache, serialised) next to the CUDA-event shares of the",
       f"bench line `{bench_json.split('/')[-1]}` (`kernel_ms_per_step`, profiled step with one frame slot).", "",
       "| kernel | launches | total us (ncu) | share of colour kernels (ncu) | share (bench, CUDA events) |", "|---|---|---|---|---|"]
colour_names = list(group) + list(ms_keys) + ["cub (radix sort / scan)"]
ncu_colour_total = sum(v for k, v in tot.items() if k in colour_names)
ms_ncu = sum(tot.get(k, 0.0) for k in ms_keys)
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if k in ms_keys:
        continue
    share = f"{100 * v / ncu_colour_total:.1f} %" if k in colour_names else "—"
    b = group.get(k, "depth_sort" if k.startswith("cub") else None)
    bshare = f"{100 * km[b] / colour_total:.1f} %" if b in km and colour_total else "—"
    out.append(f"| {k} | {cnt[k]} | {v:.0f} | {share} | {bshare} |")
out.append(f"| multisplit (count + 3 scans + scatter) | {sum(cnt[k] for k in ms_keys)} | {ms_ncu:.0f} | "
           f"{100 * ms_ncu / ncu_colour_total:.1f} % | {100 * km.get('multisplit', 0) / colour_total:.1f} % |")
out += ["", f"All launches: {sum(cnt.values())}, {total / 1e3:.2f} ms under ncu; colour kernels {ncu_colour_total / 1e3:.2f} ms "
        f"for 3 cameras = {ncu_colour_total / 3e3:.2f} ms per camera (bench: {colour_total / 200:.2f} ms per camera)."]
open(f"profiles/{tag}_launches_c3.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
