#!/usr/bin/env python
"""Turn the raw CSV pages written by ncu_colour.sh into the committed summaries:

    python profiles/recipes/summarise_ncu.py gpurun_out/r02f r02 [pairs_per_camera]

  profiles/<tag>_ncu_colour.md      per-kernel table (duration, instructions, issue / FMA / XU utilisation, occupancy, DRAM)
  profiles/<tag>_calibration.json   constants bench.py reads: executed warp instructions per (warp, Gaussian) iteration of the
                                    blend kernel (from the source page), DRAM bytes per launch of every kernel
"""
import csv
import json
import os
import sys

prefix, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rows = list(csv.reader(open(prefix + "_colour_raw.csv")))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = ["preprocess_kernel", "tree_kernel", "ms_count", "ms_scan_partial", "ms_scan_blocks", "ms_scan_apply", "ms_scatter",
        "blend_kernel", "accumulate_kernel", "tiles_kernel", "blend_tiles"]
COLS = [("gpu__time_duration.sum", "time us"), ("smsp__inst_executed.sum", "warp inst"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr % (elapsed)"),
        ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
table, dram = [], {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    short = next((k for k in KEYS if k in name), name[:40])
    vals = []
    for k, _ in COLS:
        v = r[idx[k]] if k in idx else ""
        try:
            v = f"{float(v):.4g}"
        except ValueError:
            pass
        vals.append(v)
    table.append((short, vals))
    try:
        dram[short.replace("_kernel", "")] = (float(r[idx["dram__bytes_read.sum"]]) + float(r[idx["dram__bytes_write.sum"]])) * 1e6
    except Exception:
        pass
# blend: instructions per (warp, Gaussian) iteration = total instructions / executions of the loop's first instruction
ipi = None
src = prefix + "_blend_kernel_src.csv"
if os.path.exists(src):
    srows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(srows) if r and r[0] == "Address"][0]
    h = srows[hi]
    c = h.index("Instructions Executed")
    ex = []
    for r in srows[hi + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break
        ex.append(int(r[c]))
    total = sum(ex)
    loop = max(ex)  # every instruction of the inner loop executes once per (warp, Gaussian)
    ipi = total / loop
with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_colour.md"), "w") as f:
    f.write(f"# ncu --set full, C3 (3 M Gaussians, 1280x720, SH 3), one camera — `{os.path.basename(prefix)}`\n\n")
    f.write("Recipe: `profiles/recipes/ncu_colour.sh c3 <tag>` (colour_step.py, 2 cameras, kernels of the 2nd).  Times under ncu are "
            "cold-cache and serialised: compare shares, not absolutes (bench.py reports the CUDA-event times).\n\n")
    f.write("| kernel | " + " | ".join(c for _, c in COLS) + " |\n|---|" + "---|" * len(COLS) + "\n")
    for short, vals in table:
        f.write(f"| {short} | " + " | ".join(vals) + " |\n")
    if ipi:
        f.write(f"\nblend_kernel: {total} warp instructions, inner loop executed {loop} times -> **{ipi:.1f} instructions per "
                "(warp, Gaussian) iteration** (= 128 (pixel, Gaussian) pairs); the loop body itself is 43 SASS instructions "
                "(3 LDS, 5 uniform-datapath, 4 MUFU.EX2, 16 packed FP32x2, 4 scalar FP32, 6 FMNMX, vote + branch).\n")
json.dump({"source": os.path.basename(prefix), "blend_inst_per_warp_gaussian": ipi, "dram_bytes_per_launch": dram},
          open(os.path.join(ROOT, "profiles", f"{tag}_calibration.json"), "w"), indent=1)
print("wrote", tag, "ipi", ipi)
