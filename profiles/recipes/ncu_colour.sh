#!/bin/bash
# ncu captures of the colour pipeline on a bench.py workload (run under gpurun, 1 GPU):
#   bash profiles/recipes/ncu_colour.sh c3 r02
# writes gpurun_out/<tag>_launches.csv (every launch with its device time, 3 cameras) and
# gpurun_out/<tag>_colour.ncu-rep (--set full of the hand-written kernels of the 2nd camera) + raw / source CSV pages.
WL=${1:-c3}; TAG=${2:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python profiles/recipes/colour_step.py --workload $WL --cams 3 > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'preprocess_kernel|tree_kernel|ms_count|ms_scan|ms_scatter|blend_kernel|accumulate_kernel' -s 9 -c 9 \
    -o gpurun_out/${TAG}_colour -f python profiles/recipes/colour_step.py --workload $WL --cams 2 > gpurun_out/${TAG}_colour.log 2>&1
ncu -i gpurun_out/${TAG}_colour.ncu-rep --page raw --csv > gpurun_out/${TAG}_colour_raw.csv 2>/dev/null
for K in blend_kernel preprocess_kernel ms_scatter ms_count; do
  ncu -i gpurun_out/${TAG}_colour.ncu-rep --page source --csv -k regex:$K > gpurun_out/${TAG}_${K}_src.csv 2>/dev/null
done
