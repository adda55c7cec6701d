#!/bin/bash
# ncu capture of the renderer_type=cuda back-end (csrc/s7_tiles.cu) on a bench.py workload (run under gpurun, 1 GPU):
#   bash profiles/recipes/ncu_tiles.sh c4 r02
WL=${1:-c4}; TAG=${2:-r02}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on \
    -k regex:'blend_tiles_kernel|preprocess_tiles_kernel|ms_count|ms_scatter' -s 4 -c 4 \
    -o gpurun_out/${TAG}_tiles -f python profiles/recipes/colour_step.py --workload $WL --cams 2 --renderer cuda --surface > gpurun_out/${TAG}_tiles.log 2>&1
ncu -i gpurun_out/${TAG}_tiles.ncu-rep --page raw --csv > gpurun_out/${TAG}_tiles_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_tiles.ncu-rep --page source --csv -k regex:blend_tiles_kernel > gpurun_out/${TAG}_blend_tiles_src.csv 2>/dev/null
