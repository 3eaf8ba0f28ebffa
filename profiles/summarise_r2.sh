#!/bin/bash
# Turns what profiles/refresh_r2.sh brought back in gpurun_out/ into the tracked summaries under profiles/.
set -x
cp gpurun_out/r2_bench_n1.json profiles/r2_bench_n1.json
cp gpurun_out/r2_bench_reference_arm.json profiles/r2_bench_reference_arm.json
cp gpurun_out/r2_stack_n1.json profiles/r2_stack_n1.json
python profiles/ncu_launch_table.py gpurun_out/r2_launches_raw.csv profiles/r2_launches.csv \
    --title "python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dense --no-deep --no-cnn (config 2 batch, e2e tiles, config 3 stack)"
python profiles/ncu_launch_table.py gpurun_out/r2_all_raw.csv profiles/r2_all_kernels.csv \
    --title "python profiles/all_kernels.py (every kernel of the library at BASELINE sizes)"
ncu -i gpurun_out/r2_full_tiles.ncu-rep --page raw --csv > gpurun_out/r2_full_tiles_raw.csv
ncu -i gpurun_out/r2_full_stack.ncu-rep --page raw --csv > gpurun_out/r2_full_stack_raw.csv
python profiles/ncu_tools.py raw gpurun_out/r2_full_tiles_raw.csv > profiles/r2_ncu_full_tiles.txt
python profiles/ncu_tools.py raw gpurun_out/r2_full_stack_raw.csv > profiles/r2_ncu_full_stack.txt
python profiles/traffic_r2.py
