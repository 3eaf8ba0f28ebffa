#!/bin/bash
# Regenerates the round's evidence on a B200 box (run through gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash profiles/refresh.sh'
# then, back in the container:  python profiles/launch_summary.py gpurun_out/launches_raw.csv
#                               ncu -i gpurun_out/full.ncu-rep --page raw --csv > gpurun_out/full_raw.csv
#                               python profiles/ncu_tools.py raw gpurun_out/full_raw.csv > profiles/r1_ncu_full_summary.txt
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log || exit 1
python bench.py --dense --no-cpu-baseline > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.log
python bench.py --impl reference > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.log
CMD="python bench.py --tiles 16 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_raw.csv $CMD > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'nms_peaks|assign_kernel|apply_lut' --launch-skip 9 -c 3 -f \
    -o gpurun_out/full $CMD > gpurun_out/ncu_full.log 2>&1
python profiles/kernel_times.py > gpurun_out/kernel_times.json 2> gpurun_out/kernel_times.log
python bench_stack.py --match > gpurun_out/stack_n1_512.json 2> gpurun_out/stack_n1_512.log
tail -c 400 gpurun_out/bench_n1.json
