#!/bin/bash
# Regenerates round 2's evidence on a B200 box (run through gpurun from the repo root):
#   gpurun --timeout 2400 -- 'bash profiles/refresh_r2.sh'
# then, back in the container:  bash profiles/summarise_r2.sh
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.log || exit 1
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.log
python bench_stack.py --profile --match > gpurun_out/r2_stack_n1.json 2> gpurun_out/r2_stack_n1.log
python profiles/all_kernels.py > gpurun_out/all_kernels.log 2>&1 || { tail -5 gpurun_out/all_kernels.log; exit 1; }
python bench_ortho.py > gpurun_out/r2_ortho_1024.json 2> gpurun_out/r2_ortho_1024.log
# per-launch list of the bench command (the roofline's kernel shares) and of every kernel of the library
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dense --no-deep --no-cnn"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/r2_launches_raw.csv $CMD > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/r2_all_raw.csv python profiles/all_kernels.py > gpurun_out/ncu_all.log 2>&1
# full captures: the three config-2 kernels over the whole batch, and the stack path's kernels over a 128-slice sub-block
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'nms_peaks|assign_kernel|apply_lut' --launch-skip 9 -c 3 -f \
    -o gpurun_out/r2_full_tiles $CMD > gpurun_out/ncu_full_tiles.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'median_chain|merge_lean|rle_block_emit|assign_kernel|nms_peaks|rle_block_lists|rle_block_union|rle_block_assign' --launch-skip 8 -c 8 -f \
    -o gpurun_out/r2_full_stack python bench_stack.py --depth 128 --repeat 1 > gpurun_out/ncu_full_stack.log 2>&1
ls -la gpurun_out | tail -20
