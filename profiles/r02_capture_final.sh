#!/bin/bash
# usage (on the GPU box, via gpurun): profiles/r02_capture_final.sh
# Final round-2 evidence after the scalar-mask / lean-arithmetic change: launch list and DRAM traffic of the default bench
# command, one `ncu --set full` capture of the single-segment and of the double-assembly kernel.  Each capture runs only
# after the same command has exited 0 without ncu.
T="timeout -s KILL 500"
B="python bench.py --no-extras --no-cpu-baseline"
$B --steps 2 --warmup 1 > gpurun_out/r02f_plain.log 2>&1 || exit 1
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launch_list.csv $B --steps 2 --warmup 1 > gpurun_out/r02f_launch_list.log 2>&1
$T ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:br2_step_tile_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02_dram_2000step_launch.csv $B --steps 1 --warmup 1 > gpurun_out/r02f_dram.log 2>&1
cap() {  # tag, launches to skip, warp-steps of the launch, bench args
  tag=$1; skip=$2; ws=$3; shift 3
  $B --steps 1 --warmup 1 --substeps 100 "$@" > gpurun_out/${tag}_plain.log 2>&1 || return 1
  $T ncu --set full --clock-control none --import-source on -k regex:br2_step_tile_kernel -s $skip -c 1 -f -o gpurun_out/$tag $B --steps 1 --warmup 1 --substeps 100 "$@" > gpurun_out/${tag}_ncu.log 2>&1
  python profiles/ncu_summary.py gpurun_out/$tag.ncu-rep $ws > gpurun_out/$tag.txt 2>&1
  rm -f gpurun_out/$tag.ncu-rep
}
cap r02_ncu_tile_single 1 819200                                              # 4096 instances x 2 warps x 100 steps
cap r02_ncu_tile_double 1 1638400 --workload double --batch 4096              # x 4 warps
