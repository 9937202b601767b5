T="timeout -s KILL 500"
B="python bench.py --no-extras --no-cpu-baseline"
$B --steps 1 --warmup 1 --substeps 100 > gpurun_out/r3_single_plain.log 2>&1 || exit 1
$T ncu --set full --clock-control none --import-source on -k regex:br2_step_tile_kernel -s 1 -c 1 -f -o gpurun_out/r3_single $B --steps 1 --warmup 1 --substeps 100 > gpurun_out/r3_single_ncu.log 2>&1
ls -la gpurun_out/r3_single.ncu-rep
